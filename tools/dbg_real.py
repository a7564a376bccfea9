"""Scratch: the real exp1 pair (golden fixture) through the three search modes, per iteration count."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
from oracle import oracle
z = np.load(os.path.join(ROOT, "tests", "golden", "exp1_depth_q4.npz"))
cam = synth.Camera(fx=525.0 / 4, fy=525.0 / 4, cx=319.5 / 4, cy=235.5 / 4, factor=1000.0, width=160, height=120)
ctx = s3d.Context(0)
cs, ct = ctx.from_depth(z["d1"], cam, 7.0), ctx.from_depth(z["d2"], cam, 7.0)
src, tgt = oracle.backproject(z["d1"], cam, 7.0), oracle.backproject(z["d2"], cam, 7.0)
planes = ct.segment_planes(_abi.plane_params())
got = ct.download(xyz=False, normals=True, labels=True)
nrm4 = np.c_[got["normals"], (got["labels"] >= 0).astype(np.float32)].astype(np.float32)
print("n", len(src), len(tgt), "planes", len(planes), "valid normals", int((got["labels"] >= 0).sum()))
for gate in (0.3, 0.0):
    for it in (1, 2, 3, 10):
        o = oracle.icp(src, tgt, nrm4, params=_abi.icp_params(it, max_corr_dist=gate), want_nn=True)
        line = f"gate={gate} it={it} oracle status={o['status']} inl={o['inliers']}"
        for name, mode in (("tile", _abi.SEARCH_GRID), ("lane", _abi.SEARCH_GRID_LANE), ("brute", _abi.SEARCH_BRUTE)):
            r = ctx.register(cs, ct, None, _abi.icp_params(it, max_corr_dist=gate, search=mode))
            nn = ctx.last_correspondences(len(src))
            line += f" | {name} status={r['status']} inl={r['inliers']} iters={r['iterations']} nn_mismatch={(nn != o['nn']).sum()}"
        print(line, flush=True)
