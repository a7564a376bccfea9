"""Generates the committed golden fixtures of tests/golden/ (run in the build container, where /root/reference
exists; the GPU box never reads /root/reference).

1. icp_golden.json -- known-answer vectors for the registration path from an implementation that is independent of
   both the CUDA code and the C oracle: tests/ref_numpy.py (scipy.spatial.cKDTree + numpy.linalg, float64).  Inputs
   are regenerated from seeds by slam3d_gx_b200.synth; their sha256 is stored so that generator drift is detected.
2. exp1_depth_q4.npz -- the reference's only real input for this path, the Kinect depth pair
   /root/reference/data/exp1/dep/{1,2}.png, subsampled 4x (every 4th row/column -> 160x120 uint16, 28 %/23 % holes),
   with the intrinsics of reference src/convert2PCD.cpp:19-23 scaled accordingly.  No ground-truth pose exists
   for this pair; it is a robustness / agreement input.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from slam3d_gx_b200 import synth          # noqa: E402
from ref_numpy import icp_numpy           # noqa: E402

CASES = [
    dict(name="s1_plane_10", pair=0, scale=0.25, est="plane", iters=10, kw={}),
    dict(name="s1_plane_30", pair=1, scale=0.25, est="plane", iters=30, kw={}),
    dict(name="s1_svd_10", pair=2, scale=0.25, est="svd", iters=10, kw={}),
    dict(name="s1_quant_holes", pair=3, scale=0.25, est="plane", iters=10, kw=dict(quantize=True, holes=0.25)),
    dict(name="s1_gate", pair=4, scale=0.25, est="plane", iters=8, kw={}, max_corr_dist=0.05),
    dict(name="s1_big_motion", pair=5, scale=0.2, est="plane", iters=15, kw=dict(rot_range=(0.08, 0.1), trans_range=(0.08, 0.1))),
]


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    out = []
    for c in CASES:
        cam = synth.Camera().scaled(c["scale"])
        p = synth.make_pair(c["pair"], cam=cam, **c["kw"])
        r = icp_numpy(p["src"], p["tgt"], p["tgt_normals"], c["iters"], c["est"], max_corr_dist=c.get("max_corr_dist", 0.0))
        out.append(dict(name=c["name"], pair=c["pair"], scale=c["scale"], estimator=c["est"], iterations=c["iters"], kw=c["kw"],
                        max_corr_dist=c.get("max_corr_dist", 0.0), n_src=len(p["src"]), n_tgt=len(p["tgt"]),
                        input_sha256=digest(p["src"], p["tgt"], p["tgt_normals"]),
                        T=r["T"].tolist(), inliers=r["inliers"], fitness=r["fitness"], T_gt=p["T_gt"].tolist()))
        print(c["name"], len(p["src"]), r["inliers"], synth.pose_error(r["T"], p["T_gt"]))
    with open(os.path.join(HERE, "icp_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    ref = "/root/reference/data/exp1/dep"
    if os.path.isdir(ref):
        import cv2
        d1 = cv2.imread(os.path.join(ref, "1.png"), -1)
        d2 = cv2.imread(os.path.join(ref, "2.png"), -1)
        assert d1.dtype == np.uint16 and d1.shape == (480, 640)
        np.savez_compressed(os.path.join(HERE, "exp1_depth_q4.npz"), d1=d1[::4, ::4].copy(), d2=d2[::4, ::4].copy())
        print("exp1 subsample", d1[::4, ::4].shape, (d1[::4, ::4] == 0).mean(), (d2[::4, ::4] == 0).mean())


if __name__ == "__main__":
    main()
