"""Scratch timing probe for the GPU box: the three search modes on the config-2 pair, warm."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

ctx = s3d.Context(0)
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
iters_list = [int(x) for x in (sys.argv[1:] or ["1", "2", "3", "4", "6", "10", "30"])]
ref = {}
for name, mode in (("tile", _abi.SEARCH_GRID), ("lane", _abi.SEARCH_GRID_LANE)):
    for it in iters_list:
        prm = _abi.icp_params(it, search=mode, reuse_index=1)
        for _ in range(3):
            r = ctx.register(src, tgt, None, prm)
        ts = []
        for _ in range(7):
            r = ctx.register(src, tgt, None, prm); ts.append(ctx.last_timing()["iterate_ms"])
        nn = ctx.last_correspondences(len(p["src"]))
        key = it
        same = ""
        if key in ref:
            same = f" nn_equal_to_tile={np.array_equal(nn, ref[key][0])} dT={np.abs(r['T']-ref[key][1]).max():.2e}"
        else:
            ref[key] = (nn, r["T"])
        print(f"{name} iters={it:2d} iterate={min(ts)*1e3:8.1f} us (median {np.median(ts)*1e3:8.1f})  inl={r['inliers']} status={r['status']}{same}", flush=True)
