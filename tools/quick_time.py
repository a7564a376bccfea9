"""Scratch timing probe for the GPU box (not part of the product)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

ctx = s3d.Context(0)
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
for cell in [float(x) for x in (sys.argv[1:] or ["0"])]:
    for reuse in (1, 0):
        prm = _abi.icp_params(30, grid_cell=cell, reuse_index=reuse)
        for _ in range(3):
            r = ctx.register(src, tgt, None, prm)
        ts = []
        for _ in range(5):
            t = time.perf_counter(); r = ctx.register(src, tgt, None, prm); ts.append(time.perf_counter() - t)
        tm = ctx.last_timing()
        print(f"cell={cell} reuse={reuse} wall={min(ts)*1e3:.3f} ms  index={tm['index_ms']:.3f} ms iterate={tm['iterate_ms']:.3f} ms "
              f"({tm['iterate_ms']/30*1e3:.1f} us/iter) launches={tm['total_launches']} inl={r['inliers']}", flush=True)
prm = _abi.icp_params(2, search=_abi.SEARCH_BRUTE)
r = ctx.register(src, tgt, None, prm); r = ctx.register(src, tgt, None, prm)
tm = ctx.last_timing(); print(f"brute: iterate={tm['iterate_ms']:.3f} ms for 2 iters")
t = time.perf_counter(); planes = tgt.segment_planes(_abi.plane_params()); print("segment", (time.perf_counter()-t)*1e3, "ms", len(planes))
t = time.perf_counter(); planes = tgt.segment_planes(_abi.plane_params()); print("segment", (time.perf_counter()-t)*1e3, "ms", len(planes))
