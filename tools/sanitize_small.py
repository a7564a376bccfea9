"""Driver for compute-sanitizer: a small registration (all search modes), plane extraction, filters."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
cam = synth.Camera().scaled(0.25)
p = synth.make_pair(0, cam=cam, quantize=True, holes=0.1)
ctx = s3d.Context(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
for mode in (_abi.SEARCH_GRID, _abi.SEARCH_GRID_LANE, _abi.SEARCH_BRUTE):
    r = ctx.register(src, tgt, None, _abi.icp_params(8, search=mode, reuse_index=0))
    print("mode", mode, r["status"], r["inliers"])
r = ctx.register(src, tgt, None, _abi.icp_params(6, max_corr_dist=0.05, estimator=_abi.ESTIMATOR_SVD))
print("svd gate", r["status"], r["inliers"])
res = ctx.register_batch([src] * 5, [tgt] * 5, None, _abi.icp_params(5))
res = ctx.register_batch([src] * 160, [tgt] * 160, None, _abi.icp_params(3))      # more pairs than CTAs: one CTA per pair, groups walk the list
print("batch", [x["status"] for x in res])
t2 = ctx.from_depth_normals(p["tgt_depth"], cam, 3.5, 1, 0.08)
print("planes", len(t2.segment_planes(_abi.plane_params())))
v = t2.voxel_grid(0.03); z = v.passthrough_z(0.0, 3.0); print("filters", len(v), len(z))
# the stream forms: extraction + registration enqueued behind each other, clouds released while their work is queued
for k in range(3):
    a = ctx.upload(p["src"]); b = ctx.upload(p["tgt"])
    b.segment_planes_enqueue(_abi.plane_params())
    ctx.register_enqueue(a, b, None, _abi.icp_params(4, reuse_index=0))
    a.release(); b.release()
pl = ctx.planes_drain(); rs, tm = ctx.register_drain()
print("stream", [len(x) for x in pl], [x["status"] for x in rs])
res = ctx.register_batch([src] * 20, [tgt] * 20, None, _abi.icp_params(3))      # more than 16 pairs: four groups walk the list
print("batch20", sorted(set(x["status"] for x in res)))
for c in (src, tgt, t2, v, z): c.free()
ctx.close()
