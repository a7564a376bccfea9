"""Scratch probe: per-iteration cost profile (early vs converged iterations) for several cell sizes."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

ctx = s3d.Context(0)
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
def t_iter(prm, guess=None):
    for _ in range(2): ctx.register(src, tgt, guess, prm)
    best = 1e9
    for _ in range(3):
        ctx.register(src, tgt, guess, prm); best = min(best, ctx.last_timing()["iterate_ms"])
    return best
for cell in (0.01, 0.02, 0.04, 0.08):
    cum = [t_iter(_abi.icp_params(k, grid_cell=cell)) for k in (1, 2, 3, 4, 6, 8, 30)]
    conv = t_iter(_abi.icp_params(10, grid_cell=cell), guess=p["T_gt"]) / 10
    print(f"cell={cell}: cumulative ms for k=1,2,3,4,6,8,30: " + " ".join(f"{c:.3f}" for c in cum) + f" | converged us/iter={conv*1e3:.1f}", flush=True)
