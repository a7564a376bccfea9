"""Scratch probe: late-iteration cost with and without a correspondence gate (isolates far-query stragglers)."""
import sys, os
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
for gate in (0.0, 0.3, 0.1, 0.05, 0.02):
    a = t_iter(_abi.icp_params(10, max_corr_dist=gate)); b = t_iter(_abi.icp_params(30, max_corr_dist=gate))
    c1 = t_iter(_abi.icp_params(1, max_corr_dist=gate)); c2 = t_iter(_abi.icp_params(2, max_corr_dist=gate))
    print(f"gate={gate}: it0={c1*1e3:.0f} us it1={(c2-c1)*1e3:.0f} us  late us/iter={(b-a)/20*1e3:.1f}  total30={b:.3f} ms", flush=True)
# tiny problem: fixed per-iteration overhead (launch + reduction + solve)
ps = synth.make_pair(0, cam=synth.Camera().scaled(0.05))
s2 = ctx.upload(ps["src"]); t2 = ctx.upload(ps["tgt"], ps["tgt_normals"])
for _ in range(3): ctx.register(s2, t2, None, _abi.icp_params(30))
print("tiny cloud (%d pts): us/iter=%.1f" % (len(ps["src"]), ctx.last_timing()["iterate_ms"] / 30 * 1e3))
