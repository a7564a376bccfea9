"""Per-iteration phase times of CTA 0 (S3D_PHASES build, S3D_LIBRARY=...phases.so): runs k and k+10 iterations and prints the
per-iteration delta of every phase: pass 1 (streaming), pass 2 (search), barrier (atomics + arrive + wait), totals, solve."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0); lib = ctx.lib
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
buf = (C.c_ulonglong * 32)()
def run(k):
    ctx.register(src, tgt, None, _abi.icp_params(k))
    lib.s3d_debug_stats(buf, 1)
    ctx.register(src, tgt, None, _abi.icp_params(k))
    lib.s3d_debug_stats(buf, 1)
    return np.array(list(buf), dtype=np.float64)
prev = np.zeros(32)
for k in (1, 2, 3, 4, 6, 10, 20, 30):
    cur = run(k)
    names = {8: "pass1", 9: "pass2", 10: "barrier", 11: "totals", 12: "solve"}
    print(f"iterations {k:2d}: cumulative us " + " ".join(f"{names[i]}={cur[i] / 1965.0:8.1f}" for i in sorted(names)), flush=True)
a, b = run(20), run(30)
d = (b - a) / 10 / 1965.0
print("late iteration (avg of 20..29), CTA 0, us: " + " ".join(f"{n}={d[i]:.2f}" for i, n in ((13, "p1:wait-for-staged-loads"), (14, "p1:chunks"), (15, "p1:hand-over"), (8, "p1:sync"), (9, "pass2"), (10, "barrier"), (11, "totals"), (12, "solve"))))
