"""Scratch: per-iteration search statistics from a -DS3D_STATS build of the library."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
lib = ctx.lib
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
buf = (C.c_ulonglong * 16)()
names = ["searched", "skipped", "rows", "cands", "maskloads", "rounds", "coarse", "warpmax_rows_est", "sum_rows_est_small", "big", "warps_perlane", "warps_pending"]
prev = np.zeros(16)
first=True
for k in (1, 2, 3, 4, 6, 10, 20, 30):
    lib.s3d_debug_stats(buf, 1)
    ctx.register(src, tgt, None, _abi.icp_params(k))
    lib.s3d_debug_stats(buf, 1)
    cur = np.array(list(buf), dtype=np.float64)
    n = 307200.0
    dd = cur - prev
    print(f"iters {k:2d}: delta per-query: " + " ".join(f"{names[i]}={dd[i]/n:.3f}" for i in (0, 1, 2, 3, 5, 7, 8, 9, 10, 11)))
    prev = cur
