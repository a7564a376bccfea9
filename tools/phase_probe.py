"""Scratch: per-iteration phase times of CTA 0 (S3D_PHASES build): runs k and k+10 iterations, prints the per-iteration delta."""
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
a, b = run(20), run(30)
d = (b - a) / 10 / 1965.0
print(f"late iteration (avg of 20..29), CTA 0, us: chunks(warp0)={d[8]:.2f} reduce={d[9]:.2f} barrier={d[10]:.2f} rowsum={d[11]:.2f} solve={d[12]:.2f} "
      f"| warp loop mean={d[17]/16:.2f} | searches/iter={(b[19]-a[19])/10:.1f} search us each={(b[18]-a[18])/max(1,(b[19]-a[19]))/1965:.2f}")
