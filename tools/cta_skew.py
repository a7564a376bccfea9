"""When does each CTA finish the search pass of iterations 0..7 (S3D_PHASES build)?  Spread across the 148 CTAs of a lone pair."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0); lib = ctx.lib
for seed in (0, 5):
    p = synth.make_pair(seed)
    src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
    buf = (C.c_ulonglong * (148 * 8))()
    for _ in range(3):
        ctx.register(src, tgt, None, _abi.icp_params(8))
    lib.s3d_debug_cta_times(buf)
    t = np.array(list(buf), dtype=np.float64).reshape(148, 8) * 1e-3      # us
    t0 = t[:, 0].min()
    for it in range(8):
        e = t[:, it]
        start = (t[:, it - 1].max() if it else None)
        print(f"seed {seed} it {it}: search pass ends: min {e.min() - t0:8.1f} mean {e.mean() - t0:8.1f} max {e.max() - t0:8.1f} us  spread {e.max() - e.min():6.1f} us"
              + (f"  (iteration length ~{e.max() - start:6.1f} us)" if start is not None else "") + f"  mean CTA idle {(e.max() - e).mean():5.1f} us",
              "slowest CTAs", np.argsort(-e)[:5].tolist(), "fastest", np.argsort(e)[:5].tolist())
    src.free(); tgt.free()
