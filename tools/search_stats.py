"""Per-iteration search statistics and the phase split inside tile_search (warp 0 of CTA 0) from a
-DS3D_STATS -DS3D_PHASES build (S3D_LIBRARY=.../libslam3d_b200_stats.so)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
lib = ctx.lib
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
p = synth.make_pair(seed)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
buf = (C.c_ulonglong * 32)()
prev = np.zeros(32)
n = 307200.0
mhz = 1920.0
for k in (1, 2, 3, 4, 5, 6, 8, 12, 30):
    lib.s3d_debug_stats(buf, 1)
    ctx.register(src, tgt, None, _abi.icp_params(k))
    lib.s3d_debug_stats(buf, 1)
    cur = np.array(list(buf), dtype=np.float64)
    dd = cur - prev
    passes = max(dd[5], 1)
    print(f"it<{k:2d}: searched/q={dd[0]/n:.3f} coarse/q={dd[6]/n:.3f} skipped/q={dd[1]/n:.3f} rows/pass={dd[2]/passes:.1f} "
          f"lane-cands/pass={dd[3]/passes:.0f} passes={passes:.0f} | cta0 us: p1={dd[8]/mhz:.1f} p2={dd[9]/mhz:.1f} bar={dd[10]/mhz:.1f} "
          f"tot={dd[11]/mhz:.1f} solve={dd[12]/mhz:.1f} | warp0 in search us: box={dd[21]/mhz:.1f} rowload={dd[22]/mhz:.1f} "
          f"copyissue={dd[23]/mhz:.1f} wait+compare={dd[24]/mhz:.1f} verify={dd[25]/mhz:.1f} | cta0 fills: cands={dd[26]:.0f} rows={dd[27]:.0f} passes={dd[28]:.0f} nin={dd[29]:.0f} | passes with >128 rows: {dd[13]:.0f} (rows {dd[14]:.0f}), >512 rows: {dd[15]:.0f} (rows {dd[16]:.0f}); all rows {dd[2]:.0f}; sum of largest box rows {dd[17]:.0f} | items (4 CTAs): n={dd[20]:.0f} mean={dd[19]/max(dd[20],1)/mhz:.1f} us max(all runs so far)={cur[18]/mhz:.1f} us; warp finish after pass-2 start: mean={dd[31]/(64*1)/mhz:.1f} us (per iteration sum) max so far={cur[30]/mhz:.1f} us", flush=True)
    prev = cur
