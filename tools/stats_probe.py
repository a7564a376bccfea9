"""Scratch: per-iteration search statistics from a -DS3D_STATS build of the library (S3D_LIBRARY=...stats.so)."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
lib = ctx.lib
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
buf = (C.c_ulonglong * 32)()
names = ["searched", "skipped", "rows", "cands", "maskloads", "passes", "coarse"]
mode = _abi.SEARCH_GRID_LANE if (len(sys.argv) > 1 and sys.argv[1] == "lane") else _abi.SEARCH_GRID
prev = np.zeros(32)
n = 307200.0
for k in (1, 2, 3, 4, 5, 6, 8, 10, 20, 30):
    lib.s3d_debug_stats(buf, 1)
    ctx.register(src, tgt, None, _abi.icp_params(k, search=mode))
    lib.s3d_debug_stats(buf, 1)
    cur = np.array(list(buf), dtype=np.float64)
    dd = cur - prev
    print(f"iters {k:2d}: delta: searched/q={dd[0]/n:.3f} skipped/q={dd[1]/n:.3f} rows/q={dd[2]/n:.2f} cands/q={dd[3]/n:.1f} "
          f"passes/warp={dd[5]/(n/32):.2f} coarse/q={dd[6]/n:.3f} | cta0 us: chunks={dd[8]/1965:.1f} reduce={dd[9]/1965:.1f} "
          f"barrier={dd[10]/1965:.1f} rowsum={dd[11]/1965:.1f} solve={dd[12]/1965:.1f} | warp loop: max={cur[16]/1965:.1f} mean={dd[17]/1965/16:.1f} "
          f"search: total={dd[18]/1965:.1f} n={dd[19]:.0f} max/iter={cur[20]/1965:.1f} | in search: box={dd[21]/1965:.1f} rowload={dd[22]/1965:.1f} copy={dd[23]/1965:.1f} compare={dd[24]/1965:.1f} verify={dd[25]/1965:.1f} cta0: cands={dd[26]:.0f} rows={dd[27]:.0f} passes={dd[28]:.0f} todo_lanes={dd[29]:.0f} same_pose_iters={dd[30]:.0f}", flush=True)
    prev = cur
