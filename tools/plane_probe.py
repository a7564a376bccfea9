"""Plane extraction of one 640x480 cloud: wall time of the call, CUDA-event times (s3d_last_plane_timing) and the algorithmic
bandwidth of the evaluation passes (16 B per point and pass).  Usage: python tools/plane_probe.py [reps]"""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

reps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 10
ctx = s3d.Context(0)
p = synth.make_pair(0)
c = ctx.upload(p["tgt"])
prm = _abi.plane_params(timed="timed" in sys.argv)
for _ in range(3):
    planes = c.segment_planes(prm)
ts, ev, tot = [], [], []
for _ in range(reps):
    t0 = time.perf_counter()
    planes = c.segment_planes(prm)
    ts.append((time.perf_counter() - t0) * 1e3)
    tm = ctx.last_plane_timing()
    ev.append(tm["eval_ms"]); tot.append(tm["total_ms"])
tm = ctx.last_plane_timing()
print(json.dumps({"planes": len(planes), "wall_ms_median": float(np.median(ts)), "device_total_ms_median": float(np.median(tot)),
                  "eval_ms_median": float(np.median(ev)), "rounds": tm["rounds"], "points_scanned": tm["points_scanned"],
                  "eval_GBps": tm["points_scanned"] * 16 * tm["eval_passes_per_round"] / max(float(np.median(ev)), 1e-9) / 1e-3 / 1e9}))
