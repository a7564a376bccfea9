"""Per-iteration time profile of a lone 640x480 registration: t(k iterations) for k = 1..K, CUDA events inside the library
(s3d_last_timing).  Differences of consecutive k are the cost of iteration k-1.  Usage: python tools/iter_profile.py [seed ...]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

ctx = s3d.Context(0)
seeds = [int(x) for x in sys.argv[1:]] or [0, 5]
ks = list(range(1, 15)) + [20, 30]
for seed in seeds:
    p = synth.make_pair(seed)
    src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
    t = {}
    for k in ks:
        prm = _abi.icp_params(k)
        ts = []
        for _ in range(6):
            ctx.register(src, tgt, None, prm)
            ts.append(ctx.last_timing()["iterate_ms"] * 1e3)
        t[k] = float(np.median(ts[2:]))
    per = {k: t[k] - t.get(k - 1, 0.0) for k in ks if k == 1 or (k - 1) in t}
    print(json.dumps({"seed": seed, "total_us": t, "iteration_us": {str(k - 1): round(v, 1) for k, v in per.items()},
                      "late_us": round((t[30] - t[20]) / 10, 2)}), flush=True)
    src.free(); tgt.free()
