"""The streaming (late-iteration) regime of a 64-pair batch, where the working set (64 x ~25 MB of per-query state and
clouds) does not fit L2: time per late iteration = (t(40 iterations) - t(10 iterations)) / 30, against the HBM roofline."""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = s3d.Context(0)
N = 307200
srcs, tgts = [], []
for i in range(64):
    p = synth.make_pair(i)
    srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
def t(iters):
    prm = _abi.icp_params(iters)
    ctx.register_batch(srcs, tgts, None, prm)
    best = 1e9
    for _ in range(2):
        ctx.register_batch(srcs, tgts, None, prm); best = min(best, ctx.last_timing()["iterate_ms"])
    return best
t10, t40 = t(10), t(40)
per_iter_us = (t40 - t10) / 30 * 1e3
alg = 64 * (16 * N + 32 * N)                    # 48 B per query and iteration
moved = 64 * N * 96                             # what the two passes actually read: (p, cq, xl) + (p, cq, cn)
print(json.dumps(dict(pairs=64, t10_ms=t10, t40_ms=t40, late_iteration_us=per_iter_us,
                      algorithmic_GBps=alg / per_iter_us / 1e3, frac_of_measured_hbm=alg / per_iter_us / 1e3 / peak,
                      actual_read_GBps=moved / per_iter_us / 1e3, actual_frac=moved / per_iter_us / 1e3 / peak)))
