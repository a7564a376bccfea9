"""Device time of an n-pair batch (10 iterations, indices rebuilt) under several CTA-group sizes (S3D_GROUP_CTAS is read once per
process, so one process per setting).  Usage: python tools/group_probe.py n_pairs"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = s3d.Context(0)
prm = _abi.icp_params(10, reuse_index=0)
srcs, tgts = [], []
for i in range(n):
    p = synth.make_pair(i); srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
ts = []
for _ in range(4):
    ctx.register_batch(srcs, tgts, None, prm); tm = ctx.last_timing(); ts.append(round(tm["iterate_ms"] + tm["index_ms"], 2))
print(json.dumps({"pairs": n, "group_ctas": os.environ.get("S3D_GROUP_CTAS", "rule"), "device_ms": ts, "ms_per_pair": round(min(ts) / n, 4)}), flush=True)
