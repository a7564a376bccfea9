"""Device time of one rank's config-4 shard (64 pairs x 10 iterations, indices rebuilt) for several shards on ONE GPU:
how much of the rank skew of the 8-GPU run is the workload itself.  Usage: python tools/shard_probe.py [rank ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi, sharding
ctx = s3d.Context(0)
prm = _abi.icp_params(10, reuse_index=0)
for r in [int(x) for x in sys.argv[1:]] or [0, 3, 7]:
    srcs, tgts = [], []
    for i in sharding.partition(512, 8, r):
        p = synth.make_pair(i); srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
    ts = []
    for _ in range(4):
        ctx.register_batch(srcs, tgts, None, prm); tm = ctx.last_timing(); ts.append(tm["iterate_ms"] + tm["index_ms"])
    print(json.dumps({"rank": r, "device_ms": [round(t, 2) for t in ts]}), flush=True)
    for c in srcs + tgts: c.free()
