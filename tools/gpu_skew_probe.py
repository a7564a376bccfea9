"""Which part of the rank skew of the 8-GPU config-4 run is the GPU and which the shard: every rank registers the SAME shards
(0 and 3 of the 512-pair partition: 64 pairs x 10 iterations, indices rebuilt), first all ranks at once, then one rank at a time.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/gpu_skew_probe.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi, sharding

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = s3d.Context(local)
prm = _abi.icp_params(10, reuse_index=0)
shards = {}
for sh in (0, 3):
    srcs, tgts = [], []
    for i in sharding.partition(512, 8, sh):
        p = synth.make_pair(i); srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
    shards[sh] = (srcs, tgts)

def run(sh, n=3):
    ts = []
    for _ in range(n):
        ctx.register_batch(*shards[sh], None, prm); tm = ctx.last_timing(); ts.append(round(tm["iterate_ms"] + tm["index_ms"], 2))
    return ts

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

out = {"rank": rank, "gpu": torch.cuda.get_device_properties(local).name}
run(0, 1); barrier()
out["together_shard0"] = run(0); barrier()
out["together_shard3"] = run(3); barrier()
for r in range(world):
    if r == rank:
        out["alone_shard0"] = run(0, 2); out["alone_shard3"] = run(3, 2)
    barrier()
print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
