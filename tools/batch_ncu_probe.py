"""Scratch (run under ncu): a 16-pair batch registered with 10 and with 40 iterations, twice each.  The difference of the DRAM
bytes of the two launches / 30 = DRAM traffic of one late (streaming) iteration of the batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
srcs, tgts = [], []
for i in range(16):
    p = synth.make_pair(i)
    srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
for iters in (10, 10, 40, 40):
    ctx.register_batch(srcs, tgts, None, _abi.icp_params(iters))
    print(iters, ctx.last_timing()["iterate_ms"], flush=True)
