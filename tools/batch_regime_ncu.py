"""Driver for ncu: a 16-pair batch of 640x480 pairs (working set ~400 MB > L2) registered with 10 and with 40 iterations, twice each
(one launch = the whole batch, all iterations).  The difference of the DRAM counters of the two launches is the traffic of 30 late
iterations.  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:icp_persist"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
srcs, tgts = [], []
for i in range(16):
    p = synth.make_pair(i)
    srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
for iters in (10, 40, 10, 40):
    r = ctx.register_batch(srcs, tgts, None, _abi.icp_params(iters))
    print(iters, [x["status"] for x in r][:3], ctx.last_timing())
