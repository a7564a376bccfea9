"""Driver for ncu: a 16-pair batch started from the ground-truth poses, so that nearly every iteration is a
streaming (late) iteration: the regime in which the registration kernel is bound by HBM bandwidth."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = s3d.Context(0)
srcs, tgts, gts = [], [], []
for i in range(npairs):
    p = synth.make_pair(i)
    srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"])); gts.append(p["T_gt"])
for _ in range(2):
    r = ctx.register_batch(srcs, tgts, gts, _abi.icp_params(iters))
print([x["status"] for x in r][:4], ctx.last_timing())
