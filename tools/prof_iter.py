"""Driver for ncu: one single-pair registration (config 2 shape), optionally starting converged."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
converged = len(sys.argv) > 2 and sys.argv[2] == "conv"
ctx = s3d.Context(0)
p = synth.make_pair(0)
src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
r = ctx.register(src, tgt, p["T_gt"] if converged else None, _abi.icp_params(iters))
print(r["status"], r["inliers"], ctx.last_timing())
