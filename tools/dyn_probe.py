"""Scratch: A/B of the dynamic decide pass (S3D_DYN_DIV=0 disables it; read once per process).  Prints the time of
k-iteration registrations and a hash of the results: the hashes must not depend on the setting."""
import sys, os, hashlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

ctx = s3d.Context(0)
out = []
for seed in (0, 5):
    p = synth.make_pair(seed)
    src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
    for k in (1, 2, 3, 4, 6, 10, 30):
        prm = _abi.icp_params(k)
        ts = []
        for _ in range(7):
            r = ctx.register(src, tgt, None, prm)
            ts.append(ctx.last_timing()["iterate_ms"])
        nn = ctx.last_correspondences(len(p["src"]))
        h = hashlib.sha1(nn.tobytes()).hexdigest()[:6] + "." + hashlib.sha1(np.asarray(r["T"]).tobytes()).hexdigest()[:4]
        out.append(f"{k}:{np.median(ts[2:])*1e3:.0f}us/{h}")
    src.free(); tgt.free()
print("DYN_DIV=%s FUSED_DIV=%s " % (os.environ.get("S3D_DYN_DIV", "default"), os.environ.get("S3D_FUSED_DIV", "default")) + " ".join(out), flush=True)
