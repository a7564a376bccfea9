#!/bin/bash
# Scratch: A/B of alternative builds (S3D_LIBRARY) in ONE session: lone pair (seeds 0, 5) and the 16-pair batch regime.
lone() { env "$@" python tools/iter_profile.py 0 5 2>&1 | python -c "
import json,sys
for ln in sys.stdin.read().strip().splitlines():
    d=json.loads(ln); print('  lone seed', d['seed'], 'total30', round(d['total_us']['30']), 'late', d['late_us'], [d['iteration_us'][str(k)] for k in range(0,8)])"; }
batch() { env "$@" python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi
ctx = s3d.Context(0)
srcs, tgts = [], []
for i in range(16):
    p = synth.make_pair(i); srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"]))
def t(iters):
    prm = _abi.icp_params(iters); best = 1e9
    for _ in range(4):
        ctx.register_batch(srcs, tgts, None, prm); best = min(best, ctx.last_timing()["iterate_ms"])
    return best
t10, t40 = t(10), t(40)
print("  batch16 late_iteration_us %.1f  t10 %.2f ms" % ((t40 - t10) / 30 * 1e3, t10))
PY
}
for lib in "" $LIBS; do
  echo "== lib '$lib'"
  if [ -z "$lib" ]; then lone X=1; batch X=1; else lone S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200$lib.so; batch S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200$lib.so; fi
done
