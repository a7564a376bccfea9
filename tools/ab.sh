#!/bin/bash
# A/B of alternative builds (LIBS="_suffix ..." -> S3D_LIBRARY) and of environment knobs (ENVS="K=V ...", default build) in ONE
# session on one GPU: lone pair (seeds 0, 5) and the 16-pair batch regime.  Clocks differ between boxes: only same-session rows compare.
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
for e in $ENVS; do echo "== env '$e'"; lone $e; batch $e; done
# BENCH=1: also the bench headline (avg launch of the pool's pairs) and the config-4 leg for every library
if [ -n "$BENCH" ]; then
  for lib in "" $LIBS; do
    L=""; [ -n "$lib" ] && L="S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200$lib.so"
    env $L python bench.py --no-cpu-baseline --no-batch-regime --steps 120 --warmup 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench lib \'$lib\': value', round(d['value']), 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'config4', round(d['config4']['iterations_per_s']) if d.get('config4') else None)"
  done
  for e in $ENVS; do
    env $e python bench.py --no-cpu-baseline --no-batch-regime --steps 120 --warmup 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  bench env \'$e\': value', round(d['value']), 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'config4', round(d['config4']['iterations_per_s']) if d.get('config4') else None)"
  done
fi
