#!/bin/bash
# Scratch: A/B of alternative builds (S3D_LIBRARY) on one pair (iteration profile of seeds 0 and 5).
run() { echo "== $*"; env "$@" python tools/iter_profile.py 0 5 2>&1 | python -c "
import json,sys
for ln in sys.stdin.read().strip().splitlines():
    d=json.loads(ln)
    print(' seed', d['seed'], 'total10', round(d['total_us']['10']), 'total30', round(d['total_us']['30']), 'late', d['late_us'], [d['iteration_us'][str(k)] for k in range(0,8)])"; }
run X=1
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_cpasync.so
