#!/bin/bash
# Scratch: A/B of alternative builds (S3D_LIBRARY) on one pair (iteration profile of seed 0) and on the batch regime.
run() { echo "== $*"; env "$@" python tools/iter_profile.py 0 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(' total10', round(d['total_us']['10']), 'total30', round(d['total_us']['30']), 'late', d['late_us'], [d['iteration_us'][str(k)] for k in range(0,8)])"; env "$@" python tools/batch_streaming_probe.py 2>&1 | tail -1; }
run X=1
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_u2.so
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_u3.so
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_c8.so
