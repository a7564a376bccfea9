#!/bin/bash
# Scratch: A/B sweeps of the search tunables on one pair (iteration profile of seed 0), one line per setting.
run() { echo "== $*"; env "$@" python tools/iter_profile.py 0 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(' total10', round(d['total_us']['10']), 'total30', round(d['total_us']['30']), 'late', d['late_us'], [d['iteration_us'][str(k)] for k in range(0,8)])"; }
run X=1
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_u2.so
run S3D_LIBRARY=$PWD/slam3d_gx_b200/libslam3d_b200_c8.so
for s in 1.0 1.25 2.0; do run S3D_GRID_CELL_SCALE=$s; done
for h in 0.5 0.75 1.5; do run S3D_HINT_CELLS=$h; done
for k in 0.04 0.16; do run S3D_SLACK_CELLS=$k; done
run S3D_USE_COARSE=0
run S3D_FIRST_CELLS=1.0 S3D_USE_COARSE=0
