import os, sys, subprocess, tempfile, pathlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from slam3d_gx_b200 import synth
import test_host_shell as t
tmp = pathlib.Path(tempfile.mkdtemp())
try:
    t.test_run_slam_end_to_end(tmp)
except AssertionError as e:
    print("ASSERT", str(e)[:300])
r = subprocess.run([os.path.join(t.HOST, "bin", "run_SLAM"), "12"], cwd=tmp, capture_output=True, text=True)
print(r.stdout[-3500:]); print(r.stderr[-1500:])
