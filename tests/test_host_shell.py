"""Host shell (C++): ParameterReader / Pose / g2o writer / PCD reader on CPU, and the run_SLAM driver end to end on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from slam3d_gx_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "slam3d_gx_b200", "host")

# same keys and (where it matters) the same values as the reference's stock parameters.yaml, incl. the %YAML line,
# comments (also non-ASCII) and a value with a trailing zero
STOCK_YAML = """%YAML:1.0
# part 1 data source
data_source: /tmp/some dataset
# 特征
detector_name: SIFT
descriptor_name: SIFT
start_index: 1
end_index: 2800
match_min_dist: 5
step_time: 10
optimize_step: 200
robust_kernel: Cauchy
max_pos_change: 0.25
grid_leaf: 0.03
error_threshold: 1.0
distance_threshold: 0.080   # plane threshold
plane_percent: 0.2
min_error_plane: 0.02
max_planes: 3
loop_closure_detection: yes
loopclosure_frames: 30
loop_closure_error: 1.5
loop_closure_inliers: 30
ransac_accuracy: 8.0
lost_frames: 10
use_odometry: no
error_odometry: 0.03
z_filter: 7.0
camera_fx: 517.0
camera_fy: 517.0
camera_cx: 318.6
camera_cy: 255.3
camera_factor: 5000.0
"""


def _build_host():
    subprocess.run(["make", "-C", HOST, "-s", "bin/test_host"], check=True, env={**os.environ, "CXX": "g++", "CC": "gcc"})


def test_host_pieces_cpu(tmp_path):
    _build_host()
    y = tmp_path / "parameters.yaml"
    y.write_text(STOCK_YAML, encoding="utf-8")
    r = subprocess.run([os.path.join(HOST, "bin", "test_host"), str(y), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = (tmp_path / "t.g2o").read_text().splitlines()
    assert lines[0].split() == ["VERTEX_SE3:QUAT", "0", "0", "0", "0", "0", "0", "0", "1"]
    assert lines[1] == "FIX 0"
    assert lines[2].startswith("VERTEX_SE3:QUAT 1 1 -2 0.5 ")
    e = lines[3].split()
    assert e[0] == "EDGE_SE3:QUAT" and e[1:3] == ["0", "1"] and len(e) == 3 + 7 + 21
    info = [float(x) for x in e[10:]]
    diag = [0, 6, 11, 15, 18, 20]
    assert all(info[i] == (100.0 if i in diag else 0.0) for i in range(21))      # 100 * I6 (reference GraphicEnd.cpp:330-334)


def _write_pcd(path, xyzw):
    a = np.ascontiguousarray(xyzw, dtype=np.float32)
    with open(path, "wb") as f:
        f.write((f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgba\nSIZE 4 4 4 4\nTYPE F F F U\n"
                 f"COUNT 1 1 1 1\nWIDTH {len(a)}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {len(a)}\nDATA binary\n").encode())
        f.write(a.tobytes())


def _quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.mark.gpu
def test_run_slam_end_to_end(tmp_path):
    """bin/run_SLAM on a synthetic sequence: key frames, g2o edges equal to the ground-truth relative poses."""
    subprocess.run(["make", "-C", HOST, "-s"], check=True, env={**os.environ, "CXX": "g++", "CC": "gcc"})
    cam = synth.Camera().scaled(0.25)
    n_frames = 13
    D = synth.make_T(synth.rot_axis_angle([0.1, 1.0, 0.05], -0.012), [0.03, -0.01, 0.025])    # camera motion per frame (norm ~0.05)
    poses = [synth.base_pose()]
    for k in range(1, n_frames):
        poses.append(poses[-1] @ D)
    (tmp_path / "ds" / "pcd").mkdir(parents=True)
    (tmp_path / "data").mkdir()
    for k, C in enumerate(poses):
        z, _ = synth.render_depth(C, cam, "S1", 0.002, 77, 100 + k)
        pts, _ = synth.backproject(z, cam)
        _write_pcd(tmp_path / "ds" / "pcd" / f"{k + 1}.pcd", pts)
    yaml = STOCK_YAML.replace("/tmp/some dataset", str(tmp_path / "ds")).replace("loop_closure_detection: yes", "loop_closure_detection: no")
    yaml += "icp_iterations: 20\nrandom_seed: 1\n"
    (tmp_path / "parameters.yaml").write_text(yaml, encoding="utf-8")
    r = subprocess.run([os.path.join(HOST, "bin", "run_SLAM"), str(n_frames - 1)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    kf = [tuple(int(x) for x in l.split()) for l in (tmp_path / "data" / "keyframe.txt").read_text().splitlines()]
    assert kf[0] == (0, 1) and len(kf) >= 3          # motion per frame: norm ~0.06 -> a key frame roughly every 5 frames (max_pos_change 0.25)
    assert [k[0] for k in kf] == list(range(len(kf)))
    verts, edges, fixed = {}, [], []
    for line in (tmp_path / "data" / "final.g2o").read_text().splitlines():
        t = line.split()
        if t[0] == "VERTEX_SE3:QUAT":
            verts[int(t[1])] = [float(x) for x in t[2:9]]
        elif t[0] == "FIX":
            fixed.append(int(t[1]))
        elif t[0] == "EDGE_SE3:QUAT":
            edges.append((int(t[1]), int(t[2]), [float(x) for x in t[3:10]]))
    assert fixed == [0] and len(verts) == len(kf) and len(edges) == len(kf) - 1
    frame_of = dict(kf)
    for a, b, m in edges:
        assert b == a + 1
        T = np.eye(4); T[:3, :3] = _quat_to_R(m[3:]); T[:3, 3] = m[:3]
        gt = np.linalg.inv(poses[frame_of[a] - 1]) @ poses[frame_of[b] - 1]     # pose of key frame b in key frame a
        rot, trans = synth.pose_error(T, gt)
        assert rot < 4e-3 and trans < 1e-2, (a, b, rot, trans)
    assert (tmp_path / "data" / "final_after.g2o").exists()
    log = (tmp_path / "data" / "error_of_transform.log").read_text().split()
    assert len(log) == n_frames - 1 and "9999" not in log
