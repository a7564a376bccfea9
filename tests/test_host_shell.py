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


# ---- loop closure, lost recovery, residency, the saveOutput tool -------------------------------------------------------

def _read_pcd(path):
    raw = open(path, "rb").read()
    off = raw.index(b"DATA binary\n") + 12
    n = int([ln for ln in raw[:off].decode().splitlines() if ln.startswith("POINTS")][0].split()[1])
    return np.frombuffer(raw[off:off + 16 * n], dtype=np.float32).reshape(n, 4).copy()


def _parse_g2o(path):
    verts, edges, fixed = {}, [], []
    for line in open(path).read().splitlines():
        t = line.split()
        if not t:
            continue
        if t[0] == "VERTEX_SE3:QUAT":
            verts[int(t[1])] = [float(x) for x in t[2:9]]
        elif t[0] == "FIX":
            fixed.append(int(t[1]))
        elif t[0] == "EDGE_SE3:QUAT":
            edges.append((int(t[1]), int(t[2]), [float(x) for x in t[3:10]]))
    return verts, edges, fixed


def _pose_T(m):
    T = np.eye(4); T[:3, :3] = _quat_to_R(m[3:]); T[:3, 3] = m[:3]
    return T


def _make_dataset(tmp_path, poses, cam, scene_of=None, extra_yaml=""):
    (tmp_path / "ds" / "pcd").mkdir(parents=True)
    (tmp_path / "data").mkdir()
    for k, C in enumerate(poses):
        z, _ = synth.render_depth(C, cam, "S1", 0.002, 77, 100 + k)
        pts, _ = synth.backproject(z, cam)
        _write_pcd(tmp_path / "ds" / "pcd" / f"{k + 1}.pcd", pts)
    yaml = STOCK_YAML.replace("/tmp/some dataset", str(tmp_path / "ds")) + "icp_iterations: 20\nrandom_seed: 1\n" + extra_yaml
    (tmp_path / "parameters.yaml").write_text(yaml, encoding="utf-8")


def _run_slam(tmp_path, loops):
    subprocess.run(["make", "-C", HOST, "-s"], check=True, env={**os.environ, "CXX": "g++", "CC": "gcc"})
    r = subprocess.run([os.path.join(HOST, "bin", "run_SLAM"), str(loops)], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    if os.environ.get("S3D_TEST_LOGDIR"):               # developer aid: keep the shell's output of every run
        with open(os.path.join(os.environ["S3D_TEST_LOGDIR"], tmp_path.name + ".log"), "w") as f:
            f.write(r.stdout + "\n--- stderr ---\n" + r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
def test_run_slam_loop_closure_and_save_output(tmp_path, ctx):
    """A camera that walks out and comes back over the same path with `loop_closure_detection: yes` (the stock setting):
    loopClosure() (reference src/GraphicEnd.cpp:685-762) registers the new key frame against earlier ones in one batched
    call, writes lc.txt lines `frame1 frame2 norm inliers` (:861) for the randomly picked ones and adds loop edges; every
    accepted loop edge must be the ground-truth relative pose.  Then the reference's saveOutput flow (src/saveOutput.cpp:
    29-95) on the files the run wrote: the fused map must equal the in-memory s3d_map_fuse bit for bit."""
    from test_reference_pins import parse_lc
    cam = synth.Camera().scaled(0.25)
    D = synth.make_T(synth.rot_axis_angle([0.1, 1.0, 0.05], -0.012), [0.03, -0.01, 0.025])
    out = [synth.base_pose()]
    for k in range(1, 14):
        out.append(out[-1] @ D)
    poses = out + out[-2::-1]                               # there and back again: 27 frames
    _make_dataset(tmp_path, poses, cam)
    log = _run_slam(tmp_path, len(poses) - 1)
    kf = [tuple(int(x) for x in l.split()) for l in (tmp_path / "data" / "keyframe.txt").read_text().splitlines()]
    assert len(kf) >= 5 and [k[0] for k in kf] == list(range(len(kf)))
    frame_of = dict(kf)
    verts, edges, fixed = _parse_g2o(tmp_path / "data" / "final.g2o")
    assert fixed == [0] and len(verts) == len(kf)
    chain = [(a, b) for a, b, _ in edges if b == a + 1]
    loops = [(a, b, m) for a, b, m in edges if b != a + 1]
    assert len(chain) >= len(kf) - 1 and len(loops) >= 1, (len(kf), edges)
    for a, b, m in edges:
        gt = np.linalg.inv(poses[frame_of[a] - 1]) @ poses[frame_of[b] - 1]
        rot, trans = synth.pose_error(_pose_T(m), gt)
        assert rot < 5e-3 and trans < 1.5e-2, (a, b, rot, trans)                      # odometry AND loop edges are right
    lc = parse_lc(tmp_path / "data" / "lc.txt")
    assert len(lc) >= 1
    frames = set(frame_of.values())
    for f1, f2, nrm, inl in lc:
        assert f1 in frames and f2 in frames and f1 < f2 and 0.0 <= nrm <= 1.5 and inl >= 30     # the gates of reference :703-708,739-744
    assert "Peak resident clouds" in log
    # ---- saveOutput on the run's own files: final_after.g2o carries the spanning-tree estimates
    va, _, _ = _parse_g2o(tmp_path / "data" / "final_after.g2o")
    for vid, frame in kf:
        gt = np.linalg.inv(poses[0]) @ poses[frame - 1]
        rot, trans = synth.pose_error(_pose_T(va[vid]), gt)
        assert rot < 2e-2 and trans < 5e-2, (vid, rot, trans)
    r = subprocess.run([os.path.join(HOST, "bin", "saveOutput"), "data/keyframe.txt", "data/final_after.g2o", "5.0"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = _read_pcd(tmp_path / "result.pcd")
    # the same fusion from memory: g2o text keeps 9 significant digits, so the poses are re-read through the same C++ reader
    clouds = [ctx.upload(_read_pcd(tmp_path / "ds" / "pcd" / f"{frame}.pcd")) for _, frame in kf]
    Ts = [_pose_T(va[vid]) for vid, _ in kf]
    fused = ctx.map_fuse(clouds, Ts, 0.03, 5.0)
    want = fused.download()["xyz"]
    for c in clouds + [fused]:
        c.free()
    assert got.shape[0] == want.shape[0] and got.shape[0] > 1000
    assert np.abs(got[:, :3] - want).max() < 2e-6          # python re-normalises the quaternion: poses agree to ~1e-9, points to float rounding


@pytest.mark.gpu
def test_run_slam_lost_recovery(tmp_path):
    """Frames in which the camera has turned away and sees only a far strip of floor cannot be registered (no correspondence
    within icp_max_corr_dist: status FEW -> T == Identity, the reference's failure convention).  With lost_frames: 1 the
    second such frame triggers lostRecovery() (reference src/GraphicEnd.cpp:764-838): a key frame without an edge to its
    predecessor, a line `kf_id frame_index` in lost.txt (:775-777), a sweep over all earlier key frames."""
    cam = synth.Camera().scaled(0.25)
    D = synth.make_T(synth.rot_axis_angle([0.1, 1.0, 0.05], -0.012), [0.03, -0.01, 0.025])
    turn = synth.make_T(synth.rot_axis_angle([0, 1, 0], np.pi), [0, 0, 0])
    poses = [synth.base_pose()]
    for k in range(1, 7):
        poses.append(poses[-1] @ D)                         # frames 1..7 track normally
    away = poses[-1] @ turn
    poses += [away, away @ D, away @ D @ D]                 # frames 8, 9, 10: floor only
    back = poses[6]
    for k in range(6):
        back = back @ D
        poses.append(back)                                  # frames 11..16: the walk goes on
    _make_dataset(tmp_path, poses, cam, extra_yaml="")
    y = (tmp_path / "parameters.yaml").read_text().replace("lost_frames: 10", "lost_frames: 1").replace("loop_closure_detection: yes", "loop_closure_detection: no")
    (tmp_path / "parameters.yaml").write_text(y, encoding="utf-8")
    log = _run_slam(tmp_path, len(poses) - 1)
    assert "Lost Recovery" in log and "This frame lost" in log
    lost = [tuple(int(x) for x in l.split()) for l in (tmp_path / "data" / "lost.txt").read_text().splitlines()]
    assert len(lost) >= 1 and lost[0][1] == 9               # frames 8 and 9 cannot be registered: recovery at the second one
    kf = [tuple(int(x) for x in l.split()) for l in (tmp_path / "data" / "keyframe.txt").read_text().splitlines()]
    assert (lost[0][0], 9) in kf and [k[0] for k in kf] == list(range(len(kf)))
    verts, edges, fixed = _parse_g2o(tmp_path / "data" / "final.g2o")
    assert len(verts) == len(kf)
    assert not any(b == lost[0][0] for _, b, _ in edges)   # position unknown: no edge into the recovery key frame (:791-792)
    errlog = (tmp_path / "data" / "error_of_transform.log").read_text().split()
    assert errlog.count("9999") >= 2                        # reference :176


@pytest.mark.gpu
def test_run_slam_keeps_only_key_frames_resident(tmp_path):
    """300 frames: only key frames (+ the frames in flight) stay in HBM.  Before, every frame kept its cloud and the 40 MB
    search index it got as a registration target (the stock 2800-frame sequence would have needed > 150 GB)."""
    cam = synth.Camera().scaled(0.25)
    poses = []
    for k in range(300):
        a = 2 * np.pi * k / 150.0
        poses.append(synth.base_pose() @ synth.make_T(synth.rot_axis_angle([0, 1, 0], 0.12 * np.sin(a)), [0.25 * np.sin(a), 0.0, 0.1 * (1 - np.cos(a))]))
    _make_dataset(tmp_path, poses, cam)
    log = _run_slam(tmp_path, len(poses) - 1)
    kf = [l for l in (tmp_path / "data" / "keyframe.txt").read_text().splitlines() if l.strip()]
    line = [l for l in log.splitlines() if "Peak resident clouds" in l][-1].split()
    peak_clouds, peak_bytes = int(line[3]), int(line[7])
    assert 3 <= len(kf) < 100
    assert peak_clouds <= len(kf) + 3, (peak_clouds, len(kf))
    n = cam.width * cam.height
    per_frame = n * (16 + 16 + 4) + 4096 * 4                # points + normals + labels (+ allocator granularity)
    index = 48 << 20                                         # one search index: 32 MiB cells + 4 MiB masks + sorted copies + coarse level
    assert peak_bytes <= (len(kf) + 3) * per_frame + 4 * index, (peak_bytes, len(kf))
