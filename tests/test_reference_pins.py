"""What the reference itself holds for this path, as fixtures (tests/golden/make_reference_pins.py): a verbatim excerpt of its
real PCD + the depth rows it was made from (pins the back-projection A9 and the PCD reader), numbers computed by its own
tools/evaluate_rpe.py (pin the pose-error metric and the norm formula of A8), and its shipped keyframe.txt / lc.txt (pin the
file formats the shell reads and writes).  The ICP rows A4-A6 stay unpinned: the reference has no ICP."""
import json
import os
import subprocess

import numpy as np
import pytest

from slam3d_gx_b200 import synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HOST = os.path.join(ROOT, "slam3d_gx_b200", "host")
PCD = os.path.join(GOLD, "ref_exp1_pcd1_head4096.pcd")
PCD_TOL = 5e-7          # the reference's PCD went through a 7-significant-digit ASCII round trip (measured max 4.8e-7 m)


def _build_host(target="bin/test_host"):
    subprocess.run(["make", "-C", HOST, "-s", target], check=True, env={**os.environ, "CXX": "g++", "CC": "gcc"})


def _pcd_raw():
    raw = open(PCD, "rb").read()
    off = raw.index(b"DATA binary\n") + 12
    meta = np.load(os.path.join(GOLD, "ref_exp1_dep1_rows.npz"))
    n = int(meta["n_points"])
    # the full file's data starts at byte 185; the excerpt's header is 4 characters shorter (WIDTH/POINTS 221202 -> 4096)
    assert int(meta["data_offset"]) == 185 and off == 181 and len(raw) - off - 16 * n == int(meta["trailing_bytes"]) > 0
    return np.frombuffer(raw[off:off + 16 * n], dtype=np.float32).reshape(n, 4), meta


def test_pcd_reader_on_the_reference_file_excerpt(tmp_path):
    """host/PCD.cpp on the reference's own bytes: data starting at byte 185, `rgba` column of type U, trailing padding."""
    _build_host()
    want, _ = _pcd_raw()
    out = tmp_path / "rows.bin"
    r = subprocess.run([os.path.join(HOST, "bin", "test_host"), "pcd", PCD, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf[:4], np.int32)[0])
    got = np.frombuffer(buf[4:], np.float32).reshape(n, 4)
    assert n == len(want) == 4096
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))          # bit for bit, rgba column included


def test_oracle_backprojection_against_the_reference_pcd():
    """convert2PCD (reference src/convert2PCD.cpp:54-80) produced the PCD from the depth image with the intrinsics of :19-23:
    the oracle's back-projection must give the same points in the same order."""
    want, meta = _pcd_raw()
    pts = oracle.backproject(meta["rows"], synth.Camera(height=meta["rows"].shape[0]), 0.0)
    assert len(pts) >= len(want)
    got = pts[:len(want), :3]
    assert np.abs(got - want[:, :3]).max() <= PCD_TOL
    assert np.array_equal(got[:, 2], want[:, 2])                                # z = d / 1000 is exactly representable in 7 digits
    ref_pts, _ = synth.backproject(meta["rows"], synth.Camera(height=meta["rows"].shape[0]))
    assert np.array_equal(ref_pts[:len(want), :3], got)                         # numpy restatement == C oracle, bit for bit


@pytest.mark.gpu
def test_cuda_backprojection_against_the_reference_pcd(ctx):
    want, meta = _pcd_raw()
    c = ctx.from_depth(meta["rows"], synth.Camera(height=meta["rows"].shape[0]), 0.0)
    got = c.download()["xyz"]
    c.free()
    assert np.abs(got[:len(want)] - want[:, :3]).max() <= PCD_TOL
    assert np.array_equal(got, oracle.backproject(meta["rows"], synth.Camera(height=meta["rows"].shape[0]), 0.0)[:, :3])


def test_pose_error_and_norm_against_evaluate_rpe():
    """synth.pose_error == (compute_angle, compute_distance) of reference tools/evaluate_rpe.py:162-172 applied to ominus(a, b);
    oracle_pose_norm == |min(theta, 2 pi - theta)| + 0.9 |t| of reference src/GraphicEnd.cpp:618 with theta, |t| from the same."""
    pins = json.load(open(os.path.join(GOLD, "ref_rpe_pins.json")))
    assert len(pins) >= 5
    for p in pins:
        A, B = np.array(p["A"]), np.array(p["B"])
        rot, _ = synth.pose_error(A, B)
        E = np.linalg.inv(A) @ B
        assert abs(rot - p["rel_angle"]) < 1e-7
        assert abs(np.linalg.norm(E[:3, 3]) - p["rel_distance"]) < 1e-12
        assert abs(oracle.pose_norm(A) - p["a_norm"]) < 1e-7
        I = np.eye(4)
        assert abs(synth.pose_error(I, A)[0] - p["a_angle"]) < 1e-7 and abs(synth.pose_error(I, A)[1] - p["a_distance"]) < 1e-12


def test_generate_trajectory_reads_the_reference_keyframe_file(tmp_path):
    """generateTrajectory (reference src/generateTrajectory.cpp:28-71) on the reference's own data/keyframe.txt and a g2o file
    holding those vertex ids: TUM lines `timestamp tx ty tz qx qy qz qw`, time stamp = first token of associate.txt line `frame`."""
    _build_host("bin/generateTrajectory")
    kf = [tuple(int(x) for x in ln.split()) for ln in open(os.path.join(GOLD, "ref_keyframe.txt")).read().splitlines() if ln.strip()]
    assert kf[0] == (0, 50) and all(len(k) == 2 for k in kf) and [k[0] for k in kf] == list(range(len(kf)))
    (tmp_path / "ds").mkdir()
    (tmp_path / "parameters.yaml").write_text("%YAML:1.0\ndata_source: " + str(tmp_path / "ds") + "\n")
    n_lines = max(k[1] for k in kf)
    (tmp_path / "ds" / "associate.txt").write_text("".join(f"{1000 + i}.{i % 10}5 rgb/{i}.png {1000 + i}.30 depth/{i}.png\n" for i in range(1, n_lines + 1)))
    with open(tmp_path / "final.g2o", "w") as f:
        for vid, frame in kf:
            f.write(f"VERTEX_SE3:QUAT {vid} {0.01 * vid} {-0.02 * vid} {0.5 + vid} 0 0 0.0998334166 0.995004165\n")
            if vid == 0:
                f.write("FIX 0\n")
        f.write("EDGE_SE3:QUAT 0 1 0 0 0 0 0 0 1 " + " ".join("100" if i in (0, 6, 11, 15, 18, 20) else "0" for i in range(21)) + "\n")
    r = subprocess.run([os.path.join(HOST, "bin", "generateTrajectory"), os.path.join(GOLD, "ref_keyframe.txt"), str(tmp_path / "final.g2o")],
                       cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = (tmp_path / "trajectory.txt").read_text().splitlines()
    assert len(lines) == len(kf)
    for (vid, frame), ln in zip(kf, lines):
        t = ln.split()
        assert len(t) == 8 and t[0] == f"{1000 + frame}.{frame % 10}5"
        v = [float(x) for x in t[1:]]
        assert abs(v[0] - 0.01 * vid) < 1e-5 and abs(v[2] - (0.5 + vid)) < 1e-4 and abs(v[5] - 0.0998334) < 1e-6 and abs(v[6] - 0.995004) < 1e-6


def parse_lc(path):
    """lc.txt: `frame1 frame2 norm [inliers]` (reference src/GraphicEnd.cpp:861 writes four columns; the data/lc.txt it ships, from
    an earlier revision, has the first three)."""
    rows = []
    for ln in open(path).read().splitlines():
        t = ln.split()
        if not t:
            continue
        assert len(t) in (3, 4), ln
        rows.append((int(t[0]), int(t[1]), float(t[2])) + ((int(t[3]),) if len(t) == 4 else ()))
    return rows


def test_reference_lc_file_parses():
    rows = parse_lc(os.path.join(GOLD, "ref_lc.txt"))
    assert len(rows) == 22 and rows[0][:2] == (84, 126) and abs(rows[0][2] - 0.672453) < 1e-9
    assert all(a < b and 0.0 < nrm <= 1.5 for a, b, nrm, *_ in rows)       # loop_closure_error 1.5 (parameters.yaml)
