"""Fixtures that pin the oracle / host shell to what the reference itself holds (run in the build container, where
/root/reference exists; the GPU box never reads /root/reference).  Everything written here is a small excerpt of the
reference's own DATA files or numbers computed by the reference's own code:

1. ref_exp1_pcd1_head4096.pcd -- the header of reference data/exp1/pcd/1.pcd (WIDTH/POINTS edited to 4096), its first
   4096 point records verbatim, and the trailing padding bytes the real file carries after its last record, verbatim.
   ref_exp1_dep1_rows.npz -- the rows of reference data/exp1/dep/1.png that hold those 4096 points (uint16).
   The PCD was produced by the reference's convert2PCD (src/convert2PCD.cpp:54-80) from that depth image and then went
   through a 7-significant-digit ASCII round trip, so it pins the back-projection to 5e-7 m (measured max 4.8e-7), not
   bit for bit; point count and order are exact.
2. ref_rpe_pins.json -- translation distance and rotation angle of a few SE(3) matrices computed by the reference's
   tools/evaluate_rpe.py functions ominus / compute_distance / compute_angle (:134-172; the file is Python 2, so exactly
   those function definitions are exec'd from its source text at generation time), plus the norm formula of reference
   src/GraphicEnd.cpp:618 evaluated from them.
3. ref_keyframe.txt, ref_lc.txt -- reference data/keyframe.txt and data/lc.txt verbatim (line formats of the output
   files, reference src/GraphicEnd.cpp:673-679,861).
"""
import json
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
N_HEAD = 4096


def pcd_excerpt():
    import cv2
    raw = open(os.path.join(REF, "data/exp1/pcd/1.pcd"), "rb").read()
    off = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    header = raw[:off].decode("ascii")
    n = int(re.search(r"POINTS (\d+)", header).group(1))
    tail = raw[off + n * 16:]                           # the real file is longer than POINTS * 16 bytes
    header = header.replace("WIDTH %d" % n, "WIDTH %d" % N_HEAD).replace("POINTS %d" % n, "POINTS %d" % N_HEAD)
    with open(os.path.join(HERE, "ref_exp1_pcd1_head4096.pcd"), "wb") as f:
        f.write(header.encode("ascii") + raw[off:off + N_HEAD * 16] + tail)
    d1 = cv2.imread(os.path.join(REF, "data/exp1/dep/1.png"), -1)
    assert d1.dtype == np.uint16 and d1.shape == (480, 640)
    nz = np.flatnonzero(d1.ravel())
    last_row = int(nz[N_HEAD - 1] // 640)
    np.savez_compressed(os.path.join(HERE, "ref_exp1_dep1_rows.npz"), rows=d1[:last_row + 1].copy(), n_points=N_HEAD,
                        n_points_full=n, data_offset=off, trailing_bytes=len(tail))
    print("pcd excerpt:", N_HEAD, "points, data offset", off, "trailing bytes", len(tail), "depth rows", last_row + 1)


def rpe_pins():
    src = open(os.path.join(REF, "tools/evaluate_rpe.py")).read().splitlines()
    # the three pure functions, located by name (python-2 file: cannot be imported under python 3)
    ns = {"numpy": np}
    for name in ("ominus", "compute_distance", "compute_angle"):
        i0 = next(i for i, ln in enumerate(src) if ln.startswith("def %s(" % name))
        i1 = next(i for i in range(i0 + 1, len(src)) if src[i].startswith("def "))
        exec("\n".join(src[i0:i1]), ns)
    from slam3d_gx_b200 import synth
    pins = []
    seeds = [(1, (0.01, 0.05), (0.01, 0.05)), (2, (0.3, 0.6), (0.1, 0.4)), (3, (3.0, 3.14), (1.0, 2.0)), (4, (1e-4, 2e-4), (1e-4, 2e-4)),
             (5, (1.5, 1.6), (0.0, 0.0))]
    for seed, rr, tr in seeds:
        A = synth.random_rel_pose(1000 + seed, rr, tr)
        B = synth.random_rel_pose(2000 + seed, rr, tr)
        E = ns["ominus"](A, B)
        dist_e, ang_e = float(ns["compute_distance"](E)), float(ns["compute_angle"](E))
        dist_a, ang_a = float(ns["compute_distance"](A)), float(ns["compute_angle"](A))
        pins.append(dict(A=A.tolist(), B=B.tolist(), rel_distance=dist_e, rel_angle=ang_e, a_distance=dist_a, a_angle=ang_a,
                         a_norm=abs(min(ang_a, 2 * np.pi - ang_a)) + 0.9 * abs(dist_a)))      # src/GraphicEnd.cpp:618
    json.dump(pins, open(os.path.join(HERE, "ref_rpe_pins.json"), "w"), indent=1)
    print("rpe pins:", len(pins))


def data_files():
    shutil.copyfile(os.path.join(REF, "data/keyframe.txt"), os.path.join(HERE, "ref_keyframe.txt"))
    shutil.copyfile(os.path.join(REF, "data/lc.txt"), os.path.join(HERE, "ref_lc.txt"))
    os.chmod(os.path.join(HERE, "ref_keyframe.txt"), 0o644)
    os.chmod(os.path.join(HERE, "ref_lc.txt"), 0o644)


if __name__ == "__main__":
    pcd_excerpt()
    rpe_pins()
    data_files()
