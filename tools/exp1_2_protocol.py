#!/usr/bin/env python
"""The reference's registration-error protocol (exp1_2.py:6-27 driving src/exp1/exp1_2.cpp:268-295) on synthetic pairs.

The reference draws 100 start frames, registers each against the frames 1..19 steps later and appends one line per pair to
data/exp1/error.log:   f1 f2 |t(Tr)| angle(Tr) |t(Terror)| angle(Terror) inliers      with Terror = Tr^-1 * T.
Here a "frame offset" k scales the relative motion of the pair (k times the one-step motion range of SURVEY App. B), the truth
Tr is analytic, and T comes from the CUDA path.  (`run` takes the registration as a callable, so tests/ can drive the same
protocol with the CPU oracle as a check; nothing here touches oracle/.)

    python tools/exp1_2_protocol.py --tests 10 --offsets 1 2 4 --out error_icp.log [--scale 0.25]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from slam3d_gx_b200 import synth, _abi  # noqa: E402


def error_angle(T):      # src/exp1/exp1_2.cpp:167-170
    return float(np.arccos(np.clip((np.trace(T[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)))


def cuda_register():
    import slam3d_gx_b200 as s3d
    ctx = s3d.Context(0)

    def reg(p, prm):
        src = ctx.upload(p["src"]); tgt = ctx.upload(p["tgt"], p["tgt_normals"])
        try:
            return ctx.register(src, tgt, None, prm)
        finally:
            src.free(); tgt.free()
    return reg


def run(tests, offsets, register, cam, iterations, seed0=0, out=None):
    """register(pair_dict, icp_params) -> {'T': 4x4, 'inliers': int, ...}"""
    prm = _abi.icp_params(iterations)
    lines = []
    for t in range(tests):
        for k in offsets:
            p = synth.make_pair(seed0 + t, cam=cam, rot_range=(0.01 * k, 0.05 * k), trans_range=(0.01 * k, 0.05 * k))
            r = register(p, prm)
            Tr, T = p["T_gt"], np.asarray(r["T"]).reshape(4, 4)
            Terr = np.linalg.inv(Tr) @ T
            lines.append("%d %d %.9g %.9g %.9g %.9g %d" % (t, t + k, np.linalg.norm(Tr[:3, 3]), error_angle(Tr),
                                                          np.linalg.norm(Terr[:3, 3]), error_angle(Terr), r["inliers"]))
    if out:
        with open(out, "a") as f:      # the reference appends (ofstream::app)
            f.write("\n".join(lines) + "\n")
    return lines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tests", type=int, default=10)
    ap.add_argument("--offsets", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--scale", type=float, default=1.0, help="resolution relative to 640x480 (0.25 -> 160x120)")
    ap.add_argument("--iterations", type=int, default=30)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cam = synth.Camera().scaled(a.scale)
    for line in run(a.tests, a.offsets, cuda_register(), cam, a.iterations, out=a.out):
        print(line)
