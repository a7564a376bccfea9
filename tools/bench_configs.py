"""BASELINE.json configs 1-5 on one GPU (config 4: the 64-pair shard one rank of 8 gets): iterations/s with device-resident
clouds (CUDA events inside s3d_register_batch), pose error against the analytic ground truth, and the SURVEY 8(d) roofline
fraction.  Not the bench line (bench.py is config 2); evidence for DESIGN.md.  Usage: python tools/bench_configs.py [configs...]"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam3d_gx_b200 as s3d
from slam3d_gx_b200 import synth, _abi

want = [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 5]
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = s3d.Context(0)
out = []


def run(name, srcs, tgts, iters, gts, reps=3, bytes_per_iter_pair=None):
    prm = _abi.icp_params(iters)
    for _ in range(2):
        res = ctx.register_batch(srcs, tgts, None, prm)
    best = 1e9
    for _ in range(reps):
        res = ctx.register_batch(srcs, tgts, None, prm)
        best = min(best, ctx.last_timing()["iterate_ms"])
    errs = [synth.pose_error(r["T"], g) for r, g in zip(res, gts)]
    n_pairs = len(srcs)
    its = n_pairs * iters / (best * 1e-3)
    gbs = bytes_per_iter_pair * n_pairs * iters / (best * 1e-3) / 1e9
    rec = dict(config=name, pairs=n_pairs, iterations=iters, ms=best, iterations_per_s=its, algorithmic_GBps=gbs, frac_of_measured_hbm=gbs / peak,
               status=[r["status"] for r in res][:4], max_rot_err_vs_gt=max(e[0] for e in errs), max_trans_err_vs_gt=max(e[1] for e in errs))
    out.append(rec)
    print(json.dumps(rec), flush=True)


N = 307200
if 1 in want or 2 in want:
    p = synth.make_pair(0)
    s, t = ctx.upload(p["src"]), ctx.upload(p["tgt"], p["tgt_normals"])
    if 1 in want: run("1: single pair, 10 iterations", [s], [t], 10, [p["T_gt"]], bytes_per_iter_pair=16 * N + 32 * N)
    if 2 in want: run("2: single pair, 30 iterations", [s], [t], 30, [p["T_gt"]], bytes_per_iter_pair=16 * N + 32 * N)
    s.free(); t.free()
if 3 in want:
    # loop-closure sweep: 64 sources against ONE shared target (reference src/GraphicEnd.cpp:729-761)
    base = synth.make_pair(0)
    tgt = ctx.upload(base["tgt"], base["tgt_normals"])
    C2 = synth.base_pose() @ np.linalg.inv(base["T_gt"])
    srcs, gts = [], []
    for i in range(1, 65):
        T = synth.random_rel_pose(synth.BASE_SEED + i)
        z, _ = synth.render_depth(C2 @ T, synth.Camera(), "S1", 0.002, synth.BASE_SEED + i, 11)
        pts, _ = synth.backproject(z, synth.Camera())
        srcs.append(ctx.upload(pts)); gts.append(T)
    run("3a: 64 sources, one shared target, 10 iterations", srcs, [tgt] * 64, 10, gts, reps=2, bytes_per_iter_pair=(64 * 16 * N + 32 * N) / 64)
    for c in srcs: c.free()
    tgt.free()
if 4 in want:
    # 64 independent pairs = the shard one of 8 ranks gets in config 4
    srcs, tgts, gts = [], [], []
    for i in range(64):
        p = synth.make_pair(i)
        srcs.append(ctx.upload(p["src"])); tgts.append(ctx.upload(p["tgt"], p["tgt_normals"])); gts.append(p["T_gt"])
    run("3b/4: 64 independent pairs (one rank's shard of config 4), 10 iterations", srcs, tgts, 10, gts, reps=2, bytes_per_iter_pair=16 * N + 32 * N)
    for c in srcs + tgts: c.free()
if 5 in want:
    m = synth.make_map()
    tgt = ctx.upload(m["map"], m["map_normals"]); src = ctx.upload(m["frame"])
    run("5: 1.2M-point fused map vs 307k-point frame, 50 iterations", [src], [tgt], 50, [m["T_gt"]], bytes_per_iter_pair=16 * N + 32 * len(m["map"]))
    src.free(); tgt.free()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1) if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
