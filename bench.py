#!/usr/bin/env python
"""bench.py -- ICP iterations/s of the registration hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one frame pair of BASELINE config 2: a synthetic 640x480
RGB-D pair (307 200 x 307 200 points), 30 point-to-plane ICP iterations, the target's search index rebuilt
inside the step.  A pool of POOL distinct pairs (> 126 MB of device data, i.e. larger than L2) rotates so no
step finds its inputs in cache.  With N > 1 (torchrun, one rank per GPU) every rank registers its own pairs
(frame pairs are independent units: weak scaling, no data-path collective) and the resulting pose records are
all-gathered over NCCL inside the timed region.

value      = N * K * 30 / t        device-resident inputs, CUDA events on the launching stream, max over ranks
e2e        = same metric through the C ABI from HOST (pinned) buffers: upload of both clouds, plane extraction
             of the target (the reference flow extractPlanes -> multiPnP), registration, result read-back
roofline   = correspondence+reduction kernel: algorithmic bytes (16 N + 32 M per iteration)
             over its measured launch time, against the measured HBM peak of MEASURED_PEAKS.json
cpu_baseline = the CPU oracle (PCL-1.7-equivalent restatement; the reference's own PCL path cannot be built
             here) timed on this box's host cores on a bounded sample
--impl reference = that same CPU restatement with every host thread, as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ITERS = 30
POOL = 16
N_PTS = 307200
B_ALG = 16 * N_PTS + 32 * N_PTS          # algorithmic bytes per ICP iteration
METRIC = "ICP iterations/sec on 640x480 RGB-D clouds"
UNIT = "iterations/s"
WORKLOAD = ("config2: single synthetic 640x480 RGB-D frame pair (307200 x 307200 pts), 30 point-to-plane ICP "
            "iterations, exact NN, search index rebuilt every step")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) > 8 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": "warm-up + timed steps (same workload), nvidia-smi every 20 ms"}


def config_dict(world, pool):
    """The `config` object of the JSON line: the same for the GPU arm and the reference arm (same workload)."""
    return {"workload": WORKLOAD, "pairs_per_step_per_gpu": 1, "icp_iterations": ITERS, "points": [N_PTS, N_PTS],
            "pool_pairs_per_gpu": pool, "l2": "inputs larger than L2: a pool of %d pairs (>%d MB) rotates" % (pool, pool * 40),
            "parallelism": "pairs sharded over %d GPU(s), NCCL all_gather of pose records" % world}


def host_threads():
    """Every hardware thread this process may run on.  torchrun exports OMP_NUM_THREADS=1; the CPU arm must not inherit that."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args, rank):
    """Reference arm: the CPU restatement of the reference's PCL path with all host threads (rank 0 only)."""
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)       # before the OpenMP runtime of the oracle library starts
    from slam3d_gx_b200 import synth, _abi
    from oracle import oracle
    assert threads > 1 or (os.cpu_count() or 1) == 1, "reference arm would run single-threaded on a multi-core box"
    pairs = [synth.make_pair(i) for i in range(min(2, max(1, args.steps)))]
    prm = _abi.icp_params(ITERS)
    for w in range(args.warmup):
        p = pairs[w % len(pairs)]
        oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=_abi.icp_params(3), nthreads=threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        p = pairs[s % len(pairs)]
        r = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, nthreads=threads)
        assert r["status"] == 0
    dt = time.perf_counter() - t0
    value = args.steps * ITERS / dt
    sample = f"{args.steps} step(s) x 1 pair x {ITERS} iterations (KD-tree build included), OpenMP over source points"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / max(1, args.steps) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus, args.pool),
        "note": "CPU restatement of the reference's PCL-1.7 ICP (oracle/icp_oracle.c; the reference itself cannot be built: no PCL), rank 0 only",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch-regime", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--config4-pairs", type=int, default=64, help="pairs per GPU of the config-4 leg (512 / 8)")
    ap.add_argument("--config4-steps", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 40:     # defaults sized for the GPU arm; keep the CPU arm bounded
            args.steps, args.warmup = 4, 1
        run_reference(args, rank)
        return
    # the contract is ONE JSON line on stdout: native libraries (NCCL prints its version) write to fd 1 directly, so
    # fd 1 is pointed at stderr for the run and the JSON line goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import slam3d_gx_b200 as s3d
    from slam3d_gx_b200 import synth, _abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = s3d.Context(local_rank)
    # an explicit (non-default) torch stream: its handle is what the library launches on, so the CUDA events recorded on it
    # below bracket the kernels themselves (handle 0, torch's default stream, would mean "the ctx's own stream" to the ABI)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    W = max(3, args.warmup)

    # ---- inputs: every rank owns its own pool of pairs (pair index = rank*pool + k) ---------------------
    host, src, tgt = [], [], []
    for k in range(args.pool):
        p = synth.make_pair(rank * args.pool + k)
        host.append(p)
        src.append(ctx.upload(p["src"]))
        tgt.append(ctx.upload(p["tgt"], p["tgt_normals"]))
    prm = _abi.icp_params(ITERS, reuse_index=0)
    from slam3d_gx_b200 import sharding
    rec_bytes = _abi.RESULT_BYTES

    # Pose gather over NCCL: the only collective of the path, through the product's own C ABI
    # (s3d_comm_create / s3d_gather_results: persistent device + page-locked buffers inside the ctx).  Like config 4 (a batch
    # of pairs sharded over the ranks, ONE all-gather of the pose records when the batch is done), the records of a rank's
    # steps are all-gathered in one NCCL call at the end of the timed region, inside it.  (A per-step NCCL kernel cannot run
    # beside the registration kernel, which holds every SM's register file; it would serialise the ranks on each other.)
    uid = [s3d.Context.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = ctx.comm_create(uid[0], world, rank)
    records = []

    def step(i):
        k = i % args.pool
        res = ctx.register_batch([src[k]], [tgt[k]], None, prm, raw=True)
        if world > 1:
            records.append(res[0])
        return res[0]

    def finish_gathers():
        if world == 1 or not records:
            return 0
        local = (_abi.Result * len(records))(*records)
        allr = ctx.gather_results(comm, local, world)
        records.clear()
        return len(allr)

    for k in range(args.pool):      # prime every pair once (first-use device allocations of its index), untimed
        step(k)
    # clocks / throttle reasons: nvidia-smi sampled every 20 ms from the warm-up steps (the same workload) to the end of the timed region
    sampler = ClockSampler(local_rank)          # every rank samples its own GPU
    sampler.start()
    time.sleep(0.25)                # nvidia-smi needs a moment to start; its first rows then fall into the warm-up
    for i in range(W):
        step(i)
    finish_gathers()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    launches0 = ctx.launch_count
    iter_ms, index_ms, iter_launches = 0.0, 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    last = None
    # The timed steps are ENQUEUED (s3d_register_enqueue: index rebuild + registration + result record formed on the device) and
    # drained every ASYNC_DEPTH steps (s3d_register_drain: one device-to-host copy of the records and the CUDA-event times of
    # every step): no host round trip per step, the per-launch kernel times still cover every launch of the timed region.
    from slam3d_gx_b200.binding import ASYNC_DEPTH

    def drain():
        nonlocal iter_ms, index_ms, iter_launches, last
        res, tms = ctx.register_drain(raw=True)
        for tm in tms:
            iter_ms += tm["iterate_ms"]; index_ms += tm["index_ms"]; iter_launches += tm["iter_launches"]
        for r in res:
            assert r.status == 0, "registration failed inside the timed region"
            if world > 1:
                records.append(r)
        if len(res):
            last = res[len(res) - 1]

    for i in range(args.steps):
        k = (W + i) % args.pool
        ctx.register_enqueue(src[k], tgt[k], None, prm)
        if (i + 1) % ASYNC_DEPTH == 0:
            drain()
    drain()
    if world > 1:
        assert finish_gathers() == world * args.steps
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    assert last.status == 0, "registration failed inside the timed region"

    # ---- e2e: host buffers through the C ABI (upload + plane extraction + registration + read-back) ------
    pin = []
    for k in range(args.pool):          # the same pool of pairs as the device-resident arm
        a = torch.from_numpy(host[k]["src"].copy()).pin_memory()
        b = torch.from_numpy(host[k]["tgt"].copy()).pin_memory()
        pin.append((a, b))
    plane_prm = _abi.plane_params()
    e2e_prm = _abi.icp_params(ITERS, reuse_index=0)

    # Every step uploads its own two clouds from pinned host memory, extracts the target's planes, registers, and its result
    # record and plane list are copied back to page-locked host memory right behind it.  The step is ISSUED without waiting
    # (s3d_segment_planes_enqueue + s3d_register_enqueue: labels, normals, index, 30 iterations and the record are formed on the
    # device in stream order); the host collects the results of the last E2E_DRAIN steps in one go (s3d_*_drain).  The upload
    # of step i+1 is issued (s3d_cloud_upload_async, the ctx copy stream) before the compute of step i, so the copy engine works
    # while the SMs do: one upload per step, all of them inside the timed region.
    E2E_DRAIN = 16
    e2e_t = {"upload_issue": 0.0, "plane_extraction_issue": 0.0, "register_issue": 0.0, "free": 0.0, "drain": 0.0,
             "index_build_dev": 0.0, "iterations_dev": 0.0}
    e2e_out = {"records": 0, "planes": 0}

    def e2e_upload(i):
        a, b = pin[i % len(pin)]
        t0 = time.perf_counter()
        cs, ct = ctx.upload_async(a.numpy()), ctx.upload_async(b.numpy())
        e2e_t["upload_issue"] += time.perf_counter() - t0
        return cs, ct

    def e2e_compute(cs, ct):
        t0 = time.perf_counter()
        ct.segment_planes_enqueue(plane_prm)
        t1 = time.perf_counter()
        ctx.register_enqueue(cs, ct, None, e2e_prm)
        t2 = time.perf_counter()
        cs.release(); ct.release()       # back to the context's stream-ordered pool without waiting: the next upload reuses the buffers behind this step
        t3 = time.perf_counter()
        e2e_t["plane_extraction_issue"] += t1 - t0; e2e_t["register_issue"] += t2 - t1; e2e_t["free"] += t3 - t2

    def e2e_drain():
        t0 = time.perf_counter()
        planes_all = ctx.planes_drain()
        res, tms = ctx.register_drain(raw=True)
        e2e_t["drain"] += time.perf_counter() - t0
        assert len(planes_all) == len(res) == len(tms)
        for pl, r, tm in zip(planes_all, res, tms):
            assert r.status == 0 and len(pl) == 3, "e2e: a registration failed or a plane is missing"
            e2e_t["index_build_dev"] += tm["index_ms"] * 1e-3; e2e_t["iterations_dev"] += tm["iterate_ms"] * 1e-3
        e2e_out["records"] += len(res); e2e_out["planes"] += sum(len(pl) for pl in planes_all)

    # one continuous pipeline: warm-up steps (which also let the index build capture its few launch graphs, one per recurring
    # set of pool buffers) run straight into the timed steps; every timed step issues exactly one upload (of its successor)
    e2e_steps = max(4, min(args.steps, 64))
    e2e_warm = 8
    nxt = e2e_upload(0)
    t0 = None
    for i in range(e2e_warm + e2e_steps):
        if i == e2e_warm:
            e2e_drain()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            for k in e2e_t:
                e2e_t[k] = 0.0
            e2e_out["records"] = e2e_out["planes"] = 0
            t0 = time.perf_counter()
        cs, ct = nxt
        nxt = e2e_upload(i + 1)
        e2e_compute(cs, ct)
        if i >= e2e_warm and (i - e2e_warm + 1) % E2E_DRAIN == 0:
            e2e_drain()
    e2e_drain()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    nxt[0].free(); nxt[1].free()
    assert e2e_out["records"] == e2e_steps and e2e_out["planes"] == 3 * e2e_steps
    h2d = int(pin[0][0].numel() * 4 + pin[0][1].numel() * 4)
    d2h = int(rec_bytes + 528)          # the pair's record + the final state of the plane extraction's device loop (planes, counts)

    # ---- plane extraction (A2): 16 N bytes per evaluation pass over the CUDA-event time of plane_eval_kernel --------
    planes_roofline = None
    if world == 1:
        cp = ctx.upload(host[0]["tgt"])
        timed_prm = _abi.plane_params(timed=True)        # per-pass events (the default call replays a CUDA graph and times only the whole)
        for _ in range(3):
            cp.segment_planes(timed_prm)
        ev, tot, pts = [], [], 0
        for _ in range(10):
            cp.segment_planes(timed_prm)
            tmp = ctx.last_plane_timing()
            ev.append(tmp["eval_ms"]); tot.append(tmp["total_ms"]); pts = tmp["points_scanned"] * tmp["eval_passes_per_round"]
        cp.free()
        planes_roofline = {"kernel": "plane_eval_kernel (64 RANSAC candidates per pass over the remaining points)", "bound": "hbm",
                           "algorithmic_bytes": pts * 16, "eval_us": float(np.median(ev)) * 1e3, "extraction_device_us": float(np.median(tot)) * 1e3,
                           "achieved": pts * 16 / (float(np.median(ev)) * 1e-3) / 1e9, "unit": "GB/s", "passes": int(tmp["rounds"] * tmp["eval_passes_per_round"])}

    # ---- the bandwidth-bound regime of the same kernel: late iterations of a batch whose working set exceeds L2 ----------
    # (every query keeps its correspondence: an iteration is ONE streaming pass over source, correspondence, normal + bound
    # and a flag byte: 49 B per query).  time per late iteration = (t(40 iterations) - t(10 iterations)) / 30.
    batch_regime = None
    if world == 1 and not args.no_batch_regime and args.pool >= 8:
        nb = min(args.pool, 16)

        def batch_ms(iters):
            prm_b = _abi.icp_params(iters)
            best = None
            for _ in range(3):
                rb = ctx.register_batch(src[:nb], tgt[:nb], None, prm_b, raw=True)
                ms = ctx.last_timing()["iterate_ms"]
                best = ms if best is None else min(best, ms)
            assert all(x.status == 0 for x in rb)
            return best

        t10, t40 = batch_ms(10), batch_ms(40)
        per_it_s = (t40 - t10) / 30.0 * 1e-3
        batch_regime = {"pairs": nb, "t10_ms": t10, "t40_ms": t40, "late_iteration_us": per_it_s * 1e6,
                        "algorithmic_GBps": nb * B_ALG / per_it_s / 1e9, "read_GBps": nb * N_PTS * 49 / per_it_s / 1e9,
                        "bytes_read_per_query": 49}

    # ---- config 4 (BASELINE.json configs[3]): 512*N/8 independent pairs block-partitioned over the N ranks, 10 iterations,
    # ONE s3d_register_batch_gather per step and rank: the shard goes through the persistent kernel as one batch, the pose
    # records are formed on the device in the all-gather send buffer, ncclAllGather follows on the same stream, one
    # device-to-host copy returns the world's records.  Search indices rebuilt inside the step (like the headline).
    config4 = None
    if not args.no_config4:
        C4_ITERS, per_rank = 10, args.config4_pairs
        n_total = per_rank * world
        mine = sharding.partition(n_total, world, rank)
        c4_src, c4_tgt = [], []
        for i in mine:
            if i - mine.start < len(src) and rank * args.pool + (i - mine.start) == i:
                k = i - mine.start                       # the headline pool already holds this pair (single-GPU run)
                c4_src.append(src[k]); c4_tgt.append(tgt[k])
            else:
                p = synth.make_pair(i)
                c4_src.append(ctx.upload(p["src"])); c4_tgt.append(ctx.upload(p["tgt"], p["tgt_normals"]))
        c4_prm = _abi.icp_params(C4_ITERS, reuse_index=0)
        n_slot = max(len(sharding.partition(n_total, world, r)) for r in range(world))
        c4_steps, c4_warm = args.config4_steps, 2
        for _ in range(c4_warm):
            allr = ctx.register_batch_gather(comm, c4_src, c4_tgt, world, n_slot, None, c4_prm, raw=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        c4_l0 = ctx.launch_count
        c4_sampler = ClockSampler(local_rank)
        c4_sampler.start()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        c4_dev_ms = 0.0
        for _ in range(c4_steps):
            allr = ctx.register_batch_gather(comm, c4_src, c4_tgt, world, n_slot, None, c4_prm, raw=True)
            tm = ctx.last_timing()
            c4_dev_ms += tm["iterate_ms"] + tm["index_ms"]
        c1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        c4_ms = c0.elapsed_time(c1)
        c4_clocks = c4_sampler.stop()
        ok = [allr[r * n_slot + k].status for r in range(world) for k in range(len(sharding.partition(n_total, world, r)))]
        assert len(ok) == n_total and all(st == 0 for st in ok), "config 4: a pair failed or a record is missing"
        # every rank holds every record: pair 0's pose as rank 0 computed it must be what this rank received
        config4 = {"ms": c4_ms, "dev_ms": c4_dev_ms, "launches": ctx.launch_count - c4_l0, "steps": c4_steps, "pairs_total": n_total,
                   "pairs_per_gpu": per_rank, "iterations": C4_ITERS, "T0": [allr[0].T[k] for k in range(12)], "clocks": c4_clocks}

    # ---- max over ranks --------------------------------------------------------------------------------
    t = torch.tensor([elapsed_ms, e2e_s * 1e3, iter_ms, config4["ms"] if config4 else 0.0], dtype=torch.float64, device="cuda")
    tmin = torch.tensor([elapsed_ms, config4["ms"] if config4 else 0.0, config4["dev_ms"] if config4 else 0.0], dtype=torch.float64, device="cuda")
    tmax2 = tmin.clone()
    cnt = torch.tensor([float(launches), float(config4["launches"]) if config4 else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tmax2, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        if config4:
            t0c = torch.tensor(config4["T0"], dtype=torch.float64, device="cuda")
            t0ref = t0c.clone()
            dist.broadcast(t0ref, src=0)
            assert torch.equal(t0c, t0ref), "config 4: gathered records differ between ranks"
    per_rank = [{"rank": rank, "elapsed_ms": elapsed_ms, "device_ms": iter_ms + index_ms, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons"),
                 "config4_ms": config4["ms"] if config4 else None, "config4_device_ms": config4["dev_ms"] if config4 else None,
                 "config4_sm_mhz": config4["clocks"].get("sm_mhz") if config4 else None,
                 "config4_reasons": config4["clocks"].get("reasons") if config4 else None}]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank[0])
        per_rank = gathered
    elapsed_ms, e2e_ms, iter_ms_max, c4_ms_max = [float(x) for x in t.tolist()]
    total_launches = int(cnt[0].item())

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * args.steps * ITERS / (elapsed_ms * 1e-3)
        e2e_value = world * e2e_steps * ITERS / (e2e_ms * 1e-3)
        # dominant kernel: icp_persist_kernel -- ALL iterations of a registration in two back-to-back cooperative launches of its
        # two instances (search-only for the leading full-search iterations, general for the rest); the CUDA events of every
        # registration of the timed region bracket both, so "launch" below is that pair = one registration of ITERS iterations
        launches_per_registration = iter_launches / max(1, args.steps)
        per_launch_s = (iter_ms * 1e-3) / max(1, args.steps)
        units_per_launch = float(ITERS)
        achieved = B_ALG * units_per_launch / per_launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(world, args.pool),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "host_ms_per_step": {k: v / e2e_steps * 1e3 for k, v in e2e_t.items()}, "includes": "pinned-host upload of both clouds (step i+1's copy overlaps step i's compute), RANSAC plane extraction of the target, index build, 30 iterations, copy of the result record and the planes to page-locked host memory behind every step; steps are issued without waiting (s3d_segment_planes_enqueue, s3d_register_enqueue) and collected every %d steps (s3d_*_drain)" % E2E_DRAIN},
            "gpu_launches": total_launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "icp_persist_kernel<point_to_plane> (all 30 iterations of a registration: the search-only instance for the leading full-search iterations + the general instance, back to back)", "peak_source": peak_src,
                         "algorithmic_bytes_per_iteration": B_ALG, "iterations_per_launch": units_per_launch,
                         "cooperative_launches_per_registration": launches_per_registration,
                         "avg_launch_us": per_launch_s * 1e6,
                         "note": "achieved = algorithmic bytes (16N+32M per iteration, SURVEY.md 8d) / measured kernel time of a registration (CUDA events around its launches, every registration of the timed region); the single-pair working set (~45 MB) is L2 resident, so the kernel is bound by search issue slots and the per-iteration group sum + solve, not by HBM"},
            "breakdown_ms_per_step": {"index_build": index_ms / args.steps, "iterations": iter_ms / args.steps},
            "elapsed_ms_ranks": {"min": float(tmin[0].item()), "max": float(tmax2[0].item())},
            "per_rank": per_rank,
        }
        if config4:
            c4_its = config4["pairs_total"] * config4["iterations"] * config4["steps"] / (c4_ms_max * 1e-3)
            out["config4"] = {
                "workload": "config4: %d independent 640x480 pairs block-partitioned over %d GPU(s) (%d per GPU), 10 point-to-plane ICP iterations, "
                            "one s3d_register_batch_gather per rank and step (search indices rebuilt, records packed on the device, ncclAllGather, one D2H)"
                            % (config4["pairs_total"], world, config4["pairs_per_gpu"]),
                "iterations_per_s": c4_its, "ms_per_step": c4_ms_max / config4["steps"], "steps": config4["steps"],
                "pairs_total": config4["pairs_total"], "pairs_per_gpu": config4["pairs_per_gpu"], "icp_iterations": config4["iterations"],
                "elapsed_ms_ranks": {"min": float(tmin[1].item()), "max": float(tmax2[1].item())},
                "device_ms_ranks": {"min": float(tmin[2].item()), "max": float(tmax2[2].item())},
                "gathered_records": config4["pairs_total"], "record_bytes": rec_bytes, "gpu_launches": int(cnt[1].item()),
                "roofline_frac_algorithmic": B_ALG * config4["pairs_per_gpu"] * config4["iterations"] * config4["steps"] / (float(tmax2[2].item()) * 1e-3) / 1e9 / peak,
            }
        if planes_roofline:
            planes_roofline["peak"] = peak
            planes_roofline["frac"] = planes_roofline["achieved"] / peak
            out["roofline_planes"] = planes_roofline
        if batch_regime:
            batch_regime["frac_algorithmic"] = batch_regime["algorithmic_GBps"] / peak
            batch_regime["frac_read"] = batch_regime["read_GBps"] / peak
            batch_regime["note"] = ("same kernel, %d-pair batch (working set > L2), iterations 10..40 where every query keeps its correspondence: "
                                    "the regime in which the path is HBM bound; not the headline workload" % batch_regime["pairs"])
            out["roofline_batch_regime"] = batch_regime
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle          # checker / baseline only
            p = host[0]
            t0 = time.perf_counter()
            ro = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=_abi.icp_params(ITERS), nthreads=1)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": ITERS / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": "1 pair x 30 iterations (KD-tree build included), single thread, oracle/icp_oracle.c"}
            T_gpu = np.array(list(last.T)).reshape(4, 4)
            # last timed step used pair index (W+steps-1) % pool; compare pair 0 separately
            r0 = ctx.register_batch([src[0]], [tgt[0]], None, prm)[0]
            rot, trans = synth.pose_error(r0["T"], ro["T"])
            out["parity_vs_oracle"] = {"rot_rad": rot, "trans_m": trans, "tolerance": 1e-4}
        else:
            out["cpu_baseline"] = None
        real_stdout.write(json.dumps(out) + "\n")
        real_stdout.flush()
    ctx.comm_destroy(comm)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
