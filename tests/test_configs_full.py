"""BASELINE.json configs 3, 4 and 5 at full size on the GPU, against the CPU oracle, and the multi-GPU C ABI on a 1-rank
communicator.  Integer results (correspondence counts) and poses must be the oracle's; a pair registered inside a batch must
equal its single-pair registration BIT FOR BIT (the sums are order independent)."""
import ctypes as C

import numpy as np
import pytest

from slam3d_gx_b200 import synth, _abi, sharding
from oracle import oracle
from conftest import pose_close

pytestmark = pytest.mark.gpu
ROT_TOL = TRANS_TOL = 1e-4          # north_star tolerance vs the oracle
ORACLE_SAMPLE = 8                   # pairs of a 64-pair batch that are also run through the oracle


def _same_record(a, b):
    return (np.array_equal(a["T"], b["T"]) and a["inliers"] == b["inliers"] and a["status"] == b["status"]
            and a["iterations"] == b["iterations"] and a["fitness"] == b["fitness"] and a["norm"] == b["norm"])


def test_config5_map_vs_frame_full_size(ctx):
    """Config 5: 1 228 800-point fused map (4 frames in the map frame, the shape reference src/saveOutput.cpp:76-93 builds)
    vs a 307 200-point incoming frame, 50 iterations."""
    m = synth.make_map()
    assert len(m["map"]) == 1228800 and len(m["frame"]) == 307200
    prm = _abi.icp_params(50)
    tgt = ctx.upload(m["map"], m["map_normals"])
    src = ctx.upload(m["frame"])
    r = ctx.register(src, tgt, None, prm)
    nn = ctx.last_correspondences(len(m["frame"]))
    r1 = ctx.register(src, tgt, None, _abi.icp_params(1))
    nn1 = ctx.last_correspondences(len(m["frame"]))
    src.free(); tgt.free()
    o = oracle.icp(m["frame"], m["map"], m["map_normals"], params=prm, want_nn=True, nthreads=0)
    assert r["status"] == o["status"] == 0 and r["iterations"] == 50
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert np.array_equal(r["T"], o["T"]) and r["inliers"] == o["inliers"] and r["fitness"] == o["fitness"]
    assert np.array_equal(nn, o["nn"])
    idx, _ = oracle.nn(m["frame"], m["map"], None, nthreads=0)
    assert np.array_equal(nn1, idx)                       # first-iteration correspondences against the 1.2 M-point map
    ok, err = pose_close(r["T"], m["T_gt"], 1e-3, 3e-3)   # and the analytic ground truth (noise floor of sigma = 2 mm)
    assert ok, err


def _sweep_sources(n):
    """Loop-closure sweep shape (reference src/GraphicEnd.cpp:729-761): n distinct sources against ONE shared target."""
    base = synth.make_pair(0)
    C2 = synth.base_pose() @ np.linalg.inv(base["T_gt"])
    srcs, gts = [], []
    for i in range(1, n + 1):
        T = synth.random_rel_pose(synth.BASE_SEED + i)
        z, _ = synth.render_depth(C2 @ T, synth.Camera(), "S1", 0.002, synth.BASE_SEED + i, 11)
        pts, _ = synth.backproject(z, synth.Camera())
        srcs.append(pts); gts.append(T)
    return base, srcs, gts


def test_config3_shared_target_sweep_full_size(ctx):
    """Config 3: 64 sources x one shared 307 200-point target, 10 iterations, one s3d_register_batch."""
    base, srcs, gts = _sweep_sources(64)
    prm = _abi.icp_params(10)
    tgt = ctx.upload(base["tgt"], base["tgt_normals"])
    clouds = [ctx.upload(s) for s in srcs]
    res = ctx.register_batch(clouds, [tgt] * 64, None, prm)
    singles = [ctx.register(c, tgt, None, prm) for c in clouds]
    for c in clouds:
        c.free()
    tgt.free()
    assert all(r["status"] == 0 and r["iterations"] == 10 for r in res)
    for i in range(64):
        assert _same_record(res[i], singles[i]), i                      # batch == single, bit for bit
        ok, err = pose_close(res[i]["T"], gts[i], 3e-3, 8e-3)
        assert ok, (i, err)
    for i in range(0, 64, 64 // ORACLE_SAMPLE):
        o = oracle.icp(srcs[i], base["tgt"], base["tgt_normals"], params=prm, nthreads=0)
        assert np.array_equal(res[i]["T"], o["T"]) and res[i]["inliers"] == o["inliers"] and res[i]["fitness"] == o["fitness"], i


@pytest.fixture(scope="module")
def comm1(ctx):
    uid = ctx.comm_unique_id()
    assert len(uid) == 128
    comm = ctx.comm_create(uid, 1, 0)
    yield comm
    ctx.comm_destroy(comm)


def test_config4_rank_shard_through_register_batch_gather(ctx, comm1):
    """Config 4, one rank's shard: 64 independent pairs (sharding.partition(512, 8, rank)), 10 iterations, registered and
    gathered by ONE s3d_register_batch_gather on a 1-rank NCCL communicator (records packed on the device, ncclAllGather,
    one copy back); padding slots come back ABSENT."""
    rank = 3
    mine = sharding.partition(512, 8, rank)
    assert len(mine) == 64 and mine.start == 192
    pairs = [synth.make_pair(i) for i in mine]
    prm = _abi.icp_params(10)
    srcs = [ctx.upload(p["src"]) for p in pairs]
    tgts = [ctx.upload(p["tgt"], p["tgt_normals"]) for p in pairs]
    n_slot = 66
    res = ctx.register_batch_gather(comm1, srcs, tgts, 1, n_slot, None, prm)
    assert len(res) == n_slot and [r["status"] for r in res[64:]] == [_abi.PAIR_ABSENT] * 2
    plain = ctx.register_batch(srcs, tgts, None, prm)
    singles = [ctx.register(s, t, None, prm) for s, t in zip(srcs, tgts)]
    for c in srcs + tgts:
        c.free()
    for i in range(64):
        assert res[i]["status"] == 0 and res[i]["iterations"] == 10
        assert _same_record(res[i], plain[i]) and _same_record(res[i], singles[i]), i
        ok, err = pose_close(res[i]["T"], pairs[i]["T_gt"], 3e-3, 8e-3)
        assert ok, (i, err)
    for i in range(0, 64, 64 // ORACLE_SAMPLE):
        p = pairs[i]
        o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, nthreads=0)
        assert np.array_equal(res[i]["T"], o["T"]) and res[i]["inliers"] == o["inliers"] and res[i]["norm"] == o["norm"], i


def test_gather_results_and_host_alloc_exports(ctx, comm1, small_pair):
    """s3d_gather_results on hardware (1-rank communicator: the gathered records are the local ones, byte for byte), fed
    from page-locked memory of s3d_host_alloc; s3d_host_alloc rows also serve s3d_cloud_upload_async."""
    p = small_pair
    n = len(p["src"])
    ptr = ctx.host_alloc(n * 16)
    rows = np.ctypeslib.as_array((C.c_float * (n * 4)).from_address(ptr)).reshape(n, 4)
    rows[:] = p["src"]
    cs = ctx.upload_async(rows)
    ct = ctx.upload(p["tgt"], p["tgt_normals"])
    prm = _abi.icp_params(6)
    raw = ctx.register_batch([cs, cs, cs], [ct, ct, ct], None, prm, raw=True)
    allr = ctx.gather_results(comm1, raw, 1)
    assert len(allr) == 3
    for i in range(3):
        assert bytes(allr[i]) == bytes(raw[i])
    ref = ctx.register(ctx.upload(p["src"]), ct, None, prm)
    assert np.array_equal(_abi.result_to_dict(allr[0])["T"], ref["T"])
    # an empty shard still takes part in the gather
    res = ctx.register_batch_gather(comm1, [], [], 1, 2, None, prm)
    assert [r["status"] for r in res] == [_abi.PAIR_ABSENT, _abi.PAIR_ABSENT]
    cs.free(); ct.free()
    ctx.host_free(ptr)


def test_memory_stats_follow_clouds(ctx, small_pair):
    p = small_pair
    before = ctx.memory_stats()
    c = ctx.upload(p["tgt"], p["tgt_normals"])
    s = ctx.upload(p["src"])
    ctx.register(s, c, None, _abi.icp_params(2))
    with_index = ctx.memory_stats()
    c.drop_index()
    dropped = ctx.memory_stats()
    c.free(); s.free()
    after = ctx.memory_stats()
    assert with_index["live_bytes"] > before["live_bytes"] + 2 * len(p["tgt"]) * 16
    assert dropped["live_bytes"] < with_index["live_bytes"] - (1 << 20)          # the index (>= 32 MiB of cells) went back to the pool
    assert after["live_bytes"] == before["live_bytes"] and after["peak_live_bytes"] >= with_index["live_bytes"]
