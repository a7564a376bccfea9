"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bar: bit-exact for integer/index results (correspondences, inlier counts,
labels, flags) at EVERY iteration; the north star's pose tolerance is 1e-4 rad / 1e-4 m, and because the sums
are order independent (fixed-point accumulation) and the small solvers run in strict double with the oracle's
operation order, the poses, fitness values and plane coefficients are in fact equal bit for bit."""
import numpy as np
import pytest

from slam3d_gx_b200 import synth, _abi
from oracle import oracle
from conftest import pose_close

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-4   # rad, north_star
TRANS_TOL = 1e-4  # m, north_star


def _gpu_icp(ctx, p, prm, guess=None, normals="analytic"):
    src = ctx.upload(p["src"])
    tgt = ctx.upload(p["tgt"], p["tgt_normals"] if normals == "analytic" else None)
    try:
        r = ctx.register(src, tgt, guess, prm)
        nn = ctx.last_correspondences(len(p["src"])) if prm.max_iterations > 0 else None
    finally:
        src.free(); tgt.free()
    return r, nn


@pytest.mark.parametrize("search", [_abi.SEARCH_GRID, _abi.SEARCH_BRUTE, _abi.SEARCH_GRID_LANE])
def test_correspondences_bit_exact(ctx, small_pair, search):
    p = small_pair
    prm = _abi.icp_params(1, search=search)
    r, nn = _gpu_icp(ctx, p, prm)
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, want_nn=True)
    assert np.array_equal(nn, o["nn"])
    assert r["inliers"] == o["inliers"] and r["status"] == o["status"] == 0
    assert r["fitness"] == o["fitness"] and np.array_equal(r["T"], o["T"])


@pytest.mark.parametrize("stride", [3, 4])
def test_async_upload_pipeline_matches_blocking_upload(ctx, small_pair, stride):
    """s3d_cloud_upload_async: frame k+1 is uploaded on the copy stream while frame k is registered; every result is
    bitwise the one of the blocking upload (correspondences, pose, downloaded points)."""
    import torch
    p = small_pair
    prm = _abi.icp_params(6)
    ref, ref_nn = _gpu_icp(ctx, p, prm)

    def rows(a):
        out = np.zeros((len(a), stride), np.float32)
        out[:, :3] = a[:, :3]
        if stride == 4:
            out[:, 3] = 123.0      # PCD rgba column: ignored
        return torch.from_numpy(out).pin_memory().numpy()

    hs, ht = rows(p["src"]), rows(p["tgt"])
    nxt = (ctx.upload_async(hs), ctx.upload_async(ht))
    for k in range(3):
        cs, ct = nxt
        if k < 2:
            nxt = (ctx.upload_async(hs), ctx.upload_async(ht))      # in flight during the registration below
        if k == 0:
            assert np.array_equal(cs.download()["xyz"], p["src"][:, :3])
        ct.set_normals(p["tgt_normals"])
        r = ctx.register(cs, ct, None, prm)
        nn = ctx.last_correspondences(len(p["src"]))
        cs.free(); ct.free()
        assert np.array_equal(nn, ref_nn)
        assert np.array_equal(r["T"], ref["T"]) and r["inliers"] == ref["inliers"]
    c = ctx.upload_async(hs)
    c.wait()
    c.free()                                                       # freed without ever being used


@pytest.mark.parametrize("cell", [0.0, 0.005, 0.05, 0.5, 5.0])
def test_grid_search_exact_for_any_cell_size(ctx, small_pair, cell):
    """Exactness must not depend on the grid resolution (ring expansion / clamping)."""
    p = small_pair
    prm = _abi.icp_params(1, grid_cell=cell)
    _, nn = _gpu_icp(ctx, p, prm, guess=p["T_gt"])
    idx, _ = oracle.nn(p["src"], p["tgt"], p["T_gt"])
    assert np.array_equal(nn, idx)


def test_grid_search_far_apart_clouds(ctx, small_pair):
    """Source far outside the target's bounding box: queries clamp to border cells and still find the exact NN."""
    p = small_pair
    T = synth.make_T(np.eye(3), [3.0, -2.0, 5.0])
    prm = _abi.icp_params(1)
    _, nn = _gpu_icp(ctx, p, prm, guess=T)
    idx, _ = oracle.nn(p["src"], p["tgt"], T)
    assert np.array_equal(nn, idx)


@pytest.mark.parametrize("est", [_abi.ESTIMATOR_POINT_TO_PLANE, _abi.ESTIMATOR_SVD])
@pytest.mark.parametrize("search", [_abi.SEARCH_GRID, _abi.SEARCH_BRUTE, _abi.SEARCH_GRID_LANE])
def test_icp_pose_parity_small(ctx, small_pair, est, search):
    p = small_pair
    prm = _abi.icp_params(10, estimator=est, search=search)
    r, nn = _gpu_icp(ctx, p, prm)
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, want_nn=True)
    assert r["status"] == 0 and r["iterations"] == 10
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert np.array_equal(nn, o["nn"]) and r["inliers"] == o["inliers"]
    assert np.array_equal(r["T"], o["T"]) and r["norm"] == o["norm"] and r["fitness"] == o["fitness"]


def test_grid_and_brute_agree_bitwise(ctx, small_pair):
    """Three independent exact searches: identical correspondences and identical pose bits after 6 iterations (the sums do
    not depend on the order in which a kernel adds them)."""
    p = small_pair
    a, nna = _gpu_icp(ctx, p, _abi.icp_params(6, search=_abi.SEARCH_GRID))
    b, nnb = _gpu_icp(ctx, p, _abi.icp_params(6, search=_abi.SEARCH_BRUTE))
    c, nnc = _gpu_icp(ctx, p, _abi.icp_params(6, search=_abi.SEARCH_GRID_LANE))
    assert np.array_equal(nna, nnb) and np.array_equal(nnc, nnb)
    assert np.array_equal(c["T"], b["T"]) and np.array_equal(a["T"], b["T"])
    assert a["inliers"] == b["inliers"] == c["inliers"] and a["fitness"] == b["fitness"] == c["fitness"]


@pytest.mark.parametrize("k", [2, 3, 5, 9, 16])
def test_skip_logic_is_exact(ctx, small_pair, k):
    """Iteration k of a long run (correspondences kept by the triangle-inequality test, searches started from hints)
    must equal a fresh first iteration started from the pose after k-1 iterations (every query searched from scratch):
    same correspondences, same pose bits."""
    p = small_pair
    a, nna = _gpu_icp(ctx, p, _abi.icp_params(k))
    b, _ = _gpu_icp(ctx, p, _abi.icp_params(k - 1))
    c, nnc = _gpu_icp(ctx, p, _abi.icp_params(1), guess=b["T"])
    assert np.array_equal(nna, nnc)
    assert np.array_equal(a["T"], c["T"])
    idx, _ = oracle.nn(p["src"], p["tgt"], b["T"])
    assert np.array_equal(nna, idx)


def test_enqueue_drain_matches_blocking_calls(ctx, small_cam, small_pair):
    """s3d_register_enqueue / s3d_register_drain (a stream of registrations, no host round trip per pair): records in enqueue
    order, bit-identical to s3d_register_pair and to the oracle, per-pair device timings, S3D_E_STATE beyond the depth, a
    failing pair in the middle keeps its slot, and an empty drain returns nothing."""
    import slam3d_gx_b200 as s3d
    from slam3d_gx_b200.binding import ASYNC_DEPTH
    assert ctx.register_drain() == ([], [])
    pairs = [small_pair] + [synth.make_pair(k, cam=small_cam) for k in (1, 2)]
    prm = _abi.icp_params(6, reuse_index=0)
    clouds = [(ctx.upload(p["src"]), ctx.upload(p["tgt"], p["tgt_normals"])) for p in pairs]
    few = ctx.upload(pairs[0]["src"][:2])                       # 2 points: S3D_PAIR_FEW_CORRESPONDENCES
    try:
        ref = [ctx.register(cs, ct, None, prm) for cs, ct in clouds]
        o = oracle.icp(pairs[1]["src"], pairs[1]["tgt"], pairs[1]["tgt_normals"], params=prm)
        assert np.array_equal(ref[1]["T"], o["T"]) and ref[1]["inliers"] == o["inliers"]
        order = [0, 1, 2, 1, 0]
        for k in order[:2]:
            ctx.register_enqueue(*clouds[k], None, prm)
        ctx.register_enqueue(few, clouds[0][1], None, prm)
        for k in order[2:]:
            ctx.register_enqueue(*clouds[k], None, prm)
        res, tms = ctx.register_drain()
        assert len(res) == len(tms) == 6
        got = res[:2] + res[3:]
        for k, r in zip(order, got):
            assert r["status"] == 0 and np.array_equal(r["T"], ref[k]["T"]) and r["inliers"] == ref[k]["inliers"]
            assert r["fitness"] == ref[k]["fitness"] and r["norm"] == ref[k]["norm"] and r["iterations"] == 6
        assert res[2]["status"] == _abi.PAIR_FEW and np.array_equal(res[2]["T"], np.eye(4))
        assert all(t["iterate_ms"] > 0 and t["index_ms"] > 0 and t["iter_launches"] >= 1 for t in tms)
        assert ctx.register_drain() == ([], [])
        # depth limit
        for _ in range(ASYNC_DEPTH):
            ctx.register_enqueue(*clouds[0], None, _abi.icp_params(1))
        with pytest.raises(s3d.S3DError):
            ctx.register_enqueue(*clouds[0], None, _abi.icp_params(1))
        res, _ = ctx.register_drain()
        assert len(res) == ASYNC_DEPTH and all(r["status"] == 0 for r in res)
        one = ctx.register(*clouds[0], None, _abi.icp_params(1))
        assert all(np.array_equal(r["T"], one["T"]) for r in res)
    finally:
        few.free()
        for cs, ct in clouds:
            cs.free(); ct.free()


def test_icp_is_deterministic(ctx, small_pair):
    p = small_pair
    a, _ = _gpu_icp(ctx, p, _abi.icp_params(8))
    b, _ = _gpu_icp(ctx, p, _abi.icp_params(8))
    assert np.array_equal(a["T"], b["T"]) and a["inliers"] == b["inliers"]


@pytest.mark.parametrize("kw", [dict(quantize=True), dict(holes=0.3), dict(quantize=True, holes=0.25),
                                dict(rot_range=(0.08, 0.1), trans_range=(0.08, 0.1))])
def test_icp_parity_ragged_and_quantised(ctx, small_cam, kw):
    p = synth.make_pair(7, cam=small_cam, **kw)
    assert "holes" not in kw or len(p["src"]) != len(p["tgt"])
    prm = _abi.icp_params(10, max_corr_dist=0.25)
    r, nn = _gpu_icp(ctx, p, prm)
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, want_nn=True)
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert r["status"] == o["status"]
    assert np.array_equal(nn, o["nn"]) and np.array_equal(r["T"], o["T"]) and r["inliers"] == o["inliers"]


def test_icp_guess_and_gate(ctx, small_pair):
    p = small_pair
    prm = _abi.icp_params(5, max_corr_dist=0.03)
    r, nn = _gpu_icp(ctx, p, prm, guess=p["T_gt"])
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], guess=p["T_gt"], params=prm, want_nn=True)
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert r["inliers"] == o["inliers"] and np.array_equal(nn, o["nn"]) and np.array_equal(r["T"], o["T"])
    assert (nn == -1).sum() > 0 or r["inliers"] == len(p["src"])


def test_failure_conventions(ctx, small_cam, small_pair):
    # single plane: rank-deficient normal equations -> status DEGENERATE and T exactly identity
    p0 = synth.make_pair(3, cam=small_cam, scene="S0")
    r, _ = _gpu_icp(ctx, p0, _abi.icp_params(10))
    o = oracle.icp(p0["src"], p0["tgt"], p0["tgt_normals"], params=_abi.icp_params(10))
    assert r["status"] == o["status"] == _abi.PAIR_DEGENERATE
    assert np.array_equal(r["T"], np.eye(4))
    # nothing within the gate -> FEW
    prm = _abi.icp_params(10, max_corr_dist=1e-7)
    r, _ = _gpu_icp(ctx, small_pair, prm)
    assert r["status"] == _abi.PAIR_FEW and np.array_equal(r["T"], np.eye(4)) and r["norm"] == 0.0
    # zero iterations: the guess comes back untouched
    r, _ = _gpu_icp(ctx, small_pair, _abi.icp_params(0), guess=small_pair["T_gt"])
    assert r["status"] == 0 and r["iterations"] == 0 and np.allclose(r["T"], small_pair["T_gt"], atol=0)


def test_tiny_and_empty_clouds(ctx):
    import slam3d_gx_b200 as s3d
    rng = np.random.default_rng(0)
    tgt = rng.normal(size=(5, 3)).astype(np.float32)
    nrm = rng.normal(size=(5, 3)).astype(np.float32)
    src = rng.normal(size=(2, 3)).astype(np.float32)
    cs, ct = ctx.upload(src), ctx.upload(tgt, nrm)
    r = ctx.register(cs, ct, None, _abi.icp_params(3))
    assert r["status"] == _abi.PAIR_FEW                         # 2 correspondences < 3
    ce = ctx.upload(np.zeros((0, 3), np.float32))
    r = ctx.register(ce, ct, None, _abi.icp_params(3))
    assert r["status"] == _abi.PAIR_FEW and r["inliers"] == 0
    cte = ctx.upload(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    r = ctx.register(cs, cte, None, _abi.icp_params(3))
    assert r["status"] == _abi.PAIR_FEW
    # point-to-plane without target normals is an API state error, not a silent fallback
    ctn = ctx.upload(tgt)
    with pytest.raises(s3d.S3DError):
        ctx.register(cs, ctn, None, _abi.icp_params(3))
    for c in (cs, ct, ce, cte, ctn):
        c.free()


def test_batch_shared_target_matches_single(ctx, small_cam):
    """Loop-closure sweep shape (reference src/GraphicEnd.cpp:729-761): many sources, one shared target."""
    base = synth.make_pair(0, cam=small_cam)
    tgt = ctx.upload(base["tgt"], base["tgt_normals"])
    C2 = synth.base_pose() @ np.linalg.inv(base["T_gt"])
    srcs, gts = [], []
    for i in range(1, 9):
        T = synth.random_rel_pose(synth.BASE_SEED + 100 + i)
        C1 = C2 @ T                                              # X2 = T X1  =>  C1 = C2 T
        z, _ = synth.render_depth(C1, small_cam, "S1", 0.002, synth.BASE_SEED + i, 40)
        pts, _ = synth.backproject(z, small_cam)
        srcs.append(pts); gts.append(T)
    clouds = [ctx.upload(s) for s in srcs]
    prm = _abi.icp_params(10)
    res = ctx.register_batch(clouds, [tgt] * len(clouds), None, prm)
    for i, (s, r) in enumerate(zip(srcs, res)):
        o = oracle.icp(s, base["tgt"], base["tgt_normals"], params=prm)
        ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
        assert ok, (i, err)
        assert np.array_equal(r["T"], o["T"]) and r["inliers"] == o["inliers"], i
        ok, err = pose_close(r["T"], gts[i], 3e-3, 8e-3)
        assert ok, (i, err)
    for c in clouds:
        c.free()
    tgt.free()


def test_batch_with_more_pairs_than_ctas(ctx):
    """More pairs than co-resident CTAs: the groups of the persistent kernel walk the pair list (one CTA per pair, several
    pairs per CTA one after the other); every pair must come out exactly as when it is registered alone."""
    cam = synth.Camera().scaled(0.1)                      # 64 x 48 = 3072 points per cloud
    pairs = [synth.make_pair(200 + i, cam=cam) for i in range(12)]
    srcs = [ctx.upload(p["src"]) for p in pairs]
    tgts = [ctx.upload(p["tgt"], p["tgt_normals"]) for p in pairs]
    prm = _abi.icp_params(8)
    singles = [ctx.register(s, t, None, prm) for s, t in zip(srcs, tgts)]
    n = 192                                               # > 148 co-resident CTAs, a multiple of 12
    idx = [i % 12 for i in range(n)]
    res = ctx.register_batch([srcs[i] for i in idx], [tgts[i] for i in idx], None, prm)
    assert len(res) == n
    for k, i in enumerate(idx):
        assert res[k]["status"] == singles[i]["status"] == 0
        assert res[k]["inliers"] == singles[i]["inliers"]
        assert np.array_equal(res[k]["T"], singles[i]["T"]) and res[k]["fitness"] == singles[i]["fitness"], k
    for c in srcs + tgts:
        c.free()


def test_batch_mixed_status(ctx, small_cam, small_pair):
    good = small_pair
    bad = synth.make_pair(3, cam=small_cam, scene="S0")
    cs = [ctx.upload(good["src"]), ctx.upload(bad["src"]), ctx.upload(good["src"])]
    ct = [ctx.upload(good["tgt"], good["tgt_normals"]), ctx.upload(bad["tgt"], bad["tgt_normals"])]
    res = ctx.register_batch(cs, [ct[0], ct[1], ct[0]], None, _abi.icp_params(10))
    assert [r["status"] for r in res] == [0, _abi.PAIR_DEGENERATE, 0]
    assert np.array_equal(res[0]["T"], res[2]["T"]) and np.array_equal(res[1]["T"], np.eye(4))
    for c in cs + ct:
        c.free()


def test_full_size_config1_gate(ctx, full_pair):
    """BASELINE config 1: single synthetic 640x480 pair, 10 iterations, pose-match gate vs the oracle."""
    p = full_pair
    assert len(p["src"]) == len(p["tgt"]) == 307200
    prm = _abi.icp_params(10)
    r, nn = _gpu_icp(ctx, p, prm)
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, want_nn=True, nthreads=0)
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert np.array_equal(nn, o["nn"]) and np.array_equal(r["T"], o["T"]) and r["inliers"] == o["inliers"] and r["fitness"] == o["fitness"]
    # first-iteration correspondences are bit exact at full size
    r1, nn1 = _gpu_icp(ctx, p, _abi.icp_params(1))
    idx, _ = oracle.nn(p["src"], p["tgt"], None, nthreads=0)
    assert np.array_equal(nn1, idx)


def test_full_size_skip_logic_is_exact(ctx, full_pair):
    """Same check at 640x480 for a late iteration: the 12th iteration of a run equals a from-scratch iteration at that pose,
    and its correspondences are the oracle's exact nearest neighbours."""
    p = full_pair
    a, nna = _gpu_icp(ctx, p, _abi.icp_params(12))
    b, _ = _gpu_icp(ctx, p, _abi.icp_params(11))
    c, nnc = _gpu_icp(ctx, p, _abi.icp_params(1), guess=b["T"])
    assert np.array_equal(nna, nnc) and np.array_equal(a["T"], c["T"])
    idx, _ = oracle.nn(p["src"], p["tgt"], b["T"], nthreads=0)
    assert np.array_equal(nna, idx)


def test_full_size_config2_30_iterations(ctx, full_pair):
    p = full_pair
    prm = _abi.icp_params(30)
    r, _ = _gpu_icp(ctx, p, prm)
    o = oracle.icp(p["src"], p["tgt"], p["tgt_normals"], params=prm, nthreads=0)
    ok, err = pose_close(r["T"], o["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert np.array_equal(r["T"], o["T"]) and r["inliers"] == o["inliers"] and r["norm"] == o["norm"]
    ok, err = pose_close(r["T"], p["T_gt"], 1e-3, 3e-3)
    assert ok, err


def test_full_size_brute_force_matches_grid(ctx, full_pair):
    p = full_pair
    a, nna = _gpu_icp(ctx, p, _abi.icp_params(2, search=_abi.SEARCH_GRID))
    b, nnb = _gpu_icp(ctx, p, _abi.icp_params(2, search=_abi.SEARCH_BRUTE))
    c, nnc = _gpu_icp(ctx, p, _abi.icp_params(2, search=_abi.SEARCH_GRID_LANE))
    assert np.array_equal(nna, nnb) and np.array_equal(nnc, nnb) and np.array_equal(c["T"], b["T"]) and np.array_equal(a["T"], b["T"])


def test_roundtrip_property_full_size(ctx, full_pair):
    """Size-independent property: registering a cloud against a rigidly moved copy of itself recovers the motion
    (every correspondence is exact, residual 0), for several random motions."""
    p = full_pair
    tgt = ctx.upload(p["tgt"], p["tgt_normals"])
    for k in range(3):
        T = synth.random_rel_pose(900 + k, (0.002, 0.004), (0.002, 0.004))
        Ti = np.linalg.inv(T)
        moved = p["tgt"].copy()
        moved[:, :3] = (p["tgt"][:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
        src = ctx.upload(moved)
        r = ctx.register(src, tgt, None, _abi.icp_params(20))
        ok, err = pose_close(r["T"], T, 2e-5, 2e-5)
        assert ok, err
        assert r["fitness"] < 1e-10
        src.free()
    tgt.free()


# ---- plane extraction -----------------------------------------------------------------------------------

@pytest.mark.parametrize("kw", [dict(), dict(quantize=True, holes=0.2)])
def test_plane_segmentation_parity(ctx, small_cam, kw):
    p = synth.make_pair(4, cam=small_cam, **kw)
    prm = _abi.plane_params(seed=777)
    c = ctx.upload(p["tgt"])
    planes = c.segment_planes(prm)
    got = c.download(xyz=False, normals=True, labels=True)
    c.free()
    o = oracle.segment_planes(p["tgt"], prm)
    assert len(planes) == len(o["planes"]) == 3
    for a, b in zip(planes, o["planes"]):
        assert a["hypotheses"] == b["hypotheses"]
        assert np.array_equal(a["coef"], b["coef"]) and a["inliers"] == b["inliers"]
    assert np.array_equal(got["labels"], o["labels"])
    assert np.array_equal(got["normals"], o["normals"][:, :3])


def test_plane_segmentation_full_size_and_icp_with_segmented_normals(ctx, full_pair):
    p = full_pair
    prm = _abi.plane_params(timed=True)        # per-pass events (the default call replays the launches from a CUDA graph)
    tgt = ctx.upload(p["tgt"])
    planes = tgt.segment_planes(prm)
    got = tgt.download(xyz=False, normals=True, labels=True)
    o = oracle.segment_planes(p["tgt"], prm)
    assert len(planes) == len(o["planes"]) == 3
    assert np.array_equal(got["labels"], o["labels"]) and np.array_equal(got["normals"], o["normals"][:, :3])
    for a, b in zip(planes, o["planes"]):
        assert a["hypotheses"] == b["hypotheses"] and np.array_equal(a["coef"], b["coef"]) and a["inliers"] == b["inliers"]
    # the reference flow: extract planes, then register against them
    src = ctx.upload(p["src"])
    icp = _abi.icp_params(10)
    r = ctx.register(src, tgt, None, icp)
    nrm4 = np.c_[got["normals"], (got["labels"] >= 0).astype(np.float32)].astype(np.float32)
    oi = oracle.icp(p["src"], p["tgt"], nrm4, params=icp, nthreads=0)
    ok, err = pose_close(r["T"], oi["T"], ROT_TOL, TRANS_TOL)
    assert ok, err
    assert np.array_equal(r["T"], oi["T"]) and r["inliers"] == oi["inliers"]
    tm = ctx.last_plane_timing()
    assert tm["rounds"] == 3 and tm["eval_passes_per_round"] == 1 and 307200 < tm["points_scanned"] < 3 * 307200 and 0 < tm["eval_ms"] <= tm["total_ms"]
    # the graph replay (default parameters) gives the same planes and labels, twice (capture, then replay)
    t2 = ctx.upload(p["tgt"])
    for _ in range(2):
        again = t2.segment_planes(_abi.plane_params())
        assert [a["inliers"] for a in again] == [a["inliers"] for a in planes] and all(np.array_equal(a["coef"], b["coef"]) for a, b in zip(again, planes))
        assert np.array_equal(t2.download(xyz=False, labels=True)["labels"], got["labels"])
        assert ctx.last_plane_timing()["eval_ms"] == 0.0 and ctx.last_plane_timing()["total_ms"] > 0.0
    t2.free()
    src.free(); tgt.free()


def test_plane_extraction_enqueue_drain(ctx, small_cam):
    """s3d_segment_planes_enqueue / s3d_segment_planes_drain: the same planes, labels and normals as the blocking call, and a
    registration enqueued behind the extraction sees the normals it wrote (stream order, no host round trip in between)."""
    p = synth.make_pair(3, cam=small_cam)
    prm = _abi.plane_params()
    icp = _abi.icp_params(5)
    a, b = ctx.upload(p["tgt"]), ctx.upload(p["tgt"])
    src = ctx.upload(p["src"])
    try:
        ref_planes = a.segment_planes(prm)
        ref = ctx.register(src, a, None, icp)
        assert ctx.planes_drain() == []
        b.segment_planes_enqueue(prm)
        ctx.register_enqueue(src, b, None, icp)
        b.segment_planes_enqueue(prm)                  # a second extraction of the same cloud: same answer, second slot
        # a cloud released while its work is enqueued: the work still runs on intact data, the next uploads get its buffers behind it
        c = ctx.upload(p["tgt"]); s2 = ctx.upload(p["src"])
        c.segment_planes_enqueue(prm)
        ctx.register_enqueue(s2, c, None, icp)
        c.release(); s2.release()
        junk = [ctx.upload(np.full_like(p["tgt"], 7.0)) for _ in range(4)]
        got = ctx.planes_drain()
        res, _ = ctx.register_drain()
        for j in junk:
            j.free()
        assert len(got) == 3 and len(res) == 2
        assert res[1]["status"] == 0 and np.array_equal(res[1]["T"], ref["T"]) and res[1]["inliers"] == ref["inliers"]
        for pl in got:
            assert len(pl) == len(ref_planes) > 0
            for x, y in zip(pl, ref_planes):
                assert np.array_equal(x["coef"], y["coef"]) and x["inliers"] == y["inliers"] and x["hypotheses"] == y["hypotheses"]
        da, db = a.download(normals=True, labels=True), b.download(normals=True, labels=True)
        assert np.array_equal(da["labels"], db["labels"]) and np.array_equal(da["normals"], db["normals"])
        assert res[0]["status"] == 0 and np.array_equal(res[0]["T"], ref["T"]) and res[0]["inliers"] == ref["inliers"]
    finally:
        a.free(); b.free(); src.free()


def test_plane_segmentation_edge_cases(ctx):
    prm = _abi.plane_params()
    c = ctx.upload(np.zeros((0, 3), np.float32))
    assert c.segment_planes(prm) == []
    c.free()
    c = ctx.upload(np.array([[0, 0, 1], [1, 0, 1]], np.float32))
    assert c.segment_planes(prm) == []
    c.free()
    g = np.stack(np.meshgrid(np.linspace(-1, 1, 40), np.linspace(-1, 1, 40)), -1).reshape(-1, 2)
    pts = np.c_[g, np.full(len(g), 2.0)].astype(np.float32)
    c = ctx.upload(pts)
    planes = c.segment_planes(prm)
    lab = c.download(xyz=False, labels=True)["labels"]
    c.free()
    o = oracle.segment_planes(np.c_[pts, np.ones(len(pts), np.float32)], prm)
    assert len(planes) == len(o["planes"]) == 1 and (lab == 0).all()
    assert np.array_equal(planes[0]["coef"], o["planes"][0]["coef"])


# ---- ingest and keypoint planarity ------------------------------------------------------------------------

def test_backprojection_bit_exact(ctx, small_cam):
    p = synth.make_pair(2, cam=small_cam, quantize=True, holes=0.2)
    for zmax in (0.0, 2.5):
        c = ctx.from_depth(p["tgt_depth"], small_cam, zmax)
        xyz = c.download()["xyz"]
        c.free()
        want = oracle.backproject(p["tgt_depth"], small_cam, zmax)
        assert xyz.shape[0] == want.shape[0]
        assert np.array_equal(xyz, want[:, :3])
    cam = synth.Camera()
    full = synth.make_pair(5, quantize=True, holes=0.1)
    c = ctx.from_depth(full["src_depth"], cam)
    assert np.array_equal(c.download()["xyz"], full["src"][:, :3])
    c.free()


def test_planar_keypoints_parity(ctx):
    cam = synth.Camera(cx=320.0)        # planarFeatures.cpp:13-14 hard-codes cx = 320
    z, _ = synth.render_depth(synth.base_pose(), cam, "S1", 0.004, 5, 1)
    depth = synth.quantize_depth(z, cam)
    depth[100:140, 200:260] = 0
    rng = np.random.default_rng(1)
    uv = np.stack([rng.integers(0, 640, 4000), rng.integers(0, 480, 4000)], 1).astype(np.int32)
    for thr in (0.01, 0.004):
        got = ctx.planar_keypoints(depth, cam, uv, thr, 40, 4242)
        want = oracle.planar_keypoints(depth, cam, uv, thr, 40, 4242)
        assert np.array_equal(got, want)
        assert 0 < want.sum() < len(want)
