// tile_search.cuh -- exact nearest-neighbour search: a warp serves up to TS_GROUP neighbouring queries at a
// time from a shared-memory tile of target points staged by TMA bulk copies.
//
// Semantics: PCL-1.7 CorrespondenceEstimation::determineCorrespondences (exact 1-NN of every source point in
// the target; restated in oracle/icp_oracle.c): argmin over the float32 value s3d_dist2() with the lowest
// target index winning ties.  Nothing below changes that answer; it is only about finding it with
// convergent, bandwidth-friendly work instead of per-lane dependent gathers:
//
//   * Consecutive source points (a piece of a scan line of an organised cloud) are neighbours in space, so
//     their search balls overlap.  The warp takes the bounding box (in cell space) of the balls of a group of
//     pending queries; every row of cells (fixed y,z; contiguous along x in the cell-sorted target) that
//     crosses the box is ONE contiguous range of float4 points: two cell-start loads per row, then one TMA
//     bulk copy (cp.async.bulk, completion counted by the warp's mbarrier) per row into the warp's tile.
//     (-DTS_USE_CPASYNC: per-lane 16-byte cp.async instead; the two measure the same.)
//   * The 32 lanes split the tile between them: with n queries in the group, 32/n lanes work on each query
//     (a lone pending query of a late iteration is served by all 32 lanes), partial results merged by
//     shuffles.  Shared-memory reads are broadcasts, there is no divergence, ~11 FP32-pipe instructions per
//     candidate, nearest and runner-up kept.
//   * Completeness is VERIFIED, not assumed: every target point that is not in the tile lies outside the
//     box, i.e. at least `margin` (distance from the query to the nearest box face that is not a face of the
//     whole grid) away.  A query is done when its best candidate is nearer than that; otherwise its radius is
//     replaced by the distance just found (a valid bound) and it goes round again.  Radii passed in are
//     therefore only hints.  Queries whose radius is far above the rest of their group (depth
//     discontinuities, image borders) wait for a later pass so that they do not inflate everybody's box.
//   * min(runner-up, margin) is a lower bound on the distance to every OTHER target point: the caller uses
//     it to skip the search altogether on later iterations (triangle inequality, see icp.cu).
#pragma once
#include <limits.h>
// Row copies into the tile: TMA bulk copies (cp.async.bulk + mbarrier) by default; -DTS_USE_CPASYNC selects per-lane 16-byte
// cp.async (LDGSTS) instead.  Measured on the config-2 pair: 1051 us vs 1047 us per 30-iteration registration -- a tie.
#ifndef TS_USE_CPASYNC
#define TS_USE_TMA 1
#endif
#include "context.h"
#include "common.cuh"
#include "grid.cuh"
#include "search.cuh"

#ifndef TS_CAP
#define TS_CAP 640            // candidates per warp tile (10 KiB): also 5 staged chunks of per-query state (icp.cu)
#endif
#define TS_GROUP 8            // queries per box: an aligned octet of lanes
#define TS_MAX_TRIES 64

#if defined(S3D_PHASES)
#define TS_TM_ARG , long long *tm
#define TS_TM_PASS , tm
#define TS_T0() long long ts_last = clock64()
#define TS_T(i) do { const long long n_ = clock64(); tm[i] += n_ - ts_last; ts_last = n_; } while (0)
#define TS_CNT(i, v) do { if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_stats[i], (unsigned long long)(v)); } while (0)
#else
#define TS_TM_ARG
#define TS_TM_PASS
#define TS_T0()
#define TS_T(i)
#define TS_CNT(i, v)
#endif

// ---- mbarrier / TMA bulk copy (PTX) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t ts_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ts_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint64_t *bar, uint32_t bytes)      // no arrival, only more bytes to wait for
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(ts_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ts_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ts_smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void ts_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(ts_smem_u32(dst)), "l"(src), "r"(bytes), "r"(ts_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ts_fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS): per-lane addresses, no register staging, fire and forget
__device__ __forceinline__ void ts_cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ts_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ts_cp_async16_s(uint32_t saddr, const void *src)     // destination given as a shared address
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void ts_cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// The tile pointer crosses a real call (tile_search is not inlined), where the compiler no longer knows it is
// shared memory and would emit generic LD.E (measured: the compare loop stalled on them).  Candidates are
// therefore read with explicit ld.shared through the 32-bit shared address.
__device__ __forceinline__ float4 ts_lds128(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

// position of the (s+1)-th set bit of a byte m: (c_nth8[m] >> 4 s) & 7  (one constant load instead of the loop behind __fns)
__constant__ uint32_t c_nth8[256] = {
    0x00000000u, 0x00000000u, 0x00000001u, 0x00000010u, 0x00000002u, 0x00000020u, 0x00000021u, 0x00000210u,
    0x00000003u, 0x00000030u, 0x00000031u, 0x00000310u, 0x00000032u, 0x00000320u, 0x00000321u, 0x00003210u,
    0x00000004u, 0x00000040u, 0x00000041u, 0x00000410u, 0x00000042u, 0x00000420u, 0x00000421u, 0x00004210u,
    0x00000043u, 0x00000430u, 0x00000431u, 0x00004310u, 0x00000432u, 0x00004320u, 0x00004321u, 0x00043210u,
    0x00000005u, 0x00000050u, 0x00000051u, 0x00000510u, 0x00000052u, 0x00000520u, 0x00000521u, 0x00005210u,
    0x00000053u, 0x00000530u, 0x00000531u, 0x00005310u, 0x00000532u, 0x00005320u, 0x00005321u, 0x00053210u,
    0x00000054u, 0x00000540u, 0x00000541u, 0x00005410u, 0x00000542u, 0x00005420u, 0x00005421u, 0x00054210u,
    0x00000543u, 0x00005430u, 0x00005431u, 0x00054310u, 0x00005432u, 0x00054320u, 0x00054321u, 0x00543210u,
    0x00000006u, 0x00000060u, 0x00000061u, 0x00000610u, 0x00000062u, 0x00000620u, 0x00000621u, 0x00006210u,
    0x00000063u, 0x00000630u, 0x00000631u, 0x00006310u, 0x00000632u, 0x00006320u, 0x00006321u, 0x00063210u,
    0x00000064u, 0x00000640u, 0x00000641u, 0x00006410u, 0x00000642u, 0x00006420u, 0x00006421u, 0x00064210u,
    0x00000643u, 0x00006430u, 0x00006431u, 0x00064310u, 0x00006432u, 0x00064320u, 0x00064321u, 0x00643210u,
    0x00000065u, 0x00000650u, 0x00000651u, 0x00006510u, 0x00000652u, 0x00006520u, 0x00006521u, 0x00065210u,
    0x00000653u, 0x00006530u, 0x00006531u, 0x00065310u, 0x00006532u, 0x00065320u, 0x00065321u, 0x00653210u,
    0x00000654u, 0x00006540u, 0x00006541u, 0x00065410u, 0x00006542u, 0x00065420u, 0x00065421u, 0x00654210u,
    0x00006543u, 0x00065430u, 0x00065431u, 0x00654310u, 0x00065432u, 0x00654320u, 0x00654321u, 0x06543210u,
    0x00000007u, 0x00000070u, 0x00000071u, 0x00000710u, 0x00000072u, 0x00000720u, 0x00000721u, 0x00007210u,
    0x00000073u, 0x00000730u, 0x00000731u, 0x00007310u, 0x00000732u, 0x00007320u, 0x00007321u, 0x00073210u,
    0x00000074u, 0x00000740u, 0x00000741u, 0x00007410u, 0x00000742u, 0x00007420u, 0x00007421u, 0x00074210u,
    0x00000743u, 0x00007430u, 0x00007431u, 0x00074310u, 0x00007432u, 0x00074320u, 0x00074321u, 0x00743210u,
    0x00000075u, 0x00000750u, 0x00000751u, 0x00007510u, 0x00000752u, 0x00007520u, 0x00007521u, 0x00075210u,
    0x00000753u, 0x00007530u, 0x00007531u, 0x00075310u, 0x00007532u, 0x00075320u, 0x00075321u, 0x00753210u,
    0x00000754u, 0x00007540u, 0x00007541u, 0x00075410u, 0x00007542u, 0x00075420u, 0x00075421u, 0x00754210u,
    0x00007543u, 0x00075430u, 0x00075431u, 0x00754310u, 0x00075432u, 0x00754320u, 0x00754321u, 0x07543210u,
    0x00000076u, 0x00000760u, 0x00000761u, 0x00007610u, 0x00000762u, 0x00007620u, 0x00007621u, 0x00076210u,
    0x00000763u, 0x00007630u, 0x00007631u, 0x00076310u, 0x00007632u, 0x00076320u, 0x00076321u, 0x00763210u,
    0x00000764u, 0x00007640u, 0x00007641u, 0x00076410u, 0x00007642u, 0x00076420u, 0x00076421u, 0x00764210u,
    0x00007643u, 0x00076430u, 0x00076431u, 0x00764310u, 0x00076432u, 0x00764320u, 0x00764321u, 0x07643210u,
    0x00000765u, 0x00007650u, 0x00007651u, 0x00076510u, 0x00007652u, 0x00076520u, 0x00076521u, 0x00765210u,
    0x00007653u, 0x00076530u, 0x00076531u, 0x00765310u, 0x00076532u, 0x00765320u, 0x00765321u, 0x07653210u,
    0x00007654u, 0x00076540u, 0x00076541u, 0x00765410u, 0x00076542u, 0x00765420u, 0x00765421u, 0x07654210u,
    0x00076543u, 0x00765430u, 0x00765431u, 0x07654310u, 0x00765432u, 0x07654320u, 0x07654321u, 0x76543210u
};

struct TileBest {             // nearest candidate (point + original index in .w), its d2, and the runner-up's d2
    float4 bq; float bd; float sd;
};

__device__ __forceinline__ void tile_best_init(TileBest &B)
{
    B.bq = make_float4(0.f, 0.f, 0.f, __int_as_float(INT_MAX)); B.bd = INFINITY; B.sd = INFINITY;
}

// merge of two partial results over DISJOINT candidate sets (commutative, associative)
__device__ __forceinline__ void tile_best_merge(TileBest &a, const float4 obq, float obd, float osd)
{
    const int ai = __float_as_int(a.bq.w), oi = __float_as_int(obq.w);
    const bool better = obd < a.bd || (obd == a.bd && oi < ai);
    a.sd = fminf(fminf(a.sd, osd), better ? a.bd : obd);
    if (better) { a.bd = obd; a.bq = obq; }
}

// One lane's share of the tile: candidates sub, sub+step, ... < m against the query (qx,qy,qz); result merged into W.
// EXCL: the candidate whose original index is `excl` is left out (second pass that looks for the runner-up).
#define TS_D2(c) (EXCL && __float_as_int((c).w) == excl ? INFINITY : s3d_dist2(qx, qy, qz, (c).x, (c).y, (c).z))
template <bool EXCL>
__device__ __forceinline__ void tile_compare_pass(uint32_t sbuf, int m, int sub, int step,
                                                  float qx, float qy, float qz, TileBest &W, int excl)
{
    // two independent running minima (even / odd visits) halve the dependent chain
    float pd0 = INFINITY, ps0 = INFINITY, pd1 = INFINITY, ps1 = INFINITY;
    int pk0 = -1, pk1 = -1;
    int k = sub;
    for (; k + 3 * step < m; k += 4 * step) {          // four candidates per turn, two per chain
        const float4 a = ts_lds128(sbuf + 16u * (uint32_t)k), b = ts_lds128(sbuf + 16u * (uint32_t)(k + step));
        const float4 c = ts_lds128(sbuf + 16u * (uint32_t)(k + 2 * step)), e = ts_lds128(sbuf + 16u * (uint32_t)(k + 3 * step));
        const float da = TS_D2(a), db = TS_D2(b);
        const float dc = TS_D2(c), de = TS_D2(e);
        bool la = da < pd0, lbb = db < pd1;
        ps0 = fminf(ps0, fmaxf(pd0, da)); ps1 = fminf(ps1, fmaxf(pd1, db));
        pd0 = fminf(pd0, da); pd1 = fminf(pd1, db);
        pk0 = la ? k : pk0; pk1 = lbb ? k + step : pk1;
        la = dc < pd0; lbb = de < pd1;
        ps0 = fminf(ps0, fmaxf(pd0, dc)); ps1 = fminf(ps1, fmaxf(pd1, de));
        pd0 = fminf(pd0, dc); pd1 = fminf(pd1, de);
        pk0 = la ? k + 2 * step : pk0; pk1 = lbb ? k + 3 * step : pk1;
    }
    for (; k + step < m; k += 2 * step) {
        const float4 a = ts_lds128(sbuf + 16u * (uint32_t)k), b = ts_lds128(sbuf + 16u * (uint32_t)(k + step));
        const float da = TS_D2(a), db = TS_D2(b);
        const bool la = da < pd0, lbb = db < pd1;
        ps0 = fminf(ps0, fmaxf(pd0, da)); ps1 = fminf(ps1, fmaxf(pd1, db));
        pd0 = fminf(pd0, da); pd1 = fminf(pd1, db);
        pk0 = la ? k : pk0; pk1 = lbb ? k + step : pk1;
    }
    if (k < m) {
        const float4 a = ts_lds128(sbuf + 16u * (uint32_t)k);
        const float da = TS_D2(a);
        const bool la = da < pd0;
        ps0 = fminf(ps0, fmaxf(pd0, da)); pd0 = fminf(pd0, da); pk0 = la ? k : pk0;
    }
    // combine the two chains: nearest, runner-up, and whether the minimum is attained more than once
    const float pd = fminf(pd0, pd1);
    const float ps = fminf(fminf(ps0, ps1), fmaxf(pd0, pd1));
    int pk = (pd1 < pd0) ? pk1 : pk0;
    if (pk < 0 || !(pd < INFINITY)) return;               // this lane saw no (admissible) candidate
    if (ps == pd) {
        // equal distances: the lowest original index wins (rare; rescan this lane's share)
        int bi = INT_MAX;
        for (int kk = sub; kk < m; kk += step) {
            const float4 q = ts_lds128(sbuf + 16u * (uint32_t)kk);
            const float d2 = TS_D2(q);
            const int qi = __float_as_int(q.w);
            if (d2 == pd && qi < bi) { bi = qi; pk = kk; }
        }
    }
    tile_best_merge(W, ts_lds128(sbuf + 16u * (uint32_t)pk), pd, ps);
}

// Exact NN of the pending queries of one warp.  (px,py,pz): this lane's query; r: its search radius hint in
// metres (any positive finite value; a valid upper bound on the NN distance makes the first pass final);
// `pending`: lane has a query.  gate_r: nothing farther than this can be accepted (INFINITY: no gate).
// buf: this warp's tile (TS_CAP float4); bar/parity (TMA variant): this warp's mbarrier and its phase.  Returns,
// for pending lanes: bq/bd = nearest target point among ALL target points and its squared distance (bd = INFINITY
// when the target has no point within reach), lb = lower bound on the distance from the query to every target
// point other than the winner.
struct TileCfg {              // one search level (fine grid / decimated grid); lives in shared memory
    GridParams gp; const uint32_t *cell_start; const float4 *pts; float slack; float gate_r;
};
struct TileOut { float4 bq; float bd; float lb; float4 q2; float lb3; };     // q2/lb3: runner-up and bound on all others (near ties)
#define TS_TIE_GAP 1e-5f       // a nearest/runner-up gap below this (metres) is a near tie: the runner-up is kept too

#ifdef TS_USE_TMA
#define TS_BAR_ARG , uint64_t *bar, uint32_t *parity_p
#else
#define TS_BAR_ARG
#endif

// NOT inlined on purpose: the caller keeps 29 double accumulators and a prefetched chunk in registers; as a real
// call the search spills them only around itself (the rare path of a late iteration) instead of raising the
// register pressure of the whole streaming loop.
template <int CAP>
__device__ __noinline__ TileOut tile_search(const TileCfg *cfg, float px, float py, float pz, float r, bool pending,
                                            float4 *__restrict__ buf, int lane, bool want2 TS_BAR_ARG TS_TM_ARG)
{
    TS_T0();
    const unsigned full = 0xffffffffu;
    const GridParams gp = cfg->gp;
    const uint32_t *__restrict__ cell_start = cfg->cell_start;
    const float slack = cfg->slack, gate_r = cfg->gate_r;
    TileBest B; float lbound = 0.f;
    tile_best_init(B);
    float4 Q2 = make_float4(0.f, 0.f, 0.f, __int_as_float(-1)); float LB3 = 0.f;
#ifdef TS_USE_TMA
    uint32_t parity = *parity_p;
#endif
    const float fx = grid_fcoord(px, gp.ox, gp.inv_cell), fy = grid_fcoord(py, gp.oy, gp.inv_cell), fz = grid_fcoord(pz, gp.oz, gp.inv_cell);
    const float4 *__restrict__ sp = cfg->pts;
    const uint32_t sbuf = ts_smem_u32(buf);
    bool todo = pending;
    r = fminf(r, gate_r * 1.00001f + 1e-6f);
#ifdef TS_USE_TMA
    ts_fence_proxy_async();          // the tile may have been written with ordinary stores since the last bulk copy
#endif
    for (int tries = 0; ; ++tries) {
        const unsigned todomask = __ballot_sync(full, todo);
        if (!todomask) break;
        // ---- this pass's group: the first TS_GROUP pending queries, minus those whose radius is far above the rest ----
        // (an aligned octet of lanes = 8 consecutive source points; different octets may be far apart in space)
        unsigned grp = todomask & (0xffu << ((__ffs(todomask) - 1) & 24));
        const bool ingrp = (grp >> lane) & 1u;
        float rmin = ingrp ? r : INFINITY;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) rmin = fminf(rmin, __shfl_xor_sync(full, rmin, o));
        float thr = 3.f * rmin + slack;
        if (2 * __popc(__ballot_sync(full, ingrp && r > thr)) >= __popc(grp)) thr = INFINITY;
        const bool inbox = ingrp && r <= thr;
        const unsigned inmask = __ballot_sync(full, inbox);
        const int nin = __popc(inmask);
        // ---- the box: union of the balls of the group ----
        int lox = INT_MAX, loy = INT_MAX, loz = INT_MAX, hix = INT_MIN, hiy = INT_MIN, hiz = INT_MIN;
        if (inbox) {
            const float Rc = r * gp.inv_cell * 1.00001f + GRID_MARGIN;
            lox = min(max(__float2int_rd(fx - Rc), 0), gp.nx - 1); hix = min(max(__float2int_rd(fx + Rc), 0), gp.nx - 1);
            loy = min(max(__float2int_rd(fy - Rc), 0), gp.ny - 1); hiy = min(max(__float2int_rd(fy + Rc), 0), gp.ny - 1);
            loz = min(max(__float2int_rd(fz - Rc), 0), gp.nz - 1); hiz = min(max(__float2int_rd(fz + Rc), 0), gp.nz - 1);
        }
        lox = __reduce_min_sync(full, lox); loy = __reduce_min_sync(full, loy); loz = __reduce_min_sync(full, loz);
        hix = __reduce_max_sync(full, hix); hiy = __reduce_max_sync(full, hiy); hiz = __reduce_max_sync(full, hiz);
        if (tries >= TS_MAX_TRIES) { lox = loy = loz = 0; hix = gp.nx - 1; hiy = gp.ny - 1; hiz = gp.nz - 1; }   // safety net: whole grid
        const int ny_s = hiy - loy + 1, nz_s = hiz - loz + 1;
        const int rows = gp.n_points > 0 ? ny_s * nz_s : 0;
        // ---- work split: NS query slots (power of two >= nin), 32/NS lanes per slot ----
        const int lgNS = nin > 1 ? 32 - __clz(nin - 1) : 0;             // NS = 1 << lgNS
        const int NS = 1 << lgNS, step = 32 >> lgNS, slot = lane & (NS - 1), sub = lane >> lgNS;
        const int obase = (__ffs(todomask) - 1) & 24;                  // the group's octet
        const int src = (slot < nin) ? obase + (int)((c_nth8[(inmask >> obase) & 0xffu] >> (4 * slot)) & 7u) : 0;
        const float qx = __shfl_sync(full, px, src), qy = __shfl_sync(full, py, src), qz = __shfl_sync(full, pz, src);
        const bool work = slot < nin;
        TileBest W; tile_best_init(W);
        STAT(5, lane == 0);
        TS_CNT(27, rows); TS_CNT(28, 1); TS_CNT(29, nin);
        TS_T(0);
        // ---- gather the rows into the tile by TMA, compare whenever it is full ----
        // 32 rows per round (one per lane): two cell-start loads give the row's point range, one bulk copy moves it.
        // What does not fit into the tile stays with its lane for the next turn (`carry`): rows of any length go through.
        int fill = 0, t0 = 0, nfills = 0, last_fill = 0;
        uint32_t rs = 0u, cnt = 0u;
        bool carry = false;
        while (carry || t0 < rows) {
            if (!carry) {
                const int t = t0 + lane;
                rs = 0u; cnt = 0u;
                if (t < rows) {
                    // t / ny_s without an integer division (t < 2^22, exact in float with the half-step offset)
                    const int zi = __float2int_rz(__fdividef((float)t + 0.5f, (float)ny_s));
                    const size_t base = ((size_t)(loz + zi) * gp.ny + (loy + (t - zi * ny_s))) * gp.nx;
                    rs = __ldg(&cell_start[base + lox]);
                    cnt = __ldg(&cell_start[base + hix + 1]) - rs;
                    STAT(2, 1);
                }
                t0 += 32;
            }
            TS_T(1);
            uint32_t incl = cnt;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(full, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(full, incl, 31), excl = incl - cnt;
            const uint32_t room = (uint32_t)(CAP - fill);
            uint32_t take = cnt;
            if (total > room) take = excl >= room ? 0u : min(cnt, room - excl);
            const uint32_t moved_pts = min(total, room);
#ifdef TS_USE_TMA
            if (moved_pts > 0u) {
                if (lane == 0) ts_mbar_expect_tx(bar, moved_pts * 16u);
                __syncwarp();
                if (take > 0u) ts_bulk_g2s(buf + fill + excl, sp + rs, take * 16u, bar);
            }
#else
            {
                const uint32_t dst = sbuf + 16u * ((uint32_t)fill + excl);
                const float4 *srcp = sp + rs;
                for (uint32_t k = 0; k < take; k += 4) {          // four per turn: the loop overhead was 11 % of all instructions
                    const uint32_t d0 = dst + 16u * k;
                    const float4 *s0 = srcp + k;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(s0) : "memory");
                    if (k + 1 < take) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u), "l"(s0 + 1) : "memory");
                    if (k + 2 < take) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 32u), "l"(s0 + 2) : "memory");
                    if (k + 3 < take) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d0 + 48u), "l"(s0 + 3) : "memory");
                }
            }
#endif
            rs += take; cnt -= take;
            TS_T(2);
            carry = total > room;
            fill += (int)moved_pts;
            if ((carry || t0 >= rows) && fill > 0) {
#ifdef TS_USE_TMA
                if (lane == 0) ts_mbar_arrive(bar);
                ts_mbar_wait(bar, parity);
                parity ^= 1u;
#else
                ts_cp_async_wait_all();
                __syncwarp();
#endif
                if (work) tile_compare_pass<false>(sbuf, fill, sub, step, qx, qy, qz, W, -1);
                ++nfills; last_fill = fill;
                STAT(3, work ? fill / step : 0);
                __syncwarp();
                TS_CNT(26, fill);
                fill = 0;
                TS_T(3);
            }
        }
        // ---- merge the lanes of each slot, hand the result back to the query's lane ----
        for (int o = NS; o < 32; o <<= 1) {
            const float4 obq = make_float4(__shfl_xor_sync(full, W.bq.x, o), __shfl_xor_sync(full, W.bq.y, o),
                                           __shfl_xor_sync(full, W.bq.z, o), __shfl_xor_sync(full, W.bq.w, o));
            const float obd = __shfl_xor_sync(full, W.bd, o), osd = __shfl_xor_sync(full, W.sd, o);
            tile_best_merge(W, obq, obd, osd);
        }
        const int myslot = __popc(inmask & ((1u << lane) - 1u));        // slot that served this lane's query (valid if inbox)
        TileBest R;
        R.bq = make_float4(__shfl_sync(full, W.bq.x, myslot), __shfl_sync(full, W.bq.y, myslot),
                           __shfl_sync(full, W.bq.z, myslot), __shfl_sync(full, W.bq.w, myslot));
        R.bd = __shfl_sync(full, W.bd, myslot); R.sd = __shfl_sync(full, W.sd, myslot);
        // ---- verify: everything outside the box is at least `margin` away ----
        float mc = INFINITY;
        if (lox > 0) mc = fminf(mc, fx - (float)lox);
        if (hix < gp.nx - 1) mc = fminf(mc, (float)(hix + 1) - fx);
        if (loy > 0) mc = fminf(mc, fy - (float)loy);
        if (hiy < gp.ny - 1) mc = fminf(mc, (float)(hiy + 1) - fy);
        if (loz > 0) mc = fminf(mc, fz - (float)loz);
        if (hiz < gp.nz - 1) mc = fminf(mc, (float)(hiz + 1) - fz);
        const float margin = fmaxf((mc - GRID_MARGIN) * gp.cell * 0.99999f, 0.f);     // INFINITY when the box is the whole grid
        if (inbox) {
            const bool done = (R.bd <= margin * margin) || (margin >= gate_r) || !(mc < INFINITY);
            if (done) {
                B = R;
                lbound = fmaxf(fminf(sqrtf(R.sd), margin) * 0.999998f - 5e-8f, 0.f);
                todo = false;
                LB3 = 0.f;
            } else {
                // what was found is a valid bound; a query that found nothing doubles its hint
                r = (R.bd < INFINITY) ? sqrtf(R.bd) * 1.00001f + slack : 2.f * r + slack;
                r = fminf(r, gate_r * 1.00001f + 1e-6f);
            }
        }
        // ---- near ties: a query whose runner-up is within TS_TIE_GAP of its nearest neighbour would fail the caller's
        // skip test in every later iteration (the float noise of the test is ~1e-6 m) and be searched again and again.
        // For those the runner-up itself and a bound on everything else are returned: the caller then settles the
        // order of the two by evaluating both.  One more pass over the tile (still in place when it was filled once),
        // leaving out the winner.
        {
            const bool tie = want2 && inbox && !todo && nfills == 1 && R.sd < INFINITY &&
                             sqrtf(R.sd) - sqrtf(R.bd) < TS_TIE_GAP && sqrtf(R.sd) < margin;
            if (__any_sync(full, tie)) {
                TileBest W2; tile_best_init(W2);
                const int excl = __float_as_int(W.bq.w);           // the slot's winner (all lanes of a slot hold it)
                if (work) tile_compare_pass<true>(sbuf, last_fill, sub, step, qx, qy, qz, W2, excl);
                for (int o = NS; o < 32; o <<= 1) {
                    const float4 obq = make_float4(__shfl_xor_sync(full, W2.bq.x, o), __shfl_xor_sync(full, W2.bq.y, o),
                                                   __shfl_xor_sync(full, W2.bq.z, o), __shfl_xor_sync(full, W2.bq.w, o));
                    const float obd = __shfl_xor_sync(full, W2.bd, o), osd = __shfl_xor_sync(full, W2.sd, o);
                    tile_best_merge(W2, obq, obd, osd);
                }
                const float4 r2 = make_float4(__shfl_sync(full, W2.bq.x, myslot), __shfl_sync(full, W2.bq.y, myslot),
                                              __shfl_sync(full, W2.bq.z, myslot), __shfl_sync(full, W2.bq.w, myslot));
                const float rd2 = __shfl_sync(full, W2.bd, myslot), td2 = __shfl_sync(full, W2.sd, myslot);
                if (tie && rd2 < INFINITY) {
                    const float lb3 = fminf(sqrtf(td2), margin) * 0.999998f - 5e-8f;
                    if (lb3 > sqrtf(rd2) + 1e-6f) { Q2 = r2; LB3 = lb3; }      // worth it only if the third is clearly behind
                }
                __syncwarp();
            }
        }
        TS_T(4);
    }
#ifdef TS_USE_TMA
    *parity_p = parity;
#endif
    TileOut o;
    o.bq = B.bq; o.bd = B.bd; o.lb = lbound; o.q2 = Q2; o.lb3 = LB3;
    return o;
}
