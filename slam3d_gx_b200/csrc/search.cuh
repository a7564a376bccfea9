// search.cuh -- exact nearest-neighbour search on the uniform grid, warp-cooperative.
//
// Semantics: PCL-1.7 CorrespondenceEstimation::determineCorrespondences (exact 1-NN of every source
// point in the target; restated in oracle/icp_oracle.c): argmin over the float32 value s3d_dist2()
// with the lowest target index winning ties.  Everything below is only about finding that argmin
// quickly; none of it changes the answer.
//
// How the search is organised (why: the kernel is bound by the latency of dependent gathers and by
// SIMT divergence, not by bandwidth):
//   * A query is searched by a GROUP of 8 lanes (4 queries per warp at a time).  The rows of cells
//     (fixed y,z; contiguous along x in the cell-sorted array) that intersect the search ball are dealt
//     out to the 8 lanes, so their lookups are independent loads in flight instead of one serial chain,
//     and a far query costs its warp a few rounds instead of stalling 31 idle lanes.
//   * Every search starts from an upper bound ("seed"): the query's correspondence of the previous
//     iteration, the results of neighbouring lanes that are already done (adjacent source points have
//     adjacent nearest neighbours), and on the first iteration the nearest point of a 16x decimated copy
//     of the target.  A seed only bounds the ball; the winner is still the exact argmin.
//   * Per-row occupancy bit masks (one bit per cell along x) rule out empty rows with one load and give
//     the first/last occupied cell of the x-run, so empty space inside the ball costs almost nothing.
//   * The runner-up distance is tracked as well: on return every point other than the winner is known
//     to be at least `lb` away, which lets the caller skip the search on later iterations while the
//     pose update is smaller than the slack (triangle inequality, see icp.cu).
#pragma once
#include "context.h"
#include "common.cuh"
#include "grid.cuh"

#define GRID_MARGIN 1e-3f     // cell-coordinate rounding allowance, in cells (DESIGN.md "exactness of the grid search")
#define COOP 8                // lanes per query

#if defined(S3D_STATS) || defined(S3D_PHASES)
__device__ unsigned long long g_stats[32];   // debug builds only: [0] searched [1] skipped [2] rows looked up [3] candidates
                                             // [4] mask loads [5] search rounds [6] coarse searches [8..] phase cycles of CTA 0
#endif
#ifdef S3D_STATS
#define STAT(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
#else
#define STAT(i, v)
#endif

struct Best {            // running result of a search: nearest point and the squared distance of the runner-up
    float bd; int bpos; int bidx; float sd;
};

__device__ __forceinline__ void best_init(Best &b) { b.bd = INFINITY; b.sd = INFINITY; b.bpos = -1; b.bidx = 0x7fffffff; }

__device__ __forceinline__ void nn_update(const float4 q, uint32_t k, float px, float py, float pz, Best &b)
{
    float d2 = s3d_dist2(px, py, pz, q.x, q.y, q.z);
    int qi = __float_as_int(q.w);
    if (d2 < b.bd || (d2 == b.bd && qi < b.bidx)) { b.sd = b.bd; b.bd = d2; b.bpos = (int)k; b.bidx = qi; }
    else b.sd = fminf(b.sd, d2);
}

// candidates [s,e) of the cell-sorted array, four independent loads in flight per step
__device__ __forceinline__ void scan_range(const float4 *__restrict__ sp, uint32_t s, uint32_t e, float px, float py, float pz, Best &b)
{
    STAT(3, e > s ? e - s : 0);
    for (uint32_t k = s; k < e; k += 4) {
        const uint32_t k1 = min(k + 1, e - 1), k2 = min(k + 2, e - 1), k3 = min(k + 3, e - 1);
        const float4 q0 = __ldg(&sp[k]), q1 = __ldg(&sp[k1]), q2 = __ldg(&sp[k2]), q3 = __ldg(&sp[k3]);
        nn_update(q0, k, px, py, pz, b);
        if (k + 1 < e) nn_update(q1, k1, px, py, pz, b);
        if (k + 2 < e) nn_update(q2, k2, px, py, pz, b);
        if (k + 3 < e) nn_update(q3, k3, px, py, pz, b);
    }
}

__device__ __forceinline__ float sqrt_up(float v)   // cheap upper bound of sqrt(v) for conservative pruning extents
{
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r * 1.00001f;
}

// distance (cell units, conservative) from cell-space coordinate f to the cell interval [a, b)
__device__ __forceinline__ float gap_iv(float f, int a, int b)
{
    return fmaxf(fmaxf((float)a - f, f - (float)b) - GRID_MARGIN, 0.f);
}

// merge of two partial results (commutative, associative): nearest of the two, and the runner-up
__device__ __forceinline__ void best_merge(Best &a, float obd, int obpos, int obidx, float osd)
{
    if (obpos == a.bpos) { a.sd = fminf(a.sd, osd); return; }   // same point seen by both sides (or both empty)
    if (obd < a.bd || (obd == a.bd && obidx < a.bidx)) {
        a.sd = fminf(a.bd, osd); a.bd = obd; a.bpos = obpos; a.bidx = obidx;
    } else a.sd = fminf(a.sd, obd);
}

template <int G>
__device__ __forceinline__ void best_group_merge(Best &b, unsigned gmask)
{
    #pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const float obd = __shfl_xor_sync(gmask, b.bd, o), osd = __shfl_xor_sync(gmask, b.sd, o);
        const int obpos = __shfl_xor_sync(gmask, b.bpos, o), obidx = __shfl_xor_sync(gmask, b.bidx, o);
        best_merge(b, obd, obpos, obidx, osd);
    }
}

struct GridView {        // what a search needs of one index
    const GridParams *gp; const uint32_t *cell_start; const uint32_t *rowmask; const float4 *pts;
};

// Exact NN (and runner-up distance) of (px,py,pz) among the points of `g`, given an upper bound `lim`
// (exclusive) on the squared search radius.  Executed by a group of G lanes (gmask, sub-lane s; G = 1: every
// lane searches its own query), C rows per lane per round.  All lanes of a group pass the same query and
// return the same result.  On return `lim` is min(lim_in, runner-up): every point other than the winner is
// at least sqrt(lim) away.  `active` = false makes the caller idle (no rows) while staying convergent.
// The rows of the bounding square of the ball are walked in a flat loop; per round the row-mask words of all
// C rows are loaded first, then the cell-start pairs of the non-empty rows, then the candidates: the loads of
// a round are independent of each other, and rows that are pruned or empty cost no further memory traffic.
template <int G, int C>
__device__ __forceinline__ void ball_search(const GridView &g, const GridParams &gp, float px, float py, float pz, float &lim,
                                            Best &b, unsigned gmask, int s, bool active)
{
    best_init(b);
    const float inv_cell2 = gp.inv_cell * gp.inv_cell;
    const float fxq = grid_fcoord(px, gp.ox, gp.inv_cell), fyq = grid_fcoord(py, gp.oy, gp.inv_cell), fzq = grid_fcoord(pz, gp.oz, gp.inv_cell);
    const float Rc = sqrt_up(lim * inv_cell2) + GRID_MARGIN;
    const int z0 = max(__float2int_rd(fzq - Rc), 0), z1 = min(__float2int_rd(fzq + Rc), gp.nz - 1);
    const int y0 = max(__float2int_rd(fyq - Rc), 0), y1 = min(__float2int_rd(fyq + Rc), gp.ny - 1);
    const int ny_s = y1 - y0 + 1;
    const int total = (active && gp.n_points > 0 && lim < INFINITY && ny_s > 0 && z1 >= z0) ? ny_s * (z1 - z0 + 1) : 0;
    const int lastw = gp.words - 1;
    for (int t0 = 0; t0 < total; t0 += G * C) {
        STAT(5, s == 0);
        const float r2 = lim * inv_cell2;
        int xa[C], xb[C], rowi[C];
        uint32_t m0[C], m1[C];
        #pragma unroll
        for (int c = 0; c < C; ++c) {               // stage A: prune by the ball, fetch the occupancy words
            const int t = t0 + c * G + s;
            const int zi = t / ny_s;
            const int z = min(z0 + zi, gp.nz - 1), y = y0 + (t - zi * ny_s);
            const float gy = gap_iv(fyq, y, y + 1), gz = gap_iv(fzq, z, z + 1);
            const float row2 = gy * gy + gz * gz;
            const float ext = sqrt_up(fmaxf(r2 - row2, 0.f)) + GRID_MARGIN;
            xa[c] = max(__float2int_rd(fxq - ext), 0); xb[c] = min(__float2int_rd(fxq + ext), gp.nx - 1);
            const bool valid = (t < total) && (row2 < r2) && (xa[c] <= xb[c]);
            rowi[c] = valid ? z * gp.ny + y : -1;
            const uint32_t *mrow = g.rowmask + (size_t)max(rowi[c], 0) * gp.words;
            m0[c] = 0u; m1[c] = 0u;
            if (valid) m0[c] = __ldg(&mrow[xa[c] >> 5]);                              // predicated loads, issued back to back
            if (valid && (xb[c] >> 5) > (xa[c] >> 5)) m1[c] = __ldg(&mrow[min((xa[c] >> 5) + 1, lastw)]);
            STAT(4, valid);
        }
        uint32_t rs[C], re[C];
        #pragma unroll
        for (int c = 0; c < C; ++c) {               // stage B: first/last occupied cell of the x-run, fetch its point range
            const int wa = xa[c] >> 5, wb = xb[c] >> 5;
            int first = -1, last = -1;
            if (rowi[c] >= 0) {
                if (wb - wa >= 2) { first = xa[c]; last = xb[c]; }      // very wide run: take it whole
                else {
                    uint32_t a0 = m0[c] & (0xffffffffu << (xa[c] & 31));
                    uint32_t a1 = (wb > wa) ? m1[c] : 0u;
                    const uint32_t hi = 0xffffffffu >> (31 - (xb[c] & 31));
                    if (wb > wa) a1 &= hi; else a0 &= hi;
                    if (a0) first = (wa << 5) + __ffs(a0) - 1; else if (a1) first = (wb << 5) + __ffs(a1) - 1;
                    if (a1) last = (wb << 5) + 31 - __clz(a1); else if (a0) last = (wa << 5) + 31 - __clz(a0);
                }
            }
            rs[c] = 0u; re[c] = 0u;
            if (first >= 0) {
                const uint32_t *crow = g.cell_start + (size_t)rowi[c] * gp.nx;
                rs[c] = __ldg(&crow[first]); re[c] = __ldg(&crow[last + 1]);
            }
            STAT(2, first >= 0);
        }
        #pragma unroll
        for (int c = 0; c < C; ++c)                 // stage C: candidates
            if (rs[c] < re[c]) scan_range(g.pts, rs[c], re[c], px, py, pz, b);
        if (G > 1) best_group_merge<G>(b, gmask);   // afterwards all lanes of the group hold its best and runner-up
        lim = fminf(lim, b.sd);
    }
}

// Per-lane variant (every lane of the warp searches its own query at the same time).  Lanes of a warp have
// different rows and different candidate counts, so interleaving "look up a row" with "scan its candidates"
// serialises the warp on every row.  Here the two are separated: rows are enumerated (two at a time, loads in
// flight together) and the non-empty point ranges are queued in shared memory; then ONE flat loop walks all
// queued candidates four at a time.  The queue is flushed whenever it is full, which also tightens the limit
// for the rows still to come.  `q` points to this thread's queue slots: q[j * qstride], j < RANGE_QCAP.
#define RANGE_QCAP 8
__device__ __forceinline__ void ball_search_lane(const GridView &g, const GridParams &gp, float px, float py, float pz, float &lim,
                                                 Best &b, bool active, uint2 *q, int qstride)
{
    best_init(b);
    const float inv_cell2 = gp.inv_cell * gp.inv_cell;
    const float fxq = grid_fcoord(px, gp.ox, gp.inv_cell), fyq = grid_fcoord(py, gp.oy, gp.inv_cell), fzq = grid_fcoord(pz, gp.oz, gp.inv_cell);
    const float Rc = sqrt_up(lim * inv_cell2) + GRID_MARGIN;
    const int z0 = max(__float2int_rd(fzq - Rc), 0), z1 = min(__float2int_rd(fzq + Rc), gp.nz - 1);
    const int y0 = max(__float2int_rd(fyq - Rc), 0), y1 = min(__float2int_rd(fyq + Rc), gp.ny - 1);
    const int ny_s = y1 - y0 + 1;
    const int total = (active && gp.n_points > 0 && lim < INFINITY && ny_s > 0 && z1 >= z0) ? ny_s * (z1 - z0 + 1) : 0;
    int y = y0, z = z0, t = 0;
    while (t < total) {
        int nq = 0;
        while (t < total && nq <= RANGE_QCAP - 2) {          // ---- enumerate: two rows per step
            STAT(5, 1);
            const float r2 = lim * inv_cell2;
            int xa[2], xb[2], rowi[2];
            uint32_t m0[2], m1[2];
            #pragma unroll
            for (int c = 0; c < 2; ++c) {
                const float gy = gap_iv(fyq, y, y + 1), gz = gap_iv(fzq, z, z + 1);
                const float row2 = gy * gy + gz * gz;
                const float ext = sqrt_up(fmaxf(r2 - row2, 0.f)) + GRID_MARGIN;
                xa[c] = max(__float2int_rd(fxq - ext), 0); xb[c] = min(__float2int_rd(fxq + ext), gp.nx - 1);
                const bool valid = (t < total) && (row2 < r2) && (xa[c] <= xb[c]);
                rowi[c] = valid ? z * gp.ny + y : -1;
                const uint32_t *mrow = g.rowmask + (size_t)max(rowi[c], 0) * gp.words;
                m0[c] = 0u; m1[c] = 0u;
                if (valid) m0[c] = __ldg(&mrow[xa[c] >> 5]);
                if (valid && (xb[c] >> 5) > (xa[c] >> 5)) m1[c] = __ldg(&mrow[min((xa[c] >> 5) + 1, gp.words - 1)]);
                STAT(4, valid);
                ++t; ++y;
                if (y > y1) { y = y0; ++z; }
            }
            uint32_t rs[2], re[2];
            #pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int wa = xa[c] >> 5, wb = xb[c] >> 5;
                int first = -1, last = -1;
                if (rowi[c] >= 0) {
                    if (wb - wa >= 2) { first = xa[c]; last = xb[c]; }
                    else {
                        uint32_t a0 = m0[c] & (0xffffffffu << (xa[c] & 31));
                        uint32_t a1 = (wb > wa) ? m1[c] : 0u;
                        const uint32_t hi = 0xffffffffu >> (31 - (xb[c] & 31));
                        if (wb > wa) a1 &= hi; else a0 &= hi;
                        if (a0) first = (wa << 5) + __ffs(a0) - 1; else if (a1) first = (wb << 5) + __ffs(a1) - 1;
                        if (a1) last = (wb << 5) + 31 - __clz(a1); else if (a0) last = (wa << 5) + 31 - __clz(a0);
                    }
                }
                rs[c] = 0u; re[c] = 0u;
                if (first >= 0) {
                    const uint32_t *crow = g.cell_start + (size_t)rowi[c] * gp.nx;
                    rs[c] = __ldg(&crow[first]); re[c] = __ldg(&crow[last + 1]);
                }
                STAT(2, first >= 0);
            }
            #pragma unroll
            for (int c = 0; c < 2; ++c)
                if (rs[c] < re[c]) { q[nq * qstride] = make_uint2(rs[c], re[c]); ++nq; }
        }
        // ---- scan: one flat loop over all queued candidates
        if (nq > 0) {
            int j = 0;
            uint2 r = q[0];
            STAT(3, r.y - r.x);
            uint32_t k = r.x;
            while (true) {
                const uint32_t e = r.y;
                const uint32_t k1 = min(k + 1, e - 1), k2 = min(k + 2, e - 1), k3 = min(k + 3, e - 1);
                const float4 q0 = __ldg(&g.pts[k]), q1 = __ldg(&g.pts[k1]), q2 = __ldg(&g.pts[k2]), q3 = __ldg(&g.pts[k3]);
                nn_update(q0, k, px, py, pz, b);
                if (k + 1 < e) nn_update(q1, k1, px, py, pz, b);
                if (k + 2 < e) nn_update(q2, k2, px, py, pz, b);
                if (k + 3 < e) nn_update(q3, k3, px, py, pz, b);
                k += 4;
                if (k >= e) {
                    if (++j >= nq) break;
                    r = q[j * qstride]; k = r.x;
                    STAT(3, r.y - r.x);
                }
            }
            lim = fminf(lim, b.sd);
        }
    }
}

// Nearest point of the query's own cell (if any): when the clouds are roughly aligned this is (nearly) the answer
// and tightens the bound far below what a seed that slid along the surface gives.  One cell-start pair + one cell.
__device__ __forceinline__ float home_cell_probe(const GridView &g, const GridParams &gp, float px, float py, float pz)
{
    const float fx = grid_fcoord(px, gp.ox, gp.inv_cell), fy = grid_fcoord(py, gp.oy, gp.inv_cell), fz = grid_fcoord(pz, gp.oz, gp.inv_cell);
    if (fx < 0.f || fy < 0.f || fz < 0.f || fx >= (float)gp.nx || fy >= (float)gp.ny || fz >= (float)gp.nz) return INFINITY;
    const size_t ci = ((size_t)__float2int_rd(fz) * gp.ny + __float2int_rd(fy)) * gp.nx + __float2int_rd(fx);
    const uint32_t s0 = __ldg(&g.cell_start[ci]), e0 = __ldg(&g.cell_start[ci + 1]);
    Best b; best_init(b);
    if (s0 < e0) scan_range(g.pts, s0, e0, px, py, pz, b);
    return b.bd;
}

// Extra search radius beyond the seed distance.  It buys the skip test its margin (the bound on "every
// other point"), but it widens the cap of the surface the ball cuts out by ~2*slack*r, so it shrinks for
// far queries (whose runner-up is a fraction of a millimetre behind the winner anyway).
__device__ __forceinline__ float seed_limit(float d2, float slack0, float half_cell)
{
    const float dq = sqrtf(d2);
    const float f = fminf(1.f, half_cell / fmaxf(dq, 1e-12f));
    const float r = dq * 1.00001f + slack0 * f * f + 1e-6f;
    return r * r;
}
