// icp.cu -- the registration hot path: exact nearest-neighbour correspondence + fused
// point-to-plane (or point-to-point) accumulation + on-device 6x6 / Kabsch solve.
//
// Replaces GraphicEnd::multiPnP (reference src/GraphicEnd.cpp:557-659) with the PCL-1.7
// IterativeClosestPoint semantics restated in oracle/icp_oracle.c:
//   per iteration   X = T*src ; j(i) = argmin_j |X_i - Q_j|^2 (exact, lowest index on ties)
//                   reject d^2 > max_corr_dist^2 (and, point-to-plane, targets without a normal)
//                   A += J^T J, g += J^T r  with J = [X x n ; n], r = n.(Q - X)     (LLS)
//                   or  sums for Kabsch                                            (SVD)
//                   T <- dT * T
// Three exact search back ends (bit-identical results, tests/test_gpu_parity.py):
//   S3D_SEARCH_GRID       (default) icp_persist_kernel: ONE cooperative launch runs every iteration of
//                         every pair.  A group of CTAs owns a pair; each warp owns octets of 8 consecutive source
//                         points.  Per iteration: a decide pass (triangle-inequality skip test, near-tie check,
//                         else exact search in a TMA-staged shared-memory tile gathered from the target's grid,
//                         tile_search.cuh) and a call-free accumulate pass (29 double sums in registers); the CTAs
//                         of the group exchange one row of 29 doubles through L2, meet at a group barrier, and
//                         every CTA sums the rows in the same fixed order and solves the 6x6 (or Kabsch) in
//                         double: no host round trip, no kernel boundary between iterations.  While most
//                         queries still search (first iterations) a CTA hands its octets out dynamically from a
//                         shared-memory counter; from the ~6th iteration on >99.9% of the queries keep their
//                         correspondence and an iteration is two streaming passes over 48 B/point each.
//   S3D_SEARCH_GRID_LANE  icp_iter_kernel, one launch per iteration, per-lane ball search (search.cuh);
//                         the last CTA to finish (atomic ticket) solves.  Kept as an independent check.
//   S3D_SEARCH_BRUTE      all targets streamed through shared memory by TMA bulk copies
//                         (cp.async.bulk + mbarrier), 4 sources per thread in registers; the north star's
//                         literal kernel and the verification mode (FP32-issue bound).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include "context.h"
#include "common.cuh"
#include "grid.cuh"

#define ICP_BLOCK 256
#ifndef ICP_MIN_BLOCKS
#define ICP_MIN_BLOCKS 4      // 64 registers per thread: 32 warps per SM hide the gather latency
#endif
#include "search.cuh"
#include "tile_search.cuh"

// ------------------------------------------------------------------------------------------------
// brute-force exact NN: target tiles staged in shared memory by TMA bulk copies
// ------------------------------------------------------------------------------------------------
#define BF_SRC_PER_THREAD 4
#define BF_TILE 1024          // target points per tile (16 KiB)
#define BF_STAGES 4

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

__global__ void __launch_bounds__(ICP_BLOCK) nn_brute_tma_kernel(const PairDesc *__restrict__ descs, const PairState *__restrict__ states,
                                                                 int32_t *__restrict__ nn_idx, float *__restrict__ nn_d2, int nn_stride)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *tiles = reinterpret_cast<float4 *>(smem_raw);
    __shared__ __align__(8) uint64_t full[BF_STAGES];
    const int pair = blockIdx.y;
    const PairDesc d = descs[pair];
    if (states[pair].status != 0) return;
    const int base = blockIdx.x * (ICP_BLOCK * BF_SRC_PER_THREAD);
    if (base >= d.n_src) return;
    float T[12];
    #pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = states[pair].Tf[k];
    float px[BF_SRC_PER_THREAD], py[BF_SRC_PER_THREAD], pz[BF_SRC_PER_THREAD], bd[BF_SRC_PER_THREAD];
    int bi[BF_SRC_PER_THREAD];
    #pragma unroll
    for (int s = 0; s < BF_SRC_PER_THREAD; ++s) {
        int i = base + s * ICP_BLOCK + threadIdx.x;
        float4 p = i < d.n_src ? d.src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float3 x = s3d_xform(T, p.x, p.y, p.z);
        px[s] = x.x; py[s] = x.y; pz[s] = x.z; bd[s] = INFINITY; bi[s] = -1;
    }
    const int m = d.n_tgt;
    const int ntiles = (m + BF_TILE - 1) / BF_TILE;
    if (threadIdx.x == 0) {
        for (int s = 0; s < BF_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 0; t < BF_STAGES && t < ntiles; ++t) {
            uint32_t cnt = (uint32_t)min(BF_TILE, m - t * BF_TILE);
            mbar_expect_tx(&full[t], cnt * 16u);
            tma_bulk_g2s(tiles + t * BF_TILE, d.tgt + (size_t)t * BF_TILE, cnt * 16u, &full[t]);
        }
    }
    for (int t = 0; t < ntiles; ++t) {
        const int stage = t % BF_STAGES;
        mbar_wait(&full[stage], (uint32_t)((t / BF_STAGES) & 1));
        const int cnt = min(BF_TILE, m - t * BF_TILE);
        const float4 *tile = tiles + stage * BF_TILE;
        const int jbase = t * BF_TILE;
        #pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            float4 q = tile[k]; // same address for all lanes: shared-memory broadcast
            #pragma unroll
            for (int s = 0; s < BF_SRC_PER_THREAD; ++s) {
                float d2 = s3d_dist2(px[s], py[s], pz[s], q.x, q.y, q.z);
                if (d2 < bd[s]) { bd[s] = d2; bi[s] = jbase + k; } // ascending j + strict '<' = lowest index on ties
            }
        }
        __syncthreads(); // every thread is done reading this stage
        if (threadIdx.x == 0 && t + BF_STAGES < ntiles) {
            int tn = t + BF_STAGES;
            uint32_t c2 = (uint32_t)min(BF_TILE, m - tn * BF_TILE);
            mbar_expect_tx(&full[stage], c2 * 16u);
            tma_bulk_g2s(tiles + stage * BF_TILE, d.tgt + (size_t)tn * BF_TILE, c2 * 16u, &full[stage]);
        }
    }
    #pragma unroll
    for (int s = 0; s < BF_SRC_PER_THREAD; ++s) {
        int i = base + s * ICP_BLOCK + threadIdx.x;
        if (i < d.n_src) { nn_idx[(size_t)pair * nn_stride + i] = bi[s]; nn_d2[(size_t)pair * nn_stride + i] = bd[s]; }
    }
}

// ------------------------------------------------------------------------------------------------
// small dense solvers (double, one thread) -- same algorithms as oracle/icp_oracle.c
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chol6_solve(const sd A[6][6], const sd g[6], double pivot_eps, sd x[6])
{
    // Square-root-free Cholesky (A = L D L^T), fully unrolled: every index is a compile-time constant, so L, D, y live
    // in registers, and the dependent chain is 6 reciprocals instead of 6 square roots + 21 divisions.  Strict double
    // (common.cuh): the same IEEE operations in the same order as chol6_solve of oracle/icp_oracle.c -- identical bits,
    // including the rank-deficiency verdict.
    sd L[6][6], D[6], iD[6];
    int bad = 0;
    #pragma unroll
    for (int j = 0; j < 6; ++j) {
        sd s = A[j][j];
        #pragma unroll
        for (int k = 0; k < j; ++k) s = s - (L[j][k] * L[j][k]) * D[k];
        if (!(s.v > __dmul_rn(pivot_eps, A[j][j].v)) || !(A[j][j].v > 0.0)) bad = 1;
        D[j] = s; iD[j] = sd(1.0) / s;
        #pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            sd v = A[i][j];
            #pragma unroll
            for (int k = 0; k < j; ++k) v = v - (L[i][k] * L[j][k]) * D[k];
            L[i][j] = v * iD[j];
        }
    }
    if (bad) return 1;
    sd y[6];
    #pragma unroll
    for (int i = 0; i < 6; ++i) {
        sd v = g[i];
        #pragma unroll
        for (int k = 0; k < i; ++k) v = v - L[i][k] * y[k];
        y[i] = v;
    }
    #pragma unroll
    for (int i = 5; i >= 0; --i) {
        sd v = y[i] * iD[i];
        #pragma unroll
        for (int k = i + 1; k < 6; ++k) v = v - L[k][i] * x[k];
        x[i] = v;
    }
    return 0;
}

__device__ void euler_to_T(const sd x[6], sd D[12])
{
    double sa_, ca_, sb_, cb_, sg_, cg_;
    s3d_sincos(x[0].v, &sa_, &ca_); s3d_sincos(x[1].v, &sb_, &cb_); s3d_sincos(x[2].v, &sg_, &cg_);
    const sd sa(sa_), ca(ca_), sb(sb_), cb(cb_), sg(sg_), cg(cg_);
    D[0] = cg * cb; D[1] = -sg * ca + cg * sb * sa; D[2] = sg * sa + cg * sb * ca;  D[3] = x[3];
    D[4] = sg * cb; D[5] = cg * ca + sg * sb * sa;  D[6] = -cg * sa + sg * sb * ca; D[7] = x[4];
    D[8] = -sb;     D[9] = cb * sa;                 D[10] = cb * ca;                D[11] = x[5];
}

__device__ void kabsch_rotation(sd H[3][3], sd R[3][3])
{
    sd HtH[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        sd s(0.0); for (int k = 0; k < 3; ++k) s = s + H[k][i] * H[k][j];
        HtH[i][j] = s;
    }
    s3d_jacobi3(HtH, V, w);
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (w[ord[b]] > w[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    sd Vs[3][3], U[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]];
    for (int c = 0; c < 2; ++c) {
        sd u[3] = {sd(0.0), sd(0.0), sd(0.0)};
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) u[r] = u[r] + H[r][k] * Vs[k][c];
        sd n = sd_sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (n.v < 1e-300) n = sd(1.0);
        for (int r = 0; r < 3; ++r) U[r][c] = u[r] / n;
    }
    {
        sd d = U[0][0] * U[0][1] + U[1][0] * U[1][1] + U[2][0] * U[2][1];
        for (int r = 0; r < 3; ++r) U[r][1] = U[r][1] - d * U[r][0];
        sd n = sd_sqrt(U[0][1] * U[0][1] + U[1][1] * U[1][1] + U[2][1] * U[2][1]);
        if (n.v < 1e-300) n = sd(1.0);
        for (int r = 0; r < 3; ++r) U[r][1] = U[r][1] / n;
    }
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    Vs[0][2] = Vs[1][0] * Vs[2][1] - Vs[2][0] * Vs[1][1];
    Vs[1][2] = Vs[2][0] * Vs[0][1] - Vs[0][0] * Vs[2][1];
    Vs[2][2] = Vs[0][0] * Vs[1][1] - Vs[1][0] * Vs[0][1];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        sd s(0.0); for (int k = 0; k < 3; ++k) s = s + Vs[i][k] * U[j][k];
        R[i][j] = s;
    }
}

// consumes the summed accumulators of one pair (already converted to double by s3d_fx_value), updates its state
// (runs in one thread).  Strict double throughout: bit-identical to the tail of oracle_icp's iteration.
template <int EST>
__device__ __noinline__ void solve_and_update(const double *acc, PairState *st, int min_corr, double pivot_eps)
{
    const double cnt_d = acc[S3D_ACC_COUNT];
    const int cnt = (int)(cnt_d + 0.5);
    st->inliers = cnt;
    st->fitness = cnt > 0 ? __ddiv_rn(acc[S3D_ACC_SUMD2], cnt_d) : 0.0;
    if (cnt < min_corr) { st->status = S3D_PAIR_FEW_CORRESPONDENCES; return; }
    sd D[12];
    if (EST == S3D_ESTIMATOR_POINT_TO_PLANE) {
        sd A[6][6], g[6], x[6];
        #pragma unroll
        for (int a = 0; a < 6; ++a) {
            #pragma unroll
            for (int b = a; b < 6; ++b) { const int k = a * 6 - (a * (a - 1)) / 2 + (b - a); A[a][b] = sd(acc[k]); A[b][a] = sd(acc[k]); }
        }
        #pragma unroll
        for (int a = 0; a < 6; ++a) g[a] = sd(acc[21 + a]);
        if (chol6_solve(A, g, pivot_eps, x)) { st->status = S3D_PAIR_DEGENERATE; return; }
        euler_to_T(x, D);
    } else {
        sd H[3][3], R[3][3], pb[3], qb[3];
        const sd cn(cnt_d);
        for (int a = 0; a < 3; ++a) { pb[a] = sd(acc[a]) / cn; qb[a] = sd(acc[3 + a]) / cn; }
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[a][b] = sd(acc[6 + 3 * a + b]) - (cn * pb[a]) * qb[b];
        kabsch_rotation(H, R);
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) D[4 * a + b] = R[a][b];
            D[4 * a + 3] = qb[a] - (R[a][0] * pb[0] + R[a][1] * pb[1] + R[a][2] * pb[2]);
        }
    }
    bool finite = true;
    for (int k = 0; k < 12; ++k) finite = finite && isfinite(D[k].v);
    if (!finite) { st->status = S3D_PAIR_NONFINITE; return; }
    sd T[12], O[12];
    for (int k = 0; k < 12; ++k) T[k] = sd(st->T[k]);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) O[4 * r + c] = D[4 * r] * T[c] + D[4 * r + 1] * T[4 + c] + D[4 * r + 2] * T[8 + c];
        O[4 * r + 3] = D[4 * r] * T[3] + D[4 * r + 1] * T[7] + D[4 * r + 2] * T[11] + D[4 * r + 3];
    }
    for (int k = 0; k < 12; ++k) { st->T[k] = O[k].v; st->Tf_prev[k] = st->Tf[k]; st->Tf[k] = (float)O[k].v; }
    st->iterations += 1;
}

// ------------------------------------------------------------------------------------------------
// the fused iteration kernel
// ------------------------------------------------------------------------------------------------
// One accepted correspondence into the order-independent sums (common.cuh; contract in oracle/oracle_common.h): J and r
// in float32 with the oracle's expressions (identical bits), every product exact in double, rounded once to 2^-g by
// fma(a, b, M), the raw bits added into wrapping int64 accumulators.  The caller counts the terms (S3D_ACC_COUNT is not
// touched here) and removes count * bits(M) when it hands the partial sums on (fx_unbias).
#define S3D_FX_SEGMENT 240     // queries per thread between hand-overs: 240 * 32 lanes * 2^49 < 2^63
template <int EST>
__device__ __forceinline__ void accumulate_fx(long long *acc, double M, float px, float py, float pz, float4 q, float4 nv, float d2)
{
    if (EST == S3D_ESTIMATOR_POINT_TO_PLANE) {
        float J[6];
        J[0] = __fmaf_rn(nv.z, py, -__fmul_rn(nv.y, pz));
        J[1] = __fmaf_rn(nv.x, pz, -__fmul_rn(nv.z, px));
        J[2] = __fmaf_rn(nv.y, px, -__fmul_rn(nv.x, py));
        J[3] = nv.x; J[4] = nv.y; J[5] = nv.z;
        float ex = __fsub_rn(q.x, px), ey = __fsub_rn(q.y, py), ez = __fsub_rn(q.z, pz);
        float r = __fmaf_rn(nv.z, ez, __fmaf_rn(nv.y, ey, __fmul_rn(nv.x, ex)));
        double Jd[6];
        #pragma unroll
        for (int a = 0; a < 6; ++a) Jd[a] = (double)J[a];
        const double rd = (double)r;
        int k = 0;
        #pragma unroll
        for (int a = 0; a < 6; ++a) {
            #pragma unroll
            for (int b = a; b < 6; ++b) { acc[k] += s3d_fx_bits(Jd[a], Jd[b], M); ++k; }
        }
        #pragma unroll
        for (int a = 0; a < 6; ++a) acc[21 + a] += s3d_fx_bits(Jd[a], rd, M);
    } else {
        const double p[3] = {(double)px, (double)py, (double)pz}, qq[3] = {(double)q.x, (double)q.y, (double)q.z};
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            acc[a] += s3d_fx_bits(p[a], 1.0, M); acc[3 + a] += s3d_fx_bits(qq[a], 1.0, M);
            #pragma unroll
            for (int b = 0; b < 3; ++b) acc[6 + 3 * a + b] += s3d_fx_bits(p[a], qq[b], M);
        }
    }
    acc[S3D_ACC_SUMD2] += s3d_fx_bits((double)d2, 1.0, M);
}
// slots that carry fixed-point sums for this estimator (the others stay 0)
template <int EST> __device__ __forceinline__ bool fx_slot_used(int k)
{
    return k == S3D_ACC_SUMD2 || k < (EST == S3D_ESTIMATOR_POINT_TO_PLANE ? 27 : 15);
}
// raw accumulator of `cnt` terms -> the true partial sum (an integer number of 2^-g units)
template <int EST> __device__ __forceinline__ long long fx_unbias(long long raw, int k, int cnt, unsigned long long mbits)
{
    return fx_slot_used<EST>(k) ? raw - (long long)((unsigned long long)cnt * mbits) : 0ll;
}
// the 29 totals (hi, lo) of a pair -> doubles for the solve (slot 28 is the plain count)
template <int EST> __device__ __forceinline__ double fx_total(long long hi, long long lo, int k, double scale)
{
    if (k == S3D_ACC_COUNT) return __dadd_rn(__dmul_rn(__ll2double_rn(hi), 4294967296.0), __ll2double_rn(lo));
    return fx_slot_used<EST>(k) ? s3d_fx_value(hi, lo, scale) : 0.0;
}

template <int EST, int SEARCH>
__global__ void __launch_bounds__(ICP_BLOCK, ICP_MIN_BLOCKS) icp_iter_kernel(const PairDesc *__restrict__ descs, PairState *__restrict__ states,
                                                                             int32_t *__restrict__ nn_pos, float *__restrict__ nn_lb, int nn_stride, int use_seed,
                                                                             float max_d2)
{
    __shared__ uint2 rq[RANGE_QCAP * ICP_BLOCK];   // per-thread queues of candidate ranges (search.cuh)
    const int pair = blockIdx.y;
    PairState *st = states + pair;
    if (st->status != 0) return;   // failed pairs stay failed; every CTA of the pair takes this exit together
    const PairDesc d = descs[pair];
    float T[12];
    #pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->Tf[k];
    const int tid0 = blockIdx.x * ICP_BLOCK + threadIdx.x, tstride = gridDim.x * ICP_BLOCK;
    int32_t *my_pos = nn_pos + (size_t)pair * nn_stride;

    // ---- phase 1: exact nearest neighbour of every transformed source point (grid search) ----------
    // A query keeps, between iterations, its correspondence and a lower bound `lb` on its distance to
    // every OTHER target point.  When the pose update moved the point by less than the slack between the
    // two, the old correspondence is provably still the exact argmin and the search is skipped (triangle
    // inequality).  The rest are searched 4 at a time by groups of 8 lanes (search.cuh), in an order
    // (lanes 0,8,16,24 first, then the lanes between them, ...) that lets every query start from the
    // results of its already finished neighbours.
    if (SEARCH == S3D_SEARCH_GRID) {
        const GridParams gp = *d.grid;
        const GridView fine = {d.grid, d.cell_start, d.rowmask, d.sorted_pts};
        const bool have_coarse = !use_seed && d.coarse_grid != nullptr;
        GridParams cgp = gp;
        if (have_coarse) cgp = *d.coarse_grid;
        const GridView coarse = {d.coarse_grid, d.coarse_cell_start, d.coarse_rowmask, d.coarse_pts};
        float *my_lb = nn_lb + (size_t)pair * nn_stride;
        float Tp[12];
        #pragma unroll
        for (int k = 0; k < 12; ++k) Tp[k] = st->Tf_prev[k];
        const float gate = max_d2 < INFINITY ? max_d2 * 1.00001f + 1e-30f : INFINITY;
        const float slack0 = 0.2f * gp.cell, half_cell = 0.5f * gp.cell;
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        for (int base = tid0 - lane; base < d.n_src; base += tstride) {      // warp-uniform: 32 consecutive queries
            const int i = base + lane;
            const bool in = i < d.n_src;
            float3 x = make_float3(0.f, 0.f, 0.f);
            float own_lim = INFINITY, rqx = 0.f, rqy = 0.f, rqz = 0.f, rlb = 0.f;
            int rpos = -1;
            bool pending = false, has_res = false, store_pos = false;
            if (in) {
                const float4 p = d.src[i];
                x = s3d_xform(T, p.x, p.y, p.z);
                pending = true; store_pos = true;
                if (use_seed) {
                    const int sp_ = my_pos[i];
                    if (sp_ >= 0) {
                        const float4 q = __ldg(&d.sorted_pts[sp_]);
                        const float d2q = s3d_dist2(x.x, x.y, x.z, q.x, q.y, q.z);
                        const float3 xo = s3d_xform(Tp, p.x, p.y, p.z);
                        const float moved = sqrtf(s3d_dist2(x.x, x.y, x.z, xo.x, xo.y, xo.z)) * 1.000002f + 5e-8f;
                        const float lb = my_lb[i] - moved;
                        rqx = q.x; rqy = q.y; rqz = q.z; rpos = sp_; has_res = true;   // a real target point: also a seed for the neighbours
                        if (sqrtf(d2q) * 1.000002f + 2e-7f < lb) { rlb = lb; pending = false; store_pos = false; STAT(1, 1); }   // still the exact NN
                        else own_lim = seed_limit(d2q, slack0, half_cell);
                    }
                } else if (d.n_tgt > 0) {
                    // first iteration: the target point at the same relative index is the first upper bound
                    const int j0 = (int)(((long long)i * d.n_tgt) / d.n_src);
                    const float4 q = __ldg(&d.tgt[j0]);
                    own_lim = seed_limit(s3d_dist2(x.x, x.y, x.z, q.x, q.y, q.z), slack0, half_cell);
                }
                own_lim = fminf(own_lim, gate);   // nothing beyond the correspondence gate can be accepted anyway
            }
            const unsigned pend = __ballot_sync(full, pending);
            const unsigned resmask = __ballot_sync(full, has_res);
            // Own bound tightened by the correspondences of the neighbouring lanes (adjacent source points have
            // adjacent nearest neighbours) and by the nearest point of the query's own cell.
            float lim = own_lim;
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int nl = lane + ((k & 1) ? 1 : -1) * (1 << (k >> 1));
                const bool ok = nl >= 0 && nl < 32 && ((resmask >> (nl & 31)) & 1u);
                const int nlc = ok ? nl : lane;
                const float nx = __shfl_sync(full, rqx, nlc), ny = __shfl_sync(full, rqy, nlc), nz = __shfl_sync(full, rqz, nlc);
                if (ok && pending) lim = fminf(lim, seed_limit(s3d_dist2(x.x, x.y, x.z, nx, ny, nz), slack0, half_cell));
            }
            if (pending) {
                const float hd = home_cell_probe(fine, gp, x.x, x.y, x.z);
                if (hd < INFINITY) lim = fminf(lim, seed_limit(hd, slack0, half_cell));
            }
            // Work classes.  A search whose ball spans many rows of cells ("big") would keep its lane busy long after the
            // others are done; unless most of the warp is in that state, big searches are done one after the other by
            // the whole warp (32 rows per round), the small ones all at once, one per lane.
            const float rc_est = 2.f * sqrtf(lim) * gp.inv_cell + 1.f;
            const bool big_q = pending && (!(lim < INFINITY) || rc_est * rc_est > 24.f || have_coarse);
            unsigned bigm = __ballot_sync(full, big_q);
            if (__popc(pend) <= COOP) bigm = pend;
            else if (__popc(bigm) > 16) bigm = 0u;
            const bool mine_small = pending && !((bigm >> lane) & 1u);
#ifdef S3D_STATS
            {
                float rr = mine_small ? rc_est * rc_est : 0.f;
                float mx = rr;
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(full, mx, o));
                STAT(7, lane == 0 ? (unsigned long long)mx : 0ull);
                STAT(8, (unsigned long long)rr);
                STAT(9, (bigm >> lane) & 1u);
                STAT(10, lane == 0 && (pend & ~bigm));
                STAT(11, lane == 0 && pend);
            }
#endif
            if (pend & ~bigm) {
                if (have_coarse) {
                    Best cb; float clim = lim;
                    STAT(6, mine_small);
                    ball_search_lane(coarse, cgp, x.x, x.y, x.z, clim, cb, mine_small, rq + threadIdx.x, ICP_BLOCK);
                    if (cb.bpos >= 0) lim = fminf(lim, seed_limit(cb.bd, slack0, half_cell));
                }
                Best b;
                STAT(0, mine_small);
                ball_search_lane(fine, gp, x.x, x.y, x.z, lim, b, mine_small, rq + threadIdx.x, ICP_BLOCK);
                if (mine_small) { rpos = b.bpos; rlb = sqrtf(lim) * 0.999998f - 5e-8f; }   // lim already follows the runner-up
            }
            for (unsigned pm = bigm; pm; pm &= pm - 1) {
                const int ol = __ffs(pm) - 1;
                const float qx = __shfl_sync(full, x.x, ol), qy = __shfl_sync(full, x.y, ol), qz = __shfl_sync(full, x.z, ol);
                float wl = __shfl_sync(full, lim, ol);
                if (have_coarse) {
                    Best cb; float clim = wl;
                    STAT(6, lane == 0);
                    ball_search<32, 1>(coarse, cgp, qx, qy, qz, clim, cb, full, lane, true);
                    if (cb.bpos >= 0) wl = fminf(wl, seed_limit(cb.bd, slack0, half_cell));
                }
                Best b;
                STAT(0, lane == 0);
                ball_search<32, 1>(fine, gp, qx, qy, qz, wl, b, full, lane, true);
                if (lane == ol) { rpos = b.bpos; rlb = sqrtf(wl) * 0.999998f - 5e-8f; }
            }
            if (in) { my_lb[i] = rlb; if (store_pos) my_pos[i] = rpos; }
        }
    }

}

// Phase 2 of an iteration of the per-iteration-launch modes: residuals and normal-equation sums of the accepted
// correspondences (order-independent fixed-point sums), one (hi, lo) row per CTA, the last CTA of the pair (atomic ticket)
// adds the rows and solves.  A kernel of its own: 29 int64 accumulators do not fit the 64 registers of the search kernel.
#define S3D_ROW 64            // int64 per partial row: hi sums at [k], lo sums at [32 + k]
template <int EST, int SEARCH>
__global__ void __launch_bounds__(ICP_BLOCK, 2) icp_accum_kernel(const PairDesc *__restrict__ descs, PairState *__restrict__ states,
                                                                 long long *__restrict__ partials, const int32_t *__restrict__ nn_idx,
                                                                 const int32_t *__restrict__ nn_pos, int nn_stride,
                                                                 float max_d2, int min_corr, double pivot_eps, int32_t *__restrict__ nn_out)
{
    __shared__ long long whi[ICP_BLOCK / 32][S3D_NACC], wlo[ICP_BLOCK / 32][S3D_NACC];
    __shared__ long long thi[8][S3D_NACC], tlo[8][S3D_NACC];
    __shared__ double total[S3D_NACC];
    __shared__ FxScale fxs;
    __shared__ bool is_last;
    const int pair = blockIdx.y;
    PairState *st = states + pair;
    if (st->status != 0) return;   // failed pairs stay failed; every CTA of the pair takes this exit together
    const PairDesc d = descs[pair];
    if (threadIdx.x == 0) fxs = s3d_fx_make(s3d_icp_bound(*d.src_absmax, *d.tgt_absmax, EST == S3D_ESTIMATOR_POINT_TO_PLANE ? d.tgt_absmax[1] : 1.0f, st->T));
    float T[12];
    #pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->Tf[k];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 29) { whi[warp][lane] = 0; wlo[warp][lane] = 0; }
    __syncthreads();
    const unsigned long long mbits = fxs.mbits;
    const double M = __longlong_as_double((long long)mbits);
    const int tid0 = blockIdx.x * ICP_BLOCK + threadIdx.x, tstride = gridDim.x * ICP_BLOCK;
    const int32_t *my_pos = nn_pos + (size_t)pair * nn_stride;
    int total_cnt = 0;
    for (int i0 = tid0 - lane; i0 < d.n_src; i0 += tstride * S3D_FX_SEGMENT) {       // warp-uniform segments
        long long acc[29];
        #pragma unroll
        for (int k = 0; k < 29; ++k) acc[k] = 0;
        int cnt = 0;
        for (int s = 0; s < S3D_FX_SEGMENT; ++s) {
            const int i = i0 + lane + s * tstride;
            if (i >= d.n_src) break;
            const float4 p = d.src[i];
            const float3 x = s3d_xform(T, p.x, p.y, p.z);
            int j = -1; float4 q = make_float4(0.f, 0.f, 0.f, 0.f), nv = make_float4(0.f, 0.f, 0.f, 1.f);
            if (SEARCH == S3D_SEARCH_GRID) {
                const int bpos = my_pos[i];
                if (bpos >= 0) {
                    q = __ldg(&d.sorted_pts[bpos]);
                    j = __float_as_int(q.w);
                    if (EST == S3D_ESTIMATOR_POINT_TO_PLANE) nv = __ldg(&d.sorted_nrm[bpos]);
                }
            } else {
                j = nn_idx[(size_t)pair * nn_stride + i];
                if (j >= 0) {
                    q = __ldg(&d.tgt[j]);
                    if (EST == S3D_ESTIMATOR_POINT_TO_PLANE) nv = __ldg(&d.tgt_nrm[j]);
                }
            }
            const float bd = s3d_dist2(x.x, x.y, x.z, q.x, q.y, q.z);   // same expression as in the search: identical bits
            const bool ok = (j >= 0) && (bd <= max_d2) && (nv.w != 0.f);
            if (ok) { accumulate_fx<EST>(acc, M, x.x, x.y, x.z, q, nv, bd); ++cnt; }
            if (nn_out) nn_out[i] = ok ? j : -1;
        }
        total_cnt += cnt;
        // hand the segment over: true partial sums, added across the warp (< 2^63), split into (hi, lo)
        #pragma unroll
        for (int k = 0; k < 28; ++k) {
            long long v = fx_unbias<EST>(acc[k], k, cnt, mbits);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { long long hi, lo; s3d_fx_split(v, hi, lo); whi[warp][k] += hi; wlo[warp][k] += lo; }
        }
    }
    total_cnt = warp_sum_i(total_cnt);
    if (lane == 0) wlo[warp][S3D_ACC_COUNT] = total_cnt;
    __syncthreads();
    long long *row = partials + ((size_t)pair * gridDim.x + blockIdx.x) * S3D_ROW;
    if (threadIdx.x < 29) {
        long long hi = 0, lo = 0;
        #pragma unroll
        for (int w = 0; w < ICP_BLOCK / 32; ++w) { hi += whi[w][threadIdx.x]; lo += wlo[w][threadIdx.x]; }
        __stcg(&row[threadIdx.x], hi); __stcg(&row[32 + threadIdx.x], lo);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&st->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last CTA of the pair: sum of all partial rows (integers: any order gives the same bits), then solve
    {
        const int slot = threadIdx.x & 31, part = threadIdx.x >> 5; // 8 parts
        long long hi = 0, lo = 0;
        if (slot < 29) {
            const long long *base = partials + (size_t)pair * gridDim.x * S3D_ROW;
            for (int c = part; c < (int)gridDim.x; c += 8) { hi += __ldcg(&base[(size_t)c * S3D_ROW + slot]); lo += __ldcg(&base[(size_t)c * S3D_ROW + 32 + slot]); }
        }
        thi[part][slot] = hi; tlo[part][slot] = lo;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        long long hi = 0, lo = 0;
        #pragma unroll
        for (int p8 = 0; p8 < 8; ++p8) { hi += thi[p8][threadIdx.x]; lo += tlo[p8][threadIdx.x]; }
        total[threadIdx.x] = threadIdx.x < 29 ? fx_total<EST>(hi, lo, threadIdx.x, fxs.scale) : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        solve_and_update<EST>(total, st, min_corr, pivot_eps);
        st->ticket = 0u;
    }
}


// ------------------------------------------------------------------------------------------------
// the persistent kernel: all iterations of all pairs in one cooperative launch
// ------------------------------------------------------------------------------------------------
#ifndef TS_BLOCK
#define TS_BLOCK 512
#endif
static_assert(TS_CAP * 16 >= 29 * 33 * 8, "the warp tile also carries the transposed 29 x 33 int64 reduction");
#define TS_WARPS (TS_BLOCK / 32)
#define PS_CHUNK_BYTES 1600u                        // one staged chunk of 32 queries: point, correspondence, normal+bound (3 x 512 B), 32 flag bytes, padding
#define PS_STAGE ((TS_CAP * 16) / PS_CHUNK_BYTES)   // chunks a warp's tile stages at a time (6)
#ifndef PS_P1_UNROLL
#define PS_P1_UNROLL 1                               // chunks of the streaming pass in flight per warp (instruction-level parallelism)
#endif
#define PS_HIST 64                                  // poses kept (ring): a query's position at its last search is recomputed from them
#define PS_REBASE_AGE 48                            // a kept query is re-based onto the current pose before its ring slot is reused
#define PS_FLAG_TIE 64u
#define PS_FLAG_NRM 128u

#if defined(S3D_PHASES)
__device__ unsigned long long g_cta_t[148 * 8];      // debug: %globaltimer at the end of the search pass of iterations 0..7, per CTA
extern "C" int s3d_debug_cta_times(unsigned long long *out) { cudaDeviceSynchronize(); return (int)cudaMemcpyFromSymbol(out, g_cta_t, sizeof(g_cta_t)); }
__device__ __forceinline__ unsigned long long ps_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
#if defined(S3D_STATS) || defined(S3D_PHASES)
#define PHASE_T0() long long ph_t = clock64()
#define PHASE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long n_ = clock64(); atomicAdd(&g_stats[i], (unsigned long long)(n_ - ph_t)); ph_t = n_; } } while (0)
#else
#define PHASE_T0()
#define PHASE(i)
#endif

struct PersistArgs {
    const PairDesc *descs; PairState *states;
    long long *gacc;             // [3][groups][S3D_ROW]: the group's (hi, lo) sums of an iteration, added with integer atomics (any order
                                 // gives the same bits); three buffers by barrier epoch: one being added to, one being read, one being zeroed
    unsigned *barriers;          // [groups], zero at launch, monotonic
    // per-query state between iterations (49 bytes per query are read by the streaming pass):
    float4 *cq;                  // its correspondence (x,y,z, original target index; -1: none)
    float4 *cn;                  // the correspondence's normal (nx,ny,nz) and, in .w, the lower bound found at the query's last search on
                                 // its distance to every target point other than the correspondence (near ties: other than it and the runner-up)
    float4 *cq2;                 // near ties only (PS_FLAG_TIE): the runner-up (x,y,z, original index)
    uint8_t *flags;              // bits 0-5: ring slot of the iteration of the last search, PS_FLAG_TIE, PS_FLAG_NRM (the normal is valid)
    uint32_t *pend;              // [ctas][pend_stride]: octets of the CTA with queries that need a search, (local octet << 8) | lane mask
    long long nn_stride;         // queries per pair in the per-query arrays (a multiple of 8)
    int pend_stride;
    int n_pairs, groups, group_ctas, iterations;
    float max_d2; int min_corr; double pivot_eps;
    int32_t *nn_out;             // correspondences of the last iteration (single pair) or null
    float hint_cells;            // first-guess search radius after a big pose update, in cells
    float first_cells;           // first-guess search radius of the first iteration when the decimated index is not used, in cells
    int use_coarse;              // first iteration: bound the search with the nearest point of the decimated index
    float slack_cells;           // extra search radius beyond the seed distance, in cells: buys the skip test its margin
    int item_octets;             // octets a warp takes from the pending list at a time (1, 2 or 4: one per 8 lanes)
};

// How an iteration runs (one group of CTAs per pair; octet u = 8 consecutive source points belongs to CTA u mod group_ctas):
//   pass 1  streaming: every warp walks its share of the CTA's octets once, 49 bytes per query (point, correspondence, normal + bound,
//           flag byte; staged into the warp's tile by cp.async, all loads of up to PS_STAGE chunks in flight at once).  Skip test
//           (triangle inequality): the query was at xs = T[its] * p when it was last searched and every other target point was at least
//           `lb` away from there; if |x - xs| + |x - q| < lb (with explicit rounding allowances) the stored correspondence is provably
//           still the exact nearest neighbour, and its terms go straight into the 29 order-independent sums.  Near ties (runner-up
//           within 1e-5 m: the float noise of the test would fail them every time) carry their runner-up: both are evaluated exactly
//           and swapped if their order changed.  Queries that fail are appended, per octet, to the CTA's pending list.
//   pass 2  searching: warps take octets from the pending list (dynamically: search cost varies several-fold with depth and surface
//           orientation), search them in a TMA-staged shared-memory tile (tile_search.cuh), write the new per-query state and add
//           the terms of the searched queries.  In the first iteration every octet is pending and there is no pass 1.
//   then    the CTA's sums (integers) are added to the group's words with atomics that also count the CTAs that have added: a word
//           whose count is complete holds the sum (no barrier); every CTA reads the 58 words and solves redundantly in strict
//           double: all CTAs hold bit-identical poses.
// Because the sums do not depend on the order of the additions, a query can be accumulated by whichever pass decides it.
template <int EST, int BLOCK_, int CAP_, bool SEARCH_ONLY>
__global__ void __launch_bounds__(BLOCK_, 1) icp_persist_kernel(const PersistArgs a)
{
    // BLOCK_ threads, CAP_ candidates per warp tile.  SEARCH_ONLY: the instance that runs a pair's leading full-search iterations
    // (every query searched, no streaming pass) with more, leaner warps and hands the pair over -- through its PairState and the
    // per-query arrays -- to the general instance at the first iteration that streams.
    constexpr int K_WARPS = BLOCK_ / 32, K_CAP = CAP_, K_STAGE = (CAP_ * 16) / (int)PS_CHUNK_BYTES;
    static_assert(CAP_ * 16 >= 29 * 33 * 8, "the warp tile also carries the transposed 29 x 33 int64 reduction");
    extern __shared__ __align__(16) unsigned char ts_smem[];
    float4 *tiles = reinterpret_cast<float4 *>(ts_smem);          // K_WARPS tiles of K_CAP candidates
    __shared__ PairState st;                                      // this CTA's copy of the pair state (all CTAs of a group agree bit for bit)
    __shared__ float4 hist[PS_HIST][3];                           // float poses (three rows) of the last PS_HIST iterations (ring)
    __shared__ long long wrow[K_WARPS][S3D_ROW];                 // every warp's own (hi, lo) sums of this iteration (no 64-bit shared atomics:
                                                                  // they are compare-and-swap loops and 16 warps meet on the same 58 words)
    __shared__ long long ctot[S3D_ROW];                           // the CTA's (hi, lo) sums of this iteration
    __shared__ double total[S3D_NACC];                            // the pair's 29 sums as doubles, input of the solve
    __shared__ FxScale fxs;                                       // resolution of this iteration's sums (from the pose and the data bounds)
    __shared__ TileCfg cfg[2];                                    // search levels of the current pair: [0] decimated grid, [1] full grid
    __shared__ int pend_count, pend_next;                         // pending list of this CTA: entries appended / handed out
    __shared__ int full_next;                                     // the next iteration skips the streaming pass (big pose update)
#ifdef TS_USE_TMA
    __shared__ __align__(8) uint64_t tile_bar[K_WARPS];         // one mbarrier per warp: completion of its TMA row copies
#endif
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = blockIdx.x / a.group_ctas, rank = blockIdx.x - group * a.group_ctas;
    float4 *buf = tiles + warp * K_CAP;
    const uint32_t sbuf = ts_smem_u32(buf);
#ifdef TS_USE_TMA
    uint64_t *bar_w = &tile_bar[warp];
    uint32_t parity = 0u;
    if (lane == 0) {
        ts_mbar_init(bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#endif
    unsigned epoch = 0;
    const float gate_r = a.max_d2 < INFINITY ? sqrtf(a.max_d2) : INFINITY;
    uint32_t *my_pend = a.pend + (size_t)blockIdx.x * a.pend_stride;

    for (int pair = group; pair < a.n_pairs; pair += a.groups) {
        const PairDesc d = a.descs[pair];
        const bool have_coarse = d.coarse_grid != nullptr && a.use_coarse;
        __syncthreads();
        if (threadIdx.x == 0) {
            st = a.states[pair];
            fxs = s3d_fx_make(s3d_icp_bound(*d.src_absmax, *d.tgt_absmax, EST == S3D_ESTIMATOR_POINT_TO_PLANE ? d.tgt_absmax[1] : 1.0f, st.T));
            pend_count = 0; pend_next = 0; full_next = st.iterations == 0 ? 1 : 0;      // (a pair taken over from the search-only instance streams)
            cfg[1].gp = *d.grid; cfg[1].cell_start = d.cell_start; cfg[1].pts = d.sorted_pts;
            cfg[1].slack = a.slack_cells * cfg[1].gp.cell; cfg[1].gate_r = gate_r;
            cfg[0] = cfg[1];
            if (have_coarse) {
                cfg[0].gp = *d.coarse_grid; cfg[0].cell_start = d.coarse_cell_start; cfg[0].pts = d.coarse_pts;
                cfg[0].slack = 0.2f * cfg[0].gp.cell;
            }
        }
        if (threadIdx.x < S3D_ROW) ctot[threadIdx.x] = 0;
        __syncthreads();
        // the pose ring: the current pose and (a pair taken over after it_begin iterations: every query was last searched in
        // iteration it_begin - 1) the pose of the iteration before
        const int it_begin = st.iterations;
        if (threadIdx.x < 3) hist[it_begin & (PS_HIST - 1)][threadIdx.x] = make_float4(st.Tf[4 * threadIdx.x], st.Tf[4 * threadIdx.x + 1], st.Tf[4 * threadIdx.x + 2], st.Tf[4 * threadIdx.x + 3]);
        else if (threadIdx.x < 6 && it_begin > 0) {
            const int k = threadIdx.x - 3;
            hist[(it_begin - 1) & (PS_HIST - 1)][k] = make_float4(st.Tf_prev[4 * k], st.Tf_prev[4 * k + 1], st.Tf_prev[4 * k + 2], st.Tf_prev[4 * k + 3]);
        }
        const float cell = cfg[1].gp.cell, slack = cfg[1].slack, ccell = cfg[0].gp.cell;
        float4 *my_cq = a.cq + (size_t)pair * a.nn_stride;
        float4 *my_cn = a.cn + (size_t)pair * a.nn_stride;
        float4 *my_cq2 = a.cq2 + (size_t)pair * a.nn_stride;
        uint8_t *my_fl = a.flags + (size_t)pair * a.nn_stride;
        // Octet u belongs to CTA (u mod group_ctas): every CTA samples the whole cloud evenly (the search cost varies smoothly
        // with depth).  The CTA's octets are m = 0 .. cta_units-1 (u = rank + group_ctas * m); in pass 1 chunk c = four
        // consecutive local octets 4c .. 4c+3 (one per 8 lanes) and warp w walks chunks w, w + K_WARPS, ...
        const int nunits = (d.n_src + 7) >> 3;
        const int cta_units = rank < nunits ? (nunits - rank + a.group_ctas - 1) / a.group_ctas : 0;
        const int cta_chunks = (cta_units + 3) >> 2;
        __syncthreads();

        // stage K_STAGE chunks of per-query state (49 B per query) into the warp's tile: every load in flight at once, no registers
#define PS_STAGE_ROUND(c0_) do {                                                                                                      \
            _Pragma("unroll") for (int s_ = 0; s_ < K_STAGE; ++s_) {                                                            \
                if (4 * ((c0_) + s_ * K_WARPS) >= cta_units) break;                               /* warp-uniform */                 \
                const int m_ = 4 * ((c0_) + s_ * K_WARPS) + (lane >> 3);                                                             \
                const int i_ = ((rank + a.group_ctas * m_) << 3) + (lane & 7);                                                        \
                if (m_ < cta_units && i_ < d.n_src) {                                                                                 \
                    const uint32_t dst_ = sbuf + PS_CHUNK_BYTES * (uint32_t)s_ + 16u * (uint32_t)lane;                                \
                    ts_cp_async16_s(dst_, &d.src[i_]);                                                                                \
                    ts_cp_async16_s(dst_ + 512u, &my_cq[i_]);                                                                         \
                    ts_cp_async16_s(dst_ + 1024u, &my_cn[i_]);                                                                        \
                    if ((lane & 7) == 0)                                                                                              \
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sbuf + PS_CHUNK_BYTES * (uint32_t)s_ + 1536u + (uint32_t)lane), "l"(my_fl + i_) : "memory"); \
                }                                                                                                                     \
            }                                                                                                                         \
        } while (0)

        bool pre_staged = false;                  // the first staging round of the coming streaming pass is already in flight
        for (int it = it_begin; it < a.iterations; ++it) {
            if (st.status != 0) break;            // failed pairs stop; every CTA of the group sees the same state
            if (SEARCH_ONLY && it > 0 && !full_next) break;       // the first iteration that streams: the general instance takes over
            const float *T = st.Tf;               // the pose is read from shared memory where it is used
            const bool last = (it == a.iterations - 1);
            const unsigned long long mbits = fxs.mbits;
            const double M = __longlong_as_double((long long)mbits);
            PHASE_T0();
#if defined(S3D_PHASES)
            long long tm[5] = {0, 0, 0, 0, 0};
#endif

            // 29 wrapping int64 accumulators per thread.  hand_over(): unbias, add the 32 lanes through the warp's tile
            // (transposed: lane k adds slot k of the 32 lanes), split into (hi, lo), add to the CTA's sums (smem atomics).
            long long acc[29];
            #pragma unroll
            for (int k = 0; k < 29; ++k) acc[k] = 0;
            int cnt = 0;
            wrow[warp][lane] = 0; wrow[warp][32 + lane] = 0;
            __syncwarp();
#define PS_HAND_OVER() do {                                                                                               \
                long long *tr_ = reinterpret_cast<long long *>(buf);          /* 29 x 33 int64 <= K_CAP float4 */          \
                __syncwarp();                                                                                             \
                /* raw (wrapping) sums: the bias count * bits(M) of a slot is removed once, after the 32 lanes are added */ \
                _Pragma("unroll") for (int k = 0; k < 28; ++k) tr_[k * 33 + lane] = acc[k];                               \
                tr_[28 * 33 + lane] = (long long)cnt;                                                                     \
                __syncwarp();                                                                                             \
                long long v_ = 0;                                                                                         \
                if (lane < 29) {                                                                                          \
                    _Pragma("unroll 8") for (int l = 0; l < 32; ++l) v_ += tr_[lane * 33 + l];                            \
                }                                                                                                         \
                const long long cn_ = __shfl_sync(full, v_, 28);                                                          \
                if (lane < 28) v_ = fx_slot_used<EST>(lane) ? v_ - (long long)((unsigned long long)cn_ * mbits) : 0ll;    \
                if (lane < 29 && v_ != 0) {                                                                               \
                    long long hi_, lo_; s3d_fx_split(v_, hi_, lo_);                                                       \
                    if (lane == 28) { hi_ = 0; lo_ = v_; }                                                                \
                    wrow[warp][lane] += hi_;                                                                              \
                    wrow[warp][32 + lane] += lo_;                                                                         \
                }                                                                                                         \
                __syncwarp();                                                                                             \
                _Pragma("unroll") for (int k = 0; k < 29; ++k) acc[k] = 0;                                                \
                cnt = 0;                                                                                                  \
            } while (0)

            // ---------------------------------------------------------------- pass 1: streaming skip test + accumulate
            // After a big pose update (the iterations right after the first) practically every query fails the skip test: when
            // the last update can have moved a source point by more than a quarter of a cell (decided after the solve, the same
            // in every CTA) the streaming pass is skipped and every octet goes to the search pass (searching a query that would
            // have passed returns the same exact answer).
            const bool full_search = it == 0 || full_next;
            if (full_search) { ts_cp_async_wait_all(); __syncwarp(); }      // a staging round issued ahead for nothing: the tile is the search's now
            if (!SEARCH_ONLY && !full_search) {
                int since = 0;
                for (int c0 = warp; c0 < cta_chunks; c0 += K_WARPS * K_STAGE) {
#ifndef PS_NO_PREFETCH
                    if (c0 != warp || !pre_staged) PS_STAGE_ROUND(c0);       // (the first round was issued before the previous iteration's group sums,
                                                                             //  unless this is the first iteration of a pair taken over)
#else
                    PS_STAGE_ROUND(c0);
#endif
                    ts_cp_async_wait_all();
                    if (c0 == warp) PHASE(13);
                    __syncwarp();
                    constexpr int p1_unroll = PS_P1_UNROLL;
                    #pragma unroll p1_unroll
                    for (int s = 0; s < K_STAGE; ++s) {
                        const int m = 4 * (c0 + s * K_WARPS) + (lane >> 3);
                        if (4 * (c0 + s * K_WARPS) >= cta_units) break;                                  // warp-uniform
                        const int i = ((rank + a.group_ctas * m) << 3) + (lane & 7);
                        const bool in = m < cta_units && i < d.n_src;
                        bool pending = false;
                        if (in) {
                            const uint32_t sl = sbuf + PS_CHUNK_BYTES * (uint32_t)s + 16u * (uint32_t)lane;
                            const float4 p = ts_lds128(sl);
                            float4 q = ts_lds128(sl + 512u), nv = ts_lds128(sl + 1024u);
                            unsigned f;
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(f) : "r"(sbuf + PS_CHUNK_BYTES * (uint32_t)s + 1536u + (uint32_t)lane));
                            const float4 *Tn = hist[it & (PS_HIST - 1)];
                            const float3 x = s3d_xform4(Tn[0], Tn[1], Tn[2], p.x, p.y, p.z);
                            pending = true;
                            if (__float_as_int(q.w) >= 0) {
                                // the query was at xs when it was last searched; every other target point was >= lb away from there
                                const int its = (int)(f & 63u);
                                const float4 *Ts = hist[its];
                                const float3 xs = s3d_xform4(Ts[0], Ts[1], Ts[2], p.x, p.y, p.z);
                                float d2q = s3d_dist2(x.x, x.y, x.z, q.x, q.y, q.z);
                                const float moved = sqrtf(s3d_dist2(x.x, x.y, x.z, xs.x, xs.y, xs.z)) * 1.000002f + 5e-8f;
                                const float lb = nv.w;
                                bool keep;
                                if (!(f & PS_FLAG_TIE)) keep = sqrtf(d2q) * 1.000002f + 2e-7f < lb - moved;      // still the exact nearest neighbour
                                else {
                                    // near tie (rare): the runner-up q2 was kept and everything else is >= lb away.  Both are evaluated;
                                    // the nearer (lower original index on equality) is the exact nearest neighbour if it beats the bound.
                                    const float4 q2 = __ldcg(&my_cq2[i]);
                                    const float d2b = s3d_dist2(x.x, x.y, x.z, q2.x, q2.y, q2.z);
                                    const bool second = d2b < d2q || (d2b == d2q && __float_as_int(q2.w) < __float_as_int(q.w));
                                    keep = sqrtf(second ? d2b : d2q) * 1.000002f + 2e-7f < lb - moved;
                                    if (keep && second) {                         // the two swap roles: state rewritten
                                        my_cq[i] = q2; my_cq2[i] = q;
                                        float4 n2 = make_float4(0.f, 0.f, 0.f, 1.f);
                                        if (EST == S3D_ESTIMATOR_POINT_TO_PLANE) n2 = __ldg(&d.tgt_nrm[__float_as_int(q2.w)]);
                                        f = (f & ~PS_FLAG_NRM) | (n2.w != 0.f ? PS_FLAG_NRM : 0u);
                                        my_cn[i] = make_float4(n2.x, n2.y, n2.z, lb);
                                        my_fl[i] = (uint8_t)f;
                                        q = q2; d2q = d2b; nv = make_float4(n2.x, n2.y, n2.z, lb);
                                    }
                                }
                                if (keep) {
                                    pending = false;
                                    STAT(1, 1);
                                    const bool ok = (d2q <= a.max_d2) && (f & PS_FLAG_NRM);
                                    if (ok) { accumulate_fx<EST>(acc, M, x.x, x.y, x.z, q, nv, d2q); ++cnt; }
                                    if (last && a.nn_out) a.nn_out[i] = ok ? __float_as_int(q.w) : -1;
                                    if (((it - its) & 63) >= PS_REBASE_AGE) {
                                        // re-base onto the current pose before the ring slot is reused: everything else is at least
                                        // lb - moved away from where the query is now
                                        my_cn[i] = make_float4(nv.x, nv.y, nv.z, fmaxf((lb - moved) * 0.999999f - 1e-7f, 0.f));
                                        my_fl[i] = (uint8_t)((f & ~63u) | (unsigned)(it & 63));
                                    }
                                }
                            }
                        }
                        // queries that need a search: one list entry per octet
                        const unsigned pm = __ballot_sync(full, pending);
                        if (pm) {
                            const unsigned mine = (pm >> (lane & 24)) & 0xffu;
                            if ((lane & 7) == 0 && mine) {
                                const int slot = atomicAdd(&pend_count, 1);
                                my_pend[slot] = ((uint32_t)m << 8) | mine;
                            }
                        }
                    }
                    __syncwarp();
                    since += K_STAGE;
                    if (since >= S3D_FX_SEGMENT - K_STAGE) { PS_HAND_OVER(); since = 0; }      // warp-uniform
                }
                PHASE(14);
                PS_HAND_OVER();
                PHASE(15);
                __threadfence_block();
                __syncthreads();                  // the pending list is complete
            }
            PHASE(8);

            // ---------------------------------------------------------------- pass 2: search what is pending
            // Warps take items (item_octets octets, one per 8 lanes) from the CTA's list until it is empty.
            {
                const int n_items = full_search ? cta_units : pend_count;
                // few octets left (late iterations): one per item, so that a lone search is not queued behind another one
                const int item = n_items <= 2 * K_WARPS ? 1 : a.item_octets;
                while (n_items > 0) {
                    int e0 = 0;
                    if (lane == 0) e0 = atomicAdd(&pend_next, item);
                    e0 = __shfl_sync(full, e0, 0);
                    if (e0 >= n_items) break;
                    const int e = e0 + (lane >> 3);
                    uint32_t ent = 0u;
                    if (e < n_items && lane < 8 * item) ent = full_search ? (((uint32_t)e << 8) | 0xffu) : __ldcg(&my_pend[e]);
                    const int m = (int)(ent >> 8);
                    const int i = ((rank + a.group_ctas * m) << 3) + (lane & 7);
                    const bool pending = ((ent >> (lane & 7)) & 1u) && i < d.n_src;
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
                    if (pending) {
                        p = d.src[i];
                        if (it > 0) q = __ldcg(&my_cq[i]);
                    }
                    float3 x = make_float3(0.f, 0.f, 0.f);
                    float r = a.first_cells * cell;
                    if (pending) {
                        x = s3d_xform(T, p.x, p.y, p.z);
                        if (__float_as_int(q.w) >= 0) {
                            // The old correspondence is a real target point, so its distance bounds the ball.  After a small move it is
                            // also tight; after a big pose update (first iterations) the point slid along the surface and one cell
                            // is the better first guess (tile_search verifies and widens when needed).
                            const float dq = sqrtf(s3d_dist2(x.x, x.y, x.z, q.x, q.y, q.z));
                            const float3 xo = s3d_xform(st.Tf_prev, p.x, p.y, p.z);
                            const float step_mv = sqrtf(s3d_dist2(x.x, x.y, x.z, xo.x, xo.y, xo.z));
                            r = dq * 1.00001f + slack;
                            if (step_mv > 0.25f * cell) r = fminf(r, a.hint_cells * cell + slack);
                        }
                    }
                    // first iteration: the nearest point of the decimated target (level 0) bounds the fine search (level 1)
                    TileOut b;
                    for (int level = (it == 0 && have_coarse) ? 0 : 1; level < 2; ++level) {
                        STAT(level ? 0 : 6, pending);
#ifdef TS_USE_TMA
                        b = tile_search<K_CAP>(&cfg[level], x.x, x.y, x.z, level ? r : ccell, pending, buf, lane, level == 1 && it >= 4, bar_w, &parity TS_TM_PASS);
#else
                        b = tile_search<K_CAP>(&cfg[level], x.x, x.y, x.z, level ? r : ccell, pending, buf, lane, level == 1 && it >= 4 TS_TM_PASS);
#endif
                        if (level == 0 && pending && b.bd < INFINITY) r = sqrtf(b.bd) * 1.00001f + slack;
                    }
                    if (pending) {
                        q = b.bq;
                        const float d2q = b.bd;
                        if (!(b.bd < INFINITY)) q.w = __int_as_float(-1);      // nothing within reach
                        float4 nv = make_float4(0.f, 0.f, 0.f, 1.f);
                        if (EST == S3D_ESTIMATOR_POINT_TO_PLANE && __float_as_int(q.w) >= 0) nv = __ldg(&d.tgt_nrm[__float_as_int(q.w)]);
                        const bool tie = b.lb3 > 0.f;
                        const unsigned f = (unsigned)(it & 63) | (tie ? PS_FLAG_TIE : 0u) | (nv.w != 0.f ? PS_FLAG_NRM : 0u);
                        my_cq[i] = q;
                        my_cn[i] = make_float4(nv.x, nv.y, nv.z, tie ? b.lb3 : b.lb);
                        my_fl[i] = (uint8_t)f;
                        if (tie) my_cq2[i] = b.q2;                             // near tie: keep the runner-up too
                        const bool ok = (__float_as_int(q.w) >= 0) && (d2q <= a.max_d2) && (nv.w != 0.f);
                        if (ok) { accumulate_fx<EST>(acc, M, x.x, x.y, x.z, q, nv, d2q); ++cnt; }
                        if (last && a.nn_out) a.nn_out[i] = ok ? __float_as_int(q.w) : -1;
                    }
                    PS_HAND_OVER();
                }
            }
            PHASE(9);
#if defined(S3D_PHASES)
            __syncthreads();
            if (threadIdx.x == 0 && it < 8 && blockIdx.x < 148) g_cta_t[blockIdx.x * 8 + it] = ps_globaltimer();
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                #pragma unroll
                for (int k = 0; k < 5; ++k) atomicAdd(&g_stats[21 + k], (unsigned long long)tm[k]);
            }
#endif
            // ---------------------------------------------------------------- the group's sums, solve
            __syncthreads();
            // The per-query state of the next iteration is final (pass 2 wrote it) and does not depend on the pose: the first staging
            // round of its streaming pass is issued now, so that its latency hides behind the group sum and the solve.  (Should the next
            // iteration skip the streaming pass after all, the copies are simply drained.)
#ifndef PS_NO_PREFETCH
            pre_staged = !SEARCH_ONLY && !last;
            if (pre_staged && warp < cta_chunks) PS_STAGE_ROUND(warp);
#endif
            // Self-validating group sums: there is no barrier.  Lane k of warp 0 owns slot k: it adds the CTA's (hi, lo) to the group's
            // two words of the slot with ONE atomic each, and every atomic also carries +1 in bits 48.. of the word.  A word whose
            // count has reached the number of CTAs of the group holds the complete sum: the lane polls its own two words, nothing
            // else is waited for (no arrive counter, no release/acquire round trip between the data and a flag).  The sums are
            // normalised per CTA (0 <= lo < 2^32, |hi| < 2^47 for up to 2^30 terms), so the count never meets the value.
            // Three buffers by epoch: one being added to, one being read, one being zeroed (by rank 0, when it has seen epoch e
            // complete: every CTA has then finished reading epoch e-1, whose buffer is used again at e+2; rank 0's adds of epoch
            // e+1 are release operations, so the zeros are in place before anybody can see e+1 complete and move on to e+2).
            if (warp == 0) {
                long long hi = 0, lo = 0;
                if (lane < 29) {
                    #pragma unroll
                    for (int w = 0; w < K_WARPS; ++w) { hi += wrow[w][lane]; lo += wrow[w][32 + lane]; }
                    hi += lo >> 32; lo &= 0xffffffffll;
                }
                if (a.group_ctas > 1) {
                    long long *g = a.gacc + ((size_t)(epoch % 3u) * a.groups + group) * S3D_ROW;
                    const long long one = 1ll << 48;
                    if (lane < 29) {
                        if (rank == 0) {
                            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(&g[lane]), "l"(hi + one) : "memory");
                            asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(&g[32 + lane]), "l"(lo + one) : "memory");
                        } else {
                            atomicAdd(reinterpret_cast<unsigned long long *>(&g[lane]), (unsigned long long)(hi + one));
                            atomicAdd(reinterpret_cast<unsigned long long *>(&g[32 + lane]), (unsigned long long)(lo + one));
                        }
                        const long long n = (long long)a.group_ctas;
                        long long wh, wl;
                        do {
                            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(wh) : "l"(&g[lane]) : "memory");
                            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(wl) : "l"(&g[32 + lane]) : "memory");
                        } while (((wh + (one >> 1)) >> 48) < n || (wl >> 48) < n);
                        hi = wh - (n << 48); lo = wl - (n << 48);
                    }
                    ++epoch;
                    __syncwarp();
                    PHASE(10);
                    if (rank == 0) {
                        long long *gz = a.gacc + ((size_t)((epoch + 1u) % 3u) * a.groups + group) * S3D_ROW;
                        __stcg(&gz[lane], 0ll); __stcg(&gz[32 + lane], 0ll);
                    }
                }
                total[lane] = lane < 29 ? fx_total<EST>(hi, lo, lane, fxs.scale) : 0.0;
                __syncwarp();
            } else if (a.group_ctas > 1) ++epoch;
            PHASE(11);
            if (threadIdx.x == 0) {
                solve_and_update<EST>(total, &st, a.min_corr, a.pivot_eps);
                fxs = s3d_fx_make(s3d_icp_bound(*d.src_absmax, *d.tgt_absmax, EST == S3D_ESTIMATOR_POINT_TO_PLANE ? d.tgt_absmax[1] : 1.0f, st.T));
                #pragma unroll
                for (int k = 0; k < 3; ++k) hist[(it + 1) & (PS_HIST - 1)][k] = make_float4(st.Tf[4 * k], st.Tf[4 * k + 1], st.Tf[4 * k + 2], st.Tf[4 * k + 3]);
                pend_count = 0; pend_next = 0;
                // can the update have moved a source point by more than a quarter of a cell?  |dR| * 3 P + |dt| bounds the motion
                float dr = 0.f, dt = 0.f;
                #pragma unroll
                for (int r = 0; r < 3; ++r) {
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) dr = fmaxf(dr, fabsf(st.Tf[4 * r + c] - st.Tf_prev[4 * r + c]));
                    dt = fmaxf(dt, fabsf(st.Tf[4 * r + 3] - st.Tf_prev[4 * r + 3]));
                }
                full_next = (3.f * dr * (*d.src_absmax) + dt > 0.25f * cell) ? 1 : 0;
            }
            __syncthreads();
            PHASE(12);
        }
        ts_cp_async_wait_all();                   // a staging round issued ahead of an iteration that did not run (failed pair)
        __syncwarp();
        if (rank == 0 && threadIdx.x == 0) a.states[pair] = st;
    }
}

// the two instances: general (512 threads, 640-candidate tiles) and search-only (768 threads, 512-candidate tiles)
#ifndef TS_SEARCH_BLOCK
#define TS_SEARCH_BLOCK 768
#endif
#ifndef TS_SEARCH_CAP
#define TS_SEARCH_CAP 512
#endif
template <int EST> static const void *persist_fn(bool search_only)
{
    return search_only ? (const void *)icp_persist_kernel<EST, TS_SEARCH_BLOCK, TS_SEARCH_CAP, true>
                       : (const void *)icp_persist_kernel<EST, TS_BLOCK, TS_CAP, false>;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// How a batch is spread over the co-resident CTAs of the persistent kernel (pure host arithmetic, no CUDA call).
// Up to 16 pairs: one group each, all at once.  Larger batches: 4 groups of a quarter of the chip each walk the batch in
// rounds.  One small group per pair -- every pair at once -- made the launch as long as its most expensive pair and left
// 148 mod n_pairs SMs idle: 54.1 / 64.5 ms for two 64-pair shards of config 4 against 38.6 / 40.0 ms with 4 groups of 37
// CTAs (128 pairs: 75.5 / 75.8 / 73.6 ms with groups of 9 / 18 / 37).  Larger groups pay the per-iteration group sum +
// solve on more SMs: for 16 pairs 9-CTA groups keep a late iteration at 87 us where 49-CTA groups need 129 us.
extern "C" int s3d_batch_shape(int n_pairs, int n_points_max, int resident_ctas, int *groups_out, int *group_ctas_out)
{
    if (n_pairs <= 0 || n_points_max < 0 || resident_ctas <= 0 || !groups_out || !group_ctas_out) return S3D_E_ARG;
    const int chunks_per_cta = 2 * TS_WARPS;      // at least two chunks of 32 queries per warp before a pair is spread wider
    const int useful = std::max(1, (n_points_max + 32 * chunks_per_cta - 1) / (32 * chunks_per_cta));
    const int max_groups = n_pairs <= 16 ? n_pairs : 4;
    const int group_ctas = std::max(1, std::min(useful, resident_ctas / std::min(max_groups, resident_ctas)));
    *group_ctas_out = group_ctas;
    *groups_out = std::max(1, std::min(n_pairs, resident_ctas / group_ctas));
    return S3D_OK;
}

extern "C" void s3d_icp_params_default(s3d_icp_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->max_iterations = 10; p->max_corr_dist = 0.f; p->estimator = S3D_ESTIMATOR_POINT_TO_PLANE;
    p->search = S3D_SEARCH_GRID; p->grid_cell = 0.f; p->min_correspondences = 3; p->pivot_eps = 1e-9; p->reuse_index = 1;
}

static double pose_norm(const double *T)
{
    // |min(theta, 2pi-theta)| + 0.9 |t|   (reference src/GraphicEnd.cpp:618)
    double sx = T[9] - T[6], sy = T[2] - T[8], sz = T[4] - T[1];
    double s = 0.5 * sqrt(sx * sx + sy * sy + sz * sz);
    double c = 0.5 * (T[0] + T[5] + T[10] - 1.0);
    double theta = atan2(s, c), alt = 2.0 * M_PI - theta;
    double th = fabs(theta < alt ? theta : alt);
    return th + 0.9 * sqrt(T[3] * T[3] + T[7] * T[7] + T[11] * T[11]);
}

static int ensure_batch(s3d_ctx *ctx, int n_pairs, int ctas)
{
    if (n_pairs > ctx->cap_pairs) {
        cudaFree(ctx->d_desc); cudaFreeHost(ctx->h_desc); cudaFree(ctx->d_state); cudaFreeHost(ctx->h_state);
        ctx->d_desc = nullptr; ctx->h_desc = nullptr; ctx->d_state = nullptr; ctx->h_state = nullptr; ctx->cap_pairs = 0;
        int cap = std::max(n_pairs, 64);
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_desc, sizeof(PairDesc) * cap));
        S3D_CUDA(ctx, cudaMallocHost(&ctx->h_desc, sizeof(PairDesc) * cap));
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_state, sizeof(PairState) * cap));
        S3D_CUDA(ctx, cudaMallocHost(&ctx->h_state, sizeof(PairState) * cap));
        ctx->cap_pairs = cap;
        cudaFree(ctx->d_partials); ctx->d_partials = nullptr; ctx->cap_ctas = 0;
    }
    if ((size_t)ctx->cap_pairs * ctas > (size_t)ctx->cap_ctas) {
        cudaFree(ctx->d_partials); ctx->d_partials = nullptr;
        size_t rows = (size_t)ctx->cap_pairs * ctas;
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_partials, sizeof(long long) * S3D_ROW * rows));
        ctx->cap_ctas = (int)std::min<size_t>(rows, 0x7fffffff);
    }
    return S3D_OK;
}

// one iteration of the per-iteration-launch modes: [grid search kernel] + accumulate/solve kernel
template <int EST, int SEARCH>
static int launch_iter(s3d_ctx *ctx, dim3 grid, int nn_stride, int use_seed, float max_d2, int min_corr, double pivot_eps, int32_t *nn_out)
{
    if (SEARCH == S3D_SEARCH_GRID) {
        icp_iter_kernel<EST, SEARCH><<<grid, ICP_BLOCK, 0, ctx->stream>>>(ctx->d_desc, ctx->d_state, ctx->d_nn_pos, ctx->d_nn_d2, nn_stride, use_seed, max_d2);
        S3D_LAUNCHED(ctx);
    }
    icp_accum_kernel<EST, SEARCH><<<grid, ICP_BLOCK, 0, ctx->stream>>>(ctx->d_desc, ctx->d_state, ctx->d_partials, ctx->d_nn_idx, ctx->d_nn_pos,
                                                                       nn_stride, max_d2, min_corr, pivot_eps, nn_out);
    S3D_LAUNCHED(ctx);
    return S3D_OK;
}

// Enqueues one batch on the ctx stream (index builds where needed, pair descriptors, every iteration launch) and records
// ev[0..2] around the two stages; the final per-pair state is left in ctx->d_state.  No host synchronisation.
int s3d_register_issue(s3d_ctx *ctx, const s3d_cloud *const *src, const s3d_cloud *const *tgt,
                       const double *guess, int n_pairs, const s3d_icp_params *prm, bool *built_out, int *iter_launches_out)
{
    if (!ctx || !src || !tgt || !prm || n_pairs <= 0) return s3d_fail(ctx, S3D_E_ARG, "s3d_register_batch: bad argument");
    if (prm->estimator != S3D_ESTIMATOR_POINT_TO_PLANE && prm->estimator != S3D_ESTIMATOR_SVD) return s3d_fail(ctx, S3D_E_ARG, "unknown estimator");
    if (prm->search != S3D_SEARCH_GRID && prm->search != S3D_SEARCH_BRUTE && prm->search != S3D_SEARCH_GRID_LANE)
        return s3d_fail(ctx, S3D_E_ARG, "unknown search mode");
    if (prm->max_iterations < 0) return s3d_fail(ctx, S3D_E_ARG, "max_iterations < 0");
    cudaSetDevice(ctx->device);
    const bool plane = prm->estimator == S3D_ESTIMATOR_POINT_TO_PLANE;
    const bool use_grid = prm->search != S3D_SEARCH_BRUTE;
    const bool persist = prm->search == S3D_SEARCH_GRID;
    int n_max = 0;
    for (int i = 0; i < n_pairs; ++i) {
        if (!src[i] || !tgt[i]) return s3d_fail(ctx, S3D_E_ARG, "null cloud in batch");
        if (plane && !tgt[i]->d_nrm) return s3d_fail(ctx, S3D_E_STATE, "point-to-plane needs target normals: call s3d_segment_planes or s3d_cloud_set_normals on the target first");
        n_max = std::max(n_max, src[i]->n);
        int rc = s3d_cloud_ready(ctx, src[i]);
        if (rc == S3D_OK) rc = s3d_cloud_ready(ctx, tgt[i]);
        if (rc) return rc;
    }
    const int64_t launches0 = ctx->launches;
    // launch geometry: every CTA of the launch is resident at once (ICP_MIN_BLOCKS per SM), and a lone pair is
    // spread so that each thread walks the fewest possible queries one after the other (the kernel is latency bound)
    const int resident = ctx->sm_count * ICP_MIN_BLOCKS;
    const int rounds = std::max(1, (n_max + resident * ICP_BLOCK - 1) / (resident * ICP_BLOCK));
    int ctas = std::max(1, std::min((n_max + rounds * ICP_BLOCK - 1) / (rounds * ICP_BLOCK), std::max(1, resident / n_pairs)));
    // persistent path: groups of CTAs, all co-resident (cooperative launch), one group per pair at a time
    int p_groups = 1, p_group_ctas = 1;
    const size_t p_smem = sizeof(float4) * TS_CAP * TS_WARPS, p_smem_search = sizeof(float4) * TS_SEARCH_CAP * (TS_SEARCH_BLOCK / 32);
    if (persist) {
        if (ctx->persist_resident[plane ? 0 : 1] == 0) {
            int res = 1 << 30;
            for (int so = 0; so < 2; ++so) {          // both instances must be co-resident with the same grid: one CTA per SM each
                const void *fn = plane ? persist_fn<S3D_ESTIMATOR_POINT_TO_PLANE>(so != 0) : persist_fn<S3D_ESTIMATOR_SVD>(so != 0);
                const size_t smem = so ? p_smem_search : p_smem;
                S3D_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int per_sm = 0;
                S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, so ? TS_SEARCH_BLOCK : TS_BLOCK, smem));
                if (per_sm < 1) return s3d_fail(ctx, S3D_E_CUDA, "icp_persist_kernel does not fit on an SM");
                res = std::min(res, per_sm * ctx->sm_count);
            }
            ctx->persist_resident[plane ? 0 : 1] = res;
        }
        const int p_res = ctx->persist_resident[plane ? 0 : 1];
        s3d_batch_shape(n_pairs, n_max, p_res, &p_groups, &p_group_ctas);
        {
            static const char *e = getenv("S3D_GROUP_CTAS");          // developer override of the group size (tools/group_probe.py)
            const int v = e ? atoi(e) : 0;
            if (v > 0) {
                const int useful = std::max(1, (n_max + 64 * TS_WARPS - 1) / (64 * TS_WARPS));
                p_group_ctas = std::max(1, std::min(std::min(useful, p_res), v));
                p_groups = std::max(1, std::min(n_pairs, p_res / p_group_ctas));
            }
        }
        ctas = 3;                                     // d_partials doubles as the groups' accumulators: 3 epochs x groups (<= pairs) rows
    }
    int rc = ensure_batch(ctx, n_pairs, ctas);
    if (rc) return rc;

    cudaEvent_t *evs = ctx->ev_use ? ctx->ev_use : ctx->ev;
    PairDesc *h_desc = ctx->h_desc_use ? ctx->h_desc_use : ctx->h_desc;
    PairState *h_state = ctx->h_state_use ? ctx->h_state_use : ctx->h_state;
    cudaEventRecord(evs[0], ctx->stream);
    bool built = false;
    if (use_grid) {
        for (int i = 0; i < n_pairs; ++i) {
            s3d_cloud *t = const_cast<s3d_cloud *>(tgt[i]);
            bool seen = false;
            for (int k = 0; k < i && !seen; ++k) seen = (tgt[k] == tgt[i]);
            if (seen) continue;
            bool stale = !t->grid.valid || t->grid.requested_cell != prm->grid_cell || (plane && !t->grid.has_normals) || !prm->reuse_index;
            if (stale) { rc = s3d_grid_build(ctx, t, prm->grid_cell); if (rc) return rc; built = true; }
        }
    }
    cudaEventRecord(evs[1], ctx->stream);

    if (!persist) {
        size_t need = (size_t)n_pairs * n_max;
        if (need > (size_t)ctx->cap_nn) {
            cudaFree(ctx->d_nn_idx); cudaFree(ctx->d_nn_d2); cudaFree(ctx->d_nn_pos);
            ctx->d_nn_idx = nullptr; ctx->d_nn_d2 = nullptr; ctx->d_nn_pos = nullptr;
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_nn_idx, sizeof(int32_t) * std::max<size_t>(need, 1)));
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_nn_d2, sizeof(float) * std::max<size_t>(need, 1)));
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_nn_pos, sizeof(int32_t) * std::max<size_t>(need, 1)));
            ctx->cap_nn = (int)need;
        }
    }
    const int nn_stride8 = (std::max(n_max, 1) + 7) & ~7;          // per-pair stride of the per-query arrays: octets stay aligned
    const int pend_stride = persist ? ((nn_stride8 / 8) + p_group_ctas - 1) / p_group_ctas + 1 : 0;      // octets one CTA of a group owns
    if (persist) {
        size_t need = (size_t)n_pairs * nn_stride8;
        if (need > ctx->cap_tile_nn) {
            cudaFree(ctx->d_cq); cudaFree(ctx->d_cn); cudaFree(ctx->d_flags); cudaFree(ctx->d_cq2);
            ctx->d_cq = nullptr; ctx->d_cn = nullptr; ctx->d_flags = nullptr; ctx->d_cq2 = nullptr; ctx->cap_tile_nn = 0;
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_cq, sizeof(float4) * need));
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_cn, sizeof(float4) * need));
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_cq2, sizeof(float4) * need));
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_flags, need + 16));
            ctx->cap_tile_nn = need;
        }
        const size_t need_pend = (size_t)p_groups * p_group_ctas * pend_stride;
        if (need_pend > ctx->cap_pend) {
            cudaFree(ctx->d_pend); ctx->d_pend = nullptr; ctx->cap_pend = 0;
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_pend, sizeof(uint32_t) * need_pend));
            ctx->cap_pend = need_pend;
        }
        if (p_groups > ctx->cap_barriers) {
            cudaFree(ctx->d_barriers); ctx->d_barriers = nullptr; ctx->cap_barriers = 0;
            int cap = std::max(p_groups, 1024);
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_barriers, sizeof(unsigned) * cap));
            ctx->cap_barriers = cap;
        }
    }
    int32_t *nn_out = nullptr;
    if (n_pairs == 1) {
        if (src[0]->n > ctx->cap_last_nn) {
            cudaFree(ctx->d_last_nn); ctx->d_last_nn = nullptr;
            S3D_CUDA(ctx, cudaMalloc(&ctx->d_last_nn, sizeof(int32_t) * (size_t)std::max(src[0]->n, 1)));
            ctx->cap_last_nn = src[0]->n;
        }
        nn_out = ctx->d_last_nn; ctx->last_nn_n = src[0]->n;
    } else ctx->last_nn_n = 0;

    for (int i = 0; i < n_pairs; ++i) {
        PairDesc &d = h_desc[i];
        d.src = src[i]->d_pts; d.n_src = src[i]->n;
        d.tgt = tgt[i]->d_pts; d.tgt_nrm = tgt[i]->d_nrm; d.n_tgt = tgt[i]->n;
        d.sorted_pts = tgt[i]->grid.d_sorted_pts; d.sorted_nrm = tgt[i]->grid.d_sorted_nrm;
        d.cell_start = tgt[i]->grid.d_cell_start; d.grid = tgt[i]->grid.d_params; d.rowmask = tgt[i]->grid.d_rowmask;
        const bool coarse = use_grid && tgt[i]->coarse.valid;
        d.coarse_pts = coarse ? tgt[i]->coarse.d_sorted_pts : nullptr;
        d.coarse_cell_start = coarse ? tgt[i]->coarse.d_cell_start : nullptr;
        d.coarse_rowmask = coarse ? tgt[i]->coarse.d_rowmask : nullptr;
        d.coarse_grid = coarse ? tgt[i]->coarse.d_params : nullptr;
        rc = s3d_cloud_absmax(ctx, src[i], false);
        if (rc == S3D_OK) rc = s3d_cloud_absmax(ctx, tgt[i], plane);
        if (rc) return rc;
        d.src_absmax = src[i]->d_absmax; d.tgt_absmax = tgt[i]->d_absmax;
        PairState &s = h_state[i];
        memset(&s, 0, sizeof(s));
        static const double I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        const double *g = guess ? guess + 16 * (size_t)i : I12;
        for (int k = 0; k < 12; ++k) { s.T[k] = g[k]; s.Tf[k] = (float)g[k]; }
    }
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_desc, h_desc, sizeof(PairDesc) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_state, h_state, sizeof(PairState) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));

    const float max_d2 = prm->max_corr_dist > 0.f ? prm->max_corr_dist * prm->max_corr_dist : INFINITY;
    const int min_corr = prm->min_correspondences > 0 ? prm->min_correspondences : 3;
    const double pivot_eps = prm->pivot_eps > 0 ? prm->pivot_eps : 1e-9;
    const dim3 grid(ctas, n_pairs);
    int iter_launches = 0;
    if (persist && prm->max_iterations > 0) {
        S3D_CUDA(ctx, cudaMemsetAsync(ctx->d_barriers, 0, sizeof(unsigned) * p_groups, ctx->stream));
        S3D_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, sizeof(long long) * S3D_ROW * 3 * (size_t)p_groups, ctx->stream));
        PersistArgs pa;
        pa.descs = ctx->d_desc; pa.states = ctx->d_state; pa.gacc = ctx->d_partials; pa.barriers = ctx->d_barriers;
        pa.cq = ctx->d_cq; pa.cn = ctx->d_cn; pa.cq2 = ctx->d_cq2; pa.flags = ctx->d_flags; pa.pend = ctx->d_pend;
        pa.nn_stride = nn_stride8; pa.pend_stride = pend_stride;
        pa.n_pairs = n_pairs; pa.groups = p_groups; pa.group_ctas = p_group_ctas; pa.iterations = prm->max_iterations;
        pa.max_d2 = max_d2; pa.min_corr = min_corr; pa.pivot_eps = pivot_eps; pa.nn_out = nn_out;
        { static const char *e = getenv("S3D_HINT_CELLS"); pa.hint_cells = e ? (float)atof(e) : 1.0f; }
        { static const char *e = getenv("S3D_FIRST_CELLS"); pa.first_cells = e ? (float)atof(e) : 1.5f; }
        { static const char *e = getenv("S3D_USE_COARSE"); pa.use_coarse = e ? atoi(e) : 1; }
        { static const char *e = getenv("S3D_SLACK_CELLS"); pa.slack_cells = e ? (float)atof(e) : 0.08f; }
        { static const char *e = getenv("S3D_ITEM_OCTETS"); const int v = e ? atoi(e) : 2; pa.item_octets = v >= 4 ? 4 : v >= 2 ? 2 : 1; }
        void *kargs[] = {&pa};
        // Two launches: the search-only instance (768 threads: the search gains from resident warps, and without the streaming
        // pass its 58 accumulator registers do not have to stay live) runs every pair's leading full-search iterations, the
        // general instance takes each pair over -- PairState + per-query arrays -- at its first streaming iteration.
        static const bool split = []() { const char *e = getenv("S3D_SPLIT"); return !e || atoi(e) != 0; }();
        if (split) {
            const void *fs = plane ? persist_fn<S3D_ESTIMATOR_POINT_TO_PLANE>(true) : persist_fn<S3D_ESTIMATOR_SVD>(true);
            S3D_CUDA(ctx, cudaLaunchCooperativeKernel(fs, dim3(p_groups * p_group_ctas), dim3(TS_SEARCH_BLOCK), kargs, p_smem_search, ctx->stream));
            S3D_LAUNCHED(ctx); ++iter_launches;
            S3D_CUDA(ctx, cudaMemsetAsync(ctx->d_partials, 0, sizeof(long long) * S3D_ROW * 3 * (size_t)p_groups, ctx->stream));
        }
        const void *fn = plane ? persist_fn<S3D_ESTIMATOR_POINT_TO_PLANE>(false) : persist_fn<S3D_ESTIMATOR_SVD>(false);
        S3D_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(p_groups * p_group_ctas), dim3(TS_BLOCK), kargs, p_smem, ctx->stream));
        S3D_LAUNCHED(ctx); ++iter_launches;
    }
    for (int it = 0; it < prm->max_iterations && !persist; ++it) {
        int32_t *no = (it == prm->max_iterations - 1) ? nn_out : nullptr;
        if (!use_grid) {
            dim3 g2((n_max + ICP_BLOCK * BF_SRC_PER_THREAD - 1) / (ICP_BLOCK * BF_SRC_PER_THREAD), n_pairs);
            size_t smem = sizeof(float4) * BF_TILE * BF_STAGES;
            if (!ctx->brute_attr_done) {      // the attribute is per device, hence per ctx
                S3D_CUDA(ctx, cudaFuncSetAttribute(nn_brute_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                ctx->brute_attr_done = true;
            }
            nn_brute_tma_kernel<<<g2, ICP_BLOCK, smem, ctx->stream>>>(ctx->d_desc, ctx->d_state, ctx->d_nn_idx, ctx->d_nn_d2, n_max);
            S3D_LAUNCHED(ctx); ++iter_launches;
            rc = plane ? launch_iter<S3D_ESTIMATOR_POINT_TO_PLANE, S3D_SEARCH_BRUTE>(ctx, grid, n_max, 0, max_d2, min_corr, pivot_eps, no)
                       : launch_iter<S3D_ESTIMATOR_SVD, S3D_SEARCH_BRUTE>(ctx, grid, n_max, 0, max_d2, min_corr, pivot_eps, no);
            ++iter_launches;
        } else {
            rc = plane ? launch_iter<S3D_ESTIMATOR_POINT_TO_PLANE, S3D_SEARCH_GRID>(ctx, grid, n_max, it > 0, max_d2, min_corr, pivot_eps, no)
                       : launch_iter<S3D_ESTIMATOR_SVD, S3D_SEARCH_GRID>(ctx, grid, n_max, it > 0, max_d2, min_corr, pivot_eps, no);
            iter_launches += 2;
        }
        if (rc) return rc;
    }
    cudaEventRecord(evs[2], ctx->stream);
    *built_out = built; *iter_launches_out = iter_launches;
    ctx->timing.total_launches = (int)(ctx->launches - launches0);
    return S3D_OK;
}

// after the stream has been synchronised: event times of the two stages of the last s3d_register_issue
void s3d_register_timing(s3d_ctx *ctx, bool built, int iter_launches)
{
    float ms_index = 0.f, ms_iter = 0.f;
    cudaEventElapsedTime(&ms_index, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ms_iter, ctx->ev[1], ctx->ev[2]);
    ctx->timing.index_ms = built ? ms_index : 0.f;
    ctx->timing.iterate_ms = ms_iter;
    ctx->timing.iter_launches = iter_launches;
}

// A record as it leaves the device (T, fitness, inliers, iterations, status) gets, on the host, the reference's `norm`
// (src/GraphicEnd.cpp:618) and the reference's failure convention T == Identity (:585-600,621-624).
void s3d_result_finish(s3d_result *r)
{
    if (r->status == S3D_PAIR_OK) {
        r->T[12] = r->T[13] = r->T[14] = 0.0; r->T[15] = 1.0;
        r->norm = pose_norm(r->T);
    } else {
        memset(r->T, 0, sizeof(r->T));
        r->T[0] = r->T[5] = r->T[10] = r->T[15] = 1.0;
        r->norm = 0.0;
    }
}

// PairState -> s3d_result on the device: the records are formed where the solve left them, directly in the buffer the pose
// gather sends (gather.cu), so nothing but the gathered records ever crosses PCIe.  Slots >= n_pairs are marked absent.
__global__ void result_pack_kernel(const PairState *__restrict__ states, int n_pairs, s3d_result *__restrict__ out, int n_slots)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    s3d_result r;
    memset(&r, 0, sizeof(r));
    if (i < n_pairs) {
        const PairState &s = states[i];
        #pragma unroll
        for (int k = 0; k < 12; ++k) r.T[k] = s.T[k];
        r.T[15] = 1.0;
        r.fitness = s.fitness; r.inliers = s.inliers; r.iterations = s.iterations; r.status = s.status;
    } else r.status = S3D_PAIR_ABSENT;
    out[i] = r;
}

int s3d_result_pack(s3d_ctx *ctx, int n_pairs, s3d_result *d_out, int n_slots)
{
    result_pack_kernel<<<(n_slots + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_state, n_pairs, d_out, n_slots);
    S3D_LAUNCHED(ctx);
    return S3D_OK;
}

extern "C" int s3d_register_batch(s3d_ctx *ctx, const s3d_cloud *const *src, const s3d_cloud *const *tgt,
                                  const double *guess, int n_pairs, const s3d_icp_params *prm, s3d_result *out)
{
    if (!out) return s3d_fail(ctx, S3D_E_ARG, "s3d_register_batch: bad argument");
    bool built = false; int iter_launches = 0;
    int rc = s3d_register_issue(ctx, src, tgt, guess, n_pairs, prm, &built, &iter_launches);
    if (rc) return rc;
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(PairState) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_pairs; ++i) {
        const PairState &s = ctx->h_state[i];
        s3d_result &r = out[i];
        memset(&r, 0, sizeof(r));
        r.inliers = s.inliers; r.iterations = s.iterations; r.status = s.status; r.fitness = s.fitness;
        memcpy(r.T, s.T, sizeof(double) * 12);
        s3d_result_finish(&r);
    }
    s3d_register_timing(ctx, built, iter_launches);
    return S3D_OK;
}

// ---- stream of single-pair registrations ---------------------------------------------------------------------------
static int async_ready(s3d_ctx *ctx)
{
    if (ctx->d_async) return S3D_OK;
    S3D_CUDA(ctx, cudaMallocHost(&ctx->h_desc_ring, sizeof(PairDesc) * S3D_ASYNC_DEPTH));
    S3D_CUDA(ctx, cudaMallocHost(&ctx->h_state_ring, sizeof(PairState) * S3D_ASYNC_DEPTH));
    S3D_CUDA(ctx, cudaMallocHost(&ctx->h_async, sizeof(s3d_result) * S3D_ASYNC_DEPTH));
    for (int i = 0; i < S3D_ASYNC_DEPTH; ++i) for (int k = 0; k < 3; ++k) S3D_CUDA(ctx, cudaEventCreate(&ctx->ev_ring[i][k]));
    S3D_CUDA(ctx, cudaMalloc(&ctx->d_async, sizeof(s3d_result) * S3D_ASYNC_DEPTH));
    return S3D_OK;
}

extern "C" int s3d_register_enqueue(s3d_ctx *ctx, const s3d_cloud *src, const s3d_cloud *tgt, const double *guess, const s3d_icp_params *prm)
{
    if (!ctx) return S3D_E_ARG;
    if (ctx->async_n >= S3D_ASYNC_DEPTH) return s3d_fail(ctx, S3D_E_STATE, "s3d_register_enqueue: S3D_ASYNC_DEPTH registrations outstanding, call s3d_register_drain first");
    cudaSetDevice(ctx->device);
    int rc = async_ready(ctx);
    if (rc) return rc;
    const int slot = ctx->async_n;
    const s3d_cloud *s[1] = {src}, *t[1] = {tgt};
    bool built = false; int iter_launches = 0;
    // the descriptor and the initial state are staged in this slot's own page-locked memory (the copies are asynchronous),
    // the event times go to this slot's events
    ctx->h_desc_use = ctx->h_desc_ring + slot; ctx->h_state_use = ctx->h_state_ring + slot; ctx->ev_use = ctx->ev_ring[slot];
    rc = s3d_register_issue(ctx, s, t, guess, 1, prm, &built, &iter_launches);
    ctx->h_desc_use = nullptr; ctx->h_state_use = nullptr; ctx->ev_use = nullptr;
    if (rc) return rc;
    rc = s3d_result_pack(ctx, 1, ctx->d_async + slot, 1);       // PairState -> record, on the device, before the next pair reuses d_state
    if (rc) return rc;
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_async + slot, ctx->d_async + slot, sizeof(s3d_result), cudaMemcpyDeviceToHost, ctx->stream));   // the record goes home right behind its pair
    ctx->async_built[slot] = built; ctx->async_launches[slot] = iter_launches; ctx->async_total_launches[slot] = ctx->timing.total_launches + 1;
    ctx->async_n = slot + 1;
    return S3D_OK;
}

extern "C" int s3d_register_drain(s3d_ctx *ctx, s3d_result *results_out, s3d_timing *timing_out, int capacity, int *n_out)
{
    if (!ctx || !results_out || !n_out) return s3d_fail(ctx, S3D_E_ARG, "s3d_register_drain: bad argument");
    const int n = ctx->async_n;
    if (capacity < n) return s3d_fail(ctx, S3D_E_ARG, "s3d_register_drain: capacity below the number of outstanding registrations");
    *n_out = 0;
    if (n == 0) return S3D_OK;
    cudaSetDevice(ctx->device);
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->async_n = 0;
    for (int i = 0; i < n; ++i) {
        results_out[i] = ctx->h_async[i];
        s3d_result_finish(&results_out[i]);
        float ms_index = 0.f, ms_iter = 0.f;
        cudaEventElapsedTime(&ms_index, ctx->ev_ring[i][0], ctx->ev_ring[i][1]);
        cudaEventElapsedTime(&ms_iter, ctx->ev_ring[i][1], ctx->ev_ring[i][2]);
        s3d_timing tm; memset(&tm, 0, sizeof(tm));
        tm.index_ms = ctx->async_built[i] ? ms_index : 0.f; tm.iterate_ms = ms_iter;
        tm.iter_launches = ctx->async_launches[i]; tm.total_launches = ctx->async_total_launches[i];
        if (timing_out) timing_out[i] = tm;
        ctx->timing = tm;
    }
    *n_out = n;
    return S3D_OK;
}

extern "C" int s3d_register_pair(s3d_ctx *ctx, const s3d_cloud *src, const s3d_cloud *tgt, const double *guess,
                                 const s3d_icp_params *params, s3d_result *result_out)
{
    const s3d_cloud *s[1] = {src}, *t[1] = {tgt};
    return s3d_register_batch(ctx, s, t, guess, 1, params, result_out);
}

extern "C" int s3d_last_correspondences(s3d_ctx *ctx, int32_t *idx_out, int n)
{
    if (!ctx || !idx_out || n != ctx->last_nn_n || n <= 0) return s3d_fail(ctx, S3D_E_ARG, "s3d_last_correspondences: size mismatch or no single-pair call before");
    cudaSetDevice(ctx->device);
    S3D_CUDA(ctx, cudaMemcpyAsync(idx_out, ctx->d_last_nn, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return S3D_OK;
}

extern "C" int s3d_last_timing(const s3d_ctx *ctx, s3d_timing *out)
{
    if (!ctx || !out) return S3D_E_ARG;
    *out = ctx->timing;
    return S3D_OK;
}

#if defined(S3D_STATS) || defined(S3D_PHASES)
extern "C" int s3d_debug_stats(unsigned long long *out32, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, g_stats, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_stats, z, sizeof(z)); }
    return 0;
}
#endif
