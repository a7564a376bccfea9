// planes.cu -- RANSAC plane extraction on the device.
//
// Replaces GraphicEnd::extractPlanesAndGenerateImage (reference src/GraphicEnd.cpp:353-430), i.e. the
// loop  { pcl::SACSegmentation::segment ; flip sign so d >= 0 ; ExtractIndices positive/negative }
// with the PCL-1.7 semantics restated in oracle/plane_oracle.c.  All candidate planes of a round are
// evaluated in ONE streaming pass over the remaining points (16 B per point for up to 64 candidates); the
// sequential adaptive-stop logic of pcl::RandomSampleConsensus::computeModel is then replayed over
// the counts by a single thread, so the chosen model is exactly the one the sequential loop picks.
// The per-point result (plane label + plane normal) is what the ICP uses as target normals.
//
// The whole extraction is driven from the device: the host enqueues max_planes rounds of five kernels
// (hypotheses / evaluate + replay / PCA refit / count / compact) with worst-case grids; every kernel reads the
// loop state (points left, planes found, stopped) from device memory and returns at once when the
// reference's `while (remaining > percent * n)` loop (src/GraphicEnd.cpp:372) would have ended.  One
// read-back at the end returns the planes.  The PCA sums are order-independent fixed-point sums and the
// refit runs in strict double (common.cuh), so coefficients and labels equal the oracle's bit for bit.
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "context.h"
#include "common.cuh"
#include "compact.cuh"

#define S3D_FX_SEGMENT_P 240   // points per thread between hand-overs of the fixed-point sums (see icp.cu)

#define PLANE_CANDIDATES_EXTRA 14
#define PLANE_MAX_CAND 1024
#define PLANE_CHUNK 64
#define PLANE_BLOCK 256

struct PlaneState {
    int n_rem;          // points not yet assigned to a plane (compacted, order kept)
    int n_planes;
    int stopped;        // the reference's loop has ended (condition false or one of its `break`s taken)
    int active;         // this round runs (decided by the hypothesis kernel from the three fields above)
    float4 ransac;      // model picked by the RANSAC replay
    float4 refined;     // after PCA refit + sign flip
    int best_count, iterations;
    unsigned ticket_eval, ticket_refit, ticket_write, pad;
    int rem_at_round[S3D_MAX_PLANES];     // points scanned by the evaluation pass of each round (roofline bookkeeping)
    s3d_plane planes[S3D_MAX_PLANES];
};

static_assert(sizeof(PlaneState) == 528, "bench.py counts this many bytes per extraction in e2e.d2h_bytes_per_step");

__global__ void plane_init_kernel(const float4 *__restrict__ pts, int n, float4 *__restrict__ rem, int32_t *__restrict__ labels,
                                  float4 *__restrict__ nrm, PlaneState *st)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_rem = n; st->n_planes = 0; st->stopped = 0; st->active = 0;
        st->ticket_eval = st->ticket_refit = st->ticket_write = 0u;
        for (int k = 0; k < S3D_MAX_PLANES; ++k) st->rem_at_round[k] = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        rem[i] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        labels[i] = -1;
        nrm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// Loop test of the reference (src/GraphicEnd.cpp:372,424) + candidate planes of this round.  One CTA.
__global__ void __launch_bounds__(PLANE_MAX_CAND) plane_hyp_kernel(const float4 *__restrict__ rem, int n, float percent, int max_planes,
                                                                   uint64_t seed, int round, int n_cand, float4 *__restrict__ coefs,
                                                                   int *__restrict__ valid, uint32_t *__restrict__ counts, PlaneState *st)
{
    const int n_rem = st->n_rem, n_planes = st->n_planes;
    const bool go = !st->stopped && (double)n_rem > (double)percent * (double)n && n_planes < max_planes && n_rem >= 3;
    __syncthreads();                 // everybody has read the state before thread 0 changes it
    const int c = threadIdx.x;
    if (c == 0) {
        st->active = go ? 1 : 0;
        if (!go) st->stopped = 1;
        else st->rem_at_round[round] = n_rem;
    }
    if (!go || c >= n_cand) return;
    uint32_t s[3];
    s3d_sample3(seed, (uint64_t)n_planes, (uint64_t)c, (uint32_t)n_rem, s);
    float4 a = rem[s[0]], b = rem[s[1]], d = rem[s[2]];
    float4 coef = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = s3d_plane_from3(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(d.x, d.y, d.z), coef);
    coefs[c] = coef; valid[c] = ok ? 1 : 0; counts[c] = 0u;
}

// ONE pass over the remaining points evaluates up to PLANE_CHUNK (64) candidates (chunk index = blockIdx.y; the stock
// 50 + 14 candidates are one chunk): 16 bytes per point and pass.  A lane holds ONE point at a time and walks the candidates
// (coefficients broadcast from shared memory); the inliers of a candidate among the warp's 32 points are a ballot, and lane
// k (k + 32) keeps the running count of candidate k: two counters per thread instead of 64, so the kernel runs at full
// occupancy.  Counts: shared-memory atomics per warp, one global atomic per candidate and CTA (integers: deterministic).
// The last CTA to finish replays pcl::RandomSampleConsensus::computeModel over the counts: the adaptive-stop value
// k(c) = log(1 - p) / log(1 - w(c)^3) of every candidate is computed in parallel, one thread then walks the candidates
// in order exactly like the sequential loop (the chosen model is the one PCL's loop picks).
__global__ void __launch_bounds__(PLANE_BLOCK) plane_eval_kernel(const float4 *__restrict__ rem, const float4 *__restrict__ coefs,
                                                                 const int *__restrict__ valid, int n_cand, float tau,
                                                                 uint32_t *__restrict__ counts, int max_iterations, double probability,
                                                                 PlaneState *st)
{
    if (!st->active) return;
    const int n_rem = st->n_rem;
    __shared__ float4 sc[PLANE_CHUNK];
    __shared__ uint32_t s_cnt[PLANE_CHUNK];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31;
    const int c0 = blockIdx.y * PLANE_CHUNK;
    if (threadIdx.x < PLANE_CHUNK) {
        int c = c0 + threadIdx.x;
        // invalid candidates get a plane no point can satisfy (NaN compares false)
        sc[threadIdx.x] = (c < n_cand && valid[c]) ? coefs[c] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
        s_cnt[threadIdx.x] = 0u;
    }
    __syncthreads();
    int c_lo = 0, c_hi = 0;
    for (int base = blockIdx.x * PLANE_BLOCK + threadIdx.x - lane; base < n_rem; base += gridDim.x * PLANE_BLOCK) {      // warp-uniform
        const int i = base + lane;
        const bool in = i < n_rem;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in) p = rem[i];
        #pragma unroll 16
        for (int k = 0; k < 32; ++k) {
            const float4 a = sc[k], b = sc[k + 32];
            const unsigned ba = __ballot_sync(0xffffffffu, in && fabsf(s3d_plane_eval(a.x, a.y, a.z, a.w, p.x, p.y, p.z)) < tau);
            const unsigned bb = __ballot_sync(0xffffffffu, in && fabsf(s3d_plane_eval(b.x, b.y, b.z, b.w, p.x, p.y, p.z)) < tau);
            if (lane == k) { c_lo += __popc(ba); c_hi += __popc(bb); }
        }
    }
    if (c_lo) atomicAdd(&s_cnt[lane], (uint32_t)c_lo);
    if (c_hi) atomicAdd(&s_cnt[lane + 32], (uint32_t)c_hi);
    __syncthreads();
    if (threadIdx.x < PLANE_CHUNK && s_cnt[threadIdx.x] && c0 + threadIdx.x < n_cand) atomicAdd(&counts[c0 + threadIdx.x], s_cnt[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->ticket_eval, 1u) == gridDim.x * gridDim.y - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // ---- replay of the sequential adaptive RANSAC loop over the pre-evaluated candidates
    __shared__ int s_count[PLANE_MAX_CAND];
    __shared__ double s_k[PLANE_MAX_CAND];
    {
        const double log_prob = log(1.0 - probability);
        const double one_over = n_rem > 0 ? 1.0 / (double)n_rem : 0.0;
        const double eps = 2.220446049250313e-16;
        for (int c = threadIdx.x; c < n_cand; c += PLANE_BLOCK) {
            const int cc = valid[c] ? (int)__ldcg(&counts[c]) : -1;                 // -1: invalid sample (skipped, iterations unchanged)
            s_count[c] = cc;
            double w = cc * one_over;
            double p_no = 1.0 - w * w * w;
            if (p_no < eps) p_no = eps;
            if (p_no > 1.0 - eps) p_no = 1.0 - eps;
            s_k[c] = log_prob / log(p_no);
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int iterations = 0, best = -1, best_count = -2147483647;
    double k = 1.0;
    for (int c = 0; c < n_cand && iterations < k; ++c) {
        const int cc = s_count[c];
        if (cc < 0) continue;
        if (cc > best_count) { best_count = cc; best = c; k = s_k[c]; }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    st->ticket_eval = 0u;
    st->best_count = best >= 0 ? best_count : 0; st->iterations = iterations;
    if (best < 0 || best_count == 0) { st->stopped = 1; st->active = 0; }          // reference :376-379
    else { st->ransac = coefs[best]; st->refined = coefs[best]; }
}

// PCA refit over the inliers of the RANSAC model (SampleConsensusModelPlane::optimizeModelCoefficients): nine
// order-independent fixed-point sums + the count; the last CTA converts the totals and refits in strict double.
__global__ void __launch_bounds__(PLANE_BLOCK) plane_refit_kernel(const float4 *__restrict__ rem, float tau, const float *__restrict__ absmax,
                                                                  PlaneState *st, long long *__restrict__ partials)
{
    if (!st->active) return;
    const int n_rem = st->n_rem;
    __shared__ long long whi[PLANE_BLOCK / 32][10], wlo[PLANE_BLOCK / 32][10];
    __shared__ bool is_last;
    const float4 m = st->ransac;
    const FxScale fx = s3d_fx_make(s3d_pca_bound(*absmax));
    const double M = __longlong_as_double((long long)fx.mbits);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < 10) { whi[warp][lane] = 0; wlo[warp][lane] = 0; }
    __syncwarp();
    const int tstride = gridDim.x * PLANE_BLOCK;
    int total_cnt = 0;
    for (int i0 = blockIdx.x * PLANE_BLOCK + threadIdx.x - lane; i0 < n_rem; i0 += tstride * S3D_FX_SEGMENT_P) {     // warp-uniform segments
        long long s[9];
        #pragma unroll
        for (int k = 0; k < 9; ++k) s[k] = 0;
        int cnt = 0;
        for (int q = 0; q < S3D_FX_SEGMENT_P; ++q) {
            const int i = i0 + lane + q * tstride;
            if (i >= n_rem) break;
            const float4 p = rem[i];
            if (fabsf(s3d_plane_eval(m.x, m.y, m.z, m.w, p.x, p.y, p.z)) < tau) {
                const double x = p.x, y = p.y, z = p.z;
                s[0] += s3d_fx_bits(x, 1.0, M); s[1] += s3d_fx_bits(y, 1.0, M); s[2] += s3d_fx_bits(z, 1.0, M);
                s[3] += s3d_fx_bits(x, x, M); s[4] += s3d_fx_bits(x, y, M); s[5] += s3d_fx_bits(x, z, M);
                s[6] += s3d_fx_bits(y, y, M); s[7] += s3d_fx_bits(y, z, M); s[8] += s3d_fx_bits(z, z, M);
                ++cnt;
            }
        }
        total_cnt += cnt;
        #pragma unroll
        for (int k = 0; k < 9; ++k) {
            long long v = s[k] - (long long)((unsigned long long)cnt * fx.mbits);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { long long hi, lo; s3d_fx_split(v, hi, lo); whi[warp][k] += hi; wlo[warp][k] += lo; }
        }
    }
    total_cnt = warp_sum_i(total_cnt);
    if (lane == 0) wlo[warp][9] = total_cnt;
    __syncthreads();
    if (threadIdx.x < 10) {
        long long hi = 0, lo = 0;
        for (int w = 0; w < PLANE_BLOCK / 32; ++w) { hi += whi[w][threadIdx.x]; lo += wlo[w][threadIdx.x]; }
        __stcg(&partials[(size_t)blockIdx.x * 20 + threadIdx.x], hi);
        __stcg(&partials[(size_t)blockIdx.x * 20 + 10 + threadIdx.x], lo);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->ticket_refit, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // total over the CTAs' rows: integers, any order (25 strided parts with their loads in flight together, then the parts)
    __shared__ long long phi[25][10], plo[25][10];
    __shared__ double tot[10];
    if (threadIdx.x < 250) {
        const int slot = threadIdx.x % 10, part = threadIdx.x / 10;
        long long hi = 0, lo = 0;
        for (int b = part; b < (int)gridDim.x; b += 25) { hi += __ldcg(&partials[(size_t)b * 20 + slot]); lo += __ldcg(&partials[(size_t)b * 20 + 10 + slot]); }
        phi[part][slot] = hi; plo[part][slot] = lo;
    }
    __syncthreads();
    if (threadIdx.x < 10) {
        long long hi = 0, lo = 0;
        for (int q = 0; q < 25; ++q) { hi += phi[q][threadIdx.x]; lo += plo[q][threadIdx.x]; }
        tot[threadIdx.x] = threadIdx.x < 9 ? s3d_fx_value(hi, lo, fx.scale) : __ll2double_rn(lo);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double ni = tot[9];
        float4 rc = m;
        if (ni >= 3.0) {
            // strict double: the operations of oracle/plane_oracle.c, in its order
            const sd nd(ni);
            const sd cx = sd(tot[0]) / nd, cy = sd(tot[1]) / nd, cz = sd(tot[2]) / nd;
            sd C[3][3], V[3][3], w[3];
            C[0][0] = sd(tot[3]) / nd - cx * cx; C[0][1] = C[1][0] = sd(tot[4]) / nd - cx * cy; C[0][2] = C[2][0] = sd(tot[5]) / nd - cx * cz;
            C[1][1] = sd(tot[6]) / nd - cy * cy; C[1][2] = C[2][1] = sd(tot[7]) / nd - cy * cz; C[2][2] = sd(tot[8]) / nd - cz * cz;
            s3d_jacobi3(C, V, w);
            int k = 0; if (w[1] < w[k]) k = 1; if (w[2] < w[k]) k = 2;
            sd nx = V[0][k], ny = V[1][k], nz = V[2][k];
            const sd nn = sd_sqrt(nx * nx + ny * ny + nz * nz);
            nx = nx / nn; ny = ny / nn; nz = nz / nn;
            sd d = -(nx * cx + ny * cy + nz * cz);
            if (d.v < 0) { nx = -nx; ny = -ny; nz = -nz; d = -d; }   // reference src/GraphicEnd.cpp:383-387
            rc = make_float4((float)nx.v, (float)ny.v, (float)nz.v, (float)d.v);
        } else if (rc.w < 0.f) rc = make_float4(-rc.x, -rc.y, -rc.z, -rc.w);
        st->refined = rc;
        st->ticket_refit = 0u;
    }
}

// ---- order-preserving removal of the inliers of the refined model (ExtractIndices negative, reference :419-420) ----
__device__ __forceinline__ bool plane_keep(const float4 m, const float4 p, float tau)
{
    return !(fabsf(s3d_plane_eval(m.x, m.y, m.z, m.w, p.x, p.y, p.z)) < tau);
}

__global__ void __launch_bounds__(S3D_COMPACT_BLOCK) plane_count_kernel(const float4 *__restrict__ rem, float tau, const PlaneState *st,
                                                                        uint32_t *__restrict__ block_counts)
{
    if (!st->active) return;
    const int n_rem = st->n_rem;
    if ((int)(blockIdx.x * S3D_COMPACT_BLOCK) >= n_rem) return;
    __shared__ int warp_cnt[32];
    const float4 m = st->refined;
    const int i = blockIdx.x * S3D_COMPACT_BLOCK + threadIdx.x;
    const bool keep = (i < n_rem) && plane_keep(m, rem[i], tau);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_sum_i(warp_cnt[threadIdx.x]);
        if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)v;
    }
}

// Every block adds up the counts of the blocks before it (a few hundred values) instead of waiting for a scan kernel; kept
// points go to rem_out in order, inliers get their label and the plane normal; the last block to finish closes the round.
__global__ void __launch_bounds__(S3D_COMPACT_BLOCK) plane_write_kernel(const float4 *__restrict__ rem, float4 *__restrict__ rem_out, float tau,
                                                                        const uint32_t *__restrict__ block_counts, int32_t *__restrict__ labels,
                                                                        float4 *__restrict__ nrm, PlaneState *st)
{
    if (!st->active) return;
    const int n_rem = st->n_rem, plane_id = st->n_planes;
    if ((int)(blockIdx.x * S3D_COMPACT_BLOCK) >= n_rem) return;
    const int nb = (n_rem + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
    __shared__ int warp_cnt[32];
    __shared__ uint32_t warp_pre[32];
    __shared__ uint32_t block_off, total_kept;
    __shared__ bool is_last;
    const float4 m = st->refined;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // offset of this block = sum of the counts of blocks [0, blockIdx.x); total = sum over all nb blocks
    {
        uint32_t before = 0u, all = 0u;
        for (int k = threadIdx.x; k < nb; k += S3D_COMPACT_BLOCK) { const uint32_t v = block_counts[k]; all += v; if (k < (int)blockIdx.x) before += v; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); all += __shfl_xor_sync(0xffffffffu, all, o); }
        if (lane == 0) { warp_pre[w] = before; warp_cnt[w] = (int)all; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t b = 0u, a = 0u;
            for (int k = 0; k < 32; ++k) { b += warp_pre[k]; a += (uint32_t)warp_cnt[k]; }
            block_off = b; total_kept = a;
        }
        __syncthreads();
    }
    const int i = blockIdx.x * S3D_COMPACT_BLOCK + threadIdx.x;
    const bool in = i < n_rem;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) p = rem[i];
    const bool keep = in && plane_keep(m, p, tau);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_cnt[threadIdx.x], incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (threadIdx.x >= o) incl += t;
        }
        warp_cnt[threadIdx.x] = incl - v;
    }
    __syncthreads();
    if (keep) rem_out[block_off + (uint32_t)warp_cnt[w] + (uint32_t)__popc(b & ((1u << lane) - 1u))] = p;
    else if (in) {
        const int oi = __float_as_int(p.w);
        labels[oi] = plane_id;
        nrm[oi] = make_float4(m.x, m.y, m.z, 1.0f);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->ticket_write, 1u) == (unsigned)nb - 1u);
    __syncthreads();
    if (!is_last || threadIdx.x != 0) return;
    // close the round (reference :376-429): nothing removed -> the loop ends without a new plane
    st->ticket_write = 0u;
    const int nin = n_rem - (int)total_kept;
    if (nin == 0) { st->stopped = 1; return; }
    s3d_plane &pl = st->planes[plane_id];
    pl.coef[0] = m.x; pl.coef[1] = m.y; pl.coef[2] = m.z; pl.coef[3] = m.w;
    pl.inliers = nin; pl.hypotheses = st->iterations;
    st->n_rem = (int)total_kept;
    st->n_planes = plane_id + 1;
}

extern "C" void s3d_plane_params_default(s3d_plane_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->distance_threshold = 0.08f; p->plane_percent = 0.2f; p->max_planes = 3; p->max_iterations = 50;
    p->probability = 0.99f; p->seed = 12345ull;
}

// Enqueues the whole extraction on the ctx stream followed by the copy of the device loop's final state to `hb` (page-locked);
// no host synchronisation.  eval_passes_out: evaluation passes per round (timing bookkeeping of the blocking call).
static int planes_issue(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *prm, PlaneState *hb, int *eval_passes_out, bool *timed_out)
{
    if (prm->max_planes < 0 || prm->max_planes > S3D_MAX_PLANES || prm->max_iterations < 1 ||
        prm->max_iterations + PLANE_CANDIDATES_EXTRA > PLANE_MAX_CAND || !(prm->distance_threshold > 0.f))
        return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes: parameter out of range");
    cudaSetDevice(ctx->device);
    { int rc = s3d_cloud_ready(ctx, cloud); if (rc) return rc; }
    const int n = cloud->n;
    const size_t np = (size_t)(n > 0 ? n : 1);
    if (!cloud->d_nrm) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &cloud->d_nrm, sizeof(float4) * np));
    if (!cloud->d_labels) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &cloud->d_labels, sizeof(int32_t) * np));
    cloud->grid.valid = false;
    cloud->absmax_nrm_valid = false;
    { int rc = s3d_cloud_absmax(ctx, cloud, false); if (rc) return rc; }

    const int n_cand = prm->max_iterations + PLANE_CANDIDATES_EXTRA;
    const int nblk_c = (int)((np + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK);
    const int g_wide = ctx->sm_count * 4;
    // scratch carve-up
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_remA = carve(sizeof(float4) * np), o_remB = carve(sizeof(float4) * np);
    size_t o_coef = carve(sizeof(float4) * PLANE_MAX_CAND), o_valid = carve(sizeof(int) * PLANE_MAX_CAND);
    size_t o_cnt = carve(sizeof(uint32_t) * PLANE_MAX_CAND), o_state = carve(sizeof(PlaneState));
    size_t o_part = carve(sizeof(long long) * 20 * (size_t)g_wide), o_blk = carve(sizeof(uint32_t) * (size_t)(nblk_c + 2));
    if (off > ctx->cap_seg) {
        cudaFree(ctx->d_seg); ctx->d_seg = nullptr; ctx->cap_seg = 0;
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_seg, off));
        ctx->cap_seg = off;
    }
    char *base = (char *)ctx->d_seg;
    float4 *rem = (float4 *)(base + o_remA), *rem2 = (float4 *)(base + o_remB);
    float4 *coefs = (float4 *)(base + o_coef); int *valid = (int *)(base + o_valid);
    uint32_t *counts = (uint32_t *)(base + o_cnt); PlaneState *state = (PlaneState *)(base + o_state);
    long long *partials = (long long *)(base + o_part); uint32_t *blk = (uint32_t *)(base + o_blk);
    cudaStream_t st = ctx->stream;
    const float tau = prm->distance_threshold;

    if (!ctx->ev_plane[0]) {
        for (int k = 0; k < 2; ++k) S3D_CUDA(ctx, cudaEventCreate(&ctx->ev_plane[k]));
        for (int k = 0; k < 2 * S3D_MAX_PLANES; ++k) S3D_CUDA(ctx, cudaEventCreate(&ctx->ev_eval[k]));
    }
    // worst-case grids (the first round scans all n points); kernels of rounds the device loop has left return at once
    const dim3 ge(std::max(1, std::min(ctx->sm_count * 4, (n + PLANE_BLOCK - 1) / PLANE_BLOCK)), (n_cand + PLANE_CHUNK - 1) / PLANE_CHUNK);
    const int nb = std::max(1, (n + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK);
    const bool timed = (prm->reserved & 1) != 0;         // time every evaluation pass with its own events (s3d_last_plane_timing.eval_ms)
    // the whole extraction: 1 + 5 * max_planes launches whose arguments are all known here
    auto launches = [&](bool with_events) -> int {
        float4 *ra = rem, *rb = rem2;
        plane_init_kernel<<<g_wide, 256, 0, st>>>(cloud->d_pts, n, ra, cloud->d_labels, cloud->d_nrm, state);
        S3D_LAUNCHED(ctx);
        for (int round = 0; round < prm->max_planes; ++round) {
            plane_hyp_kernel<<<1, PLANE_MAX_CAND, 0, st>>>(ra, n, prm->plane_percent, prm->max_planes, prm->seed, round, n_cand, coefs, valid, counts, state);
            S3D_LAUNCHED(ctx);
            if (with_events) cudaEventRecord(ctx->ev_eval[2 * round], st);
            plane_eval_kernel<<<ge, PLANE_BLOCK, 0, st>>>(ra, coefs, valid, n_cand, tau, counts, prm->max_iterations, (double)prm->probability, state);
            S3D_LAUNCHED(ctx);
            if (with_events) cudaEventRecord(ctx->ev_eval[2 * round + 1], st);
            plane_refit_kernel<<<ctx->sm_count, PLANE_BLOCK, 0, st>>>(ra, tau, cloud->d_absmax, state, partials);   // one CTA per SM: the kernel is mostly its reduction
            S3D_LAUNCHED(ctx);
            plane_count_kernel<<<nb, S3D_COMPACT_BLOCK, 0, st>>>(ra, tau, state, blk);
            S3D_LAUNCHED(ctx);
            plane_write_kernel<<<nb, S3D_COMPACT_BLOCK, 0, st>>>(ra, rb, tau, blk, cloud->d_labels, cloud->d_nrm, state);
            S3D_LAUNCHED(ctx);
            std::swap(ra, rb);
        }
        return S3D_OK;
    };
    cudaEventRecord(ctx->ev_plane[0], st);
    bool done = false;
    static const bool use_graph = []() { const char *e = getenv("S3D_PLANE_GRAPH"); return !e || atoi(e) != 0; }();
    if (use_graph && !timed) {
        // The launches are bound by the host's launch rate (16 small kernels): the sequence is captured once into a CUDA graph keyed
        // by every buffer and parameter it touches (the ctx pool hands the next frame's cloud the same buffers) and replayed.
        const void *ptrs[] = {cloud->d_pts, cloud->d_labels, cloud->d_nrm, cloud->d_absmax, ctx->d_seg, (const void *)st, (const void *)"planes"};
        uint64_t key = 0xcbf29ce484222325ull;
        auto mix = [&](const void *p_, size_t nb_) { const unsigned char *b = (const unsigned char *)p_; for (size_t i = 0; i < nb_; ++i) { key ^= b[i]; key *= 0x100000001b3ull; } };
        mix(ptrs, sizeof(ptrs)); mix(&n, sizeof(n)); mix(prm, sizeof(*prm));
        auto it = ctx->graphs.find(key);
        if (it == ctx->graphs.end()) {
            if (ctx->graphs.size() > 256) { for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second); ctx->graphs.clear(); }
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            const int64_t l0 = ctx->launches;
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int rc = launches(false);
                const cudaError_t e = cudaStreamEndCapture(st, &graph);
                ctx->launches = l0;                              // captured, not launched
                if (rc == S3D_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    it = ctx->graphs.emplace(key, exec).first;
                    ctx->graph_nodes[key] = 1 + 5 * prm->max_planes;
                }
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
            }
        }
        if (it != ctx->graphs.end() && cudaGraphLaunch(it->second, st) == cudaSuccess) {
            ctx->launches += ctx->graph_nodes[key];
            done = true;
        }
    }
    if (!done) { const int rc = launches(timed); if (rc) return rc; }
    cudaEventRecord(ctx->ev_plane[1], st);
    S3D_CUDA(ctx, cudaMemcpyAsync(hb, state, sizeof(PlaneState), cudaMemcpyDeviceToHost, st));
    *eval_passes_out = (int)ge.y; *timed_out = timed;
    return S3D_OK;
}

extern "C" int s3d_segment_planes(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *prm, s3d_plane *planes_out, int *n_planes_out)
{
    if (!ctx || !cloud || !prm || !planes_out || !n_planes_out) return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes: bad argument");
    *n_planes_out = 0;
    PlaneState *hb = (PlaneState *)s3d_pinned(ctx, sizeof(PlaneState));
    if (!hb) return s3d_fail(ctx, S3D_E_CUDA, "pinned alloc");
    int eval_passes = 1; bool timed = false;
    { const int rc = planes_issue(ctx, cloud, prm, hb, &eval_passes, &timed); if (rc) return rc; }
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n_planes = hb->n_planes;
    for (int k = 0; k < n_planes; ++k) planes_out[k] = hb->planes[k];
    *n_planes_out = n_planes;
    // bookkeeping for s3d_last_plane_timing: points scanned by the evaluation passes
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev_plane[0], ctx->ev_plane[1]);
    ctx->plane_timing.total_ms = ms;
    ctx->plane_timing.rounds = 0; ctx->plane_timing.points_scanned = 0;
    for (int k = 0; k < S3D_MAX_PLANES; ++k) if (hb->rem_at_round[k] > 0) { ctx->plane_timing.rounds++; ctx->plane_timing.points_scanned += hb->rem_at_round[k]; }
    ctx->plane_timing.eval_passes_per_round = eval_passes;
    ctx->plane_timing.eval_ms = 0.f;
    for (int k = 0; timed && k < ctx->plane_timing.rounds && k < prm->max_planes; ++k) {
        float e = 0.f;
        cudaEventElapsedTime(&e, ctx->ev_eval[2 * k], ctx->ev_eval[2 * k + 1]);
        ctx->plane_timing.eval_ms += e;
    }
    return S3D_OK;
}

// ---- extraction without a host round trip (the stream counterpart of s3d_register_enqueue) -----------------------------------
extern "C" int s3d_segment_planes_enqueue(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *prm)
{
    if (!ctx || !cloud || !prm) return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes_enqueue: bad argument");
    if (ctx->planes_async_n >= S3D_ASYNC_DEPTH) return s3d_fail(ctx, S3D_E_STATE, "s3d_segment_planes_enqueue: S3D_ASYNC_DEPTH extractions outstanding, call s3d_segment_planes_drain first");
    cudaSetDevice(ctx->device);
    if (!ctx->h_planes_ring) S3D_CUDA(ctx, cudaMallocHost(&ctx->h_planes_ring, sizeof(PlaneState) * S3D_ASYNC_DEPTH));
    int eval_passes = 1; bool timed = false;
    const int rc = planes_issue(ctx, cloud, prm, reinterpret_cast<PlaneState *>(ctx->h_planes_ring) + ctx->planes_async_n, &eval_passes, &timed);
    if (rc) return rc;
    ctx->planes_async_n++;
    return S3D_OK;
}

extern "C" int s3d_segment_planes_drain(s3d_ctx *ctx, s3d_plane *planes_out, int *n_planes_out, int capacity, int *n_out)
{
    if (!ctx || !planes_out || !n_planes_out || !n_out) return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes_drain: bad argument");
    const int n = ctx->planes_async_n;
    if (capacity < n) return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes_drain: capacity below the number of outstanding extractions");
    *n_out = 0;
    if (n == 0) return S3D_OK;
    cudaSetDevice(ctx->device);
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->planes_async_n = 0;
    const PlaneState *ring = reinterpret_cast<const PlaneState *>(ctx->h_planes_ring);
    for (int i = 0; i < n; ++i) {
        n_planes_out[i] = ring[i].n_planes;
        for (int k = 0; k < S3D_MAX_PLANES; ++k) planes_out[(size_t)i * S3D_MAX_PLANES + k] = ring[i].planes[k];
    }
    *n_out = n;
    return S3D_OK;
}

extern "C" int s3d_last_plane_timing(const s3d_ctx *ctx, s3d_plane_timing *out)
{
    if (!ctx || !out) return S3D_E_ARG;
    *out = ctx->plane_timing;
    return S3D_OK;
}
