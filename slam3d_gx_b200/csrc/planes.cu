// planes.cu -- RANSAC plane extraction on the device.
//
// Replaces GraphicEnd::extractPlanesAndGenerateImage (reference src/GraphicEnd.cpp:353-430), i.e. the
// loop  { pcl::SACSegmentation::segment ; flip sign so d >= 0 ; ExtractIndices positive/negative }
// with the PCL-1.7 semantics restated in oracle/plane_oracle.c.  All candidate planes of a round are
// evaluated in ONE streaming pass over the remaining points (16 B per point per 32 candidates); the
// sequential adaptive-stop logic of pcl::RandomSampleConsensus::computeModel is then replayed over
// the counts by a single thread, so the chosen model is exactly the one the sequential loop picks.
// The per-point result (plane label + plane normal) is what the ICP uses as target normals.
#include <cstring>
#include "context.h"
#include "common.cuh"
#include "compact.cuh"

#define PLANE_CANDIDATES_EXTRA 14
#define PLANE_MAX_CAND 1024
#define PLANE_CHUNK 32
#define PLANE_BLOCK 256

struct PlaneSel {
    float4 ransac;     // model picked by the RANSAC replay
    float4 refined;    // after PCA refit + sign flip
    int best;          // candidate index or -1
    int best_count;
    int iterations;
    int stop;          // 1: no plane found in this round
    unsigned ticket;
    int pad[3];
};

__global__ void plane_init_kernel(const float4 *__restrict__ pts, int n, float4 *__restrict__ rem, int32_t *__restrict__ labels,
                                  float4 *__restrict__ nrm)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        rem[i] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        labels[i] = -1;
        nrm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void plane_hyp_kernel(const float4 *__restrict__ rem, int n_rem, uint64_t seed, int round, int n_cand,
                                 float4 *__restrict__ coefs, int *__restrict__ valid, uint32_t *__restrict__ counts, PlaneSel *sel)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) { sel->ticket = 0u; sel->stop = 0; sel->best = -1; }
    if (c >= n_cand) return;
    uint32_t s[3];
    s3d_sample3(seed, (uint64_t)round, (uint64_t)c, (uint32_t)n_rem, s);
    float4 a = rem[s[0]], b = rem[s[1]], d = rem[s[2]];
    float4 coef = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = s3d_plane_from3(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(d.x, d.y, d.z), coef);
    coefs[c] = coef; valid[c] = ok ? 1 : 0; counts[c] = 0u;
}

// one pass over the remaining points evaluates PLANE_CHUNK candidates (chunk index = blockIdx.y)
__global__ void __launch_bounds__(PLANE_BLOCK) plane_eval_kernel(const float4 *__restrict__ rem, int n_rem, const float4 *__restrict__ coefs,
                                                                 const int *__restrict__ valid, int n_cand, float tau,
                                                                 uint32_t *__restrict__ counts)
{
    __shared__ float4 sc[PLANE_CHUNK];
    const int c0 = blockIdx.y * PLANE_CHUNK;
    if (threadIdx.x < PLANE_CHUNK) {
        int c = c0 + threadIdx.x;
        // invalid candidates get a plane no point can satisfy (NaN compares false)
        sc[threadIdx.x] = (c < n_cand && valid[c]) ? coefs[c] : make_float4(0.f, 0.f, 0.f, __int_as_float(0x7fc00000));
    }
    __syncthreads();
    int cnt[PLANE_CHUNK];
    #pragma unroll
    for (int k = 0; k < PLANE_CHUNK; ++k) cnt[k] = 0;
    for (int i = blockIdx.x * PLANE_BLOCK + threadIdx.x; i < n_rem; i += gridDim.x * PLANE_BLOCK) {
        const float4 p = rem[i];
        #pragma unroll
        for (int k = 0; k < PLANE_CHUNK; ++k) {
            const float4 c = sc[k];
            cnt[k] += (fabsf(s3d_plane_eval(c.x, c.y, c.z, c.w, p.x, p.y, p.z)) < tau) ? 1 : 0;
        }
    }
    // counts: warp shuffle -> shared-memory atomics -> ONE global atomic per candidate and CTA (every warp going to the 64
    // global counters serialised at ~27 cycles per same-address atomic: 65 us of a 88 us kernel)
    __shared__ uint32_t s_cnt[PLANE_CHUNK];
    if (threadIdx.x < PLANE_CHUNK) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < PLANE_CHUNK; ++k) {
        int v = warp_sum_i(cnt[k]);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], (uint32_t)v);
    }
    __syncthreads();
    if (threadIdx.x < PLANE_CHUNK && s_cnt[threadIdx.x] && c0 + threadIdx.x < n_cand) atomicAdd(&counts[c0 + threadIdx.x], s_cnt[threadIdx.x]);
}

// replay of pcl::RandomSampleConsensus::computeModel over the pre-evaluated candidates
__global__ void plane_select_kernel(const float4 *__restrict__ coefs, const int *__restrict__ valid, const uint32_t *__restrict__ counts,
                                    int n_cand, int n_points, int max_iterations, double probability, PlaneSel *sel)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int iterations = 0, best = -1, best_count = -2147483647;
    double k = 1.0;
    const double log_prob = log(1.0 - probability);
    const double one_over = n_points > 0 ? 1.0 / (double)n_points : 0.0;
    const double eps = 2.220446049250313e-16;
    for (int c = 0; c < n_cand && iterations < k; ++c) {
        if (!valid[c]) continue;
        int cc = (int)counts[c];
        if (cc > best_count) {
            best_count = cc; best = c;
            double w = best_count * one_over;
            double p_no = 1.0 - w * w * w;
            if (p_no < eps) p_no = eps;
            if (p_no > 1.0 - eps) p_no = 1.0 - eps;
            k = log_prob / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    sel->best = best; sel->best_count = best >= 0 ? best_count : 0; sel->iterations = iterations;
    sel->stop = (best < 0 || best_count == 0) ? 1 : 0;
    if (best >= 0) { sel->ransac = coefs[best]; sel->refined = coefs[best]; }
}

// PCA refit over the inliers of the RANSAC model (SampleConsensusModelPlane::optimizeModelCoefficients)
__global__ void __launch_bounds__(PLANE_BLOCK) plane_refit_kernel(const float4 *__restrict__ rem, int n_rem, float tau, PlaneSel *sel,
                                                                  double *__restrict__ partials)
{
    __shared__ double ws[PLANE_BLOCK / 32][10];
    __shared__ bool is_last;
    if (sel->stop) return;
    const float4 m = sel->ransac;
    double s[10];
    #pragma unroll
    for (int k = 0; k < 10; ++k) s[k] = 0.0;
    for (int i = blockIdx.x * PLANE_BLOCK + threadIdx.x; i < n_rem; i += gridDim.x * PLANE_BLOCK) {
        const float4 p = rem[i];
        if (fabsf(s3d_plane_eval(m.x, m.y, m.z, m.w, p.x, p.y, p.z)) < tau) {
            double x = p.x, y = p.y, z = p.z;
            s[0] += x; s[1] += y; s[2] += z;
            s[3] += x * x; s[4] += x * y; s[5] += x * z; s[6] += y * y; s[7] += y * z; s[8] += z * z;
            s[9] += 1.0;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int k = 0; k < 10; ++k) { double v = warp_sum_d(s[k]); if (lane == 0) ws[warp][k] = v; }
    __syncthreads();
    if (threadIdx.x < 10) {
        double v = 0.0;
        for (int w = 0; w < PLANE_BLOCK / 32; ++w) v += ws[w][threadIdx.x];
        __stcg(&partials[(size_t)blockIdx.x * 10 + threadIdx.x], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&sel->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // final sum over the CTAs' rows in a fixed two-level order: 25 strided parts (250 threads, loads in flight together), then the parts
    {
        __shared__ double parts[25][10];
        if (threadIdx.x < 250) {
            const int slot = threadIdx.x % 10, part = threadIdx.x / 10;
            double v = 0.0;
            for (int b = part; b < (int)gridDim.x; b += 25) v += __ldcg(&partials[(size_t)b * 10 + slot]);
            parts[part][slot] = v;
        }
        __syncthreads();
        if (threadIdx.x < 10) {
            double v = 0.0;
            for (int q = 0; q < 25; ++q) v += parts[q][threadIdx.x];
            ws[0][threadIdx.x] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double *t = ws[0];
        double ni = t[9];
        float4 rc = m;
        if (ni >= 3.0) {
            double cx = t[0] / ni, cy = t[1] / ni, cz = t[2] / ni;
            double C[3][3], V[3][3], w[3];
            C[0][0] = t[3] / ni - cx * cx; C[0][1] = C[1][0] = t[4] / ni - cx * cy; C[0][2] = C[2][0] = t[5] / ni - cx * cz;
            C[1][1] = t[6] / ni - cy * cy; C[1][2] = C[2][1] = t[7] / ni - cy * cz; C[2][2] = t[8] / ni - cz * cz;
            s3d_jacobi3(C, V, w);
            int k = 0; if (w[1] < w[k]) k = 1; if (w[2] < w[k]) k = 2;
            double nx = V[0][k], ny = V[1][k], nz = V[2][k];
            double nn = sqrt(nx * nx + ny * ny + nz * nz);
            nx /= nn; ny /= nn; nz /= nn;
            double d = -(nx * cx + ny * cy + nz * cz);
            if (d < 0) { nx = -nx; ny = -ny; nz = -nz; d = -d; }   // reference src/GraphicEnd.cpp:383-387
            rc = make_float4((float)nx, (float)ny, (float)nz, (float)d);
        } else if (rc.w < 0.f) rc = make_float4(-rc.x, -rc.y, -rc.z, -rc.w);
        sel->refined = rc;
        sel->ticket = 0u;
    }
}

struct KeepPred {   // keep = NOT an inlier of the refined model (ExtractIndices negative)
    const float4 *rem; const PlaneSel *sel; float tau;
    __device__ bool operator()(int i) const
    {
        if (sel->stop) return true;
        const float4 m = sel->refined, p = rem[i];
        return !(fabsf(s3d_plane_eval(m.x, m.y, m.z, m.w, p.x, p.y, p.z)) < tau);
    }
};
struct KeepEmit {
    const float4 *rem; float4 *out;
    __device__ void operator()(int i, uint32_t pos) const { out[pos] = rem[i]; }
};
struct InlierDrop {   // inliers leave the cloud: record their plane id and the plane normal
    const float4 *rem; const PlaneSel *sel; int32_t *labels; float4 *nrm; int plane_id;
    __device__ void operator()(int i) const
    {
        const float4 m = sel->refined;
        int oi = __float_as_int(rem[i].w);
        labels[oi] = plane_id;
        nrm[oi] = make_float4(m.x, m.y, m.z, 1.0f);
    }
};

extern "C" void s3d_plane_params_default(s3d_plane_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->distance_threshold = 0.08f; p->plane_percent = 0.2f; p->max_planes = 3; p->max_iterations = 50;
    p->probability = 0.99f; p->seed = 12345ull;
}

extern "C" int s3d_segment_planes(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *prm, s3d_plane *planes_out, int *n_planes_out)
{
    if (!ctx || !cloud || !prm || !planes_out || !n_planes_out) return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes: bad argument");
    if (prm->max_planes < 0 || prm->max_planes > S3D_MAX_PLANES || prm->max_iterations < 1 ||
        prm->max_iterations + PLANE_CANDIDATES_EXTRA > PLANE_MAX_CAND || !(prm->distance_threshold > 0.f))
        return s3d_fail(ctx, S3D_E_ARG, "s3d_segment_planes: parameter out of range");
    cudaSetDevice(ctx->device);
    *n_planes_out = 0;
    { int rc = s3d_cloud_ready(ctx, cloud); if (rc) return rc; }
    const int n = cloud->n;
    const size_t np = (size_t)(n > 0 ? n : 1);
    if (!cloud->d_nrm) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &cloud->d_nrm, sizeof(float4) * np));
    if (!cloud->d_labels) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &cloud->d_labels, sizeof(int32_t) * np));
    cloud->grid.valid = false;

    const int n_cand = prm->max_iterations + PLANE_CANDIDATES_EXTRA;
    const int nblk_c = (int)((np + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK);
    const int g_wide = ctx->sm_count * 4;
    // scratch carve-up
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_remA = carve(sizeof(float4) * np), o_remB = carve(sizeof(float4) * np);
    size_t o_coef = carve(sizeof(float4) * PLANE_MAX_CAND), o_valid = carve(sizeof(int) * PLANE_MAX_CAND);
    size_t o_cnt = carve(sizeof(uint32_t) * PLANE_MAX_CAND), o_sel = carve(sizeof(PlaneSel));
    size_t o_part = carve(sizeof(double) * 10 * (size_t)g_wide), o_blk = carve(sizeof(uint32_t) * (size_t)(nblk_c + 2));
    if (off > ctx->cap_seg) {
        cudaFree(ctx->d_seg); ctx->d_seg = nullptr; ctx->cap_seg = 0;
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_seg, off));
        ctx->cap_seg = off;
    }
    char *base = (char *)ctx->d_seg;
    float4 *rem = (float4 *)(base + o_remA), *rem2 = (float4 *)(base + o_remB);
    float4 *coefs = (float4 *)(base + o_coef); int *valid = (int *)(base + o_valid);
    uint32_t *counts = (uint32_t *)(base + o_cnt); PlaneSel *sel = (PlaneSel *)(base + o_sel);
    double *partials = (double *)(base + o_part); uint32_t *blk = (uint32_t *)(base + o_blk);
    struct HostBack { PlaneSel sel; uint32_t kept; } *hb = (HostBack *)s3d_pinned(ctx, sizeof(HostBack));
    if (!hb) return s3d_fail(ctx, S3D_E_CUDA, "pinned alloc");
    cudaStream_t st = ctx->stream;

    plane_init_kernel<<<g_wide, 256, 0, st>>>(cloud->d_pts, n, rem, cloud->d_labels, cloud->d_nrm);
    S3D_LAUNCHED(ctx);
    int n_rem = n, n_planes = 0;
    while ((double)n_rem > (double)prm->plane_percent * (double)n && n_planes < prm->max_planes) {   // :372, :424
        if (n_rem < 3) break;
        plane_hyp_kernel<<<(n_cand + 127) / 128, 128, 0, st>>>(rem, n_rem, prm->seed, n_planes, n_cand, coefs, valid, counts, sel);
        S3D_LAUNCHED(ctx);
        dim3 ge(std::max(1, std::min(ctx->sm_count * 2, (n_rem + PLANE_BLOCK - 1) / PLANE_BLOCK)), (n_cand + PLANE_CHUNK - 1) / PLANE_CHUNK);
        plane_eval_kernel<<<ge, PLANE_BLOCK, 0, st>>>(rem, n_rem, coefs, valid, n_cand, prm->distance_threshold, counts);
        S3D_LAUNCHED(ctx);
        plane_select_kernel<<<1, 32, 0, st>>>(coefs, valid, counts, n_cand, n_rem, prm->max_iterations, (double)prm->probability, sel);
        S3D_LAUNCHED(ctx);
        plane_refit_kernel<<<ctx->sm_count, PLANE_BLOCK, 0, st>>>(rem, n_rem, prm->distance_threshold, sel, partials);   // one CTA per SM: the kernel is mostly its reduction
        S3D_LAUNCHED(ctx);
        const int nb = (n_rem + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
        KeepPred pred{rem, sel, prm->distance_threshold};
        compact_count_kernel<<<nb, S3D_COMPACT_BLOCK, 0, st>>>(n_rem, pred, blk);
        S3D_LAUNCHED(ctx);
        compact_scan_kernel<<<1, 1024, 0, st>>>(blk, nb, blk + nb);
        S3D_LAUNCHED(ctx);
        compact_write_kernel<<<nb, S3D_COMPACT_BLOCK, 0, st>>>(n_rem, pred, KeepEmit{rem, rem2},
                                                                InlierDrop{rem, sel, cloud->d_labels, cloud->d_nrm, n_planes}, blk);
        S3D_LAUNCHED(ctx);
        S3D_CUDA(ctx, cudaMemcpyAsync(&hb->sel, sel, sizeof(PlaneSel), cudaMemcpyDeviceToHost, st));
        S3D_CUDA(ctx, cudaMemcpyAsync(&hb->kept, blk + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        S3D_CUDA(ctx, cudaStreamSynchronize(st));
        if (hb->sel.stop) break;                                   // :376-379
        const int nin = n_rem - (int)hb->kept;
        if (nin == 0) break;
        s3d_plane &pl = planes_out[n_planes];
        pl.coef[0] = hb->sel.refined.x; pl.coef[1] = hb->sel.refined.y; pl.coef[2] = hb->sel.refined.z; pl.coef[3] = hb->sel.refined.w;
        pl.inliers = nin; pl.hypotheses = hb->sel.iterations;
        std::swap(rem, rem2);
        n_rem = (int)hb->kept; ++n_planes;
    }
    *n_planes_out = n_planes;
    return S3D_OK;
}
