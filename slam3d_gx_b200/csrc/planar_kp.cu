// planar_kp.cu -- batched keypoint planarity test, one warp per keypoint.
//
// Replaces isPlanar (reference src/planarFeatures.cpp:88-136): 7x7 depth patch around the keypoint,
// reject if any depth is zero (:103-107), back-project the 49 pixels in double (:108-111) into
// float points, RANSAC plane (pcl::RandomSampleConsensus defaults: 1000 iterations, p = 0.99,
// threshold :121), planar iff more than min_inliers inliers (:127).  Each lane evaluates one candidate
// plane of a 32-candidate chunk; the sequential adaptive-stop rule is replayed across the lanes, so
// the answer is the one the sequential loop of oracle_planar_keypoints gives.
#include "context.h"
#include "common.cuh"

#define KP_WARPS 8

__global__ void __launch_bounds__(KP_WARPS * 32) planar_kp_kernel(const uint16_t *__restrict__ depth, int width, int height,
                                                                   double fx, double fy, double cx, double cy, double factor,
                                                                   const int2 *__restrict__ uv, int n, float thr, int min_inliers,
                                                                   uint64_t seed, uint8_t *__restrict__ flags)
{
    __shared__ float4 P[KP_WARPS][52];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kp = blockIdx.x * KP_WARPS + warp;
    if (kp >= n) return;
    const int u = uv[kp].x, v = uv[kp].y;
    if (u < 3 || v < 3 || u + 3 >= width || v + 3 >= height) { if (lane == 0) flags[kp] = 0; return; }
    bool zero = false;
    for (int t = lane; t < 49; t += 32) {
        int j = t / 7, i = t - 7 * j;
        uint16_t dd = depth[(size_t)(v + j - 3) * width + (u + i - 3)];
        zero |= (dd == 0);
        double z = __ddiv_rn((double)dd, factor);
        double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)(u + (i - 3)), cx), z), fx);
        double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)(v + (j - 3)), cy), z), fy);
        P[warp][t] = make_float4((float)x, (float)y, (float)z, 0.f);
    }
    if (__any_sync(0xffffffffu, zero)) { if (lane == 0) flags[kp] = 0; return; }
    __syncwarp();
    int iterations = 0, best = -2147483647;
    double k = 1.0;
    const double log_prob = log(1.0 - 0.99), eps = 2.220446049250313e-16;
    bool done = false;
    for (int chunk = 0; chunk < 33 && !done; ++chunk) {
        const int c = chunk * 32 + lane;
        uint32_t s[3];
        s3d_sample3(seed, (uint64_t)kp, (uint64_t)c, 49u, s);
        float4 a = P[warp][s[0]], b = P[warp][s[1]], d = P[warp][s[2]], coef;
        bool ok = s3d_plane_from3(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), make_float3(d.x, d.y, d.z), coef);
        int cnt = 0;
        if (ok) {
            for (int t = 0; t < 49; ++t) {
                float4 p = P[warp][t];
                cnt += (fabsf(s3d_plane_eval(coef.x, coef.y, coef.z, coef.w, p.x, p.y, p.z)) < thr) ? 1 : 0;
            }
        }
        for (int l = 0; l < 32; ++l) {
            if (!((double)iterations < k)) { done = true; break; }
            int vl = __shfl_sync(0xffffffffu, ok ? 1 : 0, l), cl = __shfl_sync(0xffffffffu, cnt, l);
            if (!vl) continue;
            if (cl > best) {
                best = cl;
                double w = cl / 49.0, p_no = 1.0 - w * w * w;
                if (p_no < eps) p_no = eps;
                if (p_no > 1.0 - eps) p_no = 1.0 - eps;
                k = log_prob / log(p_no);
            }
            ++iterations;
            if (iterations > 1000) { done = true; break; }
        }
    }
    if (lane == 0) flags[kp] = best > min_inliers ? 1 : 0;
}

extern "C" int s3d_planar_keypoints(s3d_ctx *ctx, const uint16_t *depth, int width, int height, const s3d_camera *cam,
                                    const int32_t *uv, int n, float threshold, int min_inliers, uint64_t seed, uint8_t *flags_out)
{
    if (!ctx || !depth || !cam || width <= 0 || height <= 0 || n < 0 || (n > 0 && (!uv || !flags_out)))
        return s3d_fail(ctx, S3D_E_ARG, "s3d_planar_keypoints: bad argument");
    if (n == 0) return S3D_OK;
    cudaSetDevice(ctx->device);
    size_t npx = (size_t)width * height;
    uint16_t *d_depth = nullptr; int2 *d_uv = nullptr; uint8_t *d_flags = nullptr;
    S3D_CUDA(ctx, cudaMalloc(&d_depth, sizeof(uint16_t) * npx));
    S3D_CUDA(ctx, cudaMalloc(&d_uv, sizeof(int2) * (size_t)n));
    S3D_CUDA(ctx, cudaMalloc(&d_flags, (size_t)n));
    S3D_CUDA(ctx, cudaMemcpyAsync(d_depth, depth, sizeof(uint16_t) * npx, cudaMemcpyHostToDevice, ctx->stream));
    S3D_CUDA(ctx, cudaMemcpyAsync(d_uv, uv, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    planar_kp_kernel<<<(n + KP_WARPS - 1) / KP_WARPS, KP_WARPS * 32, 0, ctx->stream>>>(d_depth, width, height, cam->fx, cam->fy, cam->cx,
                                                                                      cam->cy, cam->factor, d_uv, n, threshold,
                                                                                      min_inliers, seed, d_flags);
    S3D_LAUNCHED(ctx);
    S3D_CUDA(ctx, cudaMemcpyAsync(flags_out, d_flags, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_depth); cudaFree(d_uv); cudaFree(d_flags);
    return S3D_OK;
}
