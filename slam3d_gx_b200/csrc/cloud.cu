// cloud.cu -- context and device-resident clouds of libslam3d_b200.
//
// A cloud is stored as float4 (x,y,z,1) so that every point is one coalesced 16-byte load, plus an
// optional float4 normal (nx,ny,nz,valid) and an int32 plane label per point.  It replaces the
// pcl::PointCloud<PointXYZRGBA>::Ptr members of the reference front end (src/GraphicEnd.h:181-183).
#include <cstdio>
#include <cstring>
#include <cmath>
#include <iterator>
#include <algorithm>
#include "context.h"
#include "compact.cuh"

int s3d_fail(s3d_ctx *ctx, int code, const char *what, cudaError_t e)
{
    if (ctx) {
        ctx->err = what ? what : "";
        if (e != cudaSuccess) { ctx->err += ": "; ctx->err += cudaGetErrorString(e); }
    }
    return code;
}

void *s3d_pinned(s3d_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->cap_pinned) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr; ctx->cap_pinned = 0;
        size_t cap = bytes + bytes / 2 + 4096;
        if (cudaMallocHost(&ctx->h_pinned, cap) != cudaSuccess) { ctx->h_pinned = nullptr; return nullptr; }
        ctx->cap_pinned = cap;
    }
    return ctx->h_pinned;
}

// ---- caching device allocator --------------------------------------------------------------------
#define S3D_POOL_MAX_CACHED ((size_t)16 << 30)      // beyond this, freed blocks really go back to the driver

static size_t pool_round(size_t bytes)
{
    if (bytes < 4096) return 4096;
    size_t step = (size_t)1 << 12;
    while ((step << 4) <= bytes) step <<= 1;         // ~6-12 % granularity: blocks of similar clouds are interchangeable
    return (bytes + step - 1) / step * step;
}

cudaError_t s3d_dev_alloc(s3d_ctx *ctx, void **out, size_t bytes)
{
    const size_t want = pool_round(bytes);
    auto it = ctx->pool_free.lower_bound(want);
    if (it != ctx->pool_free.end() && it->first <= want + want / 4) {
        // last in, first out among blocks of one size: a caller that repeats the same sequence of uploads, registrations
        // and frees per frame (also with the next frame's upload in flight) then sees the same few sets of addresses
        // again and again, which is what lets the index build replay its captured graphs (grid.cu)
        it = std::prev(ctx->pool_free.upper_bound(it->first));
        *out = it->second;
        ctx->pool_live[it->second] = it->first;
        ctx->pool_live_bytes += it->first;
        if (ctx->pool_live_bytes > ctx->pool_peak_bytes) ctx->pool_peak_bytes = ctx->pool_live_bytes;
        ctx->pool_cached -= it->first;
        ctx->pool_free.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e != cudaSuccess && !ctx->pool_free.empty()) {   // out of memory with blocks cached: give them back and retry
        cudaGetLastError();
        s3d_dev_pool_release(ctx);
        e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) {
        ctx->pool_live[*out] = want;
        ctx->pool_live_bytes += want;
        if (ctx->pool_live_bytes > ctx->pool_peak_bytes) ctx->pool_peak_bytes = ctx->pool_live_bytes;
    }
    return e;
}

void s3d_dev_free(s3d_ctx *ctx, void *p)
{
    if (!p) return;
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) { cudaFree(p); return; }
    const size_t sz = it->second;
    ctx->pool_live.erase(it);
    ctx->pool_live_bytes -= sz;
    if (ctx->pool_cached + sz > S3D_POOL_MAX_CACHED) { cudaFree(p); return; }
    ctx->pool_free.insert({sz, p});
    ctx->pool_cached += sz;
}

void s3d_dev_pool_release(s3d_ctx *ctx)
{
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pool_free) cudaFree(kv.second);
    ctx->pool_free.clear();
    ctx->pool_cached = 0;
}

extern "C" int s3d_abi_version(void) { return S3D_ABI_VERSION; }

extern "C" int s3d_create(s3d_ctx **out, int device_id)
{
    if (!out) return S3D_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device_id < 0 || device_id >= ndev) {
        // No CPU fallback by design: the product path requires a CUDA device.
        fprintf(stderr, "slam3d_b200: no usable CUDA device %d (%s)\n", device_id,
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0 or id out of range");
        return S3D_E_CUDA;
    }
    s3d_ctx *ctx = new s3d_ctx();
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess) { delete ctx; return S3D_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) { delete ctx; return S3D_E_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return S3D_E_CUDA; }
    ctx->stream = ctx->own_stream;
    for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
    *out = ctx;
    return S3D_OK;
}

extern "C" void s3d_destroy(s3d_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second);
    cudaFree(ctx->d_desc); cudaFreeHost(ctx->h_desc);
    cudaFree(ctx->d_state); cudaFreeHost(ctx->h_state);
    cudaFree(ctx->d_partials); cudaFree(ctx->d_nn_idx); cudaFree(ctx->d_nn_d2); cudaFree(ctx->d_nn_pos); cudaFree(ctx->d_last_nn);
    cudaFree(ctx->d_cq); cudaFree(ctx->d_cn); cudaFree(ctx->d_flags); cudaFree(ctx->d_cq2); cudaFree(ctx->d_pend); cudaFree(ctx->d_barriers);
    cudaFree(ctx->d_seg);
    cudaFree(ctx->d_gather_send); cudaFree(ctx->d_gather_recv);
    if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
    if (ctx->h_desc_ring) cudaFreeHost(ctx->h_desc_ring);
    if (ctx->h_state_ring) cudaFreeHost(ctx->h_state_ring);
    if (ctx->h_async) cudaFreeHost(ctx->h_async);
    if (ctx->h_planes_ring) cudaFreeHost(ctx->h_planes_ring);
    cudaFree(ctx->d_async);
    for (int i = 0; i < S3D_ASYNC_DEPTH; ++i) for (int k = 0; k < 3; ++k) if (ctx->ev_ring[i][k]) cudaEventDestroy(ctx->ev_ring[i][k]);
    s3d_dev_pool_release(ctx);
    for (auto &kv : ctx->pool_live) cudaFree(kv.first);     // handles the caller never freed
    ctx->pool_live.clear();
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 2; ++i) if (ctx->ev_plane[i]) cudaEventDestroy(ctx->ev_plane[i]);
    for (int i = 0; i < 2 * S3D_MAX_PLANES; ++i) if (ctx->ev_eval[i]) cudaEventDestroy(ctx->ev_eval[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); cudaEventDestroy(ctx->copy_fence); }
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->aux_fork); cudaEventDestroy(ctx->aux_join); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *s3d_last_error(const s3d_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int s3d_set_stream(s3d_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return S3D_E_ARG;
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    // graphs were captured on the previous stream's topology only; they stay valid for launch on any stream
    return S3D_OK;
}

extern "C" int s3d_device_sm_count(const s3d_ctx *ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int64_t s3d_launch_count(const s3d_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------

static int cloud_alloc(s3d_ctx *ctx, int n, s3d_cloud **out)
{
    s3d_cloud *c = new s3d_cloud();
    c->n = n;
    cudaError_t e = s3d_dev_alloc_t(ctx, &c->d_pts, sizeof(float4) * (size_t)(n > 0 ? n : 1));
    if (e != cudaSuccess) { delete c; return s3d_fail(ctx, S3D_E_CUDA, "cudaMalloc cloud", e); }
    *out = c;
    return S3D_OK;
}

__global__ void pack_xyz_kernel(const float *in, int stride, int n, float4 *out) /* may run in place (stride 4) */
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + (size_t)i * stride;
    out[i] = make_float4(p[0], p[1], p[2], 1.0f);
}

__global__ void pack_normals_kernel(const float *__restrict__ in, int stride, int n, float4 *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *p = in + (size_t)i * stride;
    float x = p[0], y = p[1], z = p[2];
    bool ok = isfinite(x) && isfinite(y) && isfinite(z) && (x != 0.f || y != 0.f || z != 0.f);
    out[i] = ok ? make_float4(x, y, z, 1.0f) : make_float4(0.f, 0.f, 0.f, 0.f);
}

extern "C" int s3d_cloud_upload(s3d_ctx *ctx, const float *xyz, int stride_floats, int n, s3d_cloud **out)
{
    if (!ctx || !out || n < 0 || stride_floats < 3 || (n > 0 && !xyz)) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_upload: bad argument");
    cudaSetDevice(ctx->device);
    s3d_cloud *c = nullptr;
    int rc = cloud_alloc(ctx, n, &c);
    if (rc) return rc;
    float *tmp = nullptr;
    rc = [&]() -> int {       // every early return below leaves through the clean-up after the lambda
        if (n == 0) return S3D_OK;
        size_t bytes = sizeof(float) * (size_t)n * stride_floats;
        if (stride_floats == 4) {
            // PCD rows "x y z rgba" are already float4-shaped: copy straight, then normalise w on device
            S3D_CUDA(ctx, cudaMemcpyAsync(c->d_pts, xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
            pack_xyz_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const float *)c->d_pts, 4, n, c->d_pts);
            S3D_LAUNCHED(ctx);
        } else {
            S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &tmp, bytes));
            S3D_CUDA(ctx, cudaMemcpyAsync(tmp, xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
            pack_xyz_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(tmp, stride_floats, n, c->d_pts);
            S3D_LAUNCHED(ctx);
        }
        S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return S3D_OK;
    }();
    if (tmp) { cudaStreamSynchronize(ctx->stream); s3d_dev_free(ctx, tmp); }
    if (rc) { s3d_cloud_free(ctx, c); return rc; }
    *out = c;
    return S3D_OK;
}

// largest finite |x|, |y|, |z| over the rows (rows with w == 0 skipped when need_w): a maximum, hence independent of the
// order in which threads see the rows.  Non-negative floats order like their bit patterns.
__global__ void __launch_bounds__(256) absmax_kernel(const float4 *__restrict__ rows, int n, int need_w, float *__restrict__ out)
{
    float m = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = rows[i];
        if (need_w && p.w == 0.f) continue;
        const float a = fabsf(p.x), b = fabsf(p.y), c = fabsf(p.z);
        if (a <= 3.4028234663852886e38f && a > m) m = a;
        if (b <= 3.4028234663852886e38f && b > m) m = b;
        if (c <= 3.4028234663852886e38f && c > m) m = c;
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int *>(out), __float_as_uint(m));
}

int s3d_cloud_absmax(s3d_ctx *ctx, const s3d_cloud *cloud, bool want_normals)
{
    if (!cloud->d_absmax) {
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &cloud->d_absmax, 2 * sizeof(float)));
        cloud->absmax_pts_valid = cloud->absmax_nrm_valid = false;
    }
    const int blocks = std::max(1, std::min(ctx->sm_count * 4, (cloud->n + 255) / 256));
    if (!cloud->absmax_pts_valid) {
        S3D_CUDA(ctx, cudaMemsetAsync(cloud->d_absmax, 0, sizeof(float), ctx->stream));
        if (cloud->n > 0) { absmax_kernel<<<blocks, 256, 0, ctx->stream>>>(cloud->d_pts, cloud->n, 0, cloud->d_absmax); S3D_LAUNCHED(ctx); }
        cloud->absmax_pts_valid = true;
    }
    if (want_normals && !cloud->absmax_nrm_valid) {
        S3D_CUDA(ctx, cudaMemsetAsync(cloud->d_absmax + 1, 0, sizeof(float), ctx->stream));
        if (cloud->n > 0 && cloud->d_nrm) { absmax_kernel<<<blocks, 256, 0, ctx->stream>>>(cloud->d_nrm, cloud->n, 1, cloud->d_absmax + 1); S3D_LAUNCHED(ctx); }
        cloud->absmax_nrm_valid = true;
    }
    return S3D_OK;
}

int s3d_cloud_ready(s3d_ctx *ctx, const s3d_cloud *cloud)
{
    if (!cloud || !cloud->ready) return S3D_OK;
    S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, cloud->ready, 0));
    if (cloud->unpacked_stride) {
        const float *rows = cloud->d_stage ? cloud->d_stage : (const float *)cloud->d_pts;
        pack_xyz_kernel<<<(cloud->n + 255) / 256, 256, 0, ctx->stream>>>(rows, cloud->unpacked_stride, cloud->n, cloud->d_pts);
        S3D_LAUNCHED(ctx);
        cloud->unpacked_stride = 0;
    }
    return S3D_OK;
}

// Same result as s3d_cloud_upload, but the copy runs on the ctx copy stream and the call returns at once: the copy
// engine moves frame k+1 while the SMs register frame k (the registration kernel occupies every SM, a DMA transfer
// needs none).  Only copies go to the copy stream: a kernel there would queue behind the registration kernel and hold
// back the copies after it, so the float4 packing of the rows is left to the first use of the cloud, on the ctx stream.
// Truly asynchronous only from page-locked host memory; from pageable memory the driver stages the copy and the call is
// merely correct.
extern "C" int s3d_cloud_upload_async(s3d_ctx *ctx, const float *xyz, int stride_floats, int n, s3d_cloud **out)
{
    if (!ctx || !out || n < 0 || stride_floats < 3 || (n > 0 && !xyz)) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_upload_async: bad argument");
    cudaSetDevice(ctx->device);
    if (!ctx->copy_stream) {
        S3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        S3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_fence, cudaEventDisableTiming));
    }
    s3d_cloud *c = nullptr;
    int rc = cloud_alloc(ctx, n, &c);
    if (rc) return rc;
    // Blocks of the pool were last used on the ctx stream: the copy starts behind whatever that stream has been given
    // so far (nothing, when the caller pipelines frame k+1 behind a finished frame k-1).
    rc = [&]() -> int {
        S3D_CUDA(ctx, cudaEventRecord(ctx->copy_fence, ctx->stream));
        S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_fence, 0));
        S3D_CUDA(ctx, cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming));
        if (n > 0) {
            const size_t bytes = sizeof(float) * (size_t)n * stride_floats;
            void *dst = c->d_pts;
            if (stride_floats != 4) {
                S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &c->d_stage, bytes));
                dst = c->d_stage;
            }
            S3D_CUDA(ctx, cudaMemcpyAsync(dst, xyz, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            c->unpacked_stride = stride_floats;
        }
        S3D_CUDA(ctx, cudaEventRecord(c->ready, ctx->copy_stream));
        return S3D_OK;
    }();
    if (rc) { s3d_cloud_free(ctx, c); return rc; }
    *out = c;
    return S3D_OK;
}

extern "C" int s3d_cloud_wait(s3d_ctx *ctx, const s3d_cloud *cloud)
{
    if (!ctx || !cloud) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_wait: bad argument");
    if (cloud->ready) S3D_CUDA(ctx, cudaEventSynchronize(cloud->ready));
    return S3D_OK;
}

extern "C" int s3d_host_alloc(s3d_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return s3d_fail(ctx, S3D_E_ARG, "s3d_host_alloc: bad argument");
    cudaSetDevice(ctx->device);
    *out = nullptr;
    S3D_CUDA(ctx, cudaMallocHost(out, bytes > 0 ? bytes : 1));
    return S3D_OK;
}

extern "C" void s3d_host_free(s3d_ctx *ctx, void *p)
{
    if (ctx) cudaSetDevice(ctx->device);
    if (p) cudaFreeHost(p);
}

extern "C" int s3d_cloud_from_device(s3d_ctx *ctx, const void *d_xyzw, int n, s3d_cloud **out)
{
    if (!ctx || !out || n < 0 || (n > 0 && !d_xyzw)) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_from_device: bad argument");
    cudaSetDevice(ctx->device);
    s3d_cloud *c = nullptr;
    int rc = cloud_alloc(ctx, n, &c);
    if (rc) return rc;
    if (n > 0) {
        pack_xyz_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const float *)d_xyzw, 4, n, c->d_pts);
        S3D_LAUNCHED(ctx);
    }
    *out = c;
    return S3D_OK;
}

// ---- depth image -> cloud ----------------------------------------------------------------------
// x=(u-cx)z/fx, y=(v-cy)z/fy, z=d/factor evaluated in double and stored as float, exactly like
// reference src/convert2PCD.cpp:64-68 (IEEE double mul/div/sub are bit-reproducible on the device).
struct DepthPred {
    const uint16_t *depth; double inv_unused; double factor; float z_max;
    __device__ bool operator()(int i) const
    {
        uint16_t d = depth[i];
        if (d == 0) return false;
        if (z_max > 0.f) { float fz = (float)__ddiv_rn((double)d, factor); return fz >= 0.f && fz <= z_max; }
        return true;
    }
};
struct DepthEmit {
    const uint16_t *depth; int width; double fx, fy, cx, cy, factor; float4 *out;
    __device__ void operator()(int i, uint32_t pos) const
    {
        int m = i / width, n = i - m * width;
        double z = __ddiv_rn((double)depth[i], factor);
        double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)n, cx), z), fx);
        double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)m, cy), z), fy);
        out[pos] = make_float4((float)x, (float)y, (float)z, 1.0f);
    }
};
struct NoDrop { __device__ void operator()(int) const {} };

extern "C" int s3d_cloud_from_depth(s3d_ctx *ctx, const uint16_t *depth, int width, int height,
                                    const s3d_camera *cam, float z_max, s3d_cloud **out)
{
    if (!ctx || !out || !depth || !cam || width <= 0 || height <= 0) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_from_depth: bad argument");
    cudaSetDevice(ctx->device);
    int npx = width * height;
    int nblocks = (npx + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
    uint16_t *d_depth = nullptr; uint32_t *d_counts = nullptr; float4 *d_tmp = nullptr;
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_depth, sizeof(uint16_t) * (size_t)npx));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_counts, sizeof(uint32_t) * (size_t)(nblocks + 1)));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_tmp, sizeof(float4) * (size_t)npx));
    S3D_CUDA(ctx, cudaMemcpyAsync(d_depth, depth, sizeof(uint16_t) * (size_t)npx, cudaMemcpyHostToDevice, ctx->stream));
    DepthPred pred{d_depth, 0.0, cam->factor, z_max};
    DepthEmit emit{d_depth, width, cam->fx, cam->fy, cam->cx, cam->cy, cam->factor, d_tmp};
    compact_count_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(npx, pred, d_counts);
    S3D_LAUNCHED(ctx);
    compact_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nblocks, d_counts + nblocks);
    S3D_LAUNCHED(ctx);
    compact_write_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(npx, pred, emit, NoDrop(), d_counts);
    S3D_LAUNCHED(ctx);
    uint32_t total = 0;
    S3D_CUDA(ctx, cudaMemcpyAsync(&total, d_counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s3d_cloud *c = nullptr;
    int rc = cloud_alloc(ctx, (int)total, &c);
    if (rc == S3D_OK && total > 0)
        rc = cudaMemcpyAsync(c->d_pts, d_tmp, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToDevice, ctx->stream) == cudaSuccess
                 ? S3D_OK : s3d_fail(ctx, S3D_E_CUDA, "copy compacted cloud");
    cudaStreamSynchronize(ctx->stream);
    s3d_dev_free(ctx, d_depth); s3d_dev_free(ctx, d_counts); s3d_dev_free(ctx, d_tmp);
    if (rc) { if (c) { s3d_dev_free(ctx, c->d_pts); delete c; } return rc; }
    *out = c;
    return S3D_OK;
}

// ---- normals / download ------------------------------------------------------------------------

static int ensure_normals(s3d_ctx *ctx, s3d_cloud *c)
{
    if (!c->d_nrm) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &c->d_nrm, sizeof(float4) * (size_t)(c->n > 0 ? c->n : 1)));
    return S3D_OK;
}

extern "C" int s3d_cloud_set_normals(s3d_ctx *ctx, s3d_cloud *cloud, const float *nrm, int stride_floats, int n)
{
    if (!ctx || !cloud || !nrm || stride_floats < 3 || n != cloud->n) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_set_normals: bad argument");
    cudaSetDevice(ctx->device);
    int rc = s3d_cloud_ready(ctx, cloud);
    if (rc) return rc;
    rc = ensure_normals(ctx, cloud);
    if (rc) return rc;
    if (n > 0) {
        float *tmp = nullptr;
        size_t bytes = sizeof(float) * (size_t)n * stride_floats;
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &tmp, bytes));
        S3D_CUDA(ctx, cudaMemcpyAsync(tmp, nrm, bytes, cudaMemcpyHostToDevice, ctx->stream));
        pack_normals_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(tmp, stride_floats, n, cloud->d_nrm);
        S3D_LAUNCHED(ctx);
        S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        s3d_dev_free(ctx, tmp);
    }
    cloud->grid.valid = false; // sorted normals are stale
    cloud->absmax_nrm_valid = false;
    return S3D_OK;
}

extern "C" int s3d_cloud_set_normals_device(s3d_ctx *ctx, s3d_cloud *cloud, const void *d_nrm, int n)
{
    if (!ctx || !cloud || !d_nrm || n != cloud->n) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_set_normals_device: bad argument");
    cudaSetDevice(ctx->device);
    int rc = s3d_cloud_ready(ctx, cloud);
    if (rc) return rc;
    rc = ensure_normals(ctx, cloud);
    if (rc) return rc;
    if (n > 0) {
        pack_normals_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const float *)d_nrm, 4, n, cloud->d_nrm);
        S3D_LAUNCHED(ctx);
    }
    cloud->grid.valid = false;
    cloud->absmax_nrm_valid = false;
    return S3D_OK;
}

extern "C" int s3d_cloud_size(const s3d_cloud *cloud) { return cloud ? cloud->n : 0; }
extern "C" int s3d_cloud_has_normals(const s3d_cloud *cloud) { return cloud && cloud->d_nrm ? 1 : 0; }

extern "C" int s3d_cloud_download(s3d_ctx *ctx, const s3d_cloud *cloud, float *xyz, float *normals, int32_t *labels)
{
    if (!ctx || !cloud) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_download: bad argument");
    cudaSetDevice(ctx->device);
    int n = cloud->n;
    if (n == 0) return S3D_OK;
    { int rc = s3d_cloud_ready(ctx, cloud); if (rc) return rc; }
    std::vector<float4> tmp((size_t)n);
    if (xyz) {
        S3D_CUDA(ctx, cudaMemcpyAsync(tmp.data(), cloud->d_pts, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i) { xyz[3 * i] = tmp[i].x; xyz[3 * i + 1] = tmp[i].y; xyz[3 * i + 2] = tmp[i].z; }
    }
    if (normals) {
        if (!cloud->d_nrm) return s3d_fail(ctx, S3D_E_STATE, "cloud has no normals");
        S3D_CUDA(ctx, cudaMemcpyAsync(tmp.data(), cloud->d_nrm, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; ++i) { normals[3 * i] = tmp[i].x; normals[3 * i + 1] = tmp[i].y; normals[3 * i + 2] = tmp[i].z; }
    }
    if (labels) {
        if (!cloud->d_labels) return s3d_fail(ctx, S3D_E_STATE, "cloud has no labels");
        S3D_CUDA(ctx, cudaMemcpyAsync(labels, cloud->d_labels, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return S3D_OK;
}

extern "C" int s3d_cloud_drop_index(s3d_ctx *ctx, s3d_cloud *cloud)
{
    if (!ctx || !cloud) return S3D_E_ARG;
    // the buffers go back to the ctx pool (stream ordered: whatever still reads them was enqueued before)
    s3d_grid_free(ctx, cloud->grid);
    s3d_grid_free(ctx, cloud->coarse);
    s3d_dev_free(ctx, cloud->d_coarse_pts); cloud->d_coarse_pts = nullptr; cloud->cap_coarse_pts = 0;
    return S3D_OK;
}

extern "C" int s3d_memory_stats(const s3d_ctx *ctx, size_t *live_bytes, size_t *peak_live_bytes, size_t *cached_bytes)
{
    if (!ctx) return S3D_E_ARG;
    if (live_bytes) *live_bytes = ctx->pool_live_bytes;
    if (peak_live_bytes) *peak_live_bytes = ctx->pool_peak_bytes;
    if (cached_bytes) *cached_bytes = ctx->pool_cached;
    return S3D_OK;
}

static void cloud_retire(s3d_ctx *ctx, s3d_cloud *cloud, bool wait)
{
    if (!cloud) return;
    if (ctx) { cudaSetDevice(ctx->device); if (wait) cudaStreamSynchronize(ctx->stream); }
    if (!ctx) { delete cloud; return; }     // cannot return device memory without its context (leak rather than crash)
    if (cloud->ready) {                     // an upload may still be writing it
        if (wait) cudaEventSynchronize(cloud->ready);
        else cudaStreamWaitEvent(ctx->stream, cloud->ready, 0);      // whoever gets the buffers next (ctx stream order) comes after the upload
        cudaEventDestroy(cloud->ready);
    }
    s3d_dev_free(ctx, cloud->d_stage);
    s3d_grid_free(ctx, cloud->grid);
    s3d_grid_free(ctx, cloud->coarse);
    s3d_dev_free(ctx, cloud->d_coarse_pts);
    s3d_dev_free(ctx, cloud->d_pts); s3d_dev_free(ctx, cloud->d_nrm); s3d_dev_free(ctx, cloud->d_labels);
    s3d_dev_free(ctx, cloud->d_absmax);
    delete cloud;
}

extern "C" void s3d_cloud_free(s3d_ctx *ctx, s3d_cloud *cloud) { cloud_retire(ctx, cloud, true); }

// Without waiting: the buffers go back to the context's pool, which is ordered by the context's stream -- whatever was enqueued
// on the cloud still runs on intact data, whoever gets the buffers next is enqueued behind it (uploads on the copy stream are
// fenced behind the ctx stream, the index build's second stream forks from it).
extern "C" void s3d_cloud_release(s3d_ctx *ctx, s3d_cloud *cloud) { cloud_retire(ctx, cloud, false); }
