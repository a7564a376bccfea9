// filters.cu -- the steps either side of the registration path, on the device:
//   s3d_cloud_passthrough_z   pcl::PassThrough on "z"            (reference src/GraphicEnd.cpp:283-285, :291-292)
//   s3d_cloud_voxel_grid      pcl::VoxelGrid, cubic leaf         (reference src/GraphicEnd.cpp:287-295, src/saveOutput.cpp:44-46,76-79,90-93)
//   s3d_cloud_transform       pcl::transformPointCloud           (reference src/saveOutput.cpp:87)
//   s3d_cloud_concat          PointCloud::operator+=             (reference src/saveOutput.cpp:88)
//   s3d_map_fuse              the whole key-frame fusion loop    (reference src/saveOutput.cpp:47-95)
// Semantics restated in oracle/filter_oracle.c (PCL 1.7): the voxel of a point is
//   ijk = (int)(floor(x * inv_leaf) - (float)min_b), idx = i + j*dx + k*dx*dy   (float32 arithmetic, inv_leaf = 1.0f / leaf)
// with min_b/max_b the floor of the cloud's bounding box in leaf units; the output holds one centroid per
// occupied voxel in ascending idx order.  Centroids are summed in double in original point order (PCL sums in
// float in an unspecified order after an unstable sort): at least as accurate, and bit-identical between the
// CUDA path and the oracle.
#include "context.h"
#include "common.cuh"
#include "compact.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cstring>
#include <vector>

static int cloud_new(s3d_ctx *ctx, int n, s3d_cloud **out)
{
    s3d_cloud *c = new s3d_cloud();
    c->n = n;
    cudaError_t e = s3d_dev_alloc_t(ctx, &c->d_pts, sizeof(float4) * (size_t)(n > 0 ? n : 1));
    if (e != cudaSuccess) { delete c; return s3d_fail(ctx, S3D_E_CUDA, "device allocation for a cloud", e); }
    *out = c;
    return S3D_OK;
}

// ---- PassThrough ---------------------------------------------------------------------------------
struct ZPred {
    const float4 *pts; float z_min, z_max;
    __device__ bool operator()(int i) const
    {
        const float4 p = pts[i];
        // pcl::PassThrough drops non-finite points and keeps min <= z <= max
        return isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && p.z >= z_min && p.z <= z_max;
    }
};
struct CopyEmit {
    const float4 *pts; float4 *out;
    __device__ void operator()(int i, uint32_t pos) const { out[pos] = pts[i]; }
};
struct NoDropF { __device__ void operator()(int) const {} };

template <typename Pred>
static int compact_cloud(s3d_ctx *ctx, const s3d_cloud *in, Pred pred, s3d_cloud **out)
{
    const int n = in->n;
    const int nblocks = (std::max(n, 1) + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
    uint32_t *d_counts = nullptr; float4 *d_tmp = nullptr;
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_counts, sizeof(uint32_t) * (size_t)(nblocks + 1)));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_tmp, sizeof(float4) * (size_t)std::max(n, 1)));
    CopyEmit emit{in->d_pts, d_tmp};
    compact_count_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(n, pred, d_counts); S3D_LAUNCHED(ctx);
    compact_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nblocks, d_counts + nblocks); S3D_LAUNCHED(ctx);
    compact_write_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(n, pred, emit, NoDropF(), d_counts); S3D_LAUNCHED(ctx);
    uint32_t total = 0;
    S3D_CUDA(ctx, cudaMemcpyAsync(&total, d_counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s3d_cloud *c = nullptr;
    int rc = cloud_new(ctx, (int)total, &c);
    if (rc == S3D_OK && total > 0 &&
        cudaMemcpyAsync(c->d_pts, d_tmp, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
        rc = s3d_fail(ctx, S3D_E_CUDA, "copy compacted cloud");
    cudaStreamSynchronize(ctx->stream);
    s3d_dev_free(ctx, d_counts); s3d_dev_free(ctx, d_tmp);
    if (rc) { if (c) { s3d_dev_free(ctx, c->d_pts); delete c; } return rc; }
    *out = c;
    return S3D_OK;
}

extern "C" int s3d_cloud_passthrough_z(s3d_ctx *ctx, const s3d_cloud *cloud, float z_min, float z_max, s3d_cloud **out)
{
    if (!ctx || !cloud || !out) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_passthrough_z: bad argument");
    cudaSetDevice(ctx->device);
    { int rc = s3d_cloud_ready(ctx, cloud); if (rc) return rc; }
    return compact_cloud(ctx, cloud, ZPred{cloud->d_pts, z_min, z_max}, out);
}

// ---- VoxelGrid -----------------------------------------------------------------------------------
struct VoxelParams { int min_b[3]; int div_b[3]; float inv_leaf; int overflow; int n_valid; };

__global__ void __launch_bounds__(256) vg_bbox_kernel(const float4 *__restrict__ pts, int n, uint32_t *__restrict__ bbox)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;     // getMinMax3D skips them
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        #pragma unroll
        for (int a = 0; a < 3; ++a)
            if (mn[a] <= mx[a]) { atomicMin(&bbox[a], f2ord(mn[a])); atomicMax(&bbox[3 + a], f2ord(mx[a])); }
    }
}

__global__ void vg_setup_kernel(uint32_t *bbox, float leaf, VoxelParams *vp, int init)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (init) { for (int a = 0; a < 3; ++a) { bbox[a] = 0xFFFFFFFFu; bbox[3 + a] = 0u; } return; }
    const float inv = __fdiv_rn(1.0f, leaf);
    vp->inv_leaf = inv; vp->overflow = 0;
    const bool empty = bbox[0] == 0xFFFFFFFFu;
    long long prod = 1;
    for (int a = 0; a < 3; ++a) {
        const float lo = empty ? 0.f : ord2f(bbox[a]), hi = empty ? 0.f : ord2f(bbox[3 + a]);
        // PCL: min_b = (int)floor(min_p * inverse_leaf_size), max_b likewise, div_b = max_b - min_b + 1
        const double flo = floor((double)__fmul_rn(lo, inv)), fhi = floor((double)__fmul_rn(hi, inv));
        if (fhi - flo + 1.0 > 2147483647.0 || fabs(flo) > 2.0e9 || fabs(fhi) > 2.0e9) { vp->overflow = 1; vp->min_b[a] = 0; vp->div_b[a] = 1; continue; }
        vp->min_b[a] = (int)flo;
        vp->div_b[a] = (int)fhi - (int)flo + 1;
        prod *= vp->div_b[a];
        if (prod > 2147483647LL) vp->overflow = 1;       // PCL: "Leaf size is too small for the input dataset. Integer indices would overflow."
    }
}

__global__ void __launch_bounds__(256) vg_key_kernel(const float4 *__restrict__ pts, int n, const VoxelParams *__restrict__ vpp,
                                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const VoxelParams vp = *vpp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        uint32_t key = 0xFFFFFFFFu;                       // non-finite points sort to the end and are dropped
        if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
            const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, vp.inv_leaf)), (float)vp.min_b[0]);
            const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, vp.inv_leaf)), (float)vp.min_b[1]);
            const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, vp.inv_leaf)), (float)vp.min_b[2]);
            key = (uint32_t)(i0 + i1 * vp.div_b[0] + i2 * vp.div_b[0] * vp.div_b[1]);
        }
        keys[i] = key; vals[i] = (uint32_t)i;
    }
}

struct HeadPred {
    const uint32_t *keys;
    __device__ bool operator()(int i) const { const uint32_t k = keys[i]; return k != 0xFFFFFFFFu && (i == 0 || keys[i - 1] != k); }
};
struct HeadEmit {
    uint32_t *seg_start;
    __device__ void operator()(int i, uint32_t pos) const { seg_start[pos] = (uint32_t)i; }
};

// one thread per voxel: its points are contiguous in the sorted order; the sort is stable, so they are summed in
// original index order (the oracle's order)
__global__ void __launch_bounds__(256) vg_centroid_kernel(const float4 *__restrict__ pts, const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ vals, const uint32_t *__restrict__ seg_start,
                                                          int n_seg, int n, float4 *__restrict__ out)
{
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_seg; v += gridDim.x * blockDim.x) {
        const uint32_t s = seg_start[v];
        const uint32_t key = keys[s];
        double sx = 0.0, sy = 0.0, sz = 0.0; int cnt = 0;
        for (uint32_t k = s; k < (uint32_t)n && keys[k] == key; ++k) {
            const float4 p = pts[vals[k]];
            sx += (double)p.x; sy += (double)p.y; sz += (double)p.z; ++cnt;
        }
        const double c = (double)cnt;
        out[v] = make_float4((float)(sx / c), (float)(sy / c), (float)(sz / c), 1.0f);
    }
}

extern "C" int s3d_cloud_voxel_grid(s3d_ctx *ctx, const s3d_cloud *cloud, float leaf, s3d_cloud **out)
{
    if (!ctx || !cloud || !out || !(leaf > 0.f)) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_voxel_grid: bad argument");
    cudaSetDevice(ctx->device);
    const int n = cloud->n;
    if (n == 0) return cloud_new(ctx, 0, out);
    { int rc = s3d_cloud_ready(ctx, cloud); if (rc) return rc; }
    cudaStream_t st = ctx->stream;
    const int nblocks = (n + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
    uint32_t *d_bbox = nullptr, *d_keys = nullptr, *d_vals = nullptr, *d_keys2 = nullptr, *d_vals2 = nullptr, *d_counts = nullptr, *d_seg = nullptr;
    VoxelParams *d_vp = nullptr; float4 *d_tmp = nullptr; void *d_sort = nullptr;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 32, st);
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_bbox, sizeof(uint32_t) * 8));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_vp, sizeof(VoxelParams)));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_keys, sizeof(uint32_t) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_vals, sizeof(uint32_t) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_keys2, sizeof(uint32_t) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_vals2, sizeof(uint32_t) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_counts, sizeof(uint32_t) * (size_t)(nblocks + 1)));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_seg, sizeof(uint32_t) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_tmp, sizeof(float4) * (size_t)n));
    S3D_CUDA(ctx, s3d_dev_alloc(ctx, &d_sort, std::max<size_t>(sort_bytes, 16)));
    auto release = [&]() {
        s3d_dev_free(ctx, d_bbox); s3d_dev_free(ctx, d_vp); s3d_dev_free(ctx, d_keys); s3d_dev_free(ctx, d_vals); s3d_dev_free(ctx, d_keys2);
        s3d_dev_free(ctx, d_vals2); s3d_dev_free(ctx, d_counts); s3d_dev_free(ctx, d_seg); s3d_dev_free(ctx, d_tmp); s3d_dev_free(ctx, d_sort);
    };
    const int wide = std::min(ctx->sm_count * 8, (n + 255) / 256);
    vg_setup_kernel<<<1, 32, 0, st>>>(d_bbox, leaf, d_vp, 1); S3D_LAUNCHED(ctx);
    vg_bbox_kernel<<<std::min(ctx->sm_count * 2, (n + 255) / 256), 256, 0, st>>>(cloud->d_pts, n, d_bbox); S3D_LAUNCHED(ctx);
    vg_setup_kernel<<<1, 32, 0, st>>>(d_bbox, leaf, d_vp, 0); S3D_LAUNCHED(ctx);
    VoxelParams vp;
    S3D_CUDA(ctx, cudaMemcpyAsync(&vp, d_vp, sizeof(vp), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(ctx, cudaStreamSynchronize(st));
    if (vp.overflow) { release(); return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_voxel_grid: leaf size too small for the cloud (voxel indices would overflow)"); }
    vg_key_kernel<<<wide, 256, 0, st>>>(cloud->d_pts, n, d_vp, d_keys, d_vals); S3D_LAUNCHED(ctx);
    // key sort: CUB's device radix sort (stable), library plumbing for a step that is not on the north-star path
    int bits = 1;
    { long long cells = (long long)vp.div_b[0] * vp.div_b[1] * vp.div_b[2]; while (bits < 32 && (1LL << bits) <= cells) ++bits; }
    if (cub::DeviceRadixSort::SortPairs(d_sort, sort_bytes, d_keys, d_keys2, d_vals, d_vals2, n, 0, 32, st) != cudaSuccess) {
        release(); return s3d_fail(ctx, S3D_E_CUDA, "radix sort of voxel keys");
    }
    ctx->launches += 1;
    HeadPred pred{d_keys2};
    HeadEmit emit{d_seg};
    compact_count_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, st>>>(n, pred, d_counts); S3D_LAUNCHED(ctx);
    compact_scan_kernel<<<1, 1024, 0, st>>>(d_counts, nblocks, d_counts + nblocks); S3D_LAUNCHED(ctx);
    compact_write_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, st>>>(n, pred, emit, NoDropF(), d_counts); S3D_LAUNCHED(ctx);
    uint32_t n_seg = 0;
    S3D_CUDA(ctx, cudaMemcpyAsync(&n_seg, d_counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    S3D_CUDA(ctx, cudaStreamSynchronize(st));
    s3d_cloud *c = nullptr;
    int rc = cloud_new(ctx, (int)n_seg, &c);
    if (rc == S3D_OK && n_seg > 0) {
        vg_centroid_kernel<<<std::min(ctx->sm_count * 8, ((int)n_seg + 255) / 256), 256, 0, st>>>(cloud->d_pts, d_keys2, d_vals2, d_seg,
                                                                                                (int)n_seg, n, c->d_pts);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = s3d_fail(ctx, S3D_E_CUDA, "vg_centroid_kernel launch");
    }
    cudaStreamSynchronize(st);
    release();
    (void)bits;
    if (rc) { if (c) { s3d_dev_free(ctx, c->d_pts); delete c; } return rc; }
    *out = c;
    return S3D_OK;
}

// ---- transform / concat / map fusion -------------------------------------------------------------
__global__ void __launch_bounds__(256) xform_cloud_kernel(const float4 *__restrict__ in, int n, Pose12f T, float4 *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = in[i];
        const float3 x = s3d_xform(T.m, p.x, p.y, p.z);      // Matrix4f * point, the registration's own expression
        out[i] = make_float4(x.x, x.y, x.z, 1.0f);
    }
}

extern "C" int s3d_cloud_transform(s3d_ctx *ctx, const s3d_cloud *cloud, const double *T16, s3d_cloud **out)
{
    if (!ctx || !cloud || !T16 || !out) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_transform: bad argument");
    cudaSetDevice(ctx->device);
    s3d_cloud *c = nullptr;
    int rc = s3d_cloud_ready(ctx, cloud);
    if (rc) return rc;
    rc = cloud_new(ctx, cloud->n, &c);
    if (rc) return rc;
    Pose12f T;
    for (int k = 0; k < 12; ++k) T.m[k] = (float)T16[k];
    if (cloud->n > 0) {
        xform_cloud_kernel<<<std::min(ctx->sm_count * 8, (cloud->n + 255) / 256), 256, 0, ctx->stream>>>(cloud->d_pts, cloud->n, T, c->d_pts);
        S3D_LAUNCHED(ctx);
    }
    *out = c;
    return S3D_OK;
}

extern "C" int s3d_cloud_concat(s3d_ctx *ctx, const s3d_cloud *const *clouds, int n_clouds, s3d_cloud **out)
{
    if (!ctx || !out || n_clouds < 0 || (n_clouds > 0 && !clouds)) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_concat: bad argument");
    cudaSetDevice(ctx->device);
    long long total = 0;
    for (int i = 0; i < n_clouds; ++i) { if (!clouds[i]) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_concat: null cloud"); total += clouds[i]->n; }
    if (total > 0x7fffffffLL) return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_concat: too many points");
    s3d_cloud *c = nullptr;
    int rc = cloud_new(ctx, (int)total, &c);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; i < n_clouds; ++i) {
        if (clouds[i]->n == 0) continue;
        if ((rc = s3d_cloud_ready(ctx, clouds[i])) != S3D_OK) { s3d_cloud_free(ctx, c); return rc; }
        S3D_CUDA(ctx, cudaMemcpyAsync(c->d_pts + off, clouds[i]->d_pts, sizeof(float4) * (size_t)clouds[i]->n, cudaMemcpyDeviceToDevice, ctx->stream));
        off += (size_t)clouds[i]->n;
    }
    *out = c;
    return S3D_OK;
}

// The key-frame fusion loop of saveOutput (reference src/saveOutput.cpp:47-95): every key-frame cloud is voxel
// filtered, z-filtered, moved into the map frame by its optimised pose and appended; the sum is voxel filtered again.
extern "C" int s3d_map_fuse(s3d_ctx *ctx, const s3d_cloud *const *clouds, const double *poses16, int n_clouds, float leaf, float z_max,
                            s3d_cloud **out)
{
    if (!ctx || !out || n_clouds <= 0 || !clouds || !poses16 || !(leaf > 0.f)) return s3d_fail(ctx, S3D_E_ARG, "s3d_map_fuse: bad argument");
    std::vector<s3d_cloud *> parts;
    int rc = S3D_OK;
    for (int i = 0; i < n_clouds && rc == S3D_OK; ++i) {
        s3d_cloud *v = nullptr, *z = nullptr, *t = nullptr;
        rc = s3d_cloud_voxel_grid(ctx, clouds[i], leaf, &v);                                        // :76-79
        if (rc == S3D_OK) rc = s3d_cloud_passthrough_z(ctx, v, 0.0f, z_max, &z);                     // :81-84
        if (rc == S3D_OK) rc = s3d_cloud_transform(ctx, z, poses16 + 16 * (size_t)i, &t);            // :87
        if (v) s3d_cloud_free(ctx, v);
        if (z) s3d_cloud_free(ctx, z);
        if (rc == S3D_OK) parts.push_back(t); else if (t) s3d_cloud_free(ctx, t);
    }
    s3d_cloud *sum = nullptr;
    if (rc == S3D_OK) rc = s3d_cloud_concat(ctx, parts.data(), (int)parts.size(), &sum);             // :88
    for (s3d_cloud *p : parts) s3d_cloud_free(ctx, p);
    if (rc == S3D_OK) rc = s3d_cloud_voxel_grid(ctx, sum, leaf, out);                                // :90-93
    if (sum) s3d_cloud_free(ctx, sum);
    return rc;
}

// ---- per-point normals from the organised depth image --------------------------
// Companion of s3d_cloud_from_depth for scenes that are not made of a few big planes: the normal of pixel (u,v) is
// the normalised cross product of the central differences of the back-projected neighbours `step` pixels away,
// turned towards the camera.  A pixel gets no normal (w = 0) when a neighbour is missing (border, hole, outside
// the z filter) or lies more than `max_jump` metres away in depth (an occlusion edge).  Back-projection in double
// like reference src/convert2PCD.cpp:64-68, differences and cross product in float32 with the explicit
// round-to-nearest intrinsics so that oracle/filter_oracle.c reproduces every bit.
struct DepthNormalCtx {
    const uint16_t *depth; int width, height; double fx, fy, cx, cy, factor; float z_max; int step; float max_jump;
    __device__ bool valid(int u, int v) const
    {
        if (u < 0 || v < 0 || u >= width || v >= height) return false;
        const uint16_t d = depth[v * width + u];
        if (d == 0) return false;
        if (z_max > 0.f) { const float fz = (float)__ddiv_rn((double)d, factor); return fz >= 0.f && fz <= z_max; }
        return true;
    }
    __device__ float3 point(int u, int v) const
    {
        const double z = __ddiv_rn((double)depth[v * width + u], factor);
        const double x = __ddiv_rn(__dmul_rn(__dsub_rn((double)u, cx), z), fx);
        const double y = __ddiv_rn(__dmul_rn(__dsub_rn((double)v, cy), z), fy);
        return make_float3((float)x, (float)y, (float)z);
    }
};
struct DepthNormalPred {
    DepthNormalCtx c;
    __device__ bool operator()(int i) const { const int v = i / c.width; return c.valid(i - v * c.width, v); }
};
struct DepthNormalEmit {
    DepthNormalCtx c; float4 *pts; float4 *nrm;
    __device__ void operator()(int i, uint32_t pos) const
    {
        const int v = i / c.width, u = i - v * c.width;
        const float3 p = c.point(u, v);
        pts[pos] = make_float4(p.x, p.y, p.z, 1.0f);
        float4 n = make_float4(0.f, 0.f, 0.f, 0.f);
        const int s = c.step;
        if (c.valid(u - s, v) && c.valid(u + s, v) && c.valid(u, v - s) && c.valid(u, v + s)) {
            const float3 l = c.point(u - s, v), r = c.point(u + s, v), t = c.point(u, v - s), b = c.point(u, v + s);
            const bool jump = fabsf(__fsub_rn(l.z, p.z)) > c.max_jump || fabsf(__fsub_rn(r.z, p.z)) > c.max_jump ||
                              fabsf(__fsub_rn(t.z, p.z)) > c.max_jump || fabsf(__fsub_rn(b.z, p.z)) > c.max_jump;
            if (!jump) {
                const float ax = __fsub_rn(r.x, l.x), ay = __fsub_rn(r.y, l.y), az = __fsub_rn(r.z, l.z);
                const float bx = __fsub_rn(b.x, t.x), by = __fsub_rn(b.y, t.y), bz = __fsub_rn(b.z, t.z);
                float nx = __fmaf_rn(ay, bz, -__fmul_rn(az, by));
                float ny = __fmaf_rn(az, bx, -__fmul_rn(ax, bz));
                float nz = __fmaf_rn(ax, by, -__fmul_rn(ay, bx));
                const float l2 = __fmaf_rn(nz, nz, __fmaf_rn(ny, ny, __fmul_rn(nx, nx)));
                if (l2 > 1e-24f) {
                    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(l2));
                    nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
                    // towards the camera (the origin): n . p < 0
                    const float dp = __fmaf_rn(nz, p.z, __fmaf_rn(ny, p.y, __fmul_rn(nx, p.x)));
                    if (dp > 0.f) { nx = -nx; ny = -ny; nz = -nz; }
                    n = make_float4(nx, ny, nz, 1.0f);
                }
            }
        }
        nrm[pos] = n;
    }
};

extern "C" int s3d_cloud_from_depth_normals(s3d_ctx *ctx, const uint16_t *depth, int width, int height, const s3d_camera *cam,
                                            float z_max, int step, float max_jump, s3d_cloud **out)
{
    if (!ctx || !out || !depth || !cam || width <= 0 || height <= 0 || step < 1 || !(max_jump > 0.f))
        return s3d_fail(ctx, S3D_E_ARG, "s3d_cloud_from_depth_normals: bad argument");
    cudaSetDevice(ctx->device);
    const int npx = width * height;
    const int nblocks = (npx + S3D_COMPACT_BLOCK - 1) / S3D_COMPACT_BLOCK;
    uint16_t *d_depth = nullptr; uint32_t *d_counts = nullptr; float4 *d_tmp = nullptr, *d_tmpn = nullptr;
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_depth, sizeof(uint16_t) * (size_t)npx));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_counts, sizeof(uint32_t) * (size_t)(nblocks + 1)));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_tmp, sizeof(float4) * (size_t)npx));
    S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &d_tmpn, sizeof(float4) * (size_t)npx));
    S3D_CUDA(ctx, cudaMemcpyAsync(d_depth, depth, sizeof(uint16_t) * (size_t)npx, cudaMemcpyHostToDevice, ctx->stream));
    DepthNormalCtx c{d_depth, width, height, cam->fx, cam->fy, cam->cx, cam->cy, cam->factor, z_max, step, max_jump};
    DepthNormalPred pred{c};
    DepthNormalEmit emit{c, d_tmp, d_tmpn};
    compact_count_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(npx, pred, d_counts); S3D_LAUNCHED(ctx);
    compact_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, nblocks, d_counts + nblocks); S3D_LAUNCHED(ctx);
    compact_write_kernel<<<nblocks, S3D_COMPACT_BLOCK, 0, ctx->stream>>>(npx, pred, emit, NoDropF(), d_counts); S3D_LAUNCHED(ctx);
    uint32_t total = 0;
    S3D_CUDA(ctx, cudaMemcpyAsync(&total, d_counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s3d_cloud *cl = nullptr;
    int rc = cloud_new(ctx, (int)total, &cl);
    if (rc == S3D_OK) {
        if (s3d_dev_alloc_t(ctx, &cl->d_nrm, sizeof(float4) * (size_t)std::max<uint32_t>(total, 1)) != cudaSuccess)
            rc = s3d_fail(ctx, S3D_E_CUDA, "device allocation for normals");
    }
    if (rc == S3D_OK && total > 0) {
        if (cudaMemcpyAsync(cl->d_pts, d_tmp, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(cl->d_nrm, d_tmpn, sizeof(float4) * (size_t)total, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
            rc = s3d_fail(ctx, S3D_E_CUDA, "copy compacted cloud");
    }
    cudaStreamSynchronize(ctx->stream);
    s3d_dev_free(ctx, d_depth); s3d_dev_free(ctx, d_counts); s3d_dev_free(ctx, d_tmp); s3d_dev_free(ctx, d_tmpn);
    if (rc) { if (cl) { s3d_dev_free(ctx, cl->d_pts); s3d_dev_free(ctx, cl->d_nrm); delete cl; } return rc; }
    *out = cl;
    return S3D_OK;
}
