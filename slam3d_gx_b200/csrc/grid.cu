// grid.cu -- device-built uniform grid over a target cloud: the exact-NN search index.
//
// Plays the role of the FLANN KD-tree behind pcl::search::KdTree in PCL-1.7
// CorrespondenceEstimation (the "PCL KD-tree calls" of the north star), re-designed for the GPU:
// a counting sort of the target points by cell (x fastest), so that any run of cells along x is
// one contiguous, coalescible range of float4 points.  Everything, including the choice of the
// cell size, happens on the device: there is no host round trip between build and first query.
//
//   bbox reduce -> grid_setup (1 thread) -> zero counts -> count (atomic rank) -> exclusive scan
//   -> scatter (points + normals into cell order, original index kept in .w)
#include "context.h"
#include "common.cuh"
#include "grid.cuh"
#include "compact.cuh"
#include <cstdlib>
#include <algorithm>

#define SCAN_ITEMS 8
#define SCAN_BLOCK 1024
#define SCAN_TILE (SCAN_ITEMS * SCAN_BLOCK)

__global__ void bbox_init_kernel(uint32_t *bbox)
{
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4 *__restrict__ pts, int n, uint32_t *__restrict__ bbox)
{
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    __shared__ float smn[8][3], smx[8][3];
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
        for (int a = 0; a < 3; ++a) { smn[threadIdx.x >> 5][a] = mn[a]; smx[threadIdx.x >> 5][a] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {   // one atomic pair per axis per CTA (not per warp: 6 hot addresses serialise)
        const int a = threadIdx.x;
        float lo = smn[0][a], hi = smx[0][a];
        #pragma unroll
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, smn[w][a]); hi = fmaxf(hi, smx[w][a]); }
        if (lo <= hi) { atomicMin(&bbox[a], f2ord(lo)); atomicMax(&bbox[3 + a], f2ord(hi)); }
    }
}

__global__ void grid_setup_kernel(const uint32_t *__restrict__ bbox, int n, float requested_cell, float auto_scale,
                                  uint32_t max_cells, GridParams *__restrict__ gp)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float mn[3], ex[3];
    for (int a = 0; a < 3; ++a) {
        mn[a] = n > 0 ? ord2f(bbox[a]) : 0.f;
        float mxv = n > 0 ? ord2f(bbox[3 + a]) : 0.f;
        ex[a] = fmaxf(mxv - mn[a], 1e-6f);
    }
    float h = requested_cell;
    if (!(h > 0.f)) {
        // surfaces, not volumes: spacing ~ sqrt(area / n); area estimated from the bbox faces
        float area = ex[0] * ex[1] + ex[1] * ex[2] + ex[0] * ex[2];
        h = auto_scale * sqrtf(area / (float)(n > 0 ? n : 1));
    }
    h = fmaxf(h, 1e-4f);
    int nx, ny, nz;
    for (;;) {
        nx = (int)fminf((float)S3D_GRID_MAX_DIM, floorf(ex[0] / h) + 1.f);
        ny = (int)fminf((float)S3D_GRID_MAX_DIM, floorf(ex[1] / h) + 1.f);
        nz = (int)fminf((float)S3D_GRID_MAX_DIM, floorf(ex[2] / h) + 1.f);
        bool capped = (floorf(ex[0] / h) + 1.f > S3D_GRID_MAX_DIM) || (floorf(ex[1] / h) + 1.f > S3D_GRID_MAX_DIM) ||
                      (floorf(ex[2] / h) + 1.f > S3D_GRID_MAX_DIM);
        if (!capped && (double)nx * ny * nz <= (double)max_cells && (double)ny * nz * ((nx + 31) / 32) <= (double)(max_cells / 8)) break;
        h *= 1.25992105f;
    }
    gp->ox = mn[0]; gp->oy = mn[1]; gp->oz = mn[2];
    gp->cell = h; gp->inv_cell = 1.0f / h;
    gp->nx = nx; gp->ny = ny; gp->nz = nz; gp->ncells = nx * ny * nz; gp->n_points = n;
    gp->words = (nx + 31) / 32;
    gp->mask_words = ny * nz * gp->words;
}

__global__ void __launch_bounds__(256) grid_zero_kernel(uint32_t *__restrict__ cell_start, uint32_t *__restrict__ rowmask,
                                                        const GridParams *__restrict__ gp)
{
    int n = gp->ncells + 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cell_start[i] = 0u;
    int m = gp->mask_words;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) rowmask[i] = 0u;
}

__global__ void __launch_bounds__(256) grid_count_kernel(const float4 *__restrict__ pts, int n, const GridParams *__restrict__ gpp,
                                                         uint32_t *__restrict__ cell_start, uint32_t *__restrict__ rank,
                                                         uint32_t *__restrict__ rowmask)
{
    GridParams gp = *gpp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        const int cx = grid_clampi(grid_fcoord(p.x, gp.ox, gp.inv_cell), gp.nx);
        const int cy = grid_clampi(grid_fcoord(p.y, gp.oy, gp.inv_cell), gp.ny);
        const int cz = grid_clampi(grid_fcoord(p.z, gp.oz, gp.inv_cell), gp.nz);
        rank[i] = atomicAdd(&cell_start[(cz * gp.ny + cy) * gp.nx + cx], 1u);
        uint32_t *mw = &rowmask[((size_t)cz * gp.ny + cy) * gp.words + (cx >> 5)];
        const uint32_t bit = 1u << (cx & 31);
        if (!(*mw & bit)) atomicOr(mw, bit);
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK) grid_scan_reduce_kernel(const uint32_t *__restrict__ data, const GridParams *__restrict__ gp,
                                                                      uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t ws[32];
    int n = gp->ncells + 1;
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
    if (base + SCAN_ITEMS <= n) {
        uint4 a = *reinterpret_cast<const uint4 *>(data + base), b = *reinterpret_cast<const uint4 *>(data + base + 4);
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += data[base + k];
    }
    s = (uint32_t)warp_sum_i((int)s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = (uint32_t)warp_sum_i((int)ws[threadIdx.x]);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK) grid_scan_apply_kernel(uint32_t *__restrict__ data, const GridParams *__restrict__ gp,
                                                                     const uint32_t *__restrict__ block_offsets)
{
    __shared__ uint32_t ws[32];
    int n = gp->ncells + 1;
    int tile = blockIdx.x * SCAN_TILE;
    if (tile >= n) return;
    int base = tile + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    bool full = base + SCAN_ITEMS <= n;
    if (full) {
        uint4 a = *reinterpret_cast<const uint4 *>(data + base), b = *reinterpret_cast<const uint4 *>(data + base + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k) v[k] = (base + k < n) ? data[base + k] : 0u;
    }
    uint32_t tot = 0;
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { uint32_t t = v[k]; v[k] = tot; tot += t; }
    uint32_t incl = tot;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = ws[threadIdx.x], wi = w;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (threadIdx.x >= o) wi += t;
        }
        ws[threadIdx.x] = wi - w;
    }
    __syncthreads();
    uint32_t off = block_offsets[blockIdx.x] + ws[threadIdx.x >> 5] + incl - tot;
    if (full) {
        uint4 a = make_uint4(v[0] + off, v[1] + off, v[2] + off, v[3] + off);
        uint4 b = make_uint4(v[4] + off, v[5] + off, v[6] + off, v[7] + off);
        *reinterpret_cast<uint4 *>(data + base) = a; *reinterpret_cast<uint4 *>(data + base + 4) = b;
    } else {
        for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) data[base + k] = v[k] + off;
    }
}

__global__ void __launch_bounds__(256) grid_scatter_kernel(const float4 *__restrict__ pts, const float4 *__restrict__ nrm, int n,
                                                           const GridParams *__restrict__ gpp, const uint32_t *__restrict__ cell_start,
                                                           const uint32_t *__restrict__ rank, float4 *__restrict__ sorted_pts,
                                                           float4 *__restrict__ sorted_nrm)
{
    GridParams gp = *gpp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        uint32_t pos = cell_start[grid_cell_index(gp, p.x, p.y, p.z)] + rank[i];
        sorted_pts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        if (nrm) sorted_nrm[pos] = nrm[i];
    }
}

void s3d_grid_free(s3d_ctx *ctx, GridIndex &g)
{
    s3d_dev_free(ctx, g.d_params); s3d_dev_free(ctx, g.d_cell_start); s3d_dev_free(ctx, g.d_sorted_pts); s3d_dev_free(ctx, g.d_sorted_nrm);
    s3d_dev_free(ctx, g.d_rank); s3d_dev_free(ctx, g.d_bbox); s3d_dev_free(ctx, g.d_block_sums); s3d_dev_free(ctx, g.d_rowmask);
    g = GridIndex();
}

static float auto_scale_from_env()
{
    const char *e = getenv("S3D_GRID_CELL_SCALE");
    float v = e ? (float)atof(e) : 0.f;
    return v > 0.f ? v : 1.5f;
}

// One point out of every `stride` consecutive ones, at a pseudo-random offset inside its group: the
// decimated set behind the coarse seeding index.  (A fixed offset would keep whole image columns of an
// organised cloud and leave 16-pixel gaps between them; the hashed offset gives an isotropic subsample.)
__global__ void __launch_bounds__(256) decimate_kernel(const float4 *__restrict__ pts, int n_out, int stride, float4 *__restrict__ out)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
        const uint32_t off = (uint32_t)(s3d_mix((uint64_t)i) % (uint64_t)stride);
        out[i] = pts[(size_t)i * stride + off];
    }
}

// Buffers of an index over n device points (max_cells bounds the dense cell array); allocation only.
static int grid_ensure_buffers(s3d_ctx *ctx, GridIndex &g, bool with_normals, int n, uint32_t max_cells)
{
    size_t np = (size_t)(n > 0 ? n : 1);
    if (g.cap_points < n || g.cap_cells < max_cells || !g.d_params) {
        s3d_grid_free(ctx, g);
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_params, sizeof(GridParams)));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_cell_start, sizeof(uint32_t) * ((size_t)max_cells + 16)));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_sorted_pts, sizeof(float4) * np));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_rank, sizeof(uint32_t) * np));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_bbox, sizeof(uint32_t) * 8));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_block_sums, sizeof(uint32_t) * (max_cells / SCAN_TILE + 8)));
        S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_rowmask, sizeof(uint32_t) * ((size_t)max_cells / 8 + 8192)));
        g.cap_points = n; g.cap_cells = max_cells;
    }
    if (with_normals && !g.d_sorted_nrm) S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &g.d_sorted_nrm, sizeof(float4) * (size_t)(g.cap_points > 0 ? g.cap_points : 1)));
    return S3D_OK;
}

// The launches that build (or rebuild) the index: nothing but kernel launches on ctx->stream, so the sequence can be captured.
static int grid_launch(s3d_ctx *ctx, cudaStream_t st, GridIndex &g, const float4 *d_pts, const float4 *d_nrm, int n, float cell, float scale,
                       uint32_t max_cells)
{
    const int wide = ctx->sm_count * 8;
    const int scan_blocks = (int)(max_cells / SCAN_TILE) + 1;
    bbox_init_kernel<<<1, 32, 0, st>>>(g.d_bbox); S3D_LAUNCHED(ctx);
    if (n > 0) { bbox_kernel<<<std::min(ctx->sm_count * 2, (n + 255) / 256), 256, 0, st>>>(d_pts, n, g.d_bbox); S3D_LAUNCHED(ctx); }
    grid_setup_kernel<<<1, 32, 0, st>>>(g.d_bbox, n, cell, scale, max_cells, g.d_params); S3D_LAUNCHED(ctx);
    grid_zero_kernel<<<std::min(wide, (int)(max_cells / 1024) + 1), 256, 0, st>>>(g.d_cell_start, g.d_rowmask, g.d_params); S3D_LAUNCHED(ctx);
    if (n > 0) { grid_count_kernel<<<std::min(wide, (n + 255) / 256), 256, 0, st>>>(d_pts, n, g.d_params, g.d_cell_start, g.d_rank, g.d_rowmask); S3D_LAUNCHED(ctx); }
    grid_scan_reduce_kernel<<<scan_blocks, SCAN_BLOCK, 0, st>>>(g.d_cell_start, g.d_params, g.d_block_sums); S3D_LAUNCHED(ctx);
    compact_scan_kernel<<<1, 1024, 0, st>>>(g.d_block_sums, scan_blocks, g.d_block_sums + scan_blocks); S3D_LAUNCHED(ctx);
    grid_scan_apply_kernel<<<scan_blocks, SCAN_BLOCK, 0, st>>>(g.d_cell_start, g.d_params, g.d_block_sums); S3D_LAUNCHED(ctx);
    if (n > 0) {
        grid_scatter_kernel<<<std::min(wide, (n + 255) / 256), 256, 0, st>>>(d_pts, d_nrm, n, g.d_params, g.d_cell_start, g.d_rank,
                                                                           g.d_sorted_pts, d_nrm ? g.d_sorted_nrm : nullptr);
        S3D_LAUNCHED(ctx);
    }
    return S3D_OK;
}

static int build_launches(s3d_ctx *ctx, s3d_cloud *c, float cell, float scale, bool want_coarse, int nc)
{
    // The decimated seeding index (every S3D_COARSE_STRIDE-th point: first-iteration seeds, see icp.cu) depends on nothing the
    // full index produces: its ten small launches run on a second stream beside the nine of the full index (fork / join by
    // events; captured, they become two parallel branches of the graph).
    if (want_coarse) {
        if (!ctx->aux_stream) {
            S3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            S3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
            S3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
        }
        S3D_CUDA(ctx, cudaEventRecord(ctx->aux_fork, ctx->stream));
        S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
        decimate_kernel<<<std::min(ctx->sm_count * 8, (nc + 255) / 256), 256, 0, ctx->aux_stream>>>(c->d_pts, nc, S3D_COARSE_STRIDE, c->d_coarse_pts);
        S3D_LAUNCHED(ctx);
        int rc = grid_launch(ctx, ctx->aux_stream, c->coarse, c->d_coarse_pts, nullptr, nc, 0.f, scale, S3D_COARSE_MAX_CELLS);
        if (rc) return rc;
    }
    int rc = grid_launch(ctx, ctx->stream, c->grid, c->d_pts, c->d_nrm, c->n, cell, scale, S3D_GRID_MAX_CELLS);
    if (want_coarse) {
        S3D_CUDA(ctx, cudaEventRecord(ctx->aux_join, ctx->aux_stream));
        S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));
    }
    return rc;
}

static uint64_t fnv(uint64_t h, const void *p, size_t n)
{
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

// The build is ~19 small kernels whose parameters are all device resident: issued one by one they are bound by the host's
// launch rate (~5 us each).  The sequence is therefore captured into a CUDA graph, keyed by every pointer and size it
// touches (the caching allocator hands the same buffers to the next frame's cloud, so steady state replays one graph).
int s3d_grid_build(s3d_ctx *ctx, s3d_cloud *c, float cell)
{
    static const float scale = auto_scale_from_env();
    static const bool use_graph = []() { const char *e = getenv("S3D_INDEX_GRAPH"); return !e || atoi(e) != 0; }();
    const bool want_coarse = c->n >= S3D_COARSE_MIN_POINTS;
    const int nc = c->n / S3D_COARSE_STRIDE;
    int rc = grid_ensure_buffers(ctx, c->grid, c->d_nrm != nullptr, c->n, S3D_GRID_MAX_CELLS);
    if (rc) return rc;
    c->coarse.valid = false;
    if (want_coarse) {
        if (c->cap_coarse_pts < nc) {
            s3d_dev_free(ctx, c->d_coarse_pts); c->d_coarse_pts = nullptr;
            S3D_CUDA(ctx, s3d_dev_alloc_t(ctx, &c->d_coarse_pts, sizeof(float4) * (size_t)nc));
            c->cap_coarse_pts = nc;
        }
        rc = grid_ensure_buffers(ctx, c->coarse, false, nc, S3D_COARSE_MAX_CELLS);
        if (rc) return rc;
    }
    bool done = false;
    if (use_graph) {
        const void *ptrs[] = {c->d_pts, c->d_nrm, c->d_coarse_pts, c->grid.d_params, c->grid.d_cell_start, c->grid.d_sorted_pts,
                              c->grid.d_sorted_nrm, c->grid.d_rank, c->grid.d_bbox, c->grid.d_block_sums, c->grid.d_rowmask,
                              c->coarse.d_params, c->coarse.d_cell_start, c->coarse.d_sorted_pts, c->coarse.d_rank, c->coarse.d_bbox,
                              c->coarse.d_block_sums, c->coarse.d_rowmask, (const void *)ctx->stream};
        uint64_t key = fnv(0xcbf29ce484222325ull, ptrs, sizeof(ptrs));
        const int ints[] = {c->n, (int)want_coarse, nc};
        key = fnv(key, ints, sizeof(ints)); key = fnv(key, &cell, sizeof(cell)); key = fnv(key, &scale, sizeof(scale));
        auto it = ctx->graphs.find(key);
        if (it == ctx->graphs.end()) {
            if (ctx->graphs.size() > 256) { for (auto &kv : ctx->graphs) cudaGraphExecDestroy(kv.second); ctx->graphs.clear(); }
            cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
            const int64_t l0 = ctx->launches;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = build_launches(ctx, c, cell, scale, want_coarse, nc);
                cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
                ctx->launches = l0;                              // captured, not launched
                if (rc == S3D_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                    it = ctx->graphs.emplace(key, exec).first;
                    ctx->graph_nodes[key] = want_coarse ? 19 : 9;
                }
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                rc = S3D_OK;
            }
        }
        if (it != ctx->graphs.end() && cudaGraphLaunch(it->second, ctx->stream) == cudaSuccess) {
            ctx->launches += ctx->graph_nodes[key];
            done = true;
        }
    }
    if (!done) {
        rc = build_launches(ctx, c, cell, scale, want_coarse, nc);
        if (rc) return rc;
    }
    c->grid.valid = true; c->grid.has_normals = c->d_nrm != nullptr; c->grid.requested_cell = cell; c->grid.n = c->n;
    if (want_coarse) { c->coarse.valid = true; c->coarse.has_normals = false; c->coarse.requested_cell = 0.f; c->coarse.n = nc; }
    return S3D_OK;
}
