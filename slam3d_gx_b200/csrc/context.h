// context.h -- internal objects behind the opaque handles of include/slam3d_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/slam3d_b200.h"

#define S3D_GRID_MAX_CELLS (1u << 23)   // dense cell-start array budget per target (32 MiB)
#define S3D_GRID_MAX_DIM 2040           // per-axis cap: keeps the float cell coordinate error < 1e-3 cells
#ifndef S3D_COARSE_STRIDE
#define S3D_COARSE_STRIDE 16            // the coarse seeding index holds every 16th target point
#endif
#define S3D_COARSE_MIN_POINTS 4096      // smaller targets are searched without seeds
#define S3D_COARSE_MAX_CELLS (1u << 20)

// Device-resident description of a target's search grid; written by grid_setup_kernel, read by
// every kernel that searches (no host round trip between build and use).
struct GridParams {
    float ox, oy, oz;      // origin (bbox min)
    float inv_cell, cell;
    int nx, ny, nz;
    int ncells;
    int n_points;
    int words;             // row-occupancy masks: `words` 32-bit words per row (y,z), one bit per cell along x
    int mask_words;
};

struct GridIndex {
    bool valid = false;
    bool has_normals = false;
    float requested_cell = 0.f;
    int n = 0, cap_points = -1;
    uint32_t cap_cells = 0;
    GridParams *d_params = nullptr;   // 1
    uint32_t *d_cell_start = nullptr; // S3D_GRID_MAX_CELLS + 1
    float4 *d_sorted_pts = nullptr;   // n : (x,y,z, original index bits)
    float4 *d_sorted_nrm = nullptr;   // n : (nx,ny,nz,valid) in sorted order
    uint32_t *d_rank = nullptr;       // n : rank of the point inside its cell
    uint32_t *d_bbox = nullptr;       // 6 ordered-uint min/max
    uint32_t *d_block_sums = nullptr; // scan scratch
    uint32_t *d_rowmask = nullptr;    // occupancy bits per row of cells (empty-space skipping in the ball search)
};

struct s3d_cloud {
    int n = 0;
    float4 *d_pts = nullptr;      // (x,y,z,1)
    float4 *d_nrm = nullptr;      // (nx,ny,nz,valid) or null
    int32_t *d_labels = nullptr;  // plane id or -1, or null
    GridIndex grid;               // all points: the exact search index
    GridIndex coarse;             // every 16th point: first-iteration seeds
    float4 *d_coarse_pts = nullptr; int cap_coarse_pts = 0;
    // asynchronous upload (s3d_cloud_upload_async): recorded on the ctx copy stream behind the host-to-device copy;
    // every entry point that touches the cloud orders the ctx stream behind it first (s3d_cloud_ready)
    cudaEvent_t ready = nullptr;
    float *d_stage = nullptr;     // rows as the host holds them (stride != 4), until the cloud is freed
    mutable int unpacked_stride = 0;   // != 0: the rows still have to be packed into float4 (x,y,z,1), on the ctx stream, at first use
    // largest finite |coordinate| of the points [0] and largest finite |component| of the valid normals [1]: the data
    // bounds behind the resolution of the order-independent sums (common.cuh); computed on the device at first use
    mutable float *d_absmax = nullptr;
    mutable bool absmax_pts_valid = false, absmax_nrm_valid = false;
};

// one registration unit as the kernels see it
struct PairDesc {
    const float4 *src; int n_src;
    const float4 *tgt; const float4 *tgt_nrm; int n_tgt;            // original order (brute force)
    const float4 *sorted_pts; const float4 *sorted_nrm;             // grid order
    const uint32_t *cell_start; const GridParams *grid; const uint32_t *rowmask;
    const float4 *coarse_pts; const uint32_t *coarse_cell_start; const uint32_t *coarse_rowmask; const GridParams *coarse_grid;  // null when absent
    const float *src_absmax; const float *tgt_absmax;               // s3d_cloud::d_absmax of the two clouds
};

struct PairState {
    double T[12];
    float Tf[12];
    float Tf_prev[12];   // pose of the previous iteration (skip test of the seeded search)
    double fitness;
    int inliers;
    int iterations;
    int status;
    unsigned ticket;
};

struct s3d_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // uploads of the NEXT frame while the ctx stream registers the current one
    cudaEvent_t copy_fence = nullptr;     // orders the copy stream behind what the ctx stream has been given so far
    cudaStream_t aux_stream = nullptr;    // second branch of the index build (grid.cu)
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    std::string err;
    int64_t launches = 0;
    // batch scratch (grown on demand)
    int cap_pairs = 0, cap_ctas = 0;
    PairDesc *d_desc = nullptr; PairDesc *h_desc = nullptr;
    PairState *d_state = nullptr; PairState *h_state = nullptr;
    long long *d_partials = nullptr;  // [pairs][ctas][S3D_ROW]: (hi, lo) partial sums of the order-independent accumulation
    // brute-force scratch + last-correspondence buffer
    int cap_nn = 0;
    int32_t *d_nn_idx = nullptr; float *d_nn_d2 = nullptr; int32_t *d_nn_pos = nullptr;
    int last_nn_n = 0; int32_t *d_last_nn = nullptr; int cap_last_nn = 0;
    // persistent tile-search path: per-query correspondence cache + group barriers
    size_t cap_tile_nn = 0;
    float4 *d_cq = nullptr; float4 *d_cn = nullptr; float4 *d_cq2 = nullptr; uint8_t *d_flags = nullptr;
    uint32_t *d_pend = nullptr; size_t cap_pend = 0;    // pending lists of the CTAs (octets that need a search)
    unsigned *d_barriers = nullptr; int cap_barriers = 0;
    int persist_resident[2] = {0, 0};   // co-resident CTAs of icp_persist_kernel<EST> on this device
    bool brute_attr_done = false;       // dynamic shared-memory opt-in of nn_brute_tma_kernel done on this ctx's device
    // pose gather (gather.cu): persistent send / receive buffers on the device and a page-locked landing buffer
    s3d_result *d_gather_send = nullptr, *d_gather_recv = nullptr, *h_gather = nullptr;
    size_t cap_gather_send = 0, cap_gather_recv = 0, cap_gather_host = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    s3d_timing timing = {0, 0, 0, 0};
    // stream of single-pair registrations (s3d_register_enqueue / s3d_register_drain): per outstanding pair a page-locked
    // staging slot for its descriptor + initial state, three events, and a result record formed on the device
    PairDesc *h_desc_use = nullptr; PairState *h_state_use = nullptr; cudaEvent_t *ev_use = nullptr;   // where s3d_register_issue stages / records (null: h_desc, h_state, ev)
    PairDesc *h_desc_ring = nullptr; PairState *h_state_ring = nullptr;
    s3d_result *d_async = nullptr, *h_async = nullptr;
    cudaEvent_t ev_ring[S3D_ASYNC_DEPTH][3] = {};
    bool async_built[S3D_ASYNC_DEPTH] = {}; int async_launches[S3D_ASYNC_DEPTH] = {}; int async_total_launches[S3D_ASYNC_DEPTH] = {};
    int async_n = 0;
    void *h_planes_ring = nullptr; int planes_async_n = 0;      // s3d_segment_planes_enqueue: page-locked landing slots of the device loop's final state
    cudaEvent_t ev_plane[2] = {nullptr, nullptr};
    cudaEvent_t ev_eval[2 * S3D_MAX_PLANES] = {};       // around the evaluation pass of each RANSAC round
    s3d_plane_timing plane_timing = {0, 0, 0, 0, 0, 0};
    // plane segmentation scratch
    size_t cap_seg = 0;
    void *d_seg = nullptr;
    void *h_pinned = nullptr; size_t cap_pinned = 0;
    std::map<uint64_t, cudaGraphExec_t> graphs;     // index-build launch sequences (grid.cu), keyed by the buffers they touch
    std::map<uint64_t, int> graph_nodes;
    // caching device allocator: clouds and search indices come and go per frame, cudaMalloc/cudaFree of their
    // 10..40 MB buffers costs milliseconds; freed blocks are kept and handed out again (same stream => ordered)
    std::map<void *, size_t> pool_live;
    std::multimap<size_t, void *> pool_free;
    size_t pool_cached = 0;
    size_t pool_live_bytes = 0, pool_peak_bytes = 0;   // handed out right now / high-water mark (s3d_memory_stats)
};

int s3d_fail(s3d_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess);
#define S3D_CUDA(ctx, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return s3d_fail(ctx, S3D_E_CUDA, #call, e__); } while (0)
#define S3D_LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return s3d_fail(ctx, S3D_E_CUDA, "kernel launch", e__); } while (0)

// device memory pool (cloud.cu)
cudaError_t s3d_dev_alloc(s3d_ctx *ctx, void **out, size_t bytes);
void s3d_dev_free(s3d_ctx *ctx, void *p);
void s3d_dev_pool_release(s3d_ctx *ctx);
template <typename T> static inline cudaError_t s3d_dev_alloc_t(s3d_ctx *ctx, T **out, size_t bytes) { return s3d_dev_alloc(ctx, reinterpret_cast<void **>(out), bytes); }
// cloud.cu: orders the ctx stream behind a cloud's asynchronous upload and packs its rows (no-op for every other cloud)
int s3d_cloud_ready(s3d_ctx *ctx, const s3d_cloud *cloud);
// cloud.cu: makes cloud->d_absmax valid on the ctx stream (a small reduction kernel the first time, nothing afterwards)
int s3d_cloud_absmax(s3d_ctx *ctx, const s3d_cloud *cloud, bool want_normals);
// icp.cu: the two halves of s3d_register_batch (enqueue everything / read the event times after a synchronisation),
// the host-side completion of a record (norm, failure convention) and the device-side packing of the records
int s3d_register_issue(s3d_ctx *ctx, const s3d_cloud *const *src, const s3d_cloud *const *tgt, const double *guess, int n_pairs,
                       const s3d_icp_params *prm, bool *built_out, int *iter_launches_out);
void s3d_register_timing(s3d_ctx *ctx, bool built, int iter_launches);
void s3d_result_finish(s3d_result *r);
int s3d_result_pack(s3d_ctx *ctx, int n_pairs, s3d_result *d_out, int n_slots);
// grid.cu
int s3d_grid_build(s3d_ctx *ctx, s3d_cloud *cloud, float cell);
void s3d_grid_free(s3d_ctx *ctx, GridIndex &g);
// pinned staging
void *s3d_pinned(s3d_ctx *ctx, size_t bytes);
