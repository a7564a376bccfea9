/*
 * slam3d_b200.h -- C ABI of the B200-native registration path (libslam3d_b200.so).
 *
 * This is the drop-in boundary for the frame-to-frame registration seat of gaoxiang12/slam3d_gx.
 * The reference has no C ABI; its only plug-in mechanism is the C++ virtual override of
 * GraphicEnd (reference src/GraphicEnd.h:80-84,134, demonstrated by GraphicEnd2 at :262-275).
 * Every entry point below names the reference interface it replaces.  The C++ shell that sits
 * on top (slam3d_gx_b200/host/GraphicEnd.h) keeps the reference's class/method names.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; no exceptions cross the boundary.
 *   - every function returns S3D_OK (0) or a negative S3D_E_* code; s3d_last_error(ctx) gives text.
 *     *Algorithmic* failure of one pair (too few correspondences, rank-deficient system) is not an
 *     API error: it is reported in s3d_result.status and T is set to exactly identity, which is the
 *     reference's failure convention (src/GraphicEnd.cpp:585-600,621-624; callers compare
 *     T == Identity at :173,703,739,816,894).
 *   - T maps source (frame-1) coordinates into target (frame-2) coordinates, X2 = R*X1 + t, like
 *     RESULT_OF_MULTIPNP::T (src/GraphicEnd.h:59-69); callers keep their .inverse() (:170,709,745).
 *   - one ctx per host thread / GPU; calls on a ctx are serialised by the caller; work is issued on
 *     the ctx stream and the call returns after the results are valid in the caller's buffers.
 *   - there is NO CPU fallback: without a CUDA device s3d_create fails with S3D_E_CUDA.
 */
#ifndef SLAM3D_B200_H
#define SLAM3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_ABI_VERSION 1

enum {
    S3D_OK = 0,
    S3D_E_ARG = -1,      /* bad argument */
    S3D_E_CUDA = -2,     /* CUDA runtime error (see s3d_last_error) */
    S3D_E_STATE = -3,    /* object not in the required state (e.g. target has no normals) */
    S3D_E_NCCL = -4      /* NCCL error / library not loadable */
};

/* per-pair algorithmic status (s3d_result.status) */
enum {
    S3D_PAIR_OK = 0,
    S3D_PAIR_FEW_CORRESPONDENCES = 1, /* PCL: "Not enough correspondences found" (< min_correspondences) */
    S3D_PAIR_DEGENERATE = 2,          /* normal equations rank deficient (planar sliding) */
    S3D_PAIR_NONFINITE = 3,
    S3D_PAIR_ABSENT = -1              /* padding slot of a gathered shard (s3d_register_batch_gather): no pair here */
};

enum { S3D_ESTIMATOR_POINT_TO_PLANE = 0,   /* PCL TransformationEstimationPointToPlaneLLS */
       S3D_ESTIMATOR_SVD = 1 };            /* PCL TransformationEstimationSVD (Kabsch/Umeyama) */
enum { S3D_SEARCH_GRID = 0,                /* exact NN on a device-built uniform grid, shared-memory tiles, one
                                              persistent launch for all iterations (default) */
       S3D_SEARCH_BRUTE = 1,               /* exact NN, brute force over TMA-staged target tiles */
       S3D_SEARCH_GRID_LANE = 2 };         /* exact NN on the grid, per-lane ball search, one launch per iteration */

typedef struct s3d_ctx s3d_ctx;
typedef struct s3d_cloud s3d_cloud;

/* ---- parameters / results ------------------------------------------------------------------ */

/* ICP parameters: PCL-1.7 IterativeClosestPoint defaults in comments. */
typedef struct s3d_icp_params {
    int32_t max_iterations;       /* 10 */
    float   max_corr_dist;        /* <= 0: unlimited (PCL: sqrt(DBL_MAX)) */
    int32_t estimator;            /* S3D_ESTIMATOR_* */
    int32_t search;               /* S3D_SEARCH_* */
    float   grid_cell;            /* grid cell edge in metres; <= 0: automatic */
    int32_t min_correspondences;  /* 3 */
    double  pivot_eps;            /* relative Cholesky pivot below which the system is degenerate; <=0: 1e-9 */
    int32_t reuse_index;          /* 1: keep the target's grid cached in the cloud handle; 0: rebuild per call */
    int32_t reserved;
} s3d_icp_params;

/* Result record: what RESULT_OF_MULTIPNP{T,norm,inliers} (src/GraphicEnd.h:59-69) carries, plus
 * diagnostics.  160 bytes, POD, the unit gathered across GPUs. */
typedef struct s3d_result {
    double  T[16];        /* row-major 4x4, source -> target */
    double  norm;         /* |min(theta,2pi-theta)| + 0.9*|t|   (src/GraphicEnd.cpp:618) */
    double  fitness;      /* mean squared NN distance of the accepted correspondences, last iteration */
    int32_t inliers;      /* accepted correspondences in the last iteration */
    int32_t iterations;   /* iterations executed */
    int32_t status;       /* S3D_PAIR_* ; != 0 => T is exactly identity */
    int32_t reserved;
} s3d_result;

/* Plane segmentation parameters: src/GraphicEnd.cpp:360-364 + parameters.yaml:41-47 and the
 * PCL-1.7 SACSegmentation defaults. */
typedef struct s3d_plane_params {
    float    distance_threshold; /* 0.08  (parameters.yaml distance_threshold) */
    float    plane_percent;      /* 0.2   loop while remaining > percent*n (GraphicEnd.cpp:372) */
    int32_t  max_planes;         /* 3     (GraphicEnd.cpp:424) */
    int32_t  max_iterations;     /* 50    PCL SACSegmentation max_iterations_ */
    float    probability;        /* 0.99  PCL probability_ */
    int32_t  reserved;           /* bit 0: time every evaluation pass with its own CUDA events (s3d_last_plane_timing.eval_ms);
                                    otherwise the launches are replayed from a CUDA graph and only total_ms is measured */
    uint64_t seed;               /* hypothesis stream seed (PCL: mt19937 seeded 12345) */
} s3d_plane_params;

typedef struct s3d_plane {
    float   coef[4];   /* a,b,c,d with d >= 0 (GraphicEnd.cpp:383-387) */
    int32_t inliers;   /* points labelled with this plane */
    int32_t hypotheses;/* RANSAC iterations consumed (adaptive stop) */
} s3d_plane;

#define S3D_MAX_PLANES 16

typedef struct s3d_camera {
    double fx, fy, cx, cy, factor; /* camera_fx.. camera_factor (parameters.yaml:82-86, ParameterReader.cpp:9) */
} s3d_camera;

/* timing of the last s3d_register_batch, measured with CUDA events on the ctx stream */
typedef struct s3d_timing {
    float   index_ms;      /* target grid build (0 when cached) */
    float   iterate_ms;    /* all iteration launches of the batch */
    int32_t iter_launches; /* kernel launches inside iterate_ms */
    int32_t total_launches;/* every kernel this library launched during the call */
} s3d_timing;

/* timing of the last s3d_segment_planes (CUDA events on the ctx stream) and what its evaluation passes read:
 * points_scanned * 16 bytes is the algorithmic traffic of the RANSAC scan (16 N bytes per pass) */
typedef struct s3d_plane_timing {
    float   total_ms;               /* every kernel of the call */
    float   eval_ms;                /* the candidate-evaluation passes only (plane_eval_kernel) */
    int32_t rounds;                 /* RANSAC rounds that ran (planes found, + 1 when the last round found none) */
    int32_t eval_passes_per_round;  /* passes over the remaining points per round: 1 for <= 64 candidates */
    int64_t points_scanned;         /* sum over the rounds of the points one evaluation pass read */
    int64_t reserved;
} s3d_plane_timing;

/* ---- context ------------------------------------------------------------------------------- */

int  s3d_abi_version(void);
/* create/destroy the per-GPU context (replaces nothing in the reference: GraphicEnd owns no device). */
int  s3d_create(s3d_ctx **out, int device_id);
void s3d_destroy(s3d_ctx *ctx);
const char *s3d_last_error(const s3d_ctx *ctx);
/* run on a caller-owned cudaStream_t (pass 0 to go back to the ctx's own stream). */
int  s3d_set_stream(s3d_ctx *ctx, void *cuda_stream);
int  s3d_device_sm_count(const s3d_ctx *ctx);
/* number of kernels launched by this ctx since creation (evidence that the GPU path ran). */
int64_t s3d_launch_count(const s3d_ctx *ctx);

/* ---- clouds (replace pcl::PointCloud<PointXYZRGBA>::Ptr members _currCloud/_lastCloud,
 *      src/GraphicEnd.h:181-183, and the PCD load at src/GraphicEnd.cpp:279-281) ----------------- */

/* host xyz, stride in floats between points (>=3; 4 for PCD "x y z rgba" rows) */
int  s3d_cloud_upload(s3d_ctx *ctx, const float *xyz, int stride_floats, int n, s3d_cloud **out);
/* The same upload without waiting for it: the copy runs on a second stream of the ctx, so the copy engine moves the
 * NEXT frame while the SMs register the current one (the reference loads frame k+1 only after frame k is done,
 * src/GraphicEnd.cpp:266-281 from run():150-160).  xyz must stay valid and unchanged until the cloud has been used by
 * another call of this ABI or s3d_cloud_wait returned; page-locked memory is needed for a truly asynchronous copy.
 * Every other entry point orders itself behind the upload; nothing else changes for the caller. */
int  s3d_cloud_upload_async(s3d_ctx *ctx, const float *xyz, int stride_floats, int n, s3d_cloud **out);
/* block the host until the cloud's asynchronous upload has finished (no-op for other clouds) */
int  s3d_cloud_wait(s3d_ctx *ctx, const s3d_cloud *cloud);
/* page-locked host memory for the buffers handed to s3d_cloud_upload_async (a C/C++ host needs no CUDA headers) */
int  s3d_host_alloc(s3d_ctx *ctx, size_t bytes, void **out);
void s3d_host_free(s3d_ctx *ctx, void *p);
/* device-resident float4 array (x,y,z,ignored); copied device-to-device */
int  s3d_cloud_from_device(s3d_ctx *ctx, const void *d_xyzw, int n, s3d_cloud **out);
/* depth image -> cloud: non-zero pixels, row-major, x=(u-cx)z/fx, y=(v-cy)z/fy, z=d/factor
 * (src/convert2PCD.cpp:54-80); z_max > 0 additionally applies the PassThrough z in [0,z_max]
 * of src/GraphicEnd.cpp:283-285. depth is a HOST pointer. */
int  s3d_cloud_from_depth(s3d_ctx *ctx, const uint16_t *depth, int width, int height,
                          const s3d_camera *cam, float z_max, s3d_cloud **out);
/* depth image -> cloud WITH per-point normals from the organised image (scenes that are not a few big planes): the
 * normal of a pixel is the normalised cross product of the central differences of the back-projected neighbours
 * `step` pixels away, turned towards the camera; no normal (valid = 0) at borders, holes and where a neighbour is
 * more than max_jump metres away in depth.  Same point set and order as s3d_cloud_from_depth
 * (src/convert2PCD.cpp:54-80); the normals play the role of PLANE::coff (src/GraphicEnd.h:41-49) for the ICP. */
int  s3d_cloud_from_depth_normals(s3d_ctx *ctx, const uint16_t *depth, int width, int height, const s3d_camera *cam,
                                  float z_max, int step, float max_jump, s3d_cloud **out);
/* per-point normals from the caller (host, stride>=3); valid flag set for finite non-zero normals */
int  s3d_cloud_set_normals(s3d_ctx *ctx, s3d_cloud *cloud, const float *nrm, int stride_floats, int n);
/* same with a device float4 array (nx,ny,nz,valid!=0) */
int  s3d_cloud_set_normals_device(s3d_ctx *ctx, s3d_cloud *cloud, const void *d_nrm, int n);
int  s3d_cloud_size(const s3d_cloud *cloud);
int  s3d_cloud_has_normals(const s3d_cloud *cloud);
/* any of the outputs may be NULL; xyz: n*3 floats, normals: n*3 floats, labels: n int32 (-1 = none) */
int  s3d_cloud_download(s3d_ctx *ctx, const s3d_cloud *cloud, float *xyz, float *normals, int32_t *labels);
/* drop the cached search index of a cloud and return its buffers to the ctx pool (it is rebuilt on next use as a target) */
int  s3d_cloud_drop_index(s3d_ctx *ctx, s3d_cloud *cloud);
/* device memory of the ctx's clouds and indices: bytes handed out now, their high-water mark, bytes cached for reuse
 * (any pointer may be NULL).  A long run keeps only its key frames resident (reference: _keyframes, src/GraphicEnd.h:150). */
int  s3d_memory_stats(const s3d_ctx *ctx, size_t *live_bytes, size_t *peak_live_bytes, size_t *cached_bytes);
void s3d_cloud_free(s3d_ctx *ctx, s3d_cloud *cloud);
/* s3d_cloud_free waits for the context's stream first.  s3d_cloud_release does not: the buffers go back to the context's pool, which
 * is ordered by the context's stream, so work already enqueued on the cloud (s3d_register_enqueue, s3d_segment_planes_enqueue)
 * still runs on intact data and whoever is handed the buffers next is enqueued behind it. */
void s3d_cloud_release(s3d_ctx *ctx, s3d_cloud *cloud);

/* ---- filters and map fusion (the steps either side of the registration path) ------------------ */
/* pcl::PassThrough on "z": keeps finite points with z_min <= z <= z_max, order preserved
 * (reference src/GraphicEnd.cpp:283-285,291-292; src/saveOutput.cpp:40-46,81-84). */
int  s3d_cloud_passthrough_z(s3d_ctx *ctx, const s3d_cloud *cloud, float z_min, float z_max, s3d_cloud **out);
/* pcl::VoxelGrid with a cubic leaf: one centroid per occupied voxel, ascending voxel index
 * (reference src/GraphicEnd.cpp:287-295 "grid_leaf"; src/saveOutput.cpp:44-46,76-79,90-93).
 * S3D_E_ARG when the voxel indices would overflow 32 bits (PCL refuses the same input). */
int  s3d_cloud_voxel_grid(s3d_ctx *ctx, const s3d_cloud *cloud, float leaf, s3d_cloud **out);
/* pcl::transformPointCloud: out = T * cloud, T row-major 4x4 (reference src/saveOutput.cpp:87). */
int  s3d_cloud_transform(s3d_ctx *ctx, const s3d_cloud *cloud, const double *T16, s3d_cloud **out);
/* PointCloud::operator+= over a list (reference src/saveOutput.cpp:88). */
int  s3d_cloud_concat(s3d_ctx *ctx, const s3d_cloud *const *clouds, int n_clouds, s3d_cloud **out);
/* The key-frame fusion loop of saveOutput (reference src/saveOutput.cpp:47-95): per key frame voxel grid,
 * z pass-through [0, z_max], transform by its pose (poses16: n_clouds row-major 4x4), append; voxel grid of the sum. */
int  s3d_map_fuse(s3d_ctx *ctx, const s3d_cloud *const *clouds, const double *poses16, int n_clouds, float leaf, float z_max,
                  s3d_cloud **out);

/* ---- plane extraction (replaces GraphicEnd::extractPlanesAndGenerateImage, src/GraphicEnd.cpp:353-430,
 *      i.e. pcl::SACSegmentation + ExtractIndices; image painting is not on the path) ------------- */
/* Writes per-point plane label and plane normal into the cloud (used as target normals by ICP). */
int  s3d_segment_planes(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *params,
                        s3d_plane *planes_out /*[max_planes]*/, int *n_planes_out);

/* ---- registration (replaces GraphicEnd::multiPnP, src/GraphicEnd.cpp:557-659, and the sequential
 *      candidate loops of loopClosure()/lostRecovery(), :694-761, :810-836) ----------------------- */
/* guess: n_pairs*16 doubles (row-major 4x4) or NULL for identity. src[i]/tgt[i] may repeat
 * (loop-closure sweep: one shared target). */
int  s3d_register_batch(s3d_ctx *ctx, const s3d_cloud *const *src, const s3d_cloud *const *tgt,
                        const double *guess, int n_pairs, const s3d_icp_params *params,
                        s3d_result *results_out);
/* convenience single pair == RESULT_OF_MULTIPNP multiPnP(plane1, plane2, ...) */
int  s3d_register_pair(s3d_ctx *ctx, const s3d_cloud *src, const s3d_cloud *tgt, const double *guess,
                       const s3d_icp_params *params, s3d_result *result_out);
/* correspondences of the last iteration of the last single-pair call (n_src int32, -1 = rejected);
 * debugging / parity aid */
int  s3d_last_correspondences(s3d_ctx *ctx, int32_t *idx_out, int n);
int  s3d_last_timing(const s3d_ctx *ctx, s3d_timing *out);

/* A stream of independent single-pair registrations without a host round trip per pair -- the loop over loop-closure
 * candidates of reference src/GraphicEnd.cpp:729-761 when its results are only needed after the loop.  s3d_register_enqueue
 * issues one registration on the context's stream and forms its result record on the device; s3d_register_drain waits for
 * everything enqueued since the last drain and returns the records (and, when timing_out is not NULL, the device times of
 * each) in enqueue order.  At most S3D_ASYNC_DEPTH pairs may be outstanding (S3D_E_STATE
 * beyond that: drain first).  The clouds must stay unmodified until the drain.  Results are bit-identical to
 * s3d_register_pair.  Every record is copied to page-locked host memory right behind its pair (160 bytes); the drain only waits.
 * A cloud may be released (s3d_cloud_release, not s3d_cloud_free, which waits) once the work that uses it is enqueued.  s3d_segment_planes_enqueue / s3d_segment_planes_drain do the same for the plane extraction (labels and
 * normals are written on the device in stream order, so a registration enqueued behind it sees them); planes_out holds
 * S3D_MAX_PLANES planes per extraction, n_planes_out the number found by each. */
#define S3D_ASYNC_DEPTH 64
int  s3d_register_enqueue(s3d_ctx *ctx, const s3d_cloud *src, const s3d_cloud *tgt, const double *guess /* 16 doubles or NULL */,
                          const s3d_icp_params *params);
int  s3d_register_drain(s3d_ctx *ctx, s3d_result *results_out, s3d_timing *timing_out /* may be NULL */, int capacity, int *n_out);
/* How s3d_register_batch spreads n_pairs pairs (largest source: n_points_max points) over resident_ctas co-resident CTAs of the
 * registration kernel (one per SM: 148 on a B200): *groups_out pairs are in flight at a time, each on *group_ctas_out CTAs; the
 * groups walk the batch in rounds.  Pure host arithmetic (no device needed): lets a host size its batches. */
int  s3d_batch_shape(int n_pairs, int n_points_max, int resident_ctas, int *groups_out, int *group_ctas_out);
int  s3d_segment_planes_enqueue(s3d_ctx *ctx, s3d_cloud *cloud, const s3d_plane_params *params);
int  s3d_segment_planes_drain(s3d_ctx *ctx, s3d_plane *planes_out, int *n_planes_out, int capacity, int *n_out);
void s3d_icp_params_default(s3d_icp_params *p);
void s3d_plane_params_default(s3d_plane_params *p);
int  s3d_last_plane_timing(const s3d_ctx *ctx, s3d_plane_timing *out);

/* ---- keypoint planarity (replaces isPlanar, src/planarFeatures.cpp:88-136) ------------------- */
/* depth: HOST uint16 image; uv: n pairs (u,v) already truncated to int like :90-91;
 * flags_out[i] = 1 iff the 7x7 patch has no zero depth and RANSAC(plane, threshold) finds
 * more than min_inliers inliers (reference: threshold 0.01, min_inliers 40). */
int  s3d_planar_keypoints(s3d_ctx *ctx, const uint16_t *depth, int width, int height,
                          const s3d_camera *cam, const int32_t *uv, int n,
                          float threshold, int min_inliers, uint64_t seed, uint8_t *flags_out);

/* ---- multi-GPU: pairs sharded over ranks, one pose gather (no reference counterpart: the reference is a single
 *      process; the independence that is sharded is the candidate loop of src/GraphicEnd.cpp:729-761) ------------- */
/* NCCL communicator plumbing for hosts that have none (a C++ host needs no NCCL headers; NCCL is resolved with dlopen).
 * s3d_comm_unique_id: 128 bytes (ncclUniqueId) created on one rank and handed to the others by the caller's own means
 * (file, socket, MPI, torch.distributed ...); s3d_comm_create: ncclCommInitRank on the ctx's device. */
#define S3D_COMM_ID_BYTES 128
int  s3d_comm_unique_id(void *id_out /*[S3D_COMM_ID_BYTES]*/);
int  s3d_comm_create(s3d_ctx *ctx, const void *id /*[S3D_COMM_ID_BYTES]*/, int world, int rank, void **comm_out);
void s3d_comm_destroy(s3d_ctx *ctx, void *comm);
/* All-gather n_local result records from every rank into all_out (world*n_local records) with
 * ncclAllGather on the ctx stream. nccl_comm is an ncclComm_t (s3d_comm_create or the caller's own). */
int  s3d_gather_results(s3d_ctx *ctx, void *nccl_comm, const s3d_result *local, int n_local,
                        int world, s3d_result *all_out);
/* One rank's shard of a sharded batch, registered and gathered in one call: the n_local pairs go through the same
 * device path as s3d_register_batch, the result records are formed ON THE DEVICE in the all-gather send buffer,
 * ncclAllGather runs on the same stream right behind the last iteration, and the only device-to-host copy is the
 * gathered world*n_slot records.  n_slot >= n_local is the (rank-uniform) number of record slots per rank; slots
 * beyond a rank's n_local come back with status S3D_PAIR_ABSENT.  all_out[r*n_slot + k] = pair k of rank r. */
int  s3d_register_batch_gather(s3d_ctx *ctx, void *nccl_comm, const s3d_cloud *const *src, const s3d_cloud *const *tgt,
                               const double *guess, int n_local, int n_slot, const s3d_icp_params *params,
                               int world, s3d_result *all_out);

#ifdef __cplusplus
}
#endif
#endif /* SLAM3D_B200_H */
