// gather.cu -- multi-GPU pose gather: one ncclAllGather of fixed-size result records.
//
// The reference is a single process; frame pairs shard across GPUs with no data-path
// exchange (the independent candidate loop of reference src/GraphicEnd.cpp:729-761), and the only
// collective is this latency-bound gather of n_pairs * sizeof(s3d_result) bytes.  NCCL is resolved at
// run time (dlopen) so that the library has no link-time dependency on a particular libnccl: under
// torchrun the already-loaded torch-bundled NCCL is found first.
//
// s3d_register_batch_gather is the config-4 step of one rank: the records are formed on the device in
// the send buffer (result_pack_kernel, icp.cu), the all-gather is enqueued on the ctx stream right
// behind the last iteration, and one device-to-host copy brings every rank's records back.  Send,
// receive and landing buffers live in the ctx (no allocation per call).
#include <dlfcn.h>
#include <cstring>
#include "context.h"

struct nccl_uid { char internal[S3D_COMM_ID_BYTES]; };
typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*nccl_getuid_fn)(nccl_uid *);
typedef int (*nccl_initrank_fn)(void **, int, nccl_uid, int);
typedef int (*nccl_destroy_fn)(void *);
typedef const char *(*nccl_errstr_fn)(int);

static void *nccl_symbol(const char *name)
{
    static void *lib = nullptr;
    void *sym = dlsym(RTLD_DEFAULT, name);
    if (sym) return sym;
    if (!lib) {
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    }
    return lib ? dlsym(lib, name) : nullptr;
}

static int nccl_fail(s3d_ctx *ctx, const char *what, int rc)
{
    static nccl_errstr_fn es = (nccl_errstr_fn)nccl_symbol("ncclGetErrorString");
    std::string msg = what;
    if (es) { msg += ": "; msg += es(rc); }
    return s3d_fail(ctx, S3D_E_NCCL, msg.c_str());
}

extern "C" int s3d_comm_unique_id(void *id_out)
{
    if (!id_out) return S3D_E_ARG;
    nccl_getuid_fn fn = (nccl_getuid_fn)nccl_symbol("ncclGetUniqueId");
    if (!fn) return S3D_E_NCCL;
    return fn((nccl_uid *)id_out) == 0 ? S3D_OK : S3D_E_NCCL;
}

extern "C" int s3d_comm_create(s3d_ctx *ctx, const void *id, int world, int rank, void **comm_out)
{
    if (!ctx || !id || !comm_out || world <= 0 || rank < 0 || rank >= world) return s3d_fail(ctx, S3D_E_ARG, "s3d_comm_create: bad argument");
    nccl_initrank_fn fn = (nccl_initrank_fn)nccl_symbol("ncclCommInitRank");
    if (!fn) return s3d_fail(ctx, S3D_E_NCCL, "ncclCommInitRank not found (libnccl.so.2 not loadable)");
    cudaSetDevice(ctx->device);
    nccl_uid uid;
    memcpy(&uid, id, sizeof(uid));
    *comm_out = nullptr;
    int rc = fn(comm_out, world, uid, rank);
    if (rc != 0) return nccl_fail(ctx, "ncclCommInitRank failed", rc);
    return S3D_OK;
}

extern "C" void s3d_comm_destroy(s3d_ctx *ctx, void *comm)
{
    if (!comm) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    nccl_destroy_fn fn = (nccl_destroy_fn)nccl_symbol("ncclCommDestroy");
    if (fn) fn(comm);
}

// grows the persistent gather buffers; on failure nothing is leaked (the ctx keeps whatever was allocated and frees it in s3d_destroy)
static int ensure_gather(s3d_ctx *ctx, size_t n_send, size_t n_recv)
{
    if (n_send > ctx->cap_gather_send) {
        cudaFree(ctx->d_gather_send); ctx->d_gather_send = nullptr; ctx->cap_gather_send = 0;
        const size_t cap = std::max<size_t>(n_send, 64);
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_gather_send, sizeof(s3d_result) * cap));
        ctx->cap_gather_send = cap;
    }
    if (n_recv > ctx->cap_gather_recv) {
        cudaFree(ctx->d_gather_recv); ctx->d_gather_recv = nullptr; ctx->cap_gather_recv = 0;
        const size_t cap = std::max<size_t>(n_recv, 512);
        S3D_CUDA(ctx, cudaMalloc(&ctx->d_gather_recv, sizeof(s3d_result) * cap));
        ctx->cap_gather_recv = cap;
    }
    if (n_recv > ctx->cap_gather_host) {
        if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
        ctx->h_gather = nullptr; ctx->cap_gather_host = 0;
        const size_t cap = std::max<size_t>(n_recv, 512);
        S3D_CUDA(ctx, cudaMallocHost(&ctx->h_gather, sizeof(s3d_result) * cap));
        ctx->cap_gather_host = cap;
    }
    return S3D_OK;
}

extern "C" int s3d_gather_results(s3d_ctx *ctx, void *nccl_comm, const s3d_result *local, int n_local, int world, s3d_result *all_out)
{
    if (!ctx || !nccl_comm || !local || !all_out || n_local <= 0 || world <= 0) return s3d_fail(ctx, S3D_E_ARG, "s3d_gather_results: bad argument");
    nccl_allgather_fn ag = (nccl_allgather_fn)nccl_symbol("ncclAllGather");
    if (!ag) return s3d_fail(ctx, S3D_E_NCCL, "ncclAllGather not found (libnccl.so.2 not loadable)");
    cudaSetDevice(ctx->device);
    const size_t bytes = sizeof(s3d_result) * (size_t)n_local;
    int rc = ensure_gather(ctx, (size_t)n_local, (size_t)n_local * world);
    if (rc) return rc;
    // the caller's records may be pageable: stage them in the page-locked landing buffer (it is free until the copy back)
    memcpy(ctx->h_gather, local, bytes);
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->d_gather_send, ctx->h_gather, bytes, cudaMemcpyHostToDevice, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int nrc = ag(ctx->d_gather_send, ctx->d_gather_recv, bytes, /*ncclChar*/ 0, nccl_comm, ctx->stream);
    if (nrc != 0) return nccl_fail(ctx, "ncclAllGather failed", nrc);
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_gather, ctx->d_gather_recv, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(all_out, ctx->h_gather, bytes * world);
    return S3D_OK;
}

extern "C" int s3d_register_batch_gather(s3d_ctx *ctx, void *nccl_comm, const s3d_cloud *const *src, const s3d_cloud *const *tgt,
                                         const double *guess, int n_local, int n_slot, const s3d_icp_params *prm, int world, s3d_result *all_out)
{
    if (!ctx || !nccl_comm || !all_out || n_local < 0 || n_slot < n_local || n_slot <= 0 || world <= 0)
        return s3d_fail(ctx, S3D_E_ARG, "s3d_register_batch_gather: bad argument");
    nccl_allgather_fn ag = (nccl_allgather_fn)nccl_symbol("ncclAllGather");
    if (!ag) return s3d_fail(ctx, S3D_E_NCCL, "ncclAllGather not found (libnccl.so.2 not loadable)");
    cudaSetDevice(ctx->device);
    int rc = ensure_gather(ctx, (size_t)n_slot, (size_t)n_slot * world);
    if (rc) return rc;
    bool built = false; int iter_launches = 0;
    if (n_local > 0) {
        rc = s3d_register_issue(ctx, src, tgt, guess, n_local, prm, &built, &iter_launches);
        if (rc) return rc;
    } else {
        cudaEventRecord(ctx->ev[0], ctx->stream); cudaEventRecord(ctx->ev[1], ctx->stream); cudaEventRecord(ctx->ev[2], ctx->stream);
    }
    rc = s3d_result_pack(ctx, n_local, ctx->d_gather_send, n_slot);
    if (rc) return rc;
    const size_t bytes = sizeof(s3d_result) * (size_t)n_slot;
    int nrc = ag(ctx->d_gather_send, ctx->d_gather_recv, bytes, /*ncclChar*/ 0, nccl_comm, ctx->stream);
    if (nrc != 0) return nccl_fail(ctx, "ncclAllGather failed", nrc);
    S3D_CUDA(ctx, cudaMemcpyAsync(ctx->h_gather, ctx->d_gather_recv, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < (size_t)n_slot * world; ++i) {
        all_out[i] = ctx->h_gather[i];
        if (all_out[i].status != S3D_PAIR_ABSENT) s3d_result_finish(&all_out[i]);
    }
    s3d_register_timing(ctx, built, iter_launches);
    return S3D_OK;
}
