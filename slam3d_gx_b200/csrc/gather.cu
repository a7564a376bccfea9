// gather.cu -- multi-GPU pose gather: one ncclAllGather of fixed-size result records.
//
// The reference is a single process; frame pairs shard across GPUs with no data-path
// exchange, and the only collective is this latency-bound gather of n_pairs * sizeof(s3d_result)
// bytes.  NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on a
// particular libnccl: under torchrun the already-loaded torch-bundled NCCL is found first.
#include <dlfcn.h>
#include <cstring>
#include "context.h"

typedef int (*nccl_allgather_fn)(const void *, void *, size_t, int, void *, cudaStream_t);

static nccl_allgather_fn resolve_allgather()
{
    static nccl_allgather_fn fn = nullptr;
    static bool tried = false;
    if (tried) return fn;
    tried = true;
    void *sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (!sym) {
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !sym; ++i) {
            void *h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllGather");
        }
    }
    fn = (nccl_allgather_fn)sym;
    return fn;
}

extern "C" int s3d_gather_results(s3d_ctx *ctx, void *nccl_comm, const s3d_result *local, int n_local, int world, s3d_result *all_out)
{
    if (!ctx || !nccl_comm || !local || !all_out || n_local <= 0 || world <= 0) return s3d_fail(ctx, S3D_E_ARG, "s3d_gather_results: bad argument");
    nccl_allgather_fn ag = resolve_allgather();
    if (!ag) return s3d_fail(ctx, S3D_E_NCCL, "ncclAllGather not found (libnccl.so.2 not loadable)");
    cudaSetDevice(ctx->device);
    size_t bytes = sizeof(s3d_result) * (size_t)n_local;
    char *d_send = nullptr, *d_recv = nullptr;
    S3D_CUDA(ctx, cudaMalloc(&d_send, bytes));
    S3D_CUDA(ctx, cudaMalloc(&d_recv, bytes * world));
    S3D_CUDA(ctx, cudaMemcpyAsync(d_send, local, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = ag(d_send, d_recv, bytes, /*ncclChar*/ 0, nccl_comm, ctx->stream);
    if (rc != 0) { cudaFree(d_send); cudaFree(d_recv); return s3d_fail(ctx, S3D_E_NCCL, "ncclAllGather failed"); }
    S3D_CUDA(ctx, cudaMemcpyAsync(all_out, d_recv, bytes * world, cudaMemcpyDeviceToHost, ctx->stream));
    S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_send); cudaFree(d_recv);
    return S3D_OK;
}
