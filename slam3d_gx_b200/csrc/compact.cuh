// compact.cuh -- order-preserving stream compaction (count / scan / write), used by the depth
// back-projection (holes skipped in row-major order, reference src/convert2PCD.cpp:58-63) and by
// the plane segmentation (pcl::ExtractIndices negative, reference src/GraphicEnd.cpp:419-420).
#pragma once
#include "common.cuh"

#define S3D_COMPACT_BLOCK 1024

// Pred: __device__ bool operator()(int i) const
template <typename Pred>
__global__ void __launch_bounds__(S3D_COMPACT_BLOCK) compact_count_kernel(int n, Pred pred, uint32_t *__restrict__ block_counts)
{
    __shared__ int warp_cnt[32];
    int i = blockIdx.x * S3D_COMPACT_BLOCK + threadIdx.x;
    bool keep = (i < n) && pred(i);
    unsigned b = __ballot_sync(0xffffffffu, keep);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_sum_i(warp_cnt[threadIdx.x]);
        if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)v;
    }
}

// exclusive scan of up to any number of block counts by one block; total written to *total
static __global__ void __launch_bounds__(1024) compact_scan_kernel(uint32_t *__restrict__ counts, int nblocks, uint32_t *__restrict__ total)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = i < nblocks ? counts[i] : 0u;
        uint32_t incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_tot[threadIdx.x], wi = w;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warp_tot[threadIdx.x] = wi - w; // exclusive
        }
        __syncthreads();
        uint32_t excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
        if (i < nblocks) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// Emit: __device__ void operator()(int i, uint32_t out_pos) const   (called for kept elements)
// Drop: __device__ void operator()(int i) const                      (called for dropped elements)
template <typename Pred, typename Emit, typename Drop>
__global__ void __launch_bounds__(S3D_COMPACT_BLOCK) compact_write_kernel(int n, Pred pred, Emit emit, Drop drop,
                                                                          const uint32_t *__restrict__ block_offsets)
{
    __shared__ int warp_cnt[32];
    int i = blockIdx.x * S3D_COMPACT_BLOCK + threadIdx.x;
    bool in = i < n;
    bool keep = in && pred(i);
    unsigned b = __ballot_sync(0xffffffffu, keep);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
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
    if (keep) emit(i, block_offsets[blockIdx.x] + (uint32_t)warp_cnt[w] + (uint32_t)__popc(b & ((1u << lane) - 1u)));
    else if (in) drop(i);
}
