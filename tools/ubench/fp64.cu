// micro-benchmark: FP64 / FP32 FMA throughput per SM on this GPU (build: nvcc -arch=sm_100a -O3 fp64.cu -o fp64)
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void fma_kernel(T *out, int iters)
{
    T a[8];
    for (int k = 0; k < 8; ++k) a[k] = (T)(threadIdx.x + k);
    T b = (T)1.000001, c = (T)0.5;
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = a[k] * b + c;
    }
    T s = 0;
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T> void run(const char *name)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    T *out; cudaMalloc(&out, sizeof(T) * sms * 4 * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    fma_kernel<T><<<sms * 4, 512>>>(out, 100);
    cudaEventRecord(e0);
    fma_kernel<T><<<sms * 4, 512>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)sms * 4 * 512 * iters * 8;
    printf("%s: %.3f ms, %.2f TFLOP/s, %.2f FMA/clk/SM at 1.965 GHz\n", name, ms, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.965e9);
}
int main() { run<float>("fp32"); run<double>("fp64"); return 0; }
