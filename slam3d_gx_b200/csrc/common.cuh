// common.cuh -- device-side arithmetic shared by every kernel of libslam3d_b200.
//
// The float32 expressions here are, operation for operation, the ones written down in
// oracle/oracle_common.h (the CPU restatement used by the parity tests).  They use the explicit
// round-to-nearest intrinsics so nvcc cannot contract or reassociate them: correspondence indices,
// inlier counts and labels are bit-identical between the CUDA path and the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define S3D_NACC 32          // accumulator slots per pair (29 used)
#define S3D_ACC_SUMD2 27
#define S3D_ACC_COUNT 28

struct Pose12f { float m[12]; };

__device__ __forceinline__ float3 s3d_xform(const float *__restrict__ T, float x, float y, float z)
{
    float3 o;
    o.x = __fmaf_rn(T[2], z, __fmaf_rn(T[1], y, __fmaf_rn(T[0], x, T[3])));
    o.y = __fmaf_rn(T[6], z, __fmaf_rn(T[5], y, __fmaf_rn(T[4], x, T[7])));
    o.z = __fmaf_rn(T[10], z, __fmaf_rn(T[9], y, __fmaf_rn(T[8], x, T[11])));
    return o;
}

__device__ __forceinline__ float s3d_dist2(float px, float py, float pz, float qx, float qy, float qz)
{
    float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ float s3d_plane_eval(float a, float b, float c, float d, float x, float y, float z)
{
    return __fmaf_rn(c, z, __fmaf_rn(b, y, __fmaf_rn(a, x, d)));
}

__host__ __device__ __forceinline__ uint64_t s3d_mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t s3d_rand(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    return s3d_mix(s3d_mix(s3d_mix(seed + a) + b) + c);
}

// three distinct indices in [0,n), n >= 3 (same stream as orc_sample3)
__device__ __forceinline__ void s3d_sample3(uint64_t seed, uint64_t a, uint64_t b, uint32_t n, uint32_t *out)
{
    uint32_t got = 0;
    for (uint32_t c = 0; got < 3 && c < 64; ++c) {
        uint32_t v = (uint32_t)(s3d_rand(seed, a, b, c) % n);
        bool dup = false;
        for (uint32_t k = 0; k < got; ++k) dup |= (out[k] == v);
        if (!dup) out[got++] = v;
    }
    while (got < 3) { out[got] = (out[got - 1] + 1) % n; ++got; }
}

// plane through three points; false when degenerate
__device__ __forceinline__ bool s3d_plane_from3(float3 p0, float3 p1, float3 p2, float4 &coef)
{
    float ax = __fsub_rn(p1.x, p0.x), ay = __fsub_rn(p1.y, p0.y), az = __fsub_rn(p1.z, p0.z);
    float bx = __fsub_rn(p2.x, p0.x), by = __fsub_rn(p2.y, p0.y), bz = __fsub_rn(p2.z, p0.z);
    float nx = __fmaf_rn(ay, bz, -__fmul_rn(az, by));
    float ny = __fmaf_rn(az, bx, -__fmul_rn(ax, bz));
    float nz = __fmaf_rn(ax, by, -__fmul_rn(ay, bx));
    float l2 = __fmaf_rn(nz, nz, __fmaf_rn(ny, ny, __fmul_rn(nx, nx)));
    if (!(l2 > 1e-20f)) return false;
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(l2));
    nx = __fmul_rn(nx, inv); ny = __fmul_rn(ny, inv); nz = __fmul_rn(nz, inv);
    coef.x = nx; coef.y = ny; coef.z = nz;
    coef.w = -__fmaf_rn(nz, p0.z, __fmaf_rn(ny, p0.y, __fmul_rn(nx, p0.x)));
    return true;
}

// cyclic Jacobi on a symmetric 3x3 in double, fixed 12 sweeps (same control flow as orc_jacobi3)
__device__ inline void s3d_jacobi3(double A[3][3], double V[3][3], double w[3])
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 12; ++sweep) {
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            double apq = A[p][q];
            if (fabs(apq) < 1e-300) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; ++k) {
                double akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {
                double apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; ++k) {
                double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
            }
        }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

__device__ __forceinline__ float warp_sum(float v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// order-preserving float <-> uint mapping for atomicMin/Max on floats
__device__ __forceinline__ uint32_t f2ord(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u)
{
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}
