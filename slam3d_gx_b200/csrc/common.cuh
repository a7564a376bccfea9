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
#include <string.h>

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

// the same map with the pose held as three float4 rows (identical operations, identical bits)
__device__ __forceinline__ float3 s3d_xform4(const float4 r0, const float4 r1, const float4 r2, float x, float y, float z)
{
    float3 o;
    o.x = __fmaf_rn(r0.z, z, __fmaf_rn(r0.y, y, __fmaf_rn(r0.x, x, r0.w)));
    o.y = __fmaf_rn(r1.z, z, __fmaf_rn(r1.y, y, __fmaf_rn(r1.x, x, r1.w)));
    o.z = __fmaf_rn(r2.z, z, __fmaf_rn(r2.y, y, __fmaf_rn(r2.x, x, r2.w)));
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

// ------------------------------------------------------------------------------------------------
// Strict double: every operation is one correctly rounded IEEE operation, never contracted into an FMA and never
// reassociated, so the small dense solvers below perform exactly the operations of their restatement in
// oracle/icp_oracle.c / oracle/plane_oracle.c (compiled with -ffp-contract=off): poses and plane coefficients are
// bit-identical between the CUDA path and the oracle.
// ------------------------------------------------------------------------------------------------
struct sd {
    double v;
    __host__ __device__ sd() {}
    __host__ __device__ sd(double x) : v(x) {}
};
__device__ __forceinline__ sd operator+(sd a, sd b) { return sd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a, sd b) { return sd(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator*(sd a, sd b) { return sd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator/(sd a, sd b) { return sd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a) { return sd(-a.v); }
__device__ __forceinline__ bool operator<(sd a, sd b) { return a.v < b.v; }
__device__ __forceinline__ bool operator>(sd a, sd b) { return a.v > b.v; }
__device__ __forceinline__ bool operator>=(sd a, sd b) { return a.v >= b.v; }
__device__ __forceinline__ sd sd_sqrt(sd a) { return sd(__dsqrt_rn(a.v)); }
__device__ __forceinline__ sd sd_fabs(sd a) { return sd(fabs(a.v)); }

// cyclic Jacobi on a symmetric 3x3 in strict double, at most 12 sweeps with orc_jacobi3's convergence exit (same operations)
__device__ inline void s3d_jacobi3(sd A[3][3], sd V[3][3], sd w[3])
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = sd(i == j ? 1.0 : 0.0);
    for (int sweep = 0; sweep < 12; ++sweep) {
        for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
            sd apq = A[p][q];
            if (fabs(apq.v) < 1e-300) continue;
            sd theta = (A[q][q] - A[p][p]) / (sd(2.0) * apq);
            sd t = sd(theta.v >= 0 ? 1.0 : -1.0) / (sd_fabs(theta) + sd_sqrt(theta * theta + sd(1.0)));
            sd c = sd(1.0) / sd_sqrt(t * t + sd(1.0)), s = t * c;
            for (int k = 0; k < 3; ++k) {
                sd akp = A[k][p], akq = A[k][q];
                A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
            }
            for (int k = 0; k < 3; ++k) {
                sd apk = A[p][k], aqk = A[q][k];
                A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
            }
            for (int k = 0; k < 3; ++k) {
                sd vkp = V[k][p], vkq = V[k][q];
                V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
            }
        }
        // converged: the off-diagonal part is below 1e-22 of the diagonal (the oracle takes the same exit)
        const sd off = (sd_fabs(A[0][1]) + sd_fabs(A[0][2])) + sd_fabs(A[1][2]);
        const sd dia = (sd_fabs(A[0][0]) + sd_fabs(A[1][1])) + sd_fabs(A[2][2]);
        if (off.v <= (sd(1e-22) * dia).v) break;
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// sin and cos with the fixed operation order of orc_sincos (oracle/oracle_common.h): bit-identical to the oracle
__device__ inline void s3d_sincos(double x, double *sn, double *cs)
{
    const double k = rint(__dmul_rn(x, 0.63661977236758138));
    double r = __fma_rn(-k, 1.5707963267948966, x);
    r = __fma_rn(-k, 6.123233995736766e-17, r);
    const double z = __dmul_rn(r, r);
    double ps = -8.2206352466243295e-18;
    ps = __fma_rn(ps, z, 2.8114572543455206e-15);
    ps = __fma_rn(ps, z, -7.6471637318198164e-13);
    ps = __fma_rn(ps, z, 1.6059043836821613e-10);
    ps = __fma_rn(ps, z, -2.5052108385441720e-08);
    ps = __fma_rn(ps, z, 2.7557319223985893e-06);
    ps = __fma_rn(ps, z, -1.9841269841269841e-04);
    ps = __fma_rn(ps, z, 8.3333333333333332e-03);
    ps = __fma_rn(ps, z, -1.6666666666666666e-01);
    const double s0 = __fma_rn(__dmul_rn(ps, z), r, r);
    double pc = 4.1103176233121648e-19;
    pc = __fma_rn(pc, z, -1.5619206968586226e-16);
    pc = __fma_rn(pc, z, 4.7794773323873853e-14);
    pc = __fma_rn(pc, z, -1.1470745597729725e-11);
    pc = __fma_rn(pc, z, 2.0876756987868099e-09);
    pc = __fma_rn(pc, z, -2.7557319223985888e-07);
    pc = __fma_rn(pc, z, 2.4801587301587302e-05);
    pc = __fma_rn(pc, z, -1.3888888888888889e-03);
    pc = __fma_rn(pc, z, 4.1666666666666664e-02);
    pc = __fma_rn(pc, z, -0.5);
    const double c0 = __fma_rn(pc, z, 1.0);
    const long long q = (long long)k & 3;
    *sn = q == 0 ? s0 : q == 1 ? c0 : q == 2 ? -s0 : -c0;
    *cs = q == 0 ? c0 : q == 1 ? -s0 : q == 2 ? -c0 : s0;
}

// ------------------------------------------------------------------------------------------------
// Order-independent sums (the contract is written down in oracle/oracle_common.h): every product a*b of two float32
// values is rounded once to a multiple of 2^-g by s = fma(a, b, M), M = 1.5 * 2^(52-g); bits(s) - bits(M) is the integer
// rint(a*b*2^g) and integers add exactly.  A thread adds the raw bits of s (wrapping int64) and removes count * bits(M)
// when it hands its partial sums on; partial sums travel as (hi, lo) pairs of int64 with value hi * 2^32 + lo, which any
// number of adds in any order cannot overflow; the total becomes a double once, (double)hi * 2^32 + (double)lo, scaled by 2^-g.
// ------------------------------------------------------------------------------------------------
struct FxScale { unsigned long long mbits; double scale; int g; };

__host__ __device__ __forceinline__ FxScale s3d_fx_make(double B)      // 2^(E-1) <= B < 2^E, g = 49 - E
{
    FxScale f;
#ifdef __CUDA_ARCH__
    const unsigned long long bb = (unsigned long long)__double_as_longlong(B);
#else
    unsigned long long bb; memcpy(&bb, &B, 8);
#endif
    const int E = (int)((bb >> 52) & 0x7ff) - 1022;
    f.g = 49 - E;
    f.mbits = ((unsigned long long)(1075 - f.g) << 52) | (1ull << 51);
    const unsigned long long sb = (unsigned long long)(1023 - f.g) << 52;
#ifdef __CUDA_ARCH__
    f.scale = __longlong_as_double((long long)sb);
#else
    memcpy(&f.scale, &sb, 8);
#endif
    return f;
}
// raw bits of fma(a, b, M): what a thread accumulates (wrapping)
__device__ __forceinline__ long long s3d_fx_bits(double a, double b, double M) { return __double_as_longlong(__fma_rn(a, b, M)); }
// a partial sum v (true value fits int64) -> its (hi, lo) contribution
__device__ __forceinline__ void s3d_fx_split(long long v, long long &hi, long long &lo) { hi = v >> 32; lo = v & 0xffffffffll; }
__device__ __forceinline__ double s3d_fx_value(long long hi, long long lo, double scale)
{
    return __dmul_rn(__dadd_rn(__dmul_rn(__ll2double_rn(hi), 4294967296.0), __ll2double_rn(lo)), scale);
}
// bound on every product of one ICP iteration (orc_icp_bound): strict operations, identical bits on both sides
__device__ __forceinline__ double s3d_icp_bound(float P, float Q, float Nn, const double *T12)
{
    double Rm = 0.0, tm = 0.0;
    #pragma unroll
    for (int r = 0; r < 3; ++r) {
        #pragma unroll
        for (int c = 0; c < 3; ++c) { const double a = fabs(T12[4 * r + c]); if (a > Rm) Rm = a; }
        const double a = fabs(T12[4 * r + 3]); if (a > tm) tm = a;
    }
    double A = __dadd_rn(__dmul_rn(__dmul_rn(3.0, Rm), (double)P), tm);
    if ((double)Q > A) A = (double)Q;
    if (!(A > 1.0)) A = 1.0;
    const double N = (double)Nn > 1.0 ? (double)Nn : 1.0;
    double B = __dmul_rn(__dmul_rn(16.0, __dmul_rn(N, N)), __dmul_rn(A, A));
    if (!(B < 1e300)) B = 1e300;
    return B;
}
__device__ __forceinline__ double s3d_pca_bound(float A)
{
    const double a = (double)A > 1.0 ? (double)A : 1.0;
    double B = __dmul_rn(2.0, __dmul_rn(a, a));
    if (!(B < 1e300)) B = 1e300;
    return B;
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
