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
// bound on the coordinates of a transformed source point (orc_pose_bound): strict operations, identical bits on both sides
__device__ __forceinline__ double s3d_pose_bound(float P, const double *T12)
{
    double X = 0.0;
    #pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double a = __dadd_rn(__dmul_rn(__dadd_rn(__dadd_rn(fabs(T12[4 * r]), fabs(T12[4 * r + 1])), fabs(T12[4 * r + 2])), (double)P), fabs(T12[4 * r + 3]));
        if (a > X) X = a;
    }
    if (!(X < 1e150)) X = 1e150;
    return X;
}
// bound on the squared correspondence distance (orc_icp_bound)
__device__ __forceinline__ double s3d_icp_bound(double X, float Q)
{
    const double s = __dadd_rn(__dadd_rn(X, (double)Q), 1.0);
    double B = __dmul_rn(4.0, __dmul_rn(s, s));
    if (!(B < 1e300)) B = 1e300;
    return B;
}
// 26-bit fixed-point factors of the normal-equation sums (contract: oracle/oracle_common.h, orc_fxq): one power-of-two scale per
// factor class (a: x cross n, n: normal, r: residual, c: coordinates), from data bounds; products and sums are exact integers.
struct FxQ { int sa, sn, sr, sc; float fa, fn, fr, fc; };
__device__ __forceinline__ int s3d_exp_above(double B) { return (int)(((unsigned long long)__double_as_longlong(B) >> 52) & 0x7ff) - 1022; }
__device__ __forceinline__ int s3d_clamp_scale(int s) { return s < -100 ? -100 : (s > 100 ? 100 : s); }
__device__ __forceinline__ float s3d_pow2f(int e) { return __uint_as_float((unsigned)(e + 127) << 23); }
__device__ __forceinline__ FxQ s3d_fxq_make(double X, float Q, float Nn, float gate)
{
    FxQ f;
    const double N = (double)Nn > 1e-30 ? (double)Nn : 1e-30;
    const double ba = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, N), X), 1.0001), 1e-30);
    const double bn = __dmul_rn(N, 1.0001);
    double br = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(3.0, N), __dadd_rn(X, (double)Q)), 1.0001), 1e-30);
    if ((double)gate < 1e30) {
        const double bg = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(1.7320508075688772, N), (double)gate), 1.0001), 1e-30);
        if (bg < br) br = bg;
    }
    const double bc = __dadd_rn(__dmul_rn(X > (double)Q ? X : (double)Q, 1.0001), 1e-30);
    f.sa = s3d_clamp_scale(26 - s3d_exp_above(ba)); f.sn = s3d_clamp_scale(26 - s3d_exp_above(bn));
    f.sr = s3d_clamp_scale(26 - s3d_exp_above(br)); f.sc = s3d_clamp_scale(26 - s3d_exp_above(bc));
    f.fa = s3d_pow2f(f.sa); f.fn = s3d_pow2f(f.sn); f.fr = s3d_pow2f(f.sr); f.fc = s3d_pow2f(f.sc);
    return f;
}
// (hi, lo) total of a sum of integer products -> double, scaled by 2^-s_sum
__device__ __forceinline__ double s3d_fxq_value(long long hi, long long lo, int s_sum)
{
    return __dmul_rn(__dadd_rn(__dmul_rn(__ll2double_rn(hi), 4294967296.0), __ll2double_rn(lo)),
                     __longlong_as_double((long long)((unsigned long long)(1023 - s_sum) << 52)));
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
