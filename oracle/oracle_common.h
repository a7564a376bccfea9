/*
 * oracle_common.h -- shared float formulas and RNG of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY UNPINNED: the reference (gaoxiang12/slam3d_gx @ 99d5b1b) contains no ICP and ships no
 * tests or golden vectors for this path; PCL 1.7 (pinned by /root/reference/src/CMakeLists.txt:2)
 * is absent from the build container.  This oracle restates the published PCL-1.7 algorithms
 * (IterativeClosestPoint, TransformationEstimationPointToPlaneLLS, TransformationEstimationSVD,
 * SACSegmentation/RandomSampleConsensus/SampleConsensusModelPlane) and the first-party control
 * flow of /root/reference/src/GraphicEnd.cpp:353-430 and src/planarFeatures.cpp:88-136.  It is
 * pinned only against an independent scipy.spatial.cKDTree + numpy.linalg float64 implementation
 * (tests/test_oracle_crosscheck.py, tests/golden/).
 *
 * The float32 expressions below are the *definition* of the arithmetic; the CUDA kernels in
 * slam3d_gx_b200/csrc use the same operation order with __fmaf_rn/__fmul_rn so that integer
 * results (correspondence indices, inlier counts, labels) are bit-identical.
 */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H

#include <math.h>
#include <stdint.h>
#include <string.h>

/* p' = R p + t with T = row-major 3x4 float */
static inline void orc_xform(const float *T, float x, float y, float z, float *o)
{
    o[0] = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
    o[1] = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
    o[2] = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
}

/* squared Euclidean distance, the quantity the argmin runs over */
static inline float orc_dist2(const float *p, const float *q)
{
    float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* signed point-plane value a*x+b*y+c*z+d */
static inline float orc_plane_eval(const float *c, float x, float y, float z)
{
    return fmaf(c[2], z, fmaf(c[1], y, fmaf(c[0], x, c[3])));
}

/* splitmix64 finaliser; counter-based use: orc_rand(seed,a,b,c) */
static inline uint64_t orc_mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t orc_rand(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    return orc_mix(orc_mix(orc_mix(seed + a) + b) + c);
}

/* three distinct indices in [0,n) for hypothesis (a,b); n >= 3 */
static inline void orc_sample3(uint64_t seed, uint64_t a, uint64_t b, uint32_t n, uint32_t *out)
{
    uint32_t got = 0;
    for (uint32_t c = 0; got < 3 && c < 64; ++c) {
        uint32_t v = (uint32_t)(orc_rand(seed, a, b, c) % n);
        int dup = 0;
        for (uint32_t k = 0; k < got; ++k) dup |= (out[k] == v);
        if (!dup) out[got++] = v;
    }
    /* astronomically unlikely fallback keeps the triple distinct and deterministic */
    while (got < 3) { out[got] = (out[got - 1] + 1) % n; ++got; }
}

/* plane through three points (float32); returns 0 when the sample is degenerate */
static inline int orc_plane_from3(const float *p0, const float *p1, const float *p2, float *coef)
{
    float ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
    float bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
    float nx = fmaf(ay, bz, -(az * by));
    float ny = fmaf(az, bx, -(ax * bz));
    float nz = fmaf(ax, by, -(ay * bx));
    float l2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
    if (!(l2 > 1e-20f)) return 0;
    float inv = 1.0f / sqrtf(l2);
    nx *= inv; ny *= inv; nz *= inv;
    coef[0] = nx; coef[1] = ny; coef[2] = nz;
    coef[3] = -fmaf(nz, p0[2], fmaf(ny, p0[1], nx * p0[0]));
    return 1;
}

/* cyclic Jacobi eigen-decomposition of a symmetric 3x3 (double). A is destroyed; V columns are
 * eigenvectors, w eigenvalues (unsorted). At most 12 sweeps, left when the off-diagonal part is negligible; the test
 * uses the same strict operations on CPU and GPU, so the control flow is identical. */
static inline void orc_jacobi3(double A[3][3], double V[3][3], double w[3])
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
        /* converged: the off-diagonal part is below 1e-22 of the diagonal (quadratic convergence: typically 4-5 sweeps) */
        double off = (fabs(A[0][1]) + fabs(A[0][2])) + fabs(A[1][2]);
        double dia = (fabs(A[0][0]) + fabs(A[1][1])) + fabs(A[2][2]);
        if (off <= 1e-22 * dia) break;
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

/* ------------------------------------------------------------------------------------------------
 * Order-independent sums ("fixed-point accumulation") -- part of the arithmetic contract.
 *
 * Every sum of the path (the 27 normal-equation sums + sum d^2 of point-to-plane ICP, the 15 Kabsch sums, the 9 PCA
 * sums of the plane refit) is a sum of products a*b of two float32 values (b = 1 for plain sums).  a*b is exact in
 * double.  Each product is rounded ONCE to a multiple of 2^-g (round to nearest even) and the resulting integers are
 * added exactly, so the sum does not depend on the order of the additions: a sequential CPU loop, a GPU reduction tree
 * of any shape and a batched launch give the same bits.
 *   rounding:  s = fma(a, b, M) with M = 1.5 * 2^(52-g)  =>  bits(s) - bits(M) = rint(a*b * 2^g)   (|a*b| * 2^g < 2^49)
 *   g:         49 - E with 2^E > B, B a bound on every |a*b| computed from data bounds (below); g is data dependent but
 *              order independent (maxima of absolute values).
 *   value:     the exact integer sum S is converted once: ((double)(S >> 32) * 2^32 + (double)(S & 0xffffffff)) * 2^-g
 * Absolute resolution 2^-g: about 1e-11 for a room-sized scene in metres, i.e. at least as fine as plain double
 * accumulation of 3e5 terms.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int g; uint64_t mbits; double scale; } orc_fx;

static inline uint64_t orc_dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double orc_bitsd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

/* B: finite, > 0.  E = exponent with 2^(E-1) <= B < 2^E (from the bits of B), g = 49 - E. */
static inline orc_fx orc_fx_make(double B)
{
    orc_fx f;
    const int E = (int)((orc_dbits(B) >> 52) & 0x7ff) - 1022;
    f.g = 49 - E;
    f.mbits = ((uint64_t)(1075 - f.g) << 52) | ((uint64_t)1 << 51);
    f.scale = orc_bitsd((uint64_t)(1023 - f.g) << 52);
    return f;
}
static inline int64_t orc_fx_term(const orc_fx *f, float a, float b)
{
    return (int64_t)(orc_dbits(fma((double)a, (double)b, orc_bitsd(f->mbits))) - f->mbits);
}
static inline double orc_fx_value(const orc_fx *f, __int128 S)
{
    const int64_t H = (int64_t)(S >> 32), L = (int64_t)(S & (__int128)0xffffffff);
    return ((double)H * 4294967296.0 + (double)L) * f->scale;
}
/* largest finite |component| of n rows of 4 floats (first 3 used); rows with row[3] == 0 skipped when need_w */
static inline float orc_absmax3(const float *rows, int n, int need_w)
{
    float m = 0.f;
    for (int i = 0; i < n; ++i) {
        if (need_w && rows[4 * i + 3] == 0.0f) continue;
        for (int k = 0; k < 3; ++k) { float a = fabsf(rows[4 * i + k]); if (a <= 3.4028234663852886e38f && a > m) m = a; }
    }
    return m;
}
/* bound on every product of one ICP iteration: source coordinates <= P, target coordinates <= Q, normal components <= Nn
 * (1 for the SVD estimator), pose T (row-major 3x4, double).  |x| <= 3 Rm P + tm =: X; A = max(X, Q, 1), N = max(Nn, 1);
 * |J| <= 2 N A, |r| <= 6 N A, d^2 <= 12 A^2  =>  every product <= 12 N^2 A^2 < B = 16 N^2 A^2. */
static inline double orc_icp_bound(float P, float Q, float Nn, const double *T12)
{
    double Rm = 0.0, tm = 0.0;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { double a = fabs(T12[4 * r + c]); if (a > Rm) Rm = a; }
        double a = fabs(T12[4 * r + 3]); if (a > tm) tm = a;
    }
    double A = 3.0 * Rm * (double)P + tm;
    if ((double)Q > A) A = (double)Q;
    if (!(A > 1.0)) A = 1.0;                 /* also catches NaN */
    double N = (double)Nn > 1.0 ? (double)Nn : 1.0;
    double B = 16.0 * (N * N) * (A * A);
    if (!(B < 1e300)) B = 1e300;
    return B;
}
/* plane refit: sums of x, y, z and their products over points with |coordinate| <= A */
static inline double orc_pca_bound(float A)
{
    double a = (double)A > 1.0 ? (double)A : 1.0;
    double B = 2.0 * (a * a);
    if (!(B < 1e300)) B = 1e300;
    return B;
}

/* sin and cos with a fixed operation order (Cody-Waite reduction by pi/2 in two parts, Taylor-Horner in r^2, fma):
 * bit-identical on the CPU and on the GPU, unlike the two math libraries' sin/cos.  |x| < 2^20. */
static inline void orc_sincos(double x, double *sn, double *cs)
{
    const double k = rint(x * 0.63661977236758138);                 /* 2/pi */
    double r = fma(-k, 1.5707963267948966, x);                     /* pi/2 high */
    r = fma(-k, 6.123233995736766e-17, r);                         /* pi/2 low  */
    const double z = r * r;
    double ps = -8.2206352466243295e-18;                           /* -1/19! */
    ps = fma(ps, z, 2.8114572543455206e-15);                       /*  1/17! */
    ps = fma(ps, z, -7.6471637318198164e-13);                      /* -1/15! */
    ps = fma(ps, z, 1.6059043836821613e-10);                       /*  1/13! */
    ps = fma(ps, z, -2.5052108385441720e-08);                      /* -1/11! */
    ps = fma(ps, z, 2.7557319223985893e-06);                       /*  1/9!  */
    ps = fma(ps, z, -1.9841269841269841e-04);                      /* -1/7!  */
    ps = fma(ps, z, 8.3333333333333332e-03);                       /*  1/5!  */
    ps = fma(ps, z, -1.6666666666666666e-01);                      /* -1/3!  */
    const double s0 = fma(ps * z, r, r);
    double pc = 4.1103176233121648e-19;                            /*  1/20! */
    pc = fma(pc, z, -1.5619206968586226e-16);                      /* -1/18! */
    pc = fma(pc, z, 4.7794773323873853e-14);                       /*  1/16! */
    pc = fma(pc, z, -1.1470745597729725e-11);                      /* -1/14! */
    pc = fma(pc, z, 2.0876756987868099e-09);                       /*  1/12! */
    pc = fma(pc, z, -2.7557319223985888e-07);                      /* -1/10! */
    pc = fma(pc, z, 2.4801587301587302e-05);                       /*  1/8!  */
    pc = fma(pc, z, -1.3888888888888889e-03);                      /* -1/6!  */
    pc = fma(pc, z, 4.1666666666666664e-02);                       /*  1/4!  */
    pc = fma(pc, z, -0.5);
    const double c0 = fma(pc, z, 1.0);
    const long long q = (long long)k & 3;
    *sn = q == 0 ? s0 : q == 1 ? c0 : q == 2 ? -s0 : -c0;
    *cs = q == 0 ? c0 : q == 1 ? -s0 : q == 2 ? -c0 : s0;
}

#endif
