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
 * eigenvectors, w eigenvalues (unsorted). Fixed 12 sweeps: identical control flow on CPU/GPU. */
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
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

#endif
