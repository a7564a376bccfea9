/*
 * icp_oracle.c -- CPU restatement of PCL-1.7 IterativeClosestPoint for the registration seat
 * GraphicEnd::multiPnP (reference src/GraphicEnd.cpp:557-659).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle_common.h).  The reference has no ICP
 * (SURVEY.md section 0.1); the algorithm restated here is the published PCL 1.7 one that the
 * north star names:
 *   - pcl::IterativeClosestPoint::computeTransformation      (registration/impl/icp.hpp)
 *   - pcl::registration::CorrespondenceEstimation::determineCorrespondences (exact 1-NN, d^2 gate)
 *   - pcl::registration::TransformationEstimationPointToPlaneLLS::estimateRigidTransformation
 *   - pcl::registration::TransformationEstimationSVD (Umeyama without scale)
 * and the reference's own result plumbing: norm formula and failure convention
 * (src/GraphicEnd.cpp:617-624, :585-600).
 *
 * Exact NN uses an own KD-tree (PCL uses FLANN's).  "Exact" means argmin over the float32 value
 * orc_dist2() with lowest-target-index tie break; pruning uses double arithmetic with a 1e-6
 * relative safety margin so no float-minimal candidate is ever skipped.
 */
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include "oracle_common.h"
#include "../include/slam3d_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ KD-tree ---------------- */
typedef struct {
    int n;
    float *pts;    /* reordered xyz + original index (as int bits) : 4 floats / point */
    int n_nodes;
    int *lo, *hi;  /* point range of node */
    int *left, *right; /* children (-1 for leaf) */
    int *dim; float *split;
} kdtree;

#define KD_LEAF 12

static int kd_cmp_dim;
static const float *kd_cmp_pts;
static int kd_cmp(const void *a, const void *b)
{
    float fa = kd_cmp_pts[4 * (*(const int *)a) + kd_cmp_dim];
    float fb = kd_cmp_pts[4 * (*(const int *)b) + kd_cmp_dim];
    return (fa > fb) - (fa < fb);
}

/* nth_element on an index array by coordinate d (Hoare quickselect) */
static void kd_select(int *idx, int n, int k, const float *xyzw, int d)
{
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        float pivot = xyzw[4 * idx[(lo + hi) / 2] + d];
        int i = lo, j = hi;
        while (i <= j) {
            while (xyzw[4 * idx[i] + d] < pivot) ++i;
            while (xyzw[4 * idx[j] + d] > pivot) --j;
            if (i <= j) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; ++i; --j; }
        }
        if (k <= j) hi = j; else if (k >= i) lo = i; else return;
    }
}

static int kd_build_rec(kdtree *t, int *idx, const float *xyzw, int lo, int hi)
{
    int node = t->n_nodes++;
    t->lo[node] = lo; t->hi[node] = hi; t->left[node] = t->right[node] = -1;
    if (hi - lo <= KD_LEAF) return node;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; ++i) for (int d = 0; d < 3; ++d) {
        float v = xyzw[4 * idx[i] + d];
        if (v < mn[d]) mn[d] = v;
        if (v > mx[d]) mx[d] = v;
    }
    int d = 0;
    if (mx[1] - mn[1] > mx[d] - mn[d]) d = 1;
    if (mx[2] - mn[2] > mx[d] - mn[d]) d = 2;
    if (!(mx[d] - mn[d] > 0.f)) return node; /* all identical: keep as (big) leaf */
    int mid = (lo + hi) / 2;
    kd_select(idx + lo, hi - lo, mid - lo, xyzw, d);
    t->dim[node] = d; t->split[node] = xyzw[4 * idx[mid] + d];
    int l = kd_build_rec(t, idx, xyzw, lo, mid);
    int r = kd_build_rec(t, idx, xyzw, mid, hi);
    t->left[node] = l; t->right[node] = r;
    return node;
}

static kdtree *kd_build(const float *xyzw, int n)
{
    kdtree *t = (kdtree *)calloc(1, sizeof(kdtree));
    t->n = n;
    int cap = 2 * (n / (KD_LEAF / 2) + 2) + 8;
    t->lo = (int *)malloc(sizeof(int) * cap); t->hi = (int *)malloc(sizeof(int) * cap);
    t->left = (int *)malloc(sizeof(int) * cap); t->right = (int *)malloc(sizeof(int) * cap);
    t->dim = (int *)malloc(sizeof(int) * cap); t->split = (float *)malloc(sizeof(float) * cap);
    int *idx = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) idx[i] = i;
    if (n > 0) kd_build_rec(t, idx, xyzw, 0, n);
    t->pts = (float *)malloc(sizeof(float) * 4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        memcpy(t->pts + 4 * i, xyzw + 4 * idx[i], 3 * sizeof(float));
        memcpy(t->pts + 4 * i + 3, &idx[i], sizeof(int));
    }
    free(idx);
    (void)kd_cmp; (void)kd_cmp_dim; (void)kd_cmp_pts;
    return t;
}

static void kd_free(kdtree *t)
{
    if (!t) return;
    free(t->pts); free(t->lo); free(t->hi); free(t->left); free(t->right); free(t->dim); free(t->split);
    free(t);
}

static void kd_query(const kdtree *t, const float *p, int *best_idx, float *best_d2)
{
    int bi = -1; float bd = INFINITY;
    if (t->n == 0) { *best_idx = -1; *best_d2 = INFINITY; return; }
    int stack[128]; double bound[128]; int sp = 0;
    stack[sp] = 0; bound[sp] = 0.0; ++sp;
    while (sp > 0) {
        --sp;
        int node = stack[sp];
        if (bound[sp] > (double)bd * 1.000001) continue;
        while (t->left[node] >= 0) {
            int d = t->dim[node];
            double diff = (double)p[d] - (double)t->split[node];
            int nearc = diff < 0 ? t->left[node] : t->right[node];
            int farc = diff < 0 ? t->right[node] : t->left[node];
            if (sp < 128) { stack[sp] = farc; bound[sp] = diff * diff; ++sp; }
            node = nearc;
        }
        for (int i = t->lo[node]; i < t->hi[node]; ++i) {
            const float *q = t->pts + 4 * i;
            float d2 = orc_dist2(p, q);
            int qi; memcpy(&qi, q + 3, sizeof(int));
            if (d2 < bd || (d2 == bd && qi < bi)) { bd = d2; bi = qi; }
        }
    }
    *best_idx = bi; *best_d2 = bd;
}

/* exported for tests: exact NN of every (transformed) source point */
void oracle_nn_kdtree(const float *src_xyzw, int n, const float *tgt_xyzw, int m,
                      const float *T12 /*nullable*/, int *idx_out, float *d2_out, int nthreads)
{
    static const float I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (!T12) T12 = I12;
    kdtree *t = kd_build(tgt_xyzw, m);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        float p[3]; orc_xform(T12, src_xyzw[4 * i], src_xyzw[4 * i + 1], src_xyzw[4 * i + 2], p);
        kd_query(t, p, &idx_out[i], &d2_out[i]);
    }
    kd_free(t);
}

/* brute-force variant: the independent check on the KD-tree (small sizes only) */
void oracle_nn_brute(const float *src_xyzw, int n, const float *tgt_xyzw, int m,
                     const float *T12, int *idx_out, float *d2_out)
{
    static const float I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (!T12) T12 = I12;
    for (int i = 0; i < n; ++i) {
        float p[3]; orc_xform(T12, src_xyzw[4 * i], src_xyzw[4 * i + 1], src_xyzw[4 * i + 2], p);
        int bi = -1; float bd = INFINITY;
        for (int j = 0; j < m; ++j) {
            float d2 = orc_dist2(p, tgt_xyzw + 4 * j);
            if (d2 < bd) { bd = d2; bi = j; }
        }
        idx_out[i] = bi; d2_out[i] = bd;
    }
}

/* --------------------------------------------------------------- small dense solvers -------- */

/* Solve of the symmetric 6x6 A x = g by the square-root-free Cholesky factorisation A = L D L^T (A given as full matrix).
 * Returns 0 on success, 1 when a pivot falls below pivot_eps * A[k][k] (rank deficient: planar sliding).  Part of the
 * arithmetic contract: the CUDA solve (csrc/icp.cu) performs the same IEEE operations in the same order (no contraction
 * on either side), so poses are bit-identical. */
static int chol6_solve(const double A[6][6], const double g[6], double pivot_eps, double x[6])
{
    double L[6][6], D[6], iD[6];
    int bad = 0;
    memset(L, 0, sizeof(L));
    for (int j = 0; j < 6; ++j) {
        double s = A[j][j];
        for (int k = 0; k < j; ++k) s = s - (L[j][k] * L[j][k]) * D[k];
        if (!(s > pivot_eps * A[j][j]) || !(A[j][j] > 0.0)) bad = 1;
        D[j] = s; iD[j] = 1.0 / s;
        for (int i = j + 1; i < 6; ++i) {
            double v = A[i][j];
            for (int k = 0; k < j; ++k) v = v - (L[i][k] * L[j][k]) * D[k];
            L[i][j] = v * iD[j];
        }
    }
    if (bad) return 1;
    double y[6];
    for (int i = 0; i < 6; ++i) { double v = g[i]; for (int k = 0; k < i; ++k) v = v - L[i][k] * y[k]; y[i] = v; }
    for (int i = 5; i >= 0; --i) { double v = y[i] * iD[i]; for (int k = i + 1; k < 6; ++k) v = v - L[k][i] * x[k]; x[i] = v; }
    return 0;
}

/* R = Rz(gamma) Ry(beta) Rx(alpha) with the full sin/cos matrix (PCL constructTransformationMatrix) */
static void euler_to_T(const double x[6], double D[12])
{
    double sa, ca, sb, cb, sg, cg;      /* orc_sincos: fixed operation order, bit-identical to the CUDA path */
    orc_sincos(x[0], &sa, &ca); orc_sincos(x[1], &sb, &cb); orc_sincos(x[2], &sg, &cg);
    D[0] = cg * cb; D[1] = -sg * ca + cg * sb * sa; D[2] = sg * sa + cg * sb * ca;  D[3] = x[3];
    D[4] = sg * cb; D[5] = cg * ca + sg * sb * sa;  D[6] = -cg * sa + sg * sb * ca; D[7] = x[4];
    D[8] = -sb;     D[9] = cb * sa;                 D[10] = cb * ca;                D[11] = x[5];
}

/* one-sided Jacobi SVD of a 3x3 via eigen-decomposition of H^T H; returns R = V diag(1,1,det) U^T */
static void kabsch_rotation(double H[3][3], double R[3][3])
{
    /* H = sum (p-pbar)(q-qbar)^T ; want R minimising sum |R p - q|^2 : H = U S V^T, R = V D U^T */
    double HtH[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];
        HtH[i][j] = s;
    }
    orc_jacobi3(HtH, V, w);
    /* sort eigenvalues descending */
    int ord[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b) if (w[ord[b]] > w[ord[a]]) { int t = ord[a]; ord[a] = ord[b]; ord[b] = t; }
    double Vs[3][3], U[3][3];
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) Vs[r][c] = V[r][ord[c]];
    /* u_c = H v_c / sigma_c for the two dominant; third by cross product (handles rank-2 H) */
    for (int c = 0; c < 2; ++c) {
        double u[3] = {0, 0, 0};
        for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) u[r] += H[r][k] * Vs[k][c];
        double n = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (n < 1e-300) n = 1;
        for (int r = 0; r < 3; ++r) U[r][c] = u[r] / n;
    }
    /* re-orthogonalise u1 against u0 */
    {
        double d = U[0][0] * U[0][1] + U[1][0] * U[1][1] + U[2][0] * U[2][1];
        for (int r = 0; r < 3; ++r) U[r][1] -= d * U[r][0];
        double n = sqrt(U[0][1] * U[0][1] + U[1][1] * U[1][1] + U[2][1] * U[2][1]);
        if (n < 1e-300) n = 1;
        for (int r = 0; r < 3; ++r) U[r][1] /= n;
    }
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    /* make V right-handed too: v2 = v0 x v1 ; then det(V)=det(U)=+1 and R = V U^T is a proper
     * rotation equal to V diag(1,1,det(V U^T)) U^T of the unconstrained SVD */
    Vs[0][2] = Vs[1][0] * Vs[2][1] - Vs[2][0] * Vs[1][1];
    Vs[1][2] = Vs[2][0] * Vs[0][1] - Vs[0][0] * Vs[2][1];
    Vs[2][2] = Vs[0][0] * Vs[1][1] - Vs[1][0] * Vs[0][1];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += Vs[i][k] * U[j][k];
        R[i][j] = s;
    }
}

static void compose(const double D[12], double T[12])
{
    double O[12];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            O[4 * r + c] = D[4 * r] * T[c] + D[4 * r + 1] * T[4 + c] + D[4 * r + 2] * T[8 + c];
        O[4 * r + 3] = D[4 * r] * T[3] + D[4 * r + 1] * T[7] + D[4 * r + 2] * T[11] + D[4 * r + 3];
    }
    memcpy(T, O, sizeof(O));
}

/* |min(theta, 2pi - theta)| + 0.9 |t|   (src/GraphicEnd.cpp:618) */
double oracle_pose_norm(const double *T16)
{
    double sx = T16[9] - T16[6], sy = T16[2] - T16[8], sz = T16[4] - T16[1];
    double s = 0.5 * sqrt(sx * sx + sy * sy + sz * sz);
    double c = 0.5 * (T16[0] + T16[5] + T16[10] - 1.0);
    double theta = atan2(s, c);
    double alt = 2.0 * M_PI - theta;
    double th = fabs(theta < alt ? theta : alt);
    double tn = sqrt(T16[3] * T16[3] + T16[7] * T16[7] + T16[11] * T16[11]);
    return th + 0.9 * tn;
}

static void result_fail(s3d_result *r, int status, int iters)
{
    memset(r->T, 0, sizeof(r->T));
    r->T[0] = r->T[5] = r->T[10] = r->T[15] = 1.0;
    r->norm = 0.0; r->status = status; r->iterations = iters;
}

/*
 * The ICP loop (PCL icp.hpp computeTransformation, fixed-iteration mode: transformation_epsilon=0,
 * euclidean_fitness_epsilon=-DBL_MAX => the only stop criterion is max_iterations).
 * src/tgt: float4 rows (x,y,z,*). tgt_nrm: float4 rows (nx,ny,nz,valid) or NULL (SVD only).
 * nn_out (nullable): correspondences of the last executed iteration, -1 = rejected.
 * PCL transforms the cloud incrementally (X <- dT X); we apply the accumulated T (kept in double,
 * rounded to float once per iteration) to the original source, which is the same map with less
 * rounding drift.
 */
int oracle_icp(const float *src, int n, const float *tgt, const float *tgt_nrm, int m,
               const double *guess16, const s3d_icp_params *prm, s3d_result *res,
               int *nn_out, int nthreads)
{
    memset(res, 0, sizeof(*res));
    double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (guess16) memcpy(T, guess16, sizeof(T));
    const int est = prm->estimator;
    if (est == S3D_ESTIMATOR_POINT_TO_PLANE && !tgt_nrm) return S3D_E_STATE;
    const double pivot_eps = prm->pivot_eps > 0 ? prm->pivot_eps : 1e-9;
    const float max_d2 = prm->max_corr_dist > 0 ? prm->max_corr_dist * prm->max_corr_dist : INFINITY;
    const int min_corr = prm->min_correspondences > 0 ? prm->min_correspondences : 3;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const float absP = orc_absmax3(src, n, 0), absQ = orc_absmax3(tgt, m, 0), absN = tgt_nrm ? orc_absmax3(tgt_nrm, m, 1) : 1.0f;
    kdtree *tree = kd_build(tgt, m);
    int *nn = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    float *nd = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    float *xp = (float *)malloc(sizeof(float) * 3 * (n > 0 ? n : 1));
    int status = S3D_PAIR_OK, it = 0;
    for (; it < prm->max_iterations; ++it) {
        float Tf[12];
        for (int k = 0; k < 12; ++k) Tf[k] = (float)T[k];
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            orc_xform(Tf, src[4 * i], src[4 * i + 1], src[4 * i + 2], xp + 3 * i);
            kd_query(tree, xp + 3 * i, &nn[i], &nd[i]);
        }
        /* order-independent fixed-point sums (oracle_common.h): every product rounded once to 2^-g, integers added exactly */
        const orc_fx fx = orc_fx_make(orc_icp_bound(absP, absQ, est == S3D_ESTIMATOR_POINT_TO_PLANE ? absN : 1.0f, T));
        __int128 SA[6][6], Sg[6], Sp[3], Sq[3], Spq[3][3], Sd2 = 0;
        memset(SA, 0, sizeof(SA)); memset(Sg, 0, sizeof(Sg)); memset(Sp, 0, sizeof(Sp)); memset(Sq, 0, sizeof(Sq)); memset(Spq, 0, sizeof(Spq));
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            int j = nn[i];
            int ok = (j >= 0) && (nd[i] <= max_d2);
            if (ok && est == S3D_ESTIMATOR_POINT_TO_PLANE) ok = tgt_nrm[4 * j + 3] != 0.0f;
            if (!ok) { nn[i] = -1; continue; }
            const float *p = xp + 3 * i, *q = tgt + 4 * j;
            ++cnt; Sd2 += orc_fx_term(&fx, nd[i], 1.0f);
            if (est == S3D_ESTIMATOR_POINT_TO_PLANE) {
                const float *nv = tgt_nrm + 4 * j;
                float J[6];
                J[0] = fmaf(nv[2], p[1], -(nv[1] * p[2]));
                J[1] = fmaf(nv[0], p[2], -(nv[2] * p[0]));
                J[2] = fmaf(nv[1], p[0], -(nv[0] * p[1]));
                J[3] = nv[0]; J[4] = nv[1]; J[5] = nv[2];
                float ex = q[0] - p[0], ey = q[1] - p[1], ez = q[2] - p[2];
                float r = fmaf(nv[2], ez, fmaf(nv[1], ey, nv[0] * ex));
                for (int a = 0; a < 6; ++a) {
                    for (int b = a; b < 6; ++b) SA[a][b] += orc_fx_term(&fx, J[a], J[b]);
                    Sg[a] += orc_fx_term(&fx, J[a], r);
                }
            } else {
                for (int a = 0; a < 3; ++a) {
                    Sp[a] += orc_fx_term(&fx, p[a], 1.0f); Sq[a] += orc_fx_term(&fx, q[a], 1.0f);
                    for (int b = 0; b < 3; ++b) Spq[a][b] += orc_fx_term(&fx, p[a], q[b]);
                }
            }
        }
        double A[6][6]; double g[6]; double sp[3], sq[3], spq[3][3];
        for (int a = 0; a < 6; ++a) { for (int b = 0; b < 6; ++b) A[a][b] = orc_fx_value(&fx, SA[a][b]); g[a] = orc_fx_value(&fx, Sg[a]); }
        for (int a = 0; a < 3; ++a) { sp[a] = orc_fx_value(&fx, Sp[a]); sq[a] = orc_fx_value(&fx, Sq[a]); for (int b = 0; b < 3; ++b) spq[a][b] = orc_fx_value(&fx, Spq[a][b]); }
        const double sum_d2 = orc_fx_value(&fx, Sd2);
        res->inliers = cnt;
        res->fitness = cnt ? sum_d2 / (double)cnt : 0.0;
        if (cnt < min_corr) { status = S3D_PAIR_FEW_CORRESPONDENCES; break; }
        double D[12];
        if (est == S3D_ESTIMATOR_POINT_TO_PLANE) {
            for (int a = 0; a < 6; ++a) for (int b = 0; b < a; ++b) A[a][b] = A[b][a];
            double x[6];
            if (chol6_solve(A, g, pivot_eps, x)) { status = S3D_PAIR_DEGENERATE; break; }
            euler_to_T(x, D);
        } else {
            double H[3][3], R[3][3], pb[3], qb[3];
            const double cnt_d = (double)cnt;
            for (int a = 0; a < 3; ++a) { pb[a] = sp[a] / cnt_d; qb[a] = sq[a] / cnt_d; }
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[a][b] = spq[a][b] - (cnt_d * pb[a]) * qb[b];
            /* H here is sum p q^T (rows p, cols q); kabsch_rotation expects H = U S V^T with
             * R = V U^T mapping p to q */
            kabsch_rotation(H, R);
            for (int a = 0; a < 3; ++a) {
                for (int b = 0; b < 3; ++b) D[4 * a + b] = R[a][b];
                D[4 * a + 3] = qb[a] - (R[a][0] * pb[0] + R[a][1] * pb[1] + R[a][2] * pb[2]);
            }
        }
        int finite = 1;
        for (int k = 0; k < 12; ++k) finite &= isfinite(D[k]);
        if (!finite) { status = S3D_PAIR_NONFINITE; break; }
        compose(D, T);
    }
    if (nn_out) memcpy(nn_out, nn, sizeof(int) * n);
    if (status != S3D_PAIR_OK) result_fail(res, status, it);
    else {
        memcpy(res->T, T, sizeof(T));
        res->T[12] = res->T[13] = res->T[14] = 0; res->T[15] = 1;
        res->norm = oracle_pose_norm(res->T);
        res->iterations = it; res->status = S3D_PAIR_OK;
    }
    free(nn); free(nd); free(xp); kd_free(tree);
    return S3D_OK;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
