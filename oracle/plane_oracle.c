/*
 * plane_oracle.c -- CPU restatement of the reference's plane extraction and keypoint planarity test.
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see oracle_common.h).
 *
 * oracle_segment_planes follows the control flow of GraphicEnd::extractPlanesAndGenerateImage
 * (reference src/GraphicEnd.cpp:353-430) literally:
 *     while remaining > plane_percent*n  (:372)
 *        seg.segment                      (:375)  -> PCL-1.7 SACSegmentation, SACMODEL_PLANE, SAC_RANSAC,
 *                                                   optimizeCoefficients=true (:360-364)
 *        no inliers -> break              (:376-379)
 *        flip sign so d >= 0              (:383-387)
 *        remove inliers from the cloud    (:419-420)
 *        stop at max_planes               (:424-425)
 * with PCL's RandomSampleConsensus::computeModel adaptive loop (k = log(1-p)/log(1-w^3), cap
 * max_iterations=50), SampleConsensusModelPlane::countWithinDistance/selectWithinDistance
 * (|a x+b y+c z+d| < threshold) and optimizeModelCoefficients (centroid + covariance, eigenvector of
 * the smallest eigenvalue) followed by inlier re-selection with the refined model.
 * PCL's boost::mt19937 sample stream cannot be reproduced without PCL; hypotheses come from the
 * counter-based stream orc_sample3(seed, round, candidate) shared with the CUDA path.
 *
 * oracle_planar_keypoints follows isPlanar (src/planarFeatures.cpp:88-136).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle_common.h"
#include "../include/slam3d_b200.h"

#define PLANE_CANDIDATES_EXTRA 14   /* candidates drawn beyond max_iterations to absorb degenerate samples */

/* Replays PCL's sequential adaptive RANSAC over pre-evaluated candidates.
 * valid[c], count[c] for c < n_cand. Returns best candidate index or -1; *iters = iterations_ consumed. */
int oracle_ransac_replay(const int *valid, const int *count, int n_cand, int n_points,
                         int max_iterations, double probability, int sample_size, int *iters)
{
    int iterations = 0, best = -1, best_count = -2147483647;
    double k = 1.0;
    const double log_prob = log(1.0 - probability);
    const double one_over = n_points > 0 ? 1.0 / (double)n_points : 0.0;
    const double eps = 2.220446049250313e-16;
    for (int c = 0; c < n_cand && iterations < k; ++c) {
        if (!valid[c]) continue; /* PCL: ++skipped_count; continue (iterations_ unchanged) */
        if (count[c] > best_count) {
            best_count = count[c]; best = c;
            double w = best_count * one_over;
            double p_no = sample_size == 3 ? 1.0 - w * w * w : 1.0 - pow(w, (double)sample_size);
            if (p_no < eps) p_no = eps;
            if (p_no > 1.0 - eps) p_no = 1.0 - eps;
            k = log_prob / log(p_no);
        }
        ++iterations;
        if (iterations > max_iterations) break;
    }
    if (iters) *iters = iterations;
    return best;
}

/*
 * xyzw: n float4 rows. labels_out[n] (plane id or -1), normals_out n float4 rows (nx,ny,nz,valid).
 * planes_out[max_planes]. Returns number of planes.
 * cand_counts_out (nullable, [max_planes][max_iterations+EXTRA]) exposes the per-candidate inlier
 * counts of each round for bit-exact comparison with the CUDA evaluation kernel.
 */
int oracle_segment_planes(const float *xyzw, int n, const s3d_plane_params *prm,
                          s3d_plane *planes_out, int32_t *labels_out, float *normals_out,
                          int32_t *cand_counts_out)
{
    const int n_cand = prm->max_iterations + PLANE_CANDIDATES_EXTRA;
    const float tau = prm->distance_threshold;
    int *rem = (int *)malloc(sizeof(int) * (n > 0 ? n : 1)); /* original indices of remaining points, in order */
    int n_rem = n, n_planes = 0;
    const orc_fx fx = orc_fx_make(orc_pca_bound(orc_absmax3(xyzw, n, 0)));
    for (int i = 0; i < n; ++i) { rem[i] = i; labels_out[i] = -1; }
    if (normals_out) memset(normals_out, 0, sizeof(float) * 4 * n);
    int *valid = (int *)malloc(sizeof(int) * n_cand), *count = (int *)malloc(sizeof(int) * n_cand);
    float *coefs = (float *)malloc(sizeof(float) * 4 * n_cand);
    while ((double)n_rem > (double)prm->plane_percent * (double)n && n_planes < prm->max_planes) {
        if (n_rem < 3) break;
        for (int c = 0; c < n_cand; ++c) {
            uint32_t s[3];
            orc_sample3(prm->seed, (uint64_t)n_planes, (uint64_t)c, (uint32_t)n_rem, s);
            valid[c] = orc_plane_from3(xyzw + 4 * rem[s[0]], xyzw + 4 * rem[s[1]], xyzw + 4 * rem[s[2]], coefs + 4 * c);
            count[c] = 0;
            if (!valid[c]) continue;
            int cnt = 0;
            for (int i = 0; i < n_rem; ++i) {
                const float *p = xyzw + 4 * rem[i];
                cnt += fabsf(orc_plane_eval(coefs + 4 * c, p[0], p[1], p[2])) < tau;
            }
            count[c] = cnt;
        }
        if (cand_counts_out) memcpy(cand_counts_out + (size_t)n_planes * n_cand, count, sizeof(int) * n_cand);
        int iters = 0;
        int best = oracle_ransac_replay(valid, count, n_cand, n_rem, prm->max_iterations, prm->probability, 3, &iters);
        if (best < 0 || count[best] == 0) break; /* :376-379 */
        /* optimizeModelCoefficients: PCA over the inliers of the best model */
        const float *bc = coefs + 4 * best;
        /* order-independent fixed-point sums (oracle_common.h), resolution from the largest coordinate of the cloud */
        __int128 S1[3] = {0, 0, 0}, S2[6] = {0, 0, 0, 0, 0, 0}; int ni = 0;
        for (int i = 0; i < n_rem; ++i) {
            const float *p = xyzw + 4 * rem[i];
            if (fabsf(orc_plane_eval(bc, p[0], p[1], p[2])) < tau) {
                S1[0] += orc_fx_term(&fx, p[0], 1.0f); S1[1] += orc_fx_term(&fx, p[1], 1.0f); S1[2] += orc_fx_term(&fx, p[2], 1.0f);
                S2[0] += orc_fx_term(&fx, p[0], p[0]); S2[1] += orc_fx_term(&fx, p[0], p[1]); S2[2] += orc_fx_term(&fx, p[0], p[2]);
                S2[3] += orc_fx_term(&fx, p[1], p[1]); S2[4] += orc_fx_term(&fx, p[1], p[2]); S2[5] += orc_fx_term(&fx, p[2], p[2]);
                ++ni;
            }
        }
        double s1[3], s2[6];
        for (int k = 0; k < 3; ++k) s1[k] = orc_fx_value(&fx, S1[k]);
        for (int k = 0; k < 6; ++k) s2[k] = orc_fx_value(&fx, S2[k]);
        float rc[4] = {bc[0], bc[1], bc[2], bc[3]};
        if (ni >= 3) { /* PCL needs > 3 inliers to refit; with fewer the RANSAC model is kept */
            const double nd = (double)ni;
            double cx = s1[0] / nd, cy = s1[1] / nd, cz = s1[2] / nd;
            double C[3][3], V[3][3], w[3];
            C[0][0] = s2[0] / nd - cx * cx; C[0][1] = C[1][0] = s2[1] / nd - cx * cy; C[0][2] = C[2][0] = s2[2] / nd - cx * cz;
            C[1][1] = s2[3] / nd - cy * cy; C[1][2] = C[2][1] = s2[4] / nd - cy * cz; C[2][2] = s2[5] / nd - cz * cz;
            orc_jacobi3(C, V, w);
            int k = 0; if (w[1] < w[k]) k = 1; if (w[2] < w[k]) k = 2;
            double nx = V[0][k], ny = V[1][k], nz = V[2][k];
            double nn = sqrt(nx * nx + ny * ny + nz * nz);
            nx /= nn; ny /= nn; nz /= nn;
            double d = -(nx * cx + ny * cy + nz * cz);
            if (d < 0) { nx = -nx; ny = -ny; nz = -nz; d = -d; } /* :383-387 */
            rc[0] = (float)nx; rc[1] = (float)ny; rc[2] = (float)nz; rc[3] = (float)d;
        } else if (rc[3] < 0) { rc[0] = -rc[0]; rc[1] = -rc[1]; rc[2] = -rc[2]; rc[3] = -rc[3]; }
        /* re-select inliers with the refined model, label them, compact the rest (order kept) */
        int w_ = 0, nin = 0;
        for (int i = 0; i < n_rem; ++i) {
            int oi = rem[i];
            const float *p = xyzw + 4 * oi;
            if (fabsf(orc_plane_eval(rc, p[0], p[1], p[2])) < tau) {
                labels_out[oi] = n_planes; ++nin;
                if (normals_out) { normals_out[4 * oi] = rc[0]; normals_out[4 * oi + 1] = rc[1]; normals_out[4 * oi + 2] = rc[2]; normals_out[4 * oi + 3] = 1.0f; }
            } else rem[w_++] = oi;
        }
        if (nin == 0) break;
        memcpy(planes_out[n_planes].coef, rc, sizeof(rc));
        planes_out[n_planes].inliers = nin; planes_out[n_planes].hypotheses = iters;
        n_rem = w_; ++n_planes;
    }
    free(rem); free(valid); free(count); free(coefs);
    return n_planes;
}

/* isPlanar (src/planarFeatures.cpp:88-136): 7x7 patch, any zero depth -> false (:103-107);
 * back-projection in double (:108-111) stored as float (pcl::PointXYZ); RANSAC plane with
 * threshold, PCL SampleConsensus defaults max_iterations=1000, probability=0.99, no refit;
 * planar iff inliers > min_inliers (:127). Keypoints closer than 3 px to the border (where the
 * reference's cv::Mat::operator() would throw) are reported non-planar. */
void oracle_planar_keypoints(const uint16_t *depth, int width, int height, const s3d_camera *cam,
                             const int32_t *uv, int n, float threshold, int min_inliers,
                             uint64_t seed, uint8_t *flags_out)
{
    for (int kp = 0; kp < n; ++kp) {
        int u = uv[2 * kp], v = uv[2 * kp + 1];
        flags_out[kp] = 0;
        if (u < 3 || v < 3 || u + 3 >= width || v + 3 >= height) continue;
        float P[49][3]; int has_zero = 0;
        for (int j = 0; j < 7 && !has_zero; ++j) for (int i = 0; i < 7; ++i) {
            uint16_t dd = depth[(size_t)(v + j - 3) * width + (u + i - 3)];
            if (dd == 0) { has_zero = 1; break; }
            double z = (double)dd / cam->factor;
            double x = ((double)(u + (i - 3)) - cam->cx) * z / cam->fx;
            double y = ((double)(v + (j - 3)) - cam->cy) * z / cam->fy;
            P[j * 7 + i][0] = (float)x; P[j * 7 + i][1] = (float)y; P[j * 7 + i][2] = (float)z;
        }
        if (has_zero) continue;
        /* sequential adaptive RANSAC, candidates from the shared stream */
        int iterations = 0, best_count = -2147483647; double k = 1.0;
        const double log_prob = log(1.0 - 0.99), eps = 2.220446049250313e-16;
        for (int c = 0; c < 1056 && iterations < k; ++c) { /* 33 chunks of 32 candidates */
            uint32_t s[3]; float coef[4];
            orc_sample3(seed, (uint64_t)kp, (uint64_t)c, 49u, s);
            if (!orc_plane_from3(P[s[0]], P[s[1]], P[s[2]], coef)) continue;
            int cnt = 0;
            for (int i = 0; i < 49; ++i) cnt += fabsf(orc_plane_eval(coef, P[i][0], P[i][1], P[i][2])) < threshold;
            if (cnt > best_count) {
                best_count = cnt;
                double w = cnt / 49.0, p_no = 1.0 - w * w * w;
                if (p_no < eps) p_no = eps;
                if (p_no > 1.0 - eps) p_no = 1.0 - eps;
                k = log_prob / log(p_no);
            }
            ++iterations;
            if (iterations > 1000) break;
        }
        flags_out[kp] = best_count > min_inliers;
    }
}

/* depth -> cloud exactly like src/convert2PCD.cpp:54-80 (double arithmetic, float storage),
 * optional z pass-through (src/GraphicEnd.cpp:283-285). Returns the number of points written. */
int oracle_backproject(const uint16_t *depth, int width, int height, const s3d_camera *cam,
                       float z_max, float *xyzw_out)
{
    int k = 0;
    for (int m = 0; m < height; ++m) for (int nn = 0; nn < width; ++nn) {
        uint16_t d = depth[(size_t)m * width + nn];
        if (d == 0) continue;
        double z = (double)d / cam->factor;
        double x = ((double)nn - cam->cx) * z / cam->fx;
        double y = ((double)m - cam->cy) * z / cam->fy;
        float fz = (float)z;
        if (z_max > 0 && !(fz >= 0.0f && fz <= z_max)) continue;
        xyzw_out[4 * k] = (float)x; xyzw_out[4 * k + 1] = (float)y; xyzw_out[4 * k + 2] = fz; xyzw_out[4 * k + 3] = 1.0f;
        ++k;
    }
    return k;
}
