/*
 * filter_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the cloud filters either side of the
 * registration path, PCL-1.7 semantics as used by the reference:
 *   pcl::PassThrough  "z" in [min,max]      reference src/GraphicEnd.cpp:283-285,291-292; src/saveOutput.cpp:40-46,81-84
 *   pcl::VoxelGrid    cubic leaf            reference src/GraphicEnd.cpp:287-295; src/saveOutput.cpp:44-46,76-79,90-93
 *   pcl::transformPointCloud                reference src/saveOutput.cpp:87
 * PCL itself is not installed (parity unpinned, see oracle_common.h): this follows upstream PCL 1.7
 * filters/voxel_grid.hpp applyFilter: min_b/max_b = floor(bbox * inverse_leaf) per axis, div_b = max_b-min_b+1,
 * voxel index of a point = ijk0 + ijk1*div_b0 + ijk2*div_b0*div_b1 with ijk = (int)(floor(x*inv_leaf) - (float)min_b),
 * all in float32; one centroid per occupied voxel, output sorted by voxel index.  Deviation (documented in
 * DESIGN.md): the centroid is summed in double in original point order (PCL: float, order left to an unstable
 * sort), so the result is at least as accurate and reproducible.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_common.h"

static int finite3(const float *p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }

/* out: capacity n*4 floats; returns the number of points kept */
int oracle_passthrough_z(const float *xyzw, int n, float z_min, float z_max, float *out)
{
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const float *p = xyzw + 4 * (size_t)i;
        if (finite3(p) && p[2] >= z_min && p[2] <= z_max) {
            out[4 * (size_t)k] = p[0]; out[4 * (size_t)k + 1] = p[1]; out[4 * (size_t)k + 2] = p[2]; out[4 * (size_t)k + 3] = 1.0f;
            ++k;
        }
    }
    return k;
}

typedef struct { uint32_t key; int idx; } vg_item;
static int vg_cmp(const void *a, const void *b)
{
    const vg_item *x = (const vg_item *)a, *y = (const vg_item *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* out: capacity n*4 floats; returns the number of voxels, or -1 when the indices would overflow (PCL refuses) */
int oracle_voxel_grid(const float *xyzw, int n, float leaf, float *out)
{
    const float inv = 1.0f / leaf;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int any = 0;
    for (int i = 0; i < n; ++i) {
        const float *p = xyzw + 4 * (size_t)i;
        if (!finite3(p)) continue;
        any = 1;
        for (int a = 0; a < 3; ++a) { if (p[a] < mn[a]) mn[a] = p[a]; if (p[a] > mx[a]) mx[a] = p[a]; }
    }
    if (!any) return 0;
    int min_b[3], div_b[3];
    long long prod = 1;
    for (int a = 0; a < 3; ++a) {
        const double flo = floor((double)(mn[a] * inv)), fhi = floor((double)(mx[a] * inv));
        if (fhi - flo + 1.0 > 2147483647.0 || fabs(flo) > 2.0e9 || fabs(fhi) > 2.0e9) return -1;
        min_b[a] = (int)flo; div_b[a] = (int)fhi - (int)flo + 1;
        prod *= div_b[a];
        if (prod > 2147483647LL) return -1;
    }
    vg_item *items = (vg_item *)malloc(sizeof(vg_item) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float *p = xyzw + 4 * (size_t)i;
        if (!finite3(p)) continue;
        const int i0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
        const int i1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
        const int i2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
        items[m].key = (uint32_t)(i0 + i1 * div_b[0] + i2 * div_b[0] * div_b[1]);
        items[m].idx = i; ++m;
    }
    qsort(items, (size_t)m, sizeof(vg_item), vg_cmp);
    int nv = 0;
    for (int s = 0; s < m;) {
        int e = s;
        double sx = 0, sy = 0, sz = 0;
        while (e < m && items[e].key == items[s].key) {
            const float *p = xyzw + 4 * (size_t)items[e].idx;
            sx += (double)p[0]; sy += (double)p[1]; sz += (double)p[2]; ++e;
        }
        const double c = (double)(e - s);
        out[4 * (size_t)nv] = (float)(sx / c); out[4 * (size_t)nv + 1] = (float)(sy / c); out[4 * (size_t)nv + 2] = (float)(sz / c);
        out[4 * (size_t)nv + 3] = 1.0f;
        ++nv; s = e;
    }
    free(items);
    return nv;
}

/* out = T * p with T the float32 cast of the row-major 4x4 double matrix (Matrix4f * point), orc_xform arithmetic */
void oracle_transform(const float *xyzw, int n, const double *T16, float *out)
{
    float T[12];
    for (int k = 0; k < 12; ++k) T[k] = (float)T16[k];
    for (int i = 0; i < n; ++i) {
        const float *p = xyzw + 4 * (size_t)i;
        float o[3];
        orc_xform(T, p[0], p[1], p[2], o);
        out[4 * (size_t)i] = o[0]; out[4 * (size_t)i + 1] = o[1]; out[4 * (size_t)i + 2] = o[2]; out[4 * (size_t)i + 3] = 1.0f;
    }
}

/* ---- per-point normals from the organised depth image (same definition as csrc/filters.cu) -------------- */
#include "../include/slam3d_b200.h"

static int dn_valid(const uint16_t *depth, int w, int h, const s3d_camera *cam, float z_max, int u, int v)
{
    if (u < 0 || v < 0 || u >= w || v >= h) return 0;
    uint16_t d = depth[(size_t)v * w + u];
    if (d == 0) return 0;
    if (z_max > 0.f) { float fz = (float)((double)d / cam->factor); return fz >= 0.f && fz <= z_max; }
    return 1;
}
static void dn_point(const uint16_t *depth, int w, const s3d_camera *cam, int u, int v, float *o)
{
    double z = (double)depth[(size_t)v * w + u] / cam->factor;
    double x = ((double)u - cam->cx) * z / cam->fx;
    double y = ((double)v - cam->cy) * z / cam->fy;
    o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;
}

/* xyzw_out, nrm_out: capacity width*height*4 floats each; returns the number of points */
int oracle_backproject_normals(const uint16_t *depth, int width, int height, const s3d_camera *cam, float z_max,
                               int step, float max_jump, float *xyzw_out, float *nrm_out)
{
    int k = 0;
    for (int v = 0; v < height; ++v) for (int u = 0; u < width; ++u) {
        if (!dn_valid(depth, width, height, cam, z_max, u, v)) continue;
        float p[3]; dn_point(depth, width, cam, u, v, p);
        float *po = xyzw_out + 4 * (size_t)k, *no = nrm_out + 4 * (size_t)k;
        po[0] = p[0]; po[1] = p[1]; po[2] = p[2]; po[3] = 1.0f;
        no[0] = no[1] = no[2] = no[3] = 0.0f;
        ++k;
        const int s = step;
        if (!(dn_valid(depth, width, height, cam, z_max, u - s, v) && dn_valid(depth, width, height, cam, z_max, u + s, v) &&
              dn_valid(depth, width, height, cam, z_max, u, v - s) && dn_valid(depth, width, height, cam, z_max, u, v + s))) continue;
        float l[3], r[3], t[3], b[3];
        dn_point(depth, width, cam, u - s, v, l); dn_point(depth, width, cam, u + s, v, r);
        dn_point(depth, width, cam, u, v - s, t); dn_point(depth, width, cam, u, v + s, b);
        if (fabsf(l[2] - p[2]) > max_jump || fabsf(r[2] - p[2]) > max_jump || fabsf(t[2] - p[2]) > max_jump || fabsf(b[2] - p[2]) > max_jump) continue;
        const float ax = r[0] - l[0], ay = r[1] - l[1], az = r[2] - l[2];
        const float bx = b[0] - t[0], by = b[1] - t[1], bz = b[2] - t[2];
        float nx = fmaf(ay, bz, -(az * by)), ny = fmaf(az, bx, -(ax * bz)), nz = fmaf(ax, by, -(ay * bx));
        const float l2 = fmaf(nz, nz, fmaf(ny, ny, nx * nx));
        if (!(l2 > 1e-24f)) continue;
        const float inv = 1.0f / sqrtf(l2);
        nx *= inv; ny *= inv; nz *= inv;
        const float dp = fmaf(nz, p[2], fmaf(ny, p[1], nx * p[0]));
        if (dp > 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        no[0] = nx; no[1] = ny; no[2] = nz; no[3] = 1.0f;
    }
    return k;
}
