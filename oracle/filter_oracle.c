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
