// grid.cuh -- the cell function shared by the index build and every query
#pragma once
#include "context.h"
#include "common.cuh"

// fractional cell coordinate along one axis
__device__ __forceinline__ float grid_fcoord(float p, float o, float inv_cell) { return __fmul_rn(__fsub_rn(p, o), inv_cell); }

__device__ __forceinline__ int grid_clampi(float f, int n)
{
    int c = __float2int_rd(f);
    return min(max(c, 0), n - 1);
}

__device__ __forceinline__ int grid_cell_index(const GridParams &gp, float x, float y, float z)
{
    int cx = grid_clampi(grid_fcoord(x, gp.ox, gp.inv_cell), gp.nx);
    int cy = grid_clampi(grid_fcoord(y, gp.oy, gp.inv_cell), gp.ny);
    int cz = grid_clampi(grid_fcoord(z, gp.oz, gp.inv_cell), gp.nz);
    return (cz * gp.ny + cy) * gp.nx + cx;
}
