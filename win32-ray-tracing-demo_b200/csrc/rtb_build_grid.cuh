// rtb_build_grid.cuh -- the reference's grid builder (Tunnel.cpp:346-465) on the device (sm_100a).
//
// rtb_scene_upload builds the grid here when the caller passes a grid accelerator WITHOUT the grid arrays and a
// resolution (rtb_flat_scene::grid_build_resolution): nothing of the accelerator crosses PCIe, and the 400^3
// flat grid of the 150-segment tunnel (3.4 M triangle references in 1.25 M of 64 M cells) takes ~2 ms instead of
// 150 ms on the host (2.6 s in the reference, which allocates 64 M std::vectors).
//
// The result is the same structure the host builder emits, bit for bit (tests: the canonical structure hash):
//   sizing     bounds of all vertices -> cell size / origin / dims with the reference's float expressions
//              (evaluated on the host from the 6 reduced floats: one tiny D2H);
//   binning    a triangle goes to every cell its AABB overlaps (the exact test is disabled in the reference,
//              Tunnel.cpp:435-445): count per triangle -> exclusive scan -> emit (cell << 32 | triangle) keys;
//   ordering   radix sort of the keys: cells ascending, and inside a cell triangles ascending -- the order in
//              which the reference's push_back loop fills each cell's list;
//   directory  run heads -> cell_start[]; occupancy bit per cell; popcount prefix per 32-cell word.
#pragma once
#include <cub/cub.cuh>
#include "rtb_device.cuh"
#include "rtb_sat.h"

namespace rtb {

struct GridSizing { float origin[3], cell[3]; int dims[3]; int exact; };

// float atomics through the ordered-integer trick (all values finite)
__device__ __forceinline__ void atomicMinFloat(float *addr, float v)
{
    if (v >= 0) atomicMin((int *)addr, __float_as_int(v));
    else atomicMax((unsigned int *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxFloat(float *addr, float v)
{
    if (v >= 0) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned int *)addr, __float_as_uint(v));
}

// bounds[0..2] = min, bounds[3..5] = max over the vertices a, b, c of all triangles ([n][12] floats: a, b, c, normal)
__global__ void k_grid_bounds(const float *__restrict__ tri, int n, float *__restrict__ bounds)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float *t = tri + 12ull * i;
#pragma unroll
        for (int v = 0; v < 3; v++)
#pragma unroll
            for (int a = 0; a < 3; a++)
            {
                mn[a] = fminf(mn[a], t[3 * v + a]);
                mx[a] = fmaxf(mx[a], t[3 * v + a]);
            }
    }
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMinFloat(bounds + a, mn[a]);
            atomicMaxFloat(bounds + 3 + a, mx[a]);
        }
    }
}

// the cell range a triangle's AABB overlaps: (int)((v - origin) / cell) per axis, Tunnel.cpp:421-432
__device__ __forceinline__ void triCellRange(const float *t, const GridSizing &G, int lo[3], int hi[3])
{
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        const float mn = fminf(fminf(t[a], t[3 + a]), t[6 + a]), mx = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
        lo[a] = f2i((mn - G.origin[a]) / G.cell[a]);
        hi[a] = f2i((mx - G.origin[a]) / G.cell[a]);
    }
}

// exact binning (GridSizing::exact): the cell the reference would construct for (x, y, z) -- position origin + index * size
// per axis, size per axis (Tunnel.cpp:437) -- against the triangle, rtb_sat.h
__device__ __forceinline__ bool cellTakes(const float *t, const GridSizing &G, int x, int y, int z)
{
    if (!G.exact) return true;
    const float pos[3] = {G.origin[0] + x * G.cell[0], G.origin[1] + y * G.cell[1], G.origin[2] + z * G.cell[2]};
    return rtb_sat::triangleOverlapsCell(t, pos, G.cell);
}

__global__ void k_grid_count(const float *__restrict__ tri, int n, const __grid_constant__ GridSizing G, unsigned int *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo[3], hi[3];
    triCellRange(tri + 12ull * i, G, lo, hi);
    if (!G.exact)
    {
        count[i] = (unsigned int)(hi[0] - lo[0] + 1) * (unsigned int)(hi[1] - lo[1] + 1) * (unsigned int)(hi[2] - lo[2] + 1);
        return;
    }
    float t[12];
    for (int k = 0; k < 12; k++) t[k] = tri[12ull * i + k];
    unsigned int c = 0;
    for (int x = lo[0]; x <= hi[0]; x++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int z = lo[2]; z <= hi[2]; z++) c += cellTakes(t, G, x, y, z) ? 1u : 0u;
    count[i] = c;
}

__global__ void k_grid_emit(const float *__restrict__ tri, int n, const __grid_constant__ GridSizing G, const unsigned int *__restrict__ offset,
                            unsigned long long *__restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo[3], hi[3];
    triCellRange(tri + 12ull * i, G, lo, hi);
    float t[12];
    for (int k = 0; k < 12; k++) t[k] = tri[12ull * i + k];
    unsigned int e = offset[i];
    for (int x = lo[0]; x <= hi[0]; x++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int z = lo[2]; z <= hi[2]; z++)
            {
                if (!cellTakes(t, G, x, y, z)) continue;
                const unsigned long long cell = (unsigned long long)((x * G.dims[1] + y) * G.dims[2] + z); // Tunnel.h:63-66
                keys[e++] = (cell << 32) | (unsigned int)i;
            }
}

// sorted keys -> triangle reference list, run heads, occupancy bits
__global__ void k_grid_heads(const unsigned long long *__restrict__ keys, unsigned int n, uint32_t *__restrict__ cell_tris,
                             unsigned int *__restrict__ head, uint2 *__restrict__ words)
{
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long k = keys[e];
    const unsigned int cell = (unsigned int)(k >> 32);
    cell_tris[e] = (uint32_t)k;
    const bool first = e == 0 || (unsigned int)(keys[e - 1] >> 32) != cell;
    head[e] = first ? 1u : 0u;
    if (first) atomicOr(&words[cell >> 5].x, 1u << (cell & 31));
}

// cell_start[rank of the run] = first entry of the run; cell_start[n_runs] = n
__global__ void k_grid_starts(const unsigned int *__restrict__ head, const unsigned int *__restrict__ rank, unsigned int n,
                              uint32_t *__restrict__ cell_start, unsigned int n_runs)
{
    const unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) cell_start[n_runs] = n;
    if (e >= n) return;
    if (head[e]) cell_start[rank[e]] = e;
}

__global__ void k_grid_word_pop(const uint2 *__restrict__ words, long long n, unsigned int *__restrict__ pop)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n) pop[w] = __popc(words[w].x);
}
__global__ void k_grid_word_rank(uint2 *__restrict__ words, long long n, const unsigned int *__restrict__ rank)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n) words[w].y = rank[w];
}

} // namespace rtb
