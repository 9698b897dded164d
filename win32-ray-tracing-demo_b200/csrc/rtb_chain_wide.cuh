// rtb_chain_wide.cuh -- Whitted reflection chains, ONE WARP PER PIXEL (sm_100a).
//
// Same arithmetic and same results as k_whitted_chain (rtb_kernels.cuh): the k-d walk is reference
// Tunnel.cpp:1163-1297, the grid walk Tunnel.cpp:819-970, shading MainWindow.cpp:69-143.  What changes is
// who does the work.  The pixels at the vanishing point of a tunnel frame run 21-ray chains of up to
// ~3,600 sequential accelerator steps and triangle tests (regular grid: 21,000, cells list up to 193
// triangles); that single dependency chain, not throughput, bounds a small frame or a multi-GPU shard
// (profiles/r01_cost_map.md).  Nothing in a chain is parallel except the triangle tests of one leaf / cell:
// here all 32 lanes of a warp hold the same ray, walk the accelerator in lock step (no divergence at all)
// and test 32 triangles of the current list at once (nearestInList<.., WIDE = true>, rtb_device.cuh).
// The critical path of a ray drops from (steps + tests) to (steps + lists / 32) serial latencies.
// 32 x the warps per tile, so only the few heaviest tiles of the heaviest-first order -- the first
// F.n_wide entries (rtb_abi.cu: wideCount) -- are rendered this way, on the high-priority side stream.
#pragma once
#include "rtb_kernels.cuh"

namespace rtb {

// 128-thread CTAs (4 pixel-warps): they fit wherever a CTA of the throughput kernels retires.  Whole-SM CTAs
// (1024 threads at 64 registers, no co-residents) measured ~15 % faster when they get their SMs at once and
// twice as slow whenever the other kernels of the frame reach the SMs first -- not kept.
template <class Probe, int FOLD>
__global__ void __launch_bounds__(RTB_CTA_THREADS, 4)
k_whitted_chain_wide(const __grid_constant__ DScene S, const __grid_constant__ FrameParams F, float *__restrict__ out,
                     Counters *__restrict__ counters)
{
    const long long t_start = clock64();
    const unsigned int w = blockIdx.x * (unsigned int)F.warps_per_cta + (threadIdx.x >> 5);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned int item = w >> 5; // 32 warps per tile, warp (w & 31) renders pixel (w & 31) of the tile
    if (item >= F.n_wide || item >= (unsigned int)F.n_tiles) return;
    const unsigned int tile = __ldg(F.order + item);
    const int ty = tile / F.tiles_x, tx = tile - ty * F.tiles_x;
    const int lr = ty * RTB_TILE_H + (int)((w >> 3) & 3u);
    int x, y;
    if (!localToGlobal(F, tx * RTB_TILE_W + (int)(w & 7u), lr, x, y)) return; // warp-uniform
    unsigned int rays = 0;
    Probe prTop, prWalk; // prTop: work every lane repeats (top-level geometries); prWalk: the tunnel walk
    const V3 c = chainWith<FOLD>(S, F, x, y, rays, prTop, [&](const Ray &r, Hit &h) {
        return sceneIntersectWith(S, r, h, prTop, [&](int &tri, float &t, V3 &n) {
            if (S.accel == RTB_ACCEL_REGULAR_GRID || S.accel == RTB_ACCEL_FLAT_GRID) return gridIntersect<true>(S, r, tri, t, n, prWalk);
            return kdIntersect<true>(S, r, tri, t, n, prWalk);
        });
    });
    if (lane == 0) storePixel(F, out, x, lr, y, c, t_start, rays, prWalk);
    // counters: accelerator steps, rays and top-level tests are identical on all lanes (count lane 0's);
    // the triangle tests of the walk are distinct per lane (sum them)
    unsigned int tris = ProbeCounts<Probe>::tris(prWalk) + (lane == 0 ? ProbeCounts<Probe>::tris(prTop) : 0u);
    tris = __reduce_add_sync(0xffffffffu, tris);
    if (lane == 0)
    {
        // no cost is recorded: the tile keeps the cost the throughput kernel measured (FrameParams::record_cost)
        const unsigned int steps = ProbeCounts<Probe>::steps(prWalk);
        if (rays) atomicAdd(&counters->rays, (unsigned long long)rays);
        if (tris) atomicAdd(&counters->tris, (unsigned long long)tris);
        if (steps) atomicAdd(&counters->steps, (unsigned long long)steps);
        if ((w & 31u) == 0u) atomicAdd(&counters->tiles, 1ull); // 32 warps per tile
    }
}

} // namespace rtb
