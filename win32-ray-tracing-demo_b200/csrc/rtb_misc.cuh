// rtb_misc.cuh -- the non-template kernels of the library (sm_100a): tile-order counting sort, parity hooks
// (primary rays with traversal recording, ray batches), the PerformanceTest bounce loop, frame un-sharding.
// Included by exactly one translation unit (rtb_abi.cu).
#pragma once
#include "rtb_kernels.cuh"

namespace rtb {

// ---- heaviest-first tile order from the recorded costs: a counting sort over 128 log-scale buckets ----
__device__ __forceinline__ int costBucket(unsigned int c)
{ // 4 buckets per power of two, bucket 127 = heaviest
    if (c < 4) return (int)c;
    const int e = 31 - __clz(c);
    return e * 4 + (int)((c >> (e - 2)) & 3u) - 4;
}

// Sorting unit u: tile u, or -- group mode (FrameParams::group4) -- four adjacent tiles which one CTA renders together: it is
// busy for as long as its slowest tile, so that is the group's cost.  Groups count as four tiles in the histogram and take four
// consecutive entries of the order; all offsets stay in tiles.  group 1: tiles 4u .. 4u + 3 of a tile row; group 2: tiles
// (4 gy + k) * tiles_x + tx, k = 0 .. 3, of a tile column (u = gy * tiles_x + tx).
__device__ __forceinline__ uint4 unitTiles(int u, int group, int tiles_x)
{
    if (group == 2)
    {
        const int gy = u / tiles_x, tx = u - gy * tiles_x;
        const unsigned int t = (unsigned int)(4 * gy * tiles_x + tx), sx = (unsigned int)tiles_x;
        return make_uint4(t, t + sx, t + 2u * sx, t + 3u * sx);
    }
    return make_uint4(4u * u, 4u * u + 1u, 4u * u + 2u, 4u * u + 3u);
}
__device__ __forceinline__ unsigned int unitCost(const unsigned int *__restrict__ cost, int u, int group, int tiles_x)
{
    if (!group) return cost[u];
    const uint4 t = unitTiles(u, group, tiles_x);
    return max(max(cost[t.x], cost[t.y]), max(cost[t.z], cost[t.w]));
}

__global__ void k_cost_histogram(const unsigned int *__restrict__ cost, int n, unsigned int *__restrict__ hist, int group, int tiles_x)
{
    __shared__ unsigned int h[RTB_COST_BUCKETS];
    for (int i = threadIdx.x; i < RTB_COST_BUCKETS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&h[costBucket(unitCost(cost, i, group, tiles_x))], group ? 4u : 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < RTB_COST_BUCKETS; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}

// one thread: hist[] -> descending exclusive offsets in cursor[], hist[] cleared for the next frame.
// Also sizes the two latency-critical sets at the head of the order:
//   (host)      the `wide_count` heaviest tiles: one warp per PIXEL with warp-parallel leaf tests
//               (rtb_chain_wide.cuh).  Their recorded costs are scaled up generously (FrameParams::cost_scale),
//               so a tile that is in the set stays in it until it really becomes light;
//   n_heavy[0]  those plus the tiles within `heavy_buckets` quarter-octaves of the heaviest tile that is NOT in
//               the first set, at most heavy_limit of them: resumable
//               traversal (rtb_chain_sm.cuh).
// n_heavy[1] = FLOOR BUCKET of the order: every tile in a bucket at or below it is "light" and is placed in one
// common bucket, i.e. (nearly) in raster order instead of by cost.  With floor_delta > 0 the floor lies that many
// quarter-octaves above the median tile (9: tiles up to ~4.8x the median are light).  Heaviest-first only matters
// for the tiles that can stretch the end of the kernel; among the light ones raster order keeps neighbouring tiles
// -- neighbouring framebuffer rows -- in flight together.  When the kernels store straight into the caller's
// page-locked host frame that is what lets the 96-byte row segments of adjacent tiles combine into long PCIe
// writes: 4K SAH frame 6.95 -> 5.79 ms with the floor, against 5.48 ms into HBM (profiles/r01_e2e_floor_bucket.log).
// Frames rendered into device memory keep floor_delta = 0 (pure heaviest-first is 3 % faster there).
// heavy_alpha > 0 selects the latency-critical set by TIME instead of by distance from the heaviest tile: with every tile's
// recorded cost c (cycles it took in the throughput kernel, on a saturated SM) the kernel's balanced finishing time is about
// T* = sum(c) / resident warp slots; a tile whose own c exceeds heavy_alpha * T* would still be running when everything else
// is done, wherever it is started -- those go to the latency tier (at most heavy_limit of them).  Unlike "within 2x of the
// heaviest tile" this does not shrink the set of a shard that happens to hold one outlier tile.
__global__ void k_cost_offsets(unsigned int *__restrict__ hist, unsigned int *__restrict__ cursor, unsigned int *__restrict__ n_heavy,
                               int n_tiles, int heavy_buckets, int heavy_limit, int wide_count, int floor_delta, float heavy_alpha,
                               int warp_slots)
{
    // launched with RTB_COST_BUCKETS threads: the histogram is fetched (and cleared for the next frame) in parallel,
    // the short serial pass below then runs out of shared memory (18 -> ~4 us; it sits at the end of every frame)
    __shared__ unsigned int h[RTB_COST_BUCKETS];
    for (int b = threadIdx.x; b < RTB_COST_BUCKETS; b += blockDim.x) { h[b] = hist[b]; hist[b] = 0; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int floorBucket = 0;
        if (floor_delta > 0)
        {
            unsigned int below = 0;
            int median = 0;
            for (int b = 0; b < RTB_COST_BUCKETS; b++)
            {
                below += h[b];
                if (2u * below >= (unsigned int)n_tiles) { median = b; break; }
            }
            floorBucket = min(median + floor_delta, RTB_COST_BUCKETS - 1);
        }
        const unsigned int wide = (unsigned int)min(wide_count, n_tiles);
        unsigned int run = 0, heavy = wide;
        int top2 = -1; // bucket of the heaviest tile outside the wide set
        for (int b = RTB_COST_BUCKETS - 1; b >= 0; b--)
        {
            if (b >= floorBucket) cursor[b] = run; // buckets below the floor are placed through cursor[floorBucket]
            run += h[b];
            if (top2 < 0 && run > wide) top2 = b;
            if (top2 >= 0 && b >= top2 - heavy_buckets) heavy = run;
        }
        if (heavy_alpha > 0.f)
        {
            // centre of bucket b (costBucket: 4 buckets per power of two): 2^e * (1 + (q + 0.5) / 4), e = (b + 4) / 4, q = (b + 4) % 4
            double total = 0;
            for (int b = 4; b < RTB_COST_BUCKETS; b++) total += (double)h[b] * ldexp(1.0 + (((b + 4) & 3) + 0.5) * 0.25, (b + 4) >> 2);
            const double limit = (double)heavy_alpha * total / (double)max(warp_slots, 1);
            unsigned int r2 = 0;
            heavy = wide;
            for (int b = RTB_COST_BUCKETS - 1; b >= 4; b--)
            {
                if (ldexp(1.0 + (((b + 4) & 3) + 0.5) * 0.25, (b + 4) >> 2) < limit) break;
                r2 += h[b];
                heavy = max(heavy, r2);
            }
        }
        if (heavy < wide) heavy = wide;
        // a bucket may be cut: with whole buckets only, a well-filled bucket at the top (flat cost distributions)
        // would leave no latency-critical set at all
        if (heavy - wide > (unsigned int)heavy_limit) heavy = wide + (unsigned int)heavy_limit;
        n_heavy[0] = heavy;
        n_heavy[1] = (unsigned int)floorBucket;
    }
}

// Each block owns a contiguous chunk of tiles: count per bucket in shared memory, reserve the chunk's
// range of every bucket with ONE global atomic per bucket, then place the tiles with shared-memory
// atomics (the naive per-tile global atomic took 146 us on the 345,600 tiles of a 4K frame).
__global__ void k_cost_scatter(const unsigned int *__restrict__ cost, int n, unsigned int *__restrict__ cursor,
                               unsigned int *__restrict__ order, const unsigned int *__restrict__ n_heavy, int group, int tiles_x)
{
    const unsigned int span = group ? 4u : 1u;
    const int floor_bucket = (int)n_heavy[1]; // written by k_cost_offsets just before this launch
    __shared__ unsigned int cnt[RTB_COST_BUCKETS], base[RTB_COST_BUCKETS];
    const int per = (n + gridDim.x - 1) / gridDim.x;
    const int begin = blockIdx.x * per, end = min(begin + per, n);
    for (int i = threadIdx.x; i < RTB_COST_BUCKETS; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x) atomicAdd(&cnt[max(costBucket(unitCost(cost, i, group, tiles_x)), floor_bucket)], span);
    __syncthreads();
    for (int i = threadIdx.x; i < RTB_COST_BUCKETS; i += blockDim.x)
    {
        base[i] = cnt[i] ? atomicAdd(&cursor[i], cnt[i]) : 0u;
        cnt[i] = 0;
    }
    __syncthreads();
    for (int i = begin + threadIdx.x; i < end; i += blockDim.x)
    {
        const int b = max(costBucket(unitCost(cost, i, group, tiles_x)), floor_bucket);
        const unsigned int at = base[b] + atomicAdd(&cnt[b], span);
        if (group) reinterpret_cast<uint4 *>(order)[at >> 2] = unitTiles(i, group, tiles_x); // `at` is a multiple of 4
        else order[at] = (unsigned int)i;
    }
}

// ---- k-d leaf lists on even positions of the reference array (upload time) --------------------------------------------------
// The list scans test two list positions per iteration out of the pair stream (pre2[p] = positions 2p, 2p + 1); a list that
// starts on an odd position costs half an iteration more on average -- one of the ~17 pair iterations per ray on the SAH tree.
// A leaf carries its own count, so the array can simply be re-laid out with every list on an even position: the slot behind an
// odd-length list is padding that lies outside every [first, first + count) and is never tested.  size[i] = padded length of
// leaf i (inner nodes 0); first = exclusive scan; the caller's layout is kept for rtb_scene_kd_download.
__global__ void k_kd_pad_sizes(const uint2 *__restrict__ nodes, int n, unsigned int *__restrict__ size)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    unsigned int v = 0;
    if (i < n)
    {
        const uint2 nd = nodes[i];
        if ((nd.y & 3u) == 3u) v = ((nd.y >> 2) + 1u) & ~1u;
    }
    size[i] = v; // size[n] = 0: the total of the scan lands in first[n]
}

__global__ void k_kd_relayout(const uint2 *__restrict__ nodes, int n, const unsigned int *__restrict__ first, const uint32_t *__restrict__ refs,
                              uint2 *__restrict__ nodes_out, uint32_t *__restrict__ refs_out, unsigned int cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint2 nd = nodes[i];
    if ((nd.y & 3u) == 3u)
    {
        const unsigned int count = nd.y >> 2, from = nd.x, to = first[i];
        if ((unsigned long long)to + ((count + 1u) & ~1u) <= cap) // (the capacity is the exact total: always true)
        {
            for (unsigned int j = 0; j < count; j++) refs_out[to + j] = refs[from + j];
            if (count & 1u) refs_out[to + count] = refs[from + count - 1u]; // padding: any valid index
            nd.x = to;
        }
        else nd.y = 3u; // never reached; an empty leaf rather than a list outside the array
    }
    nodes_out[i] = nd;
}

// ---------------------------------------------------------------------------------------------
// Parity hooks: primary rays with traversal recording, and arbitrary ray batches
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RTB_CTA_THREADS)
k_trace_primary(const __grid_constant__ DScene S, const __grid_constant__ FrameParams F, int *__restrict__ hit_id,
                float *__restrict__ hit_t, int *__restrict__ seq_len, unsigned long long *__restrict__ seq_hash,
                int *__restrict__ seq_buf, int seq_cap)
{
    int x, lr, y;
    unsigned int tile;
    if (!pixelOfThread(F, x, lr, y, tile)) return;
    const float dx = 1.0f / F.height, dy = 1.0f / F.height;
    const float sx = (x + 0.5f) * dx, sy = 1 - (y + 0.5f) * dy;
    const Ray r = generateRay(F.cam, sx, sy);
    const size_t p = (size_t)y * F.width + x;
    SeqProbe pr;
    pr.buf = seq_buf ? seq_buf + p * seq_cap : nullptr;
    pr.cap = seq_buf ? seq_cap : 0;
    Hit h;
    const bool ok = sceneIntersect(S, r, h, pr);
    if (hit_id) hit_id[p] = ok ? h.id : -1;
    if (hit_t) hit_t[p] = ok ? h.t : -1.0f;
    if (seq_len) seq_len[p] = pr.len;
    if (seq_hash) seq_hash[p] = pr.hash;
    if (seq_buf)
        for (int i = pr.len; i < seq_cap; i++) seq_buf[p * seq_cap + i] = -1;
}

__global__ void __launch_bounds__(RTB_CTA_THREADS)
k_intersect_rays(const __grid_constant__ DScene S, long long n, const float *__restrict__ rays, int *__restrict__ hit_id,
                 float *__restrict__ hit_t, float *__restrict__ position, float *__restrict__ normal)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r;
    r.o = v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    r.d = v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
    NoProbe pr;
    Hit h;
    const bool ok = sceneIntersect(S, r, h, pr);
    if (hit_id) hit_id[i] = ok ? h.id : -1;
    if (hit_t) hit_t[i] = ok ? h.t : -1.0f;
    if (position) { position[3 * i] = ok ? h.pos.x : 0; position[3 * i + 1] = ok ? h.pos.y : 0; position[3 * i + 2] = ok ? h.pos.z : 0; }
    if (normal) { normal[3 * i] = ok ? h.n.x : 0; normal[3 * i + 1] = ok ? h.n.y : 0; normal[3 * i + 2] = ok ? h.n.z : 0; }
}

// ---------------------------------------------------------------------------------------------
// PerformanceTest workload (reference src/PerformanceTest/main.cpp:29-59): mirror-reflect each ray until
// it hits a PLANE geometry (the plane closing the tunnel exit), misses, or exceeds max_depth.
// ---------------------------------------------------------------------------------------------
// WIDE = false: one lane per ray (batches that fill the device).  WIDE = true: one WARP per ray -- the 32 lanes walk the
// accelerator in lock step and test 32 triangles of a leaf / cell list at once (nearestInList<.., WIDE>), as in the
// warp-per-pixel render tier: the reference program's own batch is 1000 rays, which as 1000 threads occupies 8 CTAs and
// takes as long as its longest 200-bounce chain walked by one lane; as 1000 warps it fills the device.  Results are
// identical (same arithmetic; ties in a list go to the earliest position).  k-d trees and grids only.
template <bool WIDE>
__global__ void __launch_bounds__(RTB_CTA_THREADS)
k_bounce_rays(const __grid_constant__ DScene S, long long n, const float *__restrict__ rays, int max_depth,
              int *__restrict__ reached, int *__restrict__ depth_out, int *__restrict__ last_id, float *__restrict__ last_pos,
              unsigned long long *__restrict__ total_rays, unsigned long long *__restrict__ next_ray)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool writer = !WIDE || (threadIdx.x & 31) == 0;
    unsigned int traced = 0;
    // One lane per ray: the chains are 1 .. 200 bounces long (35 on average), so a lane that has finished its ray takes the
    // next one from a queue (`next_ray`, one atomic per ray) instead of idling until the longest chain of its warp ends.
    // The launch is sized to the device, not to the batch.  Rays are independent: the order changes no result.
    for (long long i = WIDE ? (t >> 5) : (long long)atomicAdd(next_ray, 1ull); i < n; i = WIDE ? n : (long long)atomicAdd(next_ray, 1ull))
    {
        Ray r;
        r.o = v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        r.d = v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        NoProbe pr;
        int depth = 0, ok = 0, id = -1;
        V3 pos = v3(0, 0, 0);
        RayCtx ctx; // newRay.context = r.context, PerformanceTest/main.cpp:57
        ctx.inTunnel = 0; ctx.segment = -1;
        while (true)
        {
            Hit h;
            traced++;
            bool hit;
            if (WIDE)
                hit = sceneIntersectWith(S, r, h, pr, [&](int &tri, float &tt, V3 &nn) {
                    if (S.accel == RTB_ACCEL_REGULAR_GRID || S.accel == RTB_ACCEL_FLAT_GRID) return gridIntersect<true>(S, r, tri, tt, nn, pr);
                    return kdIntersect<true>(S, r, tri, tt, nn, pr);
                });
            else hit = sceneIntersect(S, r, h, pr, ctx);
            if (!hit) { id = -1; break; }
            id = h.id; pos = h.pos;
            const V3 nl = (dot(h.n, r.d) < 0) ? h.n : h.n * -1;
            if (++depth > max_depth) break;
            if (h.prim_type == RTB_PRIM_PLANE) { ok = 1; break; }
            const V3 v = r.d - nl * 2 * dot(nl, r.d);
            r.o = h.pos;
            r.d = v;
        }
        if (writer)
        {
            if (reached) reached[i] = ok;
            if (depth_out) depth_out[i] = depth;
            if (last_id) last_id[i] = id;
            if (last_pos) { last_pos[3 * i] = pos.x; last_pos[3 * i + 1] = pos.y; last_pos[3 * i + 2] = pos.z; }
        }
    }
    if (!writer) traced = 0;
    traced = __reduce_add_sync(0xffffffffu, traced);
    if ((threadIdx.x & 31) == 0 && traced) atomicAdd(total_rays, (unsigned long long)traced);
}

// ---------------------------------------------------------------------------------------------
// Un-shard: after the NCCL all-gather the frame lives as [world][rows_per_rank][width][3] in
// block-cyclic row order; this puts it back into image row order [height][width][3].
// Pure 128-bit copies when width*3 floats is a multiple of 4 (every 4:3 frame is).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_unshard(const float *__restrict__ gathered, float *__restrict__ image, int width, int height, int world, int row_block,
          int rows_per_rank)
{
    const int y = blockIdx.y;
    const int b = y / row_block, rank = b % world, lb = b / world;
    const int lr = lb * row_block + (y - b * row_block);
    const size_t rowFloats = (size_t)width * 3;
    const float *src = gathered + ((size_t)rank * rows_per_rank + lr) * rowFloats;
    float *dst = image + (size_t)y * rowFloats;
    if ((rowFloats & 3) == 0)
    {
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rowFloats / 4; i += (size_t)gridDim.x * blockDim.x) d4[i] = s4[i];
    }
    else
        for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rowFloats; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Column-block shards (localToGlobal): gathered = [world][height][width / world][3]; one thread per float4 of a
// column block row segment (col_block is a multiple of 8 pixels = 96 bytes)
__global__ void __launch_bounds__(256)
k_unshard_cols(const float *__restrict__ gathered, float *__restrict__ image, int width, int height, int world, int row_block,
               int col_block)
{
    const int y = blockIdx.y;
    const int by = y / row_block;
    const int localWidth = width / world;
    const int segFloat4 = col_block * 3 / 4, nbx = width / col_block;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbx * segFloat4; i += gridDim.x * blockDim.x)
    {
        const int bx = i / segFloat4, q = i - bx * segFloat4;
        const int rank = (bx + by) % world;
        const int xl0 = (bx / world) * col_block;
        const float4 *src = reinterpret_cast<const float4 *>(gathered + (((size_t)rank * height + y) * localWidth + xl0) * 3);
        float4 *dst = reinterpret_cast<float4 *>(image + ((size_t)y * width + (size_t)bx * col_block) * 3);
        dst[q] = src[q];
    }
}

// A rank's local shard image -> its place in the whole frame (RTB_LAYOUT_GLOBAL addressing), as ONE streaming pass of
// 128-bit copies: the alternative to storing tile by tile into a peer-mapped frame while rendering (rtb_scatter_shard_device).
// One thread per float4; a local row is [local_width][3] floats; column-block shards move whole col_block segments.
__global__ void __launch_bounds__(256)
k_scatter_shard(const float *__restrict__ local, float *__restrict__ frame, int width, int height, int world, int rank, int row_block,
                int col_block, int local_width, int n_local_rows)
{
    const int lr = blockIdx.y;
    if (lr >= n_local_rows) return;
    int y;
    if (col_block) y = lr;
    else { const int lb = lr / row_block; y = (lb * world + rank) * row_block + (lr - lb * row_block); }
    if (y >= height) return;
    const int rowF4 = local_width * 3 / 4; // local_width % 4 == 0 (checked by the caller)
    const float4 *src = reinterpret_cast<const float4 *>(local + (size_t)lr * local_width * 3);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rowF4; i += gridDim.x * blockDim.x)
    {
        int gx4 = i; // float4 index inside the frame row
        if (col_block)
        {
            const int segF4 = col_block * 3 / 4, c = i / segF4, q = i - c * segF4;
            int shift = (rank - y / row_block) % world;
            if (shift < 0) shift += world;
            gx4 = (c * world + shift) * segF4 + q;
        }
        reinterpret_cast<float4 *>(frame + (size_t)y * width * 3)[gx4] = src[i];
    }
}

} // namespace rtb
