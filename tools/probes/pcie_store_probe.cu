// Store patterns of a 3840x2880 float RGB frame written straight into page-locked host memory (what rtb_render does when the
// caller's frame is page-locked), without any rendering: how much of the host-store penalty of the render kernels is the PCIe
// write pattern?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/pcie_store_probe tools/probes/pcie_store_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int W = 3840, H = 2880, TW = W / 8, TH = H / 4;

// one warp = one 8x4 tile: four 96-byte row segments as 128-bit stores (24 lanes)
__global__ void k_tile(float4 *out, const unsigned *order, int n_tiles, int spin)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tiles) return;
    const unsigned tile = order ? order[warp] : warp;
    const int tx = tile % TW, ty = tile / TW;
    float acc = lane;
    for (int i = 0; i < spin; i++) acc = acc * 1.0001f + 0.5f; // stand-in for the render time of the tile
    if (lane < 24)
    {
        const int row = lane / 6, seg = lane % 6;
        const size_t f4 = ((size_t)(ty * 4 + row) * W + tx * 8) * 3 / 4 + seg;
        out[f4] = make_float4(acc, 1.f, 2.f, 3.f);
    }
}
// one CTA (4 warps) = four horizontally adjacent tiles: four 384-byte row segments, 96 threads x 128-bit
__global__ void k_group(float4 *out, const unsigned *order, int n_groups, int spin)
{
    const unsigned g = order ? order[blockIdx.x] : blockIdx.x;
    const int gx = g % (TW / 4), gy = g / (TW / 4);
    float acc = threadIdx.x;
    for (int i = 0; i < spin; i++) acc = acc * 1.0001f + 0.5f;
    __syncthreads();
    if (threadIdx.x < 96)
    {
        const int row = threadIdx.x / 24, seg = threadIdx.x % 24;
        const size_t f4 = ((size_t)(gy * 4 + row) * W + gx * 32) * 3 / 4 + seg;
        out[f4] = make_float4(acc, 1.f, 2.f, 3.f);
    }
}
int main()
{
    const size_t bytes = (size_t)W * H * 12;
    float4 *host = nullptr, *dev = nullptr;
    CK(cudaHostAlloc((void **)&host, bytes, cudaHostAllocPortable));
    CK(cudaMalloc((void **)&dev, bytes));
    const int n_tiles = TW * TH, n_groups = n_tiles / 4;
    std::vector<unsigned> perm(n_tiles), gperm(n_groups);
    for (int i = 0; i < n_tiles; i++) perm[i] = i;
    for (int i = 0; i < n_groups; i++) gperm[i] = i;
    std::mt19937 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    std::shuffle(gperm.begin(), gperm.end(), rng);
    unsigned *d_perm, *d_gperm;
    CK(cudaMalloc(&d_perm, n_tiles * 4)); CK(cudaMalloc(&d_gperm, n_groups * 4));
    CK(cudaMemcpy(d_perm, perm.data(), n_tiles * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_gperm, gperm.data(), n_groups * 4, cudaMemcpyHostToDevice));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    auto report = [&](const char *what, float ms) { printf("%-58s %7.3f ms  %6.1f GB/s\n", what, ms, bytes / ms * 1e-6); };
    for (int spin : {0, 12000})
    {
        printf("-- spin %d (a tile's stand-in compute; 0 = stores only)\n", spin);
        for (int target = 0; target < 2; target++)
        {
            float4 *out = target ? host : dev;
            const char *tn = target ? "host" : "HBM ";
            for (int mode = 0; mode < 4; mode++)
            {
                float best = 1e9f;
                for (int rep = 0; rep < 5; rep++)
                {
                    CK(cudaEventRecord(a));
                    if (mode == 0) k_tile<<<n_tiles / 4, 128>>>(out, nullptr, n_tiles, spin);
                    if (mode == 1) k_tile<<<n_tiles / 4, 128>>>(out, d_perm, n_tiles, spin);
                    if (mode == 2) k_group<<<n_groups, 128>>>(out, nullptr, n_groups, spin);
                    if (mode == 3) k_group<<<n_groups, 128>>>(out, d_gperm, n_groups, spin);
                    CK(cudaEventRecord(b));
                    CK(cudaEventSynchronize(b));
                    float ms; CK(cudaEventElapsedTime(&ms, a, b));
                    best = std::min(best, ms);
                }
                char what[128];
                const char *names[4] = {"96 B row segments, warp per tile, raster order", "96 B row segments, warp per tile, random order",
                                        "384 B row segments, CTA per 4 tiles, raster order", "384 B row segments, CTA per 4 tiles, random order"};
                snprintf(what, sizeof(what), "%s %s", tn, names[mode]);
                report(what, best);
            }
        }
    }
    float best = 1e9f;
    for (int rep = 0; rep < 5; rep++)
    {
        CK(cudaEventRecord(a));
        CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    report("cudaMemcpyAsync device -> host (copy engine)", best);
    return 0;
}
