// rtb_build_kd.cuh -- the reference's SAH k-d tree builder (Tunnel.cpp:546-638 buildKdTree, 671-784 splitSAH) on the
// device (sm_100a).
//
// rtb_scene_upload builds the tree here when the caller passes RTB_ACCEL_KD_SAH WITHOUT node arrays and with build
// parameters (rtb_flat_scene::kd_build_max_depth > 0): nothing of the accelerator crosses PCIe, and the 150-segment tunnel's
// tree (68,955 nodes, 379,888 leaf references) takes milliseconds instead of 0.11 s on the host's cores (7-9.5 s in the
// reference, single-threaded).  The result is the array the host builder emits, bit for bit (tests: node and leaf arrays
// compared directly), because every float expression is the reference's:
//   candidates   min[axis] + (max[axis] - min[axis]) * i / N, i = 1 .. N - 1 (N = 100), per axis;
//   counts       leftCount = #{triangles with a vertex < s} = #{lo < s}, rightCount = #{hi >= s}: each triangle is binned
//                once per axis and side by binary search over the candidates (they are non-decreasing in i) and two
//                prefix sums give the counts of all 99 planes -- the "segmented reduction" over the node's list;
//   cost         (lw h + lw d + h d) leftCount + (rw h + rw d + h d) rightCount, first strict minimum over axis 0 -> 2,
//                i ascending;
//   children     left takes lo < split, right takes hi >= split, both in list order (stable compaction); child boxes
//                inherit the parent's with max[axis] / min[axis] = split; leaf iff list <= leafSize or depth > maxDepth.
// The build is level-synchronous: one CTA per node of the level evaluates its candidates (k_kd_eval), a scan sizes the
// next level, one CTA per node partitions (k_kd_partition).  Afterwards two sweeps over the levels compute subtree sizes
// and the pre-order position of every node and leaf list, and k_kd_emit writes the 8-byte node layout of include/rtb.h.
#pragma once
#include <cub/cub.cuh>
#include "rtb_device.cuh"

namespace rtb {

#define RTB_KDB_MAX_CAND 128 // candidates per axis the shared arrays hold (reference: N = 100)
#define RTB_KDB_THREADS 256
#define RTB_KDB_MAX_LEVELS 32

struct KdBuildNode
{
    float mn[3], mx[3]; // the node's box (Tunnel.h:86-87)
    int begin, count;   // its list: a segment of the level's reference array
    int depth;
    int axis;           // 0..2 inner, 3 leaf
    float split;
    int left, right;    // children (indices into the node table, which is in level order)
    int lc, rc;         // sizes of the children's lists
    int level, poolOff; // leaf: its list, copied to the level's leaf chunk at this offset
    int subNodes, subRefs; // nodes / leaf references in the subtree
    int pre, leafOff;      // pre-order index; position of the subtree's first leaf reference
};

struct KdLeafChunks { const uint32_t *p[RTB_KDB_MAX_LEVELS]; };

// lo / hi per triangle and axis (SoA: [axis][n]) -- min / max over the vertices a, b, c as Tunnel.cpp:748-758 compares them
__global__ void k_kd_extents(const float *__restrict__ tri, int n, float *__restrict__ lo, float *__restrict__ hi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *t = tri + 12ull * i;
#pragma unroll
    for (int a = 0; a < 3; a++)
    {
        lo[(size_t)a * n + i] = fminf(fminf(t[a], t[3 + a]), t[6 + a]);
        hi[(size_t)a * n + i] = fmaxf(fmaxf(t[a], t[3 + a]), t[6 + a]);
    }
}

__global__ void k_kd_root(KdBuildNode *__restrict__ nodes, const float *__restrict__ bounds, int n, uint32_t *__restrict__ refs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) refs[i] = (uint32_t)i; // Tunnel.cpp:474-484: the list starts in surface order
    if (i == 0)
    {
        KdBuildNode r;
        for (int a = 0; a < 3; a++) { r.mn[a] = bounds[a]; r.mx[a] = bounds[3 + a]; }
        r.begin = 0; r.count = n; r.depth = 0; r.axis = 3; r.split = 0; r.left = r.right = -1; r.lc = r.rc = 0;
        r.level = 0; r.poolOff = 0; r.subNodes = r.subRefs = 0; r.pre = 0; r.leafOff = 0;
        nodes[0] = r;
    }
}

// first index i in [1, N) with cand[i] > v, N if there is none (std::upper_bound over cand[1 .. N))
__device__ __forceinline__ int kdUpperBound(const float *cand, int N, float v)
{
    int lo = 1, hi = N;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (cand[mid] > v) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// One CTA per node of the level: leaf decision, or splitSAH (Tunnel.cpp:671-784)
__global__ void __launch_bounds__(RTB_KDB_THREADS)
k_kd_eval(KdBuildNode *__restrict__ nodes, int first, const uint32_t *__restrict__ refs, const float *__restrict__ lo,
          const float *__restrict__ hi, int n, int leafSize, int maxDepth, int N)
{
    KdBuildNode &nd = nodes[first + blockIdx.x];
    const int tid = threadIdx.x, count = nd.count;
    if (count <= leafSize || nd.depth > maxDepth)
    { // Tunnel.cpp:550
        if (tid == 0) { nd.axis = 3; nd.lc = nd.rc = 0; }
        return;
    }
    __shared__ float cand[3][RTB_KDB_MAX_CAND];
    __shared__ int leftFrom[3][RTB_KDB_MAX_CAND + 1], rightTo[3][RTB_KDB_MAX_CAND + 1];
    __shared__ float bestCost[3], bestSplit[3];
    __shared__ int bestL[3], bestR[3];
    for (int k = tid; k < 3 * (N + 1); k += blockDim.x)
    {
        const int a = k / (N + 1), i = k - a * (N + 1);
        leftFrom[a][i] = 0; rightTo[a][i] = 0;
        if (i >= 1 && i < N) cand[a][i] = nd.mn[a] + (nd.mx[a] - nd.mn[a]) * i / N; // Tunnel.cpp:685-686
    }
    __syncthreads();
    for (int j = tid; j < count; j += blockDim.x)
    {
        const uint32_t t = refs[nd.begin + j];
#pragma unroll
        for (int a = 0; a < 3; a++)
        { // the triangle is "left" of every candidate > lo and "right" of every candidate <= hi
            atomicAdd(&leftFrom[a][kdUpperBound(cand[a], N, lo[(size_t)a * n + t])], 1);
            atomicAdd(&rightTo[a][kdUpperBound(cand[a], N, hi[(size_t)a * n + t])], 1);
        }
    }
    __syncthreads();
    if (tid < 3)
    {
        const int axis = tid, nextAxis = (axis + 1) % 3, prevAxis = (axis + 2) % 3; // Tunnel.cpp:735-736
        const float height = nd.mx[nextAxis] - nd.mn[nextAxis], depth = nd.mx[prevAxis] - nd.mn[prevAxis]; // Vector(min, max)
        float minSAH = FLT_MAX, minSplit = 0;
        int leftCount = 0, rightCount = count, lBest = 0, rBest = 0;
        for (int i = 1; i < N; i++)
        {
            leftCount += leftFrom[axis][i];
            rightCount -= rightTo[axis][i];
            const float splitValue = cand[axis][i];
            const float leftWidth = splitValue - nd.mn[axis], rightWidth = nd.mx[axis] - splitValue;
            const float SAH = (leftWidth * height + leftWidth * depth + height * depth) * leftCount +
                              (rightWidth * height + rightWidth * depth + height * depth) * rightCount; // Tunnel.cpp:769-771
            if (SAH < minSAH) { minSAH = SAH; minSplit = splitValue; lBest = leftCount; rBest = rightCount; }
        }
        bestCost[axis] = minSAH; bestSplit[axis] = minSplit; bestL[axis] = lBest; bestR[axis] = rBest;
    }
    __syncthreads();
    if (tid == 0)
    {
        float minSAH = FLT_MAX, split = 0;
        int axis = -1, lc = 0, rc = 0;
        for (int a = 0; a < 3; a++)
            if (bestCost[a] < minSAH) { minSAH = bestCost[a]; split = bestSplit[a]; axis = a; lc = bestL[a]; rc = bestR[a]; }
        if (axis < 0)
        { // no candidate beat FLT_MAX (degenerate boxes / NaN): the reference leaves axis and split unset here; the host
          // builder of this library takes axis = depth % 3, split = 0 -- the same here, counted directly
            axis = nd.depth % 3; split = 0; lc = 0; rc = 0;
            for (int j = 0; j < count; j++)
            {
                const uint32_t t = refs[nd.begin + j];
                lc += lo[(size_t)axis * n + t] < split ? 1 : 0;
                rc += hi[(size_t)axis * n + t] >= split ? 1 : 0;
            }
        }
        nd.axis = axis; nd.split = split; nd.lc = lc; nd.rc = rc;
    }
}

// One block: exclusive scans over the level's nodes of (children list sizes, leaf list sizes, inner flags); totals -> out[0..2]
__global__ void __launch_bounds__(1024)
k_kd_level_scan(const KdBuildNode *__restrict__ nodes, int first, int count, int *__restrict__ childOff, int *__restrict__ leafOff,
                int *__restrict__ innerRank, int *__restrict__ totals)
{
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    int run0 = 0, run1 = 0, run2 = 0;
    for (int base = 0; base < count; base += 1024)
    {
        const int k = base + threadIdx.x;
        int a = 0, b = 0, c = 0;
        if (k < count)
        {
            const KdBuildNode &nd = nodes[first + k];
            if (nd.axis == 3) b = nd.count;
            else { a = nd.lc + nd.rc; c = 1; }
        }
        int ea, eb, ec, ta, tb, tc;
        Scan(tmp).ExclusiveSum(a, ea, ta); __syncthreads();
        Scan(tmp).ExclusiveSum(b, eb, tb); __syncthreads();
        Scan(tmp).ExclusiveSum(c, ec, tc); __syncthreads();
        if (k < count) { childOff[k] = run0 + ea; leafOff[k] = run1 + eb; innerRank[k] = run2 + ec; }
        run0 += ta; run1 += tb; run2 += tc;
    }
    if (threadIdx.x == 0) { totals[0] = run0; totals[1] = run1; totals[2] = run2; }
}

// One CTA per node of the level: a leaf copies its list to the level's leaf chunk; an inner node compacts its list, in
// order, into the children's lists (Tunnel.cpp:616-634) and creates the children (596-608)
__global__ void __launch_bounds__(RTB_KDB_THREADS)
k_kd_partition(KdBuildNode *__restrict__ nodes, int first, int level, int childBase, const uint32_t *__restrict__ refs,
               uint32_t *__restrict__ next, uint32_t *__restrict__ leafChunk, const int *__restrict__ childOff,
               const int *__restrict__ leafOff, const int *__restrict__ innerRank, const float *__restrict__ lo,
               const float *__restrict__ hi, int n)
{
    typedef cub::BlockScan<int, RTB_KDB_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    const int k = blockIdx.x, tid = threadIdx.x;
    KdBuildNode &nd = nodes[first + k];
    const int count = nd.count;
    if (nd.axis == 3)
    {
        for (int j = tid; j < count; j += blockDim.x) leafChunk[leafOff[k] + j] = refs[nd.begin + j];
        if (tid == 0) { nd.level = level; nd.poolOff = leafOff[k]; nd.left = nd.right = -1; }
        return;
    }
    const int axis = nd.axis;
    const float split = nd.split;
    const int outL = childOff[k], outR = outL + nd.lc;
    int runL = 0, runR = 0;
    for (int base = 0; base < count; base += blockDim.x)
    {
        const int j = base + tid;
        uint32_t t = 0;
        int fl = 0, fr = 0;
        if (j < count)
        {
            t = refs[nd.begin + j];
            fl = lo[(size_t)axis * n + t] < split ? 1 : 0;   // a vertex < median
            fr = hi[(size_t)axis * n + t] >= split ? 1 : 0;  // a vertex >= median: straddlers go to both sides
        }
        int el, er, tl, tr;
        Scan(tmp).ExclusiveSum(fl, el, tl); __syncthreads();
        Scan(tmp).ExclusiveSum(fr, er, tr); __syncthreads();
        if (fl) next[outL + runL + el] = t;
        if (fr) next[outR + runR + er] = t;
        runL += tl; runR += tr;
    }
    if (tid == 0)
    {
        const int L = childBase + 2 * innerRank[k], R = L + 1;
        nd.left = L; nd.right = R;
        KdBuildNode c;
        for (int a = 0; a < 3; a++) { c.mn[a] = nd.mn[a]; c.mx[a] = nd.mx[a]; }
        c.depth = nd.depth + 1; c.axis = 3; c.split = 0; c.left = c.right = -1; c.lc = c.rc = 0; c.level = 0; c.poolOff = 0;
        c.subNodes = c.subRefs = 0; c.pre = 0; c.leafOff = 0;
        KdBuildNode l = c, r = c;
        l.mx[axis] = split; l.begin = outL; l.count = nd.lc; // node->left->max[axis] = median
        r.mn[axis] = split; r.begin = outR; r.count = nd.rc; // node->right->min[axis] = median
        nodes[L] = l; nodes[R] = r;
    }
}

// bottom-up, one level per launch: nodes / leaf references of every subtree
__global__ void k_kd_sizes(KdBuildNode *__restrict__ nodes, int first, int count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    KdBuildNode &nd = nodes[first + k];
    if (nd.axis == 3) { nd.subNodes = 1; nd.subRefs = nd.count; }
    else { nd.subNodes = 1 + nodes[nd.left].subNodes + nodes[nd.right].subNodes; nd.subRefs = nodes[nd.left].subRefs + nodes[nd.right].subRefs; }
}

// top-down, one level per launch: pre-order index (left child = index + 1) and first leaf reference of the children
__global__ void k_kd_place(KdBuildNode *__restrict__ nodes, int first, int count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const KdBuildNode &nd = nodes[first + k];
    if (nd.axis == 3) return;
    KdBuildNode &l = nodes[nd.left], &r = nodes[nd.right];
    l.pre = nd.pre + 1; l.leafOff = nd.leafOff;
    r.pre = nd.pre + 1 + l.subNodes; r.leafOff = nd.leafOff + l.subRefs;
}

// the 8-byte node layout of include/rtb.h, nodes in pre-order, leaf lists in the order the leaves are met
__global__ void k_kd_emit(const KdBuildNode *__restrict__ nodes, int total, const __grid_constant__ KdLeafChunks chunks,
                          uint2 *__restrict__ out, uint32_t *__restrict__ leafTris)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const KdBuildNode &nd = nodes[k];
    if (nd.axis == 3)
    {
        out[nd.pre] = make_uint2((unsigned int)nd.leafOff, ((unsigned int)nd.count << 2) | 3u);
        const uint32_t *src = chunks.p[nd.level] + nd.poolOff;
        for (int j = 0; j < nd.count; j++) leafTris[nd.leafOff + j] = src[j];
    }
    else out[nd.pre] = make_uint2(__float_as_uint(nd.split), ((unsigned int)nodes[nd.right].pre << 2) | (unsigned int)nd.axis);
}

} // namespace rtb
