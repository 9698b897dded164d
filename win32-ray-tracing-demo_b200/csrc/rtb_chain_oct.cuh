// rtb_chain_oct.cuh -- Whitted reflection chains, EIGHT LANES PER PIXEL (sm_100a).
//
// Same arithmetic and same results as k_whitted_chain (rtb_kernels.cuh): the k-d walk is reference
// Tunnel.cpp:1163-1297, the grid walk Tunnel.cpp:819-970, shading MainWindow.cpp:69-143.
//
// A warp is four groups of eight lanes; a group owns one pixel at a time.  The eight lanes of a group hold the
// same ray and the same traversal state (they repeat the scalar work: accelerator steps, shading) and test
// EIGHT triangles of the current leaf / cell list at once, one per lane; a 3-step butterfly inside the group
// finds the nearest accepted hit, ties going to the earliest list position as in the reference's scan.
// The four groups run the resumable state machine of rtb_chain_sm.cuh independently (NEXT pixel / accelerator
// STEP / list round LEAF / SHADE), so a group that finishes a ray starts its next one at once.
//
// Why: nothing in a reflection chain is parallel except the triangle tests of one list.  With one lane per
// pixel a chain is (steps + tests) serial latencies; here it is (steps + tests / 8), for ~2x the warp
// instructions of the per-ray kernel on a k-d tree (the eight lanes repeat the node walk) -- against 8x the
// latency gain and ~4-8x the instructions of the warp-per-pixel kernel (rtb_chain_wide.cuh).  That makes it the
// right tool for the MANY moderately heavy tiles of a multi-GPU shard: on the tunnel frames ~1 % of the tiles
// are within 2x of the heaviest one, far too many for a warp per pixel, and a shard is bound by their latency.
// Measured (one B200, kernel ms, gpurun_out/sweep5.log / sweep6.log): worst 1/8 shard of the 4K frame
// SAH 1.99 -> 1.42, k-d median 4.48 -> 3.41, preset 4 SAH 1.80 -> 1.44; 1280x960 SAH 1.78 -> 1.44.
// Rejected on measurement: as the tier of a whole 4K frame (throughput-bound: 6.01 -> 6.85 ms) and, with eight
// pixels per group and one warp per tile, as the throughput kernel of the regular grid (14.4 -> 60 ms).
#pragma once
#include "rtb_kernels.cuh"
#include "rtb_chain_sm.cuh"

namespace rtb {

enum { OCT_NEXT = 4 };

#ifndef RTB_OCT_MIN_CTAS
#define RTB_OCT_MIN_CTAS 5
#endif
// LANES per pixel group (template parameter): 8 = four pixels per warp, eight warps per tile; 16 = two pixels per warp,
// sixteen warps per tile.  The tier is bound by the latency of its chains: fewer groups per warp wait less for each other's
// phases.  Slowest 1/8 shard of the 4K frame, kernel ms with 4 / 8 / 16 lanes: SAH 1.68 / 1.17 / 1.09, preset 4 SAH
// - / 1.22 / 1.16, k-d median 3.57 / 2.16 / 2.51 (its many short leaves keep 8 lanes busy; 16 double the instructions for
// nothing) -- profiles/r02_sweep_tiers.log.  So: 16 for SAH trees, 8 (with PRE) for median trees.
#define RTB_OCT_GROUPS (32 / LANES)          // pixel groups per warp
#define RTB_OCT_WARPS_PER_TILE LANES         // 32 pixels per tile / groups per warp
#define RTB_OCT_GROUP_BITS ((LANES == 32) ? 0xffffffffu : ((1u << LANES) - 1u))

// Eight warps per tile: group g of warp (w & 7) renders pixel (w & 7) * 4 + g of the tile.
// PRE: the list rounds go through the packed rejection test, eight PAIRS per round.  Measured on the 1/8 shards of the 4K
// frame (one B200, slowest shard, kernel ms): k-d median 2.66 -> 2.14 (10.9 leaves per ray: the rounds are a large share
// of the chain), k-d SAH 1.21 -> 1.37 (4 leaves per ray; a round became two dependent phases, test + exact test, and the
// tier is bound by the chain's latency, not by issue slots).  So: median trees only (rtb_abi.cu: launchRender).
template <class Probe, bool GRID, int FOLD, bool PRE, int LANES>
__global__ void __launch_bounds__(RTB_CTA_THREADS, RTB_OCT_MIN_CTAS)
k_whitted_chain_oct(const __grid_constant__ DScene S, const __grid_constant__ FrameParams F, float *__restrict__ out,
                    Counters *__restrict__ counters)
{
    const long long t_start = clock64();
    const unsigned int w = blockIdx.x * (RTB_CTA_THREADS / 32) + (threadIdx.x >> 5);
    const unsigned int lane = threadIdx.x & 31u, g = lane / LANES, sub = lane % LANES;
    const unsigned int item = (w / RTB_OCT_WARPS_PER_TILE) + F.item_base;
    if (item >= F.item_end || item >= (unsigned int)F.n_tiles || item >= heavyCount(F)) return; // warp-uniform
    const unsigned int tile = F.order ? __ldg(F.order + item) : item;
    const int ty = tile / F.tiles_x, tx = tile - ty * F.tiles_x;

    unsigned int rays = 0;
    Probe prTop, pr; // prTop: work all eight lanes repeat (top-level geometries); pr: the tunnel walk

    const V3 zero = v3(0, 0, 0);
    float4 fold[FOLD];
    int4 stack[GRID ? 1 : RTB_KD_STACK];
    int nfold = 0, depth = 0;
    V3 c = zero;
    Ray r;
    r.o = zero; r.d = v3(0, 0, 1);
    int x = 0, lr = 0, y = 0, k = -1;

    int st = OCT_NEXT;
    // k-d (see kdIntersect)
    float enT = 0, exT = 0;
    V3 enP = zero, exP = zero;
    int enPt = 0, exPt = 1, exNode = -1, exPrev = 0, cur = 0;
    // grid (see gridIntersect)
    int ci = 0, cj = 0, ck = 0;
    float cd = 0;
    V3 cp = zero;
    bool px = false, py = false, pz = false;
    // list [li, lend) of the reference array; lp = next pair of the pair stream; constants of the rejection test
    unsigned int li = 0, lend = 0, lp = 0;
    float lo = 0, hi = 0, minD = FLT_MAX;
    float dmx = 0, Lp = rtb_pre::lowBound(-FLT_MAX), Hp = rtb_pre::highBound(FLT_MAX, FLT_MAX);
    int hitTri = -1;
    bool tunnelHit = false;

    auto beginRay = [&]() {
        rays++;
        tunnelHit = false;
        hitTri = -1;
        minD = FLT_MAX;
        dmx = rtb_pre::dirMax(r.d.x, r.d.y, r.d.z);
        if (GRID)
        {
            if (r.o.x < S.g_origin.x || r.o.x > S.g_far.x || r.o.y < S.g_origin.y || r.o.y > S.g_far.y ||
                r.o.z < S.g_origin.z || r.o.z > S.g_far.z)
            {
                float entry, exit;
                if (!boxIntersect(S.g_origin, S.g_extent, r, entry, exit)) { st = SM_SHADE; return; }
                cd = entry;
                cp = at(r, entry);
                indexInGrid(S, cp, ci, cj, ck);
            }
            else
            {
                cp = r.o;
                cd = 0;
                indexInGrid(S, r.o, ci, cj, ck);
            }
            px = r.d.x > 0; py = r.d.y > 0; pz = r.d.z > 0;
            st = SM_STEP;
        }
        else
        {
            float a, b;
            if (!boxIntersect(S.kd_min, S.kd_size, r, a, b)) { st = SM_SHADE; return; }
            enT = a;
            enP = (a >= 0) ? (r.o + r.d * a) : r.o;
            enPt = 0;
            exPt = 1; exT = b; exP = r.o + r.d * b; exNode = -1; exPrev = 0;
            stack[1] = make_int4(-1, __float_as_int(b), 0, 3);
            cur = 0;
            st = SM_STEP;
        }
    };
    auto gridAdvance = [&]() { // Tunnel.cpp:885-966
        float ddx, ddy, ddz;
        if (px) ddx = ((S.g_origin.x + (ci + 1) * S.g_cell.x) - cp.x) / r.d.x;
        else ddx = (cp.x - (S.g_origin.x + ci * S.g_cell.x)) / -r.d.x;
        if (py) ddy = ((S.g_origin.y + (cj + 1) * S.g_cell.y) - cp.y) / r.d.y;
        else ddy = (cp.y - (S.g_origin.y + cj * S.g_cell.y)) / -r.d.y;
        if (pz) ddz = ((S.g_origin.z + (ck + 1) * S.g_cell.z) - cp.z) / r.d.z;
        else ddz = (cp.z - (S.g_origin.z + ck * S.g_cell.z)) / -r.d.z;
        if (ddx < ddy && ddx < ddz) { ci += px ? 1 : -1; cd += ddx; }
        else if (ddy < ddz) { cj += py ? 1 : -1; cd += ddy; }
        else { ck += pz ? 1 : -1; cd += ddz; }
        cp = at(r, cd);
        if (ci < 0 || ci > S.nx - 1 || cj < 0 || cj > S.ny - 1 || ck < 0 || ck > S.nz - 1) st = SM_SHADE;
        else st = SM_STEP;
    };
    auto kdPop = [&]() { // Tunnel.cpp:1285-1292
        enPt = exPt; enT = exT; enP = exP;
        cur = exNode;
        if (cur == -1) { st = SM_SHADE; return; }
        exPt = exPrev;
        const int4 e = stack[exPt];
        exNode = e.x; exT = __int_as_float(e.y); exPrev = e.w >> 2;
        exP = kdPoint(r, exT, __int_as_float(e.z), e.w & 3);
        st = SM_STEP;
    };
    auto listDone = [&]() {
        if (hitTri >= 0) { tunnelHit = true; st = SM_SHADE; return; }
        if (GRID) gridAdvance();
        else kdPop();
    };

    while (true)
    {
        // all 32 lanes are here together: the loop condition below is warp-uniform
        const unsigned int leafMask = __ballot_sync(0xffffffffu, st == SM_LEAF);
        if (__all_sync(0xffffffffu, st == SM_DONE)) break;

        if (st == OCT_NEXT)
        { // the group's next pixel
            k++;
            if (k >= 1) st = SM_DONE;
            else
            {
                const unsigned int p = (w % RTB_OCT_WARPS_PER_TILE) * RTB_OCT_GROUPS + g;
                lr = ty * RTB_TILE_H + (int)(p >> 3);
                if (localToGlobal(F, tx * RTB_TILE_W + (int)(p & 7u), lr, x, y))
                {
                    const float dx = 1.0f / F.height, dy = 1.0f / F.height;
                    const float sx = (x + 0.5f) * dx, sy = 1 - (y + 0.5f) * dy;
                    r = generateRay(F.cam, sx, sy);
                    nfold = 0; depth = 0; c = zero;
                    beginRay();
                }
            }
        }
        if (st == SM_STEP)
        {
#pragma unroll 1
            while (st == SM_STEP)
            {
                if (GRID)
                {
                    const int cell = (ci * S.ny + cj) * S.nz + ck;
                    pr.step(cell);
                    const uint2 wd = __ldg(S.g_words + ((unsigned int)cell >> 5));
                    const unsigned int bit = 1u << (cell & 31);
                    if (wd.x & bit)
                    {
                        const unsigned int rk = wd.y + __popc(wd.x & (bit - 1));
                        li = __ldg(S.g_start + rk);
                        lend = __ldg(S.g_start + rk + 1);
                        minD = FLT_MAX; hitTri = -1; lp = li >> 1;
                        Hp = rtb_pre::highBound(FLT_MAX, FLT_MAX);
                        if (li == lend) gridAdvance(); // cannot happen with a well-formed directory
                        else st = SM_LEAF;
                    }
                    else gridAdvance();
                }
                else
                {
                    const uint2 nd = __ldg(S.kd_nodes + cur);
                    pr.step(cur);
                    if ((nd.y & 3u) == 3u)
                    {
                        li = nd.x;
                        lend = nd.x + (nd.y >> 2);
                        lo = enT - 0.001f; hi = exT + 0.001f;
                        Lp = rtb_pre::lowBound(lo); Hp = rtb_pre::highBound(hi, FLT_MAX);
                        minD = FLT_MAX; hitTri = -1; lp = li >> 1;
                        if (li == lend) kdPop();
                        else st = SM_LEAF;
                    }
                    else
                    {
                        const float splitVal = __uint_as_float(nd.x);
                        const int axis = (int)(nd.y & 3u);
                        const int right = (int)(nd.y >> 2), left = cur + 1;
                        const float en = compSel(enP, axis), ex = compSel(exP, axis);
                        int farChild = -1;
                        bool both = false;
                        if (en <= splitVal)
                        {
                            if (ex <= splitVal) cur = left;
                            else { farChild = right; cur = left; both = true; }
                        }
                        else
                        {
                            if (splitVal < ex) cur = right;
                            else { farChild = left; cur = right; both = true; }
                        }
                        if (both)
                        {
                            const float t = (splitVal - compSel(r.o, axis)) / compSel(r.d, axis);
                            const int tmp = exPt++;
                            if (exPt == enPt) exPt += 1;
                            exPrev = tmp; exT = t; exNode = farChild;
                            exP = kdPoint(r, t, splitVal, axis);
                            stack[exPt] = make_int4(farChild, __float_as_int(t), __float_as_int(splitVal), axis | (tmp << 2));
                        }
                    }
                }
            }
        }
        if (PRE && ((leafMask >> lane) & 1u))
        { // one round: eight PAIRS of the list, one per lane (the group was in LEAF at the top of this iteration), through the
          // packed rejection test; the entries it cannot reject take the exact test, the group's lanes side by side (see
          // nearestInList<.., WIDE>: same scheme with 8 lanes instead of 32)
            const unsigned int gmask = RTB_OCT_GROUP_BITS << (g * LANES); // the group's lanes share one state: they are all here
            const unsigned int pEnd = (lend + 1u) >> 1;
            const unsigned int p = lp + sub, j0 = 2u * p, j1 = j0 + 1u;
            bool c0 = false, c1 = false;
            if (p < pEnd)
            {
                const rtb_pre::PreTri2 P = loadPreTri2(S.pre2, p);
                const bool in0 = j0 >= li, in1 = j1 < lend;
                if (in0) pr.tri();
                if (in1) pr.tri();
                bool r0, r1;
                rtb_pre::sureReject2<true>(P, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, dmx, Lp, Hp, r0, r1);
                c0 = in0 && !r0; c1 = in1 && !r1;
            }
            if (__ballot_sync(gmask, c0 || c1))
            {
                const uint32_t *refs = GRID ? S.g_tris : S.kd_tris;
                unsigned int key = 0xffffffffu, idx = 0;
                if (c0)
                {
                    const uint32_t i0 = __ldg(refs + j0);
                    const TriData T = loadTri(S.tri, i0);
                    float t;
                    bool ok = triIntersectT<true>(T, r, t);
                    if (!GRID) ok = ok && (t >= lo && t <= hi);
                    if (ok && t < FLT_MAX) { key = __float_as_uint(t); idx = i0; } // positive floats order like their bit patterns
                }
                if (c1)
                {
                    const uint32_t i1 = __ldg(refs + j1);
                    const TriData T = loadTri(S.tri, i1);
                    float t;
                    bool ok = triIntersectT<true>(T, r, t);
                    if (!GRID) ok = ok && (t >= lo && t <= hi);
                    if (ok && t < FLT_MAX && __float_as_uint(t) < key) { key = __float_as_uint(t); idx = i1; } // strict <: j0 keeps a tie
                }
                unsigned int best = key;
                best = min(best, __shfl_xor_sync(gmask, best, 1));
                best = min(best, __shfl_xor_sync(gmask, best, 2));
                if (LANES >= 8) best = min(best, __shfl_xor_sync(gmask, best, 4));
                if (LANES >= 16) best = min(best, __shfl_xor_sync(gmask, best, 8));
                const unsigned int eq = (__ballot_sync(gmask, key != 0xffffffffu && key == best) >> (g * LANES)) & RTB_OCT_GROUP_BITS;
                const unsigned int winner = eq ? (unsigned int)(__ffs(eq) - 1) : 0u; // earliest list position among equal distances
                const unsigned int idxW = __shfl_sync(gmask, idx, g * LANES + winner);
                if (eq && __uint_as_float(best) < minD)
                { // strict <: an equal distance in a later round does not replace the earlier one
                    minD = __uint_as_float(best);
                    hitTri = (int)idxW;
                    Hp = rtb_pre::highBound(GRID ? FLT_MAX : hi, minD);
                }
            }
            lp += LANES;
            if (lp >= pEnd) listDone();
        }
        if (!PRE && ((leafMask >> lane) & 1u))
        { // one round: eight triangles of the list, one per lane (the group was in LEAF at the top of this iteration)
            const unsigned int i = li + sub;
            float t = FLT_MAX;
            unsigned int idx = 0;
            bool ok = false;
            if (i < lend)
            {
                idx = __ldg((GRID ? S.g_tris : S.kd_tris) + i);
                const TriData T = loadTri(S.tri, idx);
                pr.tri();
                ok = triIntersectT<true>(T, r, t);
                if (!GRID) ok = ok && (t >= lo && t <= hi);
                ok = ok && t < FLT_MAX; // accepted distances are >= 0.0005; inf / NaN never become a hit in the reference either
            }
            const unsigned int key = ok ? __float_as_uint(t) : 0xffffffffu; // positive floats order like their bit patterns
            unsigned int best = key;
            best = min(best, __shfl_xor_sync(leafMask, best, 1));
            best = min(best, __shfl_xor_sync(leafMask, best, 2));
            if (LANES >= 8) best = min(best, __shfl_xor_sync(leafMask, best, 4));
            if (LANES >= 16) best = min(best, __shfl_xor_sync(leafMask, best, 8));
            const unsigned int eq = (__ballot_sync(leafMask, ok && key == best) >> (g * LANES)) & RTB_OCT_GROUP_BITS;
            const unsigned int winner = eq ? (unsigned int)(__ffs(eq) - 1) : 0u; // earliest list position among equal distances
            const unsigned int idxW = __shfl_sync(leafMask, idx, g * LANES + winner);
            if (eq && __uint_as_float(best) < minD)
            { // strict <: an equal distance in a later round does not replace the earlier one
                minD = __uint_as_float(best);
                hitTri = (int)idxW;
            }
            li += LANES;
            if (li >= lend) listDone();
        }
        if (st == SM_SHADE)
        {
            Hit h;
            const bool any = sceneIntersectWith(S, r, h, prTop, [&](int &tri, float &t, V3 &n) {
                if (!tunnelHit) return false;
                tri = hitTri;
                t = minD;
                const float4 q2 = __ldg(S.tri + 3ull * (unsigned int)hitTri + 2);
                n = v3(q2.y, q2.z, q2.w);
                return true;
            });
            bool spawned = false;
            if (any)
            {
                const rtb_material &m = S.mats[h.mat];
                const V3 nl = (dot(h.n, r.d) < 0) ? h.n : h.n * -1;
                const V3 local = matLocal(m, r, h.pos, h.n);
                if (++depth <= F.setting.max_depth && depth <= RTB_MAX_DEPTH)
                {
                    const V3 diffusive = (m.diffusiveness > 0) ? local : zero;
                    const V3 term = diffusive * m.diffusiveness;
                    if (m.reflectiveness > 0)
                    {
                        fold[nfold++] = make_float4(term.x, term.y, term.z, m.reflectiveness);
                        const V3 v = r.d - nl * 2 * dot(nl, r.d);
                        r.o = h.pos;
                        r.d = v;
                        spawned = true;
                    }
                    else c = term + zero * m.reflectiveness + zero * m.refractiveness;
                }
            }
            if (spawned) beginRay();
            else
            { // fold innermost-first, store, next pixel
                while (nfold > 0)
                {
                    const float4 f = fold[--nfold];
                    c = v3(f.x, f.y, f.z) + c * f.w + zero * 0.0f;
                }
                if (sub == 0) storePixel(F, out, x, lr, y, c, t_start, rays, pr);
                st = OCT_NEXT;
            }
        }
    }
    // counters: rays, accelerator steps and top-level tests are identical on a group's eight lanes (sub 0 counts);
    // the list tests are distinct per lane
    unsigned int tris = ProbeCounts<Probe>::tris(pr) + (sub == 0 ? ProbeCounts<Probe>::tris(prTop) : 0u);
    unsigned int steps = sub == 0 ? ProbeCounts<Probe>::steps(pr) : 0u;
    unsigned int nr = sub == 0 ? rays : 0u;
    tris = __reduce_add_sync(0xffffffffu, tris);
    steps = __reduce_add_sync(0xffffffffu, steps);
    nr = __reduce_add_sync(0xffffffffu, nr);
    if (lane == 0)
    {
        // no cost is recorded: the tile keeps the cost the throughput kernel measured (FrameParams::record_cost)
        if (nr) atomicAdd(&counters->rays, (unsigned long long)nr);
        if (tris) atomicAdd(&counters->tris, (unsigned long long)tris);
        if (steps) atomicAdd(&counters->steps, (unsigned long long)steps);
        if ((w % RTB_OCT_WARPS_PER_TILE) == 0u) atomicAdd(&counters->tiles, 1ull); // one of the tile's warps reports it
    }
}

#undef RTB_OCT_GROUPS
#undef RTB_OCT_WARPS_PER_TILE
#undef RTB_OCT_GROUP_BITS

} // namespace rtb
