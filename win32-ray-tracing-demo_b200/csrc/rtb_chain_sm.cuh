// rtb_chain_sm.cuh -- Whitted reflection chains with a RESUMABLE accelerator traversal (sm_100a).
//
// Same arithmetic and same results as k_whitted_chain (rtb_kernels.cuh) -- the k-d walk is reference
// Tunnel.cpp:1163-1297, the grid walk Tunnel.cpp:819-970, shading MainWindow.cpp:69-143 -- but the
// control flow is one per-lane loop over the states
//      STEP   advance the accelerator (k-d: one inner node or leaf entry; grid: one cell)
//      LEAF   scan the list of the current leaf / cell with the conservative rejection test (rtb_pretest.h)
//      EXACT  the reference's exact test for the candidates LEAF could not reject (see nearestInList)
//      SHADE  close the ray (top-level geometries + tunnel result), shade, fold, spawn the reflection
// A lane that has finished a ray shades it and starts the next ray of ITS OWN chain at once instead of
// waiting, at a per-ray reconvergence point, for the slowest traversal in the warp.  On the tunnel
// frames the tiles at the vanishing point run 21-ray chains whose per-ray costs are heavy-tailed; with
// per-ray reconvergence the warp pays the maximum over its lanes for every bounce
// (profiles/r01_cost_map.md), which is the critical path of the 8-GPU frame.
#pragma once
#include "rtb_kernels.cuh"

namespace rtb {

#ifndef RTB_SM_STEP_BURST
#define RTB_SM_STEP_BURST 1000 // accelerator steps taken before the warp looks at the other states
#endif
#ifndef RTB_SM_LEAF_BURST
#define RTB_SM_LEAF_BURST 1000 // triangle tests per visit of the LEAF state
#endif

enum { SM_STEP = 0, SM_LEAF = 1, SM_SHADE = 2, SM_DONE = 3, SM_EXACT = 4 };

#ifndef RTB_SM_MIN_CTAS
#define RTB_SM_MIN_CTAS 6
#endif

// MINCTAS: resident CTAs per SM the register budget is set for.  The regular grid's throughput launches take 8 (64 registers),
// the tier launches of the other accelerators 6; the regular grid's SHARDS (<= 2 Mpixel) take 5: there the warp-per-pixel tier
// runs next to this kernel on the same SMs and its chains finish sooner with fewer warps competing for the schedulers
// (slowest 1/8 shard of the 4K frame 2.76 -> 2.40 ms; a whole 4K frame prefers 8: 10.5 vs 11.9 ms).
template <class Probe, bool GRID, int FOLD, int MINCTAS>
__global__ void __launch_bounds__(RTB_CTA_THREADS, MINCTAS)
k_whitted_chain_sm(const __grid_constant__ DScene S, const __grid_constant__ FrameParams F, float *__restrict__ out,
                   Counters *__restrict__ counters)
{
    int x, lr, y;
    unsigned int tile;
    const long long t_start = clock64();
    const bool active = pixelOfThread(F, x, lr, y, tile);
    unsigned int rays = 0;
    Probe pr;
    // Who runs this kernel?  Regular grid: every tile (measured 1.75x faster than the per-ray walk on the
    // whole frame).  k-d trees and the flat grid: the per-ray walk (k_whitted_chain) is faster on coherent
    // rays (~15 % fewer instructions; 2x on the flat grid's long runs of empty cells), so it renders the bulk
    // of the frame; this kernel is launched next to it on a second stream and takes only the latency-critical
    // tiles, i.e. the first *n_heavy entries of the heaviest-first order (F.skip_heavy).
    if (F.skip_heavy)
    {
        const unsigned int warpIndex = blockIdx.x * (unsigned int)F.warps_per_cta + (threadIdx.x >> 5);
        if ((F.split4 ? (warpIndex >> 2) : warpIndex) + (F.after_wide ? F.n_wide : 0u) >= heavyCount(F)) return;
    }
    __shared__ __align__(16) uint32_t stage[RTB_CTA_THREADS / 32][RTB_TILE_STAGE_WORDS];
    V3 cOut = v3(0, 0, 0);
    if (active)
    {
        const float dx = 1.0f / F.height, dy = 1.0f / F.height;
        const float sx = (x + 0.5f) * dx, sy = 1 - (y + 0.5f) * dy;
        Ray r = generateRay(F.cam, sx, sy);
        float4 fold[FOLD];
        int4 stack[GRID ? 1 : RTB_KD_STACK];
        int nfold = 0, depth = 0;
        V3 c = v3(0, 0, 0);
        const V3 zero = v3(0, 0, 0);

        // ---- traversal state ----
        int st;
        // k-d (see kdIntersect)
        float enT = 0, exT = 0;
        V3 enP = zero, exP = zero;
        int enPt = 0, exPt = 1, exNode = -1, exPrev = 0, cur = 0;
        // grid (see gridIntersect)
        int ci = 0, cj = 0, ck = 0;
        float cd = 0;
        V3 cp = zero;
        bool px = false, py = false, pz = false;
        // leaf / cell list
        unsigned int li = 0, lend = 0;
        float lo = 0, hi = 0, minD = FLT_MAX;
        int hitTri = -1;
        // conservative rejection test: per-ray / per-list constants, pending candidate, end of the exact scan
        float dmx = 0, Lp = rtb_pre::lowBound(-FLT_MAX), Hp = rtb_pre::highBound(FLT_MAX, FLT_MAX);
        unsigned int pend = 0xffffffffu, jend = 0, lp = 0, counted = 0; // lp: current pair of the list [li, lend)
        bool tunnelHit = false;

        // start the accelerator walk of ray r (reference Tunnel.cpp:1171-1197 / 833-860)
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

        // grid: leave the current cell (Tunnel.cpp:885-966)
        auto gridAdvance = [&]() {
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
            if (ci < 0 || ci > S.nx - 1 || cj < 0 || cj > S.ny - 1 || ck < 0 || ck > S.nz - 1) st = SM_SHADE; // left the grid: miss
            else st = SM_STEP;
        };

        // k-d: the current leaf gave no hit -> pop (Tunnel.cpp:1285-1292)
        auto kdPop = [&]() {
            enPt = exPt; enT = exT; enP = exP;
            cur = exNode;
            if (cur == -1) { st = SM_SHADE; return; } // ray leaves the tree: miss
            exPt = exPrev;
            const int4 e = stack[exPt];
            exNode = e.x; exT = __int_as_float(e.y); exPrev = e.w >> 2;
            exP = kdPoint(r, exT, __int_as_float(e.z), e.w & 3);
            st = SM_STEP;
        };

        // the list of the current leaf / cell is exhausted
        auto listDone = [&]() {
            if (hitTri >= 0) { tunnelHit = true; st = SM_SHADE; return; }
            if (GRID) gridAdvance();
            else kdPop();
        };

        beginRay();
        while (st != SM_DONE)
        {
            if (st == SM_STEP)
            {
#pragma unroll 1
                for (int burst = 0; burst < RTB_SM_STEP_BURST && st == SM_STEP; burst++)
                {
                    if (GRID)
                    {
                        const int cell = (ci * S.ny + cj) * S.nz + ck;
                        pr.step(cell);
                        const uint2 w = __ldg(S.g_words + ((unsigned int)cell >> 5));
                        const unsigned int bit = 1u << (cell & 31);
                        if (w.x & bit)
                        {
                            const unsigned int rk = w.y + __popc(w.x & (bit - 1));
                            li = __ldg(S.g_start + rk);
                            lend = __ldg(S.g_start + rk + 1);
                            minD = FLT_MAX; hitTri = -1; pend = 0xffffffffu; lp = li >> 1;
                            if (li == lend) gridAdvance(); // cannot happen with a well-formed directory; the pair loop needs a non-empty list
                            else st = SM_LEAF;
                        }
                        else gridAdvance();
                    }
                    else
                    {
                        const uint2 nd = __ldg(S.kd_nodes + cur);
                        pr.step(cur);
                        if ((nd.y & 3u) == 3u)
                        { // leaf: scan its list with the +-0.001 window (Tunnel.cpp:1269-1280)
                            li = nd.x;
                            lend = nd.x + (nd.y >> 2);
                            lo = enT - 0.001f; hi = exT + 0.001f;
                            Lp = rtb_pre::lowBound(lo); Hp = rtb_pre::highBound(hi, FLT_MAX);
                            minD = FLT_MAX; hitTri = -1; pend = 0xffffffffu; lp = li >> 1;
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
                                else { farChild = right; cur = left; both = true; } // (ex == split is covered by <=, line 1223)
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
            if (st == SM_LEAF)
            { // two list entries per iteration from the pair stream (see nearestInList)
#pragma unroll 1
                for (int burst = 0; burst < RTB_SM_LEAF_BURST && st == SM_LEAF; burst++)
                {
                    const rtb_pre::PreTri2 P = loadPreTri2(S.pre2, lp);
                    const unsigned int j0 = 2u * lp, j1 = j0 + 1u;
                    const bool in0 = j0 >= li, in1 = j1 < lend;
                    if (in0) pr.tri();
                    if (in1) pr.tri();
                    bool r0, r1;
                    rtb_pre::sureReject2<!GRID>(P, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, dmx, Lp, Hp, r0, r1);
                    const bool c0 = in0 && !r0, c1 = in1 && !r1;
                    if (c0 || c1)
                    { // candidate: its exact test waits for the end of the list, unless one is waiting already
                        if (pend != 0xffffffffu || (c0 && c1))
                        {
                            if (pend == 0xffffffffu) pend = j0;
                            counted = in1 ? j1 + 1u : j1;
                            jend = lend;
                            st = SM_EXACT;
                        }
                        else pend = c0 ? j0 : j1;
                    }
                    if (st == SM_LEAF && ++lp == ((lend + 1u) >> 1))
                    {
                        if (pend != 0xffffffffu) { counted = lend; jend = pend + 1; st = SM_EXACT; }
                        else listDone();
                    }
                }
            }
            if (st == SM_EXACT)
            { // [pend, jend) in list order; entries the fast scan has not reached are counted here
                for (unsigned int j = pend; j < jend; j++)
                {
                    const uint32_t idx = __ldg((GRID ? S.g_tris : S.kd_tris) + j);
                    const TriData T = loadTri(S.tri, idx);
                    if (j >= counted) pr.tri();
                    float t;
                    if (triIntersectT<true>(T, r, t) && (GRID || (t >= lo && t <= hi)) && t < minD)
                    {
                        minD = t;
                        hitTri = (int)idx;
                    }
                }
                listDone();
            }
            if (st == SM_SHADE)
            {
                // GeometrySet::intersect with the tunnel's result plugged in at the tunnel's position
                Hit h;
                const bool any = sceneIntersectWith(S, r, h, pr, [&](int &tri, float &t, V3 &n) {
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
                else st = SM_DONE;
            }
        }
        while (nfold > 0)
        {
            const float4 f = fold[--nfold];
            c = v3(f.x, f.y, f.z) + c * f.w + zero * 0.0f;
        }
        cOut = c;
    }
    // split4 launches give a tile to four warps of 8 lanes: per-pixel stores there
    // (group stores only in the launch that renders the bulk of a frame: there no warp of a CTA has left above)
    const bool grouped = !F.split4 && !F.skip_heavy && storeGroup(F, out, tile, active, cOut, stage);
    if (!grouped && (F.split4 || !storeTile(F, out, tile, active, cOut, stage[threadIdx.x >> 5])) && active) storePixel(F, out, x, lr, y, cOut, t_start, rays, pr);
    finishWarp(F, counters, tile, t_start, rays, ProbeCounts<Probe>::tris(pr), ProbeCounts<Probe>::steps(pr));
}

} // namespace rtb
