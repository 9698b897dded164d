// rtb_k_sm.cu -- k_whitted_chain_sm: the chain with a resumable accelerator walk (rtb_chain_sm.cuh).
#include "rtb_launch.h"
#include "rtb_chain_sm.cuh"

namespace rtb {

#define RTB_SMALL_FRAME_TILES_SM 16384 // frames up to 0.5 Mpixel keep 8 (measured policy of the small-frame tier)

#define RTB_SM_SHARD_TILES 65536 // regular-grid launches up to 2 Mpixel are shards / small frames: 5 CTAs per SM (rtb_chain_sm.cuh)

template <class Probe, bool GRID, int FOLD> static void go(const Launch &L)
{
    // ... when a warp-per-pixel tier runs next to it (n_wide > 0: grids with long cell lists; preset 4's grid, 4 triangles per cell and
    // no such tier, loses throughput with 5: 9.8 -> 12.2 ms)
    if (GRID && L.S->accel == RTB_ACCEL_REGULAR_GRID && L.F->n_wide > 0 && L.F->n_tiles <= RTB_SM_SHARD_TILES && L.F->n_tiles > RTB_SMALL_FRAME_TILES_SM)
        k_whitted_chain_sm<Probe, GRID, FOLD, 5><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
    else
        k_whitted_chain_sm<Probe, GRID, FOLD, GRID ? RTB_CHAIN_MIN_CTAS : RTB_SM_MIN_CTAS><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
}
template <class Probe, bool GRID> static void byFold(const Launch &L)
{
    if (shortFold(*L.F)) go<Probe, GRID, RTB_FOLD_SHORT>(L);
    else go<Probe, GRID, RTB_FOLD_LONG>(L);
}

void launchChainSm(const Launch &L, bool grid)
{
    if (L.count) { if (grid) byFold<CountProbe, true>(L); else byFold<CountProbe, false>(L); }
    else { if (grid) byFold<NoProbe, true>(L); else byFold<NoProbe, false>(L); }
}

} // namespace rtb
