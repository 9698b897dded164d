// rtb_k_oct.cu -- k_whitted_chain_oct: eight lanes per pixel (rtb_chain_oct.cuh).
#include "rtb_launch.h"
#include "rtb_chain_oct.cuh"

namespace rtb {

// Variant per tree kind (rtb_chain_oct.cuh): SAH trees -- 16 lanes per pixel, exact list rounds; k-d median trees -- 8 lanes
// per pixel, list rounds through the packed rejection test; grids (only with RTB_OCT_TIER forced: they keep the resumable
// walk by default) -- 8 lanes, exact rounds.
static bool medianTree(const Launch &L) { return L.S->accel == RTB_ACCEL_KD_MEDIAN; }
// 16 lanes per pixel pay on SAH shards up to 2 Mpixel (1/8 and 1/16 shards of a 4K frame, 1280x960 frames: 1.17 -> 1.09 ms,
// 1.20 -> 1.06 ms); a 1920x1440 frame is better off with 8 (1.68 vs 1.86 ms: twice the warps for its larger tier)
int octWarpsPerTile(int accel, int n_tiles) { return (accel == RTB_ACCEL_KD_SAH && n_tiles <= 65536) ? 16 : 8; }

template <class Probe, bool GRID, int FOLD> static void go(const Launch &L)
{
    if (GRID) k_whitted_chain_oct<Probe, GRID, FOLD, false, 8><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
    else if (medianTree(L)) k_whitted_chain_oct<Probe, false, FOLD, true, 8><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
    else if (octWarpsPerTile(L.S->accel, L.F->n_tiles) == 16) k_whitted_chain_oct<Probe, false, FOLD, false, 16><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
    else k_whitted_chain_oct<Probe, false, FOLD, false, 8><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
}
template <class Probe, bool GRID> static void byFold(const Launch &L)
{
    if (shortFold(*L.F)) go<Probe, GRID, RTB_FOLD_SHORT>(L);
    else go<Probe, GRID, RTB_FOLD_LONG>(L);
}

void launchChainOct(const Launch &L, bool grid)
{
    if (L.count) { if (grid) byFold<CountProbe, true>(L); else byFold<CountProbe, false>(L); }
    else { if (grid) byFold<NoProbe, true>(L); else byFold<NoProbe, false>(L); }
}

} // namespace rtb
