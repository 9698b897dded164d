// rtb_k_oct.cu -- k_whitted_chain_oct: eight lanes per pixel (rtb_chain_oct.cuh).
#include "rtb_launch.h"
#include "rtb_chain_oct.cuh"

namespace rtb {

template <class Probe, bool GRID, int FOLD> static void go(const Launch &L)
{
    // packed rejection test in the list rounds: k-d median trees (many short leaves per ray); see rtb_chain_oct.cuh
    if (!GRID && L.S->accel == RTB_ACCEL_KD_MEDIAN) k_whitted_chain_oct<Probe, GRID, FOLD, true><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
    else k_whitted_chain_oct<Probe, GRID, FOLD, false><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
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
