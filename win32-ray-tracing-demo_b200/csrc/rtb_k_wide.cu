// rtb_k_wide.cu -- k_whitted_chain_wide: one warp per pixel (rtb_chain_wide.cuh).
#include "rtb_launch.h"
#include "rtb_chain_wide.cuh"

namespace rtb {

template <class Probe, int FOLD> static void go(const Launch &L)
{
    k_whitted_chain_wide<Probe, FOLD><<<L.grid, RTB_CTA_THREADS, 0, L.stream>>>(*L.S, *L.F, L.out, L.counters);
}

void launchChainWide(const Launch &L)
{
    const bool s = shortFold(*L.F);
    if (L.count) { if (s) go<CountProbe, RTB_FOLD_SHORT>(L); else go<CountProbe, RTB_FOLD_LONG>(L); }
    else { if (s) go<NoProbe, RTB_FOLD_SHORT>(L); else go<NoProbe, RTB_FOLD_LONG>(L); }
}

} // namespace rtb
