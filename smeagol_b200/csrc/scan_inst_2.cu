// Explicit instantiations of the fused scan kernel (part 2).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part2(int wb, bool dual)
{
    if (wb == 12 && dual == true) return &smg::launch_scan_regtab<12, true>;
    if (wb == 13 && dual == true) return &smg::launch_scan_regtab<13, true>;
    if (wb == 14 && dual == true) return &smg::launch_scan_regtab<14, true>;
    return nullptr;
}
