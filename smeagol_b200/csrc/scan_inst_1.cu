// Explicit instantiations of the fused scan kernel (part 1).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part1(int wb, bool dual)
{
    if (wb == 9 && dual == true) return &smg::launch_scan_regtab<9, true>;
    if (wb == 10 && dual == true) return &smg::launch_scan_regtab<10, true>;
    if (wb == 11 && dual == true) return &smg::launch_scan_regtab<11, true>;
    return nullptr;
}
