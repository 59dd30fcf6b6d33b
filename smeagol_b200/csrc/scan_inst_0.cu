// Explicit instantiations of the fused scan kernel (part 0).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part0(int wb, bool dual)
{
    if (wb == 2 && dual == true) return &smg::launch_scan_regtab<2, true>;
    if (wb == 3 && dual == true) return &smg::launch_scan_regtab<3, true>;
    if (wb == 4 && dual == true) return &smg::launch_scan_regtab<4, true>;
    if (wb == 5 && dual == true) return &smg::launch_scan_regtab<5, true>;
    if (wb == 6 && dual == true) return &smg::launch_scan_regtab<6, true>;
    if (wb == 7 && dual == true) return &smg::launch_scan_regtab<7, true>;
    if (wb == 8 && dual == true) return &smg::launch_scan_regtab<8, true>;
    return nullptr;
}
