// Explicit instantiations of the fused scan kernel (part 3).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part3(int wb, bool dual)
{
    if (wb == 15 && dual == true) return &smg::launch_scan_regtab<15, true>;
    if (wb == 16 && dual == true) return &smg::launch_scan_regtab<16, true>;
    if (wb == 17 && dual == false) return &smg::launch_scan_regtab<17, false>;
    if (wb == 18 && dual == false) return &smg::launch_scan_regtab<18, false>;
    if (wb == 19 && dual == false) return &smg::launch_scan_regtab<19, false>;
    return nullptr;
}
