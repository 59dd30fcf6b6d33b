// Explicit instantiations of the fused scan kernel (part 4).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part4(int wb, bool dual)
{
    if (wb == 20 && dual == false) return &smg::launch_scan_regtab<20, false>;
    if (wb == 21 && dual == false) return &smg::launch_scan_regtab<21, false>;
    if (wb == 22 && dual == false) return &smg::launch_scan_regtab<22, false>;
    if (wb == 23 && dual == false) return &smg::launch_scan_regtab<23, false>;
    if (wb == 24 && dual == false) return &smg::launch_scan_regtab<24, false>;
    if (wb == 25 && dual == false) return &smg::launch_scan_regtab<25, false>;
    if (wb == 26 && dual == false) return &smg::launch_scan_regtab<26, false>;
    return nullptr;
}
