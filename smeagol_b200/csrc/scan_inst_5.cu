// Explicit instantiations of the fused scan kernel (part 5).  Generated layout, edited by hand.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part5(int wb, bool dual)
{
    if (wb == 27 && dual == false) return &smg::launch_scan_regtab<27, false>;
    if (wb == 28 && dual == false) return &smg::launch_scan_regtab<28, false>;
    if (wb == 29 && dual == false) return &smg::launch_scan_regtab<29, false>;
    if (wb == 30 && dual == false) return &smg::launch_scan_regtab<30, false>;
    if (wb == 31 && dual == false) return &smg::launch_scan_regtab<31, false>;
    if (wb == 32 && dual == false) return &smg::launch_scan_regtab<32, false>;
    return nullptr;
}
