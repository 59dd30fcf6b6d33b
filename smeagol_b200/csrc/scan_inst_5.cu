// Explicit instantiations of the fused scan kernel (part 5).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part5(int wb, int nm)
{
    if (wb == 27 && nm == 1) return &smg::launch_scan_regtab<27, 1>;
    if (wb == 28 && nm == 1) return &smg::launch_scan_regtab<28, 1>;
    if (wb == 29 && nm == 1) return &smg::launch_scan_regtab<29, 1>;
    if (wb == 30 && nm == 1) return &smg::launch_scan_regtab<30, 1>;
    if (wb == 31 && nm == 1) return &smg::launch_scan_regtab<31, 1>;
    if (wb == 32 && nm == 1) return &smg::launch_scan_regtab<32, 1>;
    return nullptr;
}
