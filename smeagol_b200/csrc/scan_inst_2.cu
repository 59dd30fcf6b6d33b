// Explicit instantiations of the fused scan kernel (part 2).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part2(int wb, int nm)
{
    if (wb == 13 && nm == 1) return &smg::launch_scan_regtab<13, 1>;
    if (wb == 14 && nm == 1) return &smg::launch_scan_regtab<14, 1>;
    if (wb == 15 && nm == 1) return &smg::launch_scan_regtab<15, 1>;
    if (wb == 16 && nm == 1) return &smg::launch_scan_regtab<16, 1>;
    return nullptr;
}
