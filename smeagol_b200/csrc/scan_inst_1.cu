// Explicit instantiations of the fused scan kernel (part 1).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part1(int wb, int nm)
{
    if (wb == 9 && nm == 1) return &smg::launch_scan_regtab<9, 1>;
    if (wb == 10 && nm == 1) return &smg::launch_scan_regtab<10, 1>;
    if (wb == 11 && nm == 1) return &smg::launch_scan_regtab<11, 1>;
    if (wb == 12 && nm == 1) return &smg::launch_scan_regtab<12, 1>;
    return nullptr;
}
