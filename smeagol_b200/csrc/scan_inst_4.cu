// Explicit instantiations of the fused scan kernel (part 4).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part4(int wb, int nm)
{
    if (wb == 22 && nm == 1) return &smg::launch_scan_regtab<22, 1>;
    if (wb == 23 && nm == 1) return &smg::launch_scan_regtab<23, 1>;
    if (wb == 24 && nm == 1) return &smg::launch_scan_regtab<24, 1>;
    if (wb == 25 && nm == 1) return &smg::launch_scan_regtab<25, 1>;
    if (wb == 26 && nm == 1) return &smg::launch_scan_regtab<26, 1>;
    return nullptr;
}
