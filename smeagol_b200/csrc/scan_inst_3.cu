// Explicit instantiations of the fused scan kernel (part 3).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part3(int wb, int nm)
{
    if (wb == 17 && nm == 1) return &smg::launch_scan_regtab<17, 1>;
    if (wb == 18 && nm == 1) return &smg::launch_scan_regtab<18, 1>;
    if (wb == 19 && nm == 1) return &smg::launch_scan_regtab<19, 1>;
    if (wb == 20 && nm == 1) return &smg::launch_scan_regtab<20, 1>;
    if (wb == 21 && nm == 1) return &smg::launch_scan_regtab<21, 1>;
    return nullptr;
}
