// Explicit instantiations of the fused scan kernel (part 0).
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part0(int wb, int nm)
{
    if (wb == 2 && nm == 1) return &smg::launch_scan_regtab<2, 1>;
    if (wb == 3 && nm == 1) return &smg::launch_scan_regtab<3, 1>;
    if (wb == 4 && nm == 1) return &smg::launch_scan_regtab<4, 1>;
    if (wb == 5 && nm == 1) return &smg::launch_scan_regtab<5, 1>;
    if (wb == 6 && nm == 1) return &smg::launch_scan_regtab<6, 1>;
    if (wb == 7 && nm == 1) return &smg::launch_scan_regtab<7, 1>;
    if (wb == 8 && nm == 1) return &smg::launch_scan_regtab<8, 1>;
    return nullptr;
}
