// Instantiations with two PWMs per lane (part 6).  Only built with -DSMG_BUILD_PAIRS: measured slower than one PWM
// per lane at every width (fewer resident warps, doubled hit path), see DESIGN.md.
#include "launchers.h"
#ifdef SMG_BUILD_PAIRS
#include "scan_regtab.cuh"
#endif

smg_launch_fn smg_get_launcher_part6(int wb, int nm)
{
#ifdef SMG_BUILD_PAIRS
    if (wb == 2 && nm == 2) return &smg::launch_scan_regtab<2, 2>;
    if (wb == 3 && nm == 2) return &smg::launch_scan_regtab<3, 2>;
    if (wb == 4 && nm == 2) return &smg::launch_scan_regtab<4, 2>;
    if (wb == 5 && nm == 2) return &smg::launch_scan_regtab<5, 2>;
    if (wb == 6 && nm == 2) return &smg::launch_scan_regtab<6, 2>;
    if (wb == 7 && nm == 2) return &smg::launch_scan_regtab<7, 2>;
    if (wb == 8 && nm == 2) return &smg::launch_scan_regtab<8, 2>;
#endif
    (void)wb; (void)nm;
    return nullptr;
}
