// Instantiations with two PWMs per lane (part 7).  Only built with -DSMG_BUILD_PAIRS: measured slower than one PWM
// per lane at every width (fewer resident warps, doubled hit path), see DESIGN.md.
#include "launchers.h"
#ifdef SMG_BUILD_PAIRS
#include "scan_regtab.cuh"
#endif

smg_launch_fn smg_get_launcher_part7(int wb, int nm)
{
#ifdef SMG_BUILD_PAIRS
    if (wb == 9 && nm == 2) return &smg::launch_scan_regtab<9, 2>;
    if (wb == 10 && nm == 2) return &smg::launch_scan_regtab<10, 2>;
    if (wb == 11 && nm == 2) return &smg::launch_scan_regtab<11, 2>;
    if (wb == 12 && nm == 2) return &smg::launch_scan_regtab<12, 2>;
#endif
    (void)wb; (void)nm;
    return nullptr;
}
