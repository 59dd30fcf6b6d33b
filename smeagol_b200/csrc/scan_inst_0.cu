// Explicit instantiations of the fused scan kernel (part 0): both strands per lane, padded widths 2..8.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part0(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 2) return &smg::launch_scan_regtab<2, true>;
    if (wb == 3) return &smg::launch_scan_regtab<3, true>;
    if (wb == 4) return &smg::launch_scan_regtab<4, true>;
    if (wb == 5) return &smg::launch_scan_regtab<5, true>;
    if (wb == 6) return &smg::launch_scan_regtab<6, true>;
    if (wb == 7) return &smg::launch_scan_regtab<7, true>;
    if (wb == 8) return &smg::launch_scan_regtab<8, true>;
    return nullptr;
}
