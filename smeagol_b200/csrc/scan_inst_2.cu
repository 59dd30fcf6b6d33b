// Explicit instantiations of the fused scan kernel (part 2): both strands per lane, padded widths 13..16.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part2(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 13) return &smg::launch_scan_regtab<13, true>;
    if (wb == 14) return &smg::launch_scan_regtab<14, true>;
    if (wb == 15) return &smg::launch_scan_regtab<15, true>;
    if (wb == 16) return &smg::launch_scan_regtab<16, true>;
    return nullptr;
}
