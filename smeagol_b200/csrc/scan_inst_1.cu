// Explicit instantiations of the fused scan kernel (part 1): both strands per lane, padded widths 9..12.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part1(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 9) return &smg::launch_scan_regtab<9, true>;
    if (wb == 10) return &smg::launch_scan_regtab<10, true>;
    if (wb == 11) return &smg::launch_scan_regtab<11, true>;
    if (wb == 12) return &smg::launch_scan_regtab<12, true>;
    return nullptr;
}
