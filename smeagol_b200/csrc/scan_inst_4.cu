// Explicit instantiations of the fused scan kernel (part 4): both strands per lane, padded widths 22..26.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part4(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 22) return &smg::launch_scan_regtab<22, true>;
    if (wb == 23) return &smg::launch_scan_regtab<23, true>;
    if (wb == 24) return &smg::launch_scan_regtab<24, true>;
    if (wb == 25) return &smg::launch_scan_regtab<25, true>;
    if (wb == 26) return &smg::launch_scan_regtab<26, true>;
    return nullptr;
}
