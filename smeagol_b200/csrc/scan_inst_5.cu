// Explicit instantiations of the fused scan kernel (part 5): both strands per lane, padded widths 27..32.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part5(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 27) return &smg::launch_scan_regtab<27, true>;
    if (wb == 28) return &smg::launch_scan_regtab<28, true>;
    if (wb == 29) return &smg::launch_scan_regtab<29, true>;
    if (wb == 30) return &smg::launch_scan_regtab<30, true>;
    if (wb == 31) return &smg::launch_scan_regtab<31, true>;
    if (wb == 32) return &smg::launch_scan_regtab<32, true>;
    return nullptr;
}
