// Explicit instantiations of the fused scan kernel (part 3): both strands per lane, padded widths 17..21.
#include "scan_regtab.cuh"
#include "launchers.h"

smg_launch_fn smg_get_launcher_part3(int wb, bool dual)
{
    if (!dual) return nullptr;  // one-strand-per-lane groups are not built any more (see api.cu)
    if (wb == 17) return &smg::launch_scan_regtab<17, true>;
    if (wb == 18) return &smg::launch_scan_regtab<18, true>;
    if (wb == 19) return &smg::launch_scan_regtab<19, true>;
    if (wb == 20) return &smg::launch_scan_regtab<20, true>;
    if (wb == 21) return &smg::launch_scan_regtab<21, true>;
    return nullptr;
}
