// Explicit instantiations of the tensor-core prefilter kernel.
#include "scan_hmma.cuh"
#include "hmma_launchers.h"

smg_hmma_launch_fn smg_get_hmma_launcher(int ks, int nb)
{
    if (ks == 1 && nb == 4) return &smg::launch_scan_hmma<1, 4>;
    if (ks == 2 && nb == 4) return &smg::launch_scan_hmma<2, 4>;
    if (ks == 3 && nb == 4) return &smg::launch_scan_hmma<3, 4>;
    if (ks == 4 && nb == 4) return &smg::launch_scan_hmma<4, 4>;
    if (ks == 5 && nb == 4) return &smg::launch_scan_hmma<5, 4>;
    if (ks == 6 && nb == 4) return &smg::launch_scan_hmma<6, 4>;
    if (ks == 2 && nb == 3) return &smg::launch_scan_hmma<2, 3>;
    if (ks == 7 && nb == 3) return &smg::launch_scan_hmma<7, 3>;
    if (ks == 8 && nb == 3) return &smg::launch_scan_hmma<8, 3>;
    return nullptr;
}
