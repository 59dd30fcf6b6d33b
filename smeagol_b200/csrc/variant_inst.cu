// Explicit instantiations of the batched SNV scoring kernel.
#include "variant_regtab.cuh"
#include "variant_launchers.h"

smg_variant_launch_fn smg_get_variant_launcher(int wb)
{
    switch (wb) {
    case 2: return &smg::launch_variant_regtab<2>;
    case 3: return &smg::launch_variant_regtab<3>;
    case 4: return &smg::launch_variant_regtab<4>;
    case 5: return &smg::launch_variant_regtab<5>;
    case 6: return &smg::launch_variant_regtab<6>;
    case 7: return &smg::launch_variant_regtab<7>;
    case 8: return &smg::launch_variant_regtab<8>;
    case 9: return &smg::launch_variant_regtab<9>;
    case 10: return &smg::launch_variant_regtab<10>;
    case 11: return &smg::launch_variant_regtab<11>;
    case 12: return &smg::launch_variant_regtab<12>;
    case 13: return &smg::launch_variant_regtab<13>;
    case 14: return &smg::launch_variant_regtab<14>;
    case 15: return &smg::launch_variant_regtab<15>;
    case 16: return &smg::launch_variant_regtab<16>;
    default: return nullptr;
    }
}
