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
    case 17: return &smg::launch_variant_regtab<17>;
    case 18: return &smg::launch_variant_regtab<18>;
    case 19: return &smg::launch_variant_regtab<19>;
    case 20: return &smg::launch_variant_regtab<20>;
    case 21: return &smg::launch_variant_regtab<21>;
    case 22: return &smg::launch_variant_regtab<22>;
    case 23: return &smg::launch_variant_regtab<23>;
    case 24: return &smg::launch_variant_regtab<24>;
    case 25: return &smg::launch_variant_regtab<25>;
    case 26: return &smg::launch_variant_regtab<26>;
    case 27: return &smg::launch_variant_regtab<27>;
    case 28: return &smg::launch_variant_regtab<28>;
    case 29: return &smg::launch_variant_regtab<29>;
    case 30: return &smg::launch_variant_regtab<30>;
    case 31: return &smg::launch_variant_regtab<31>;
    case 32: return &smg::launch_variant_regtab<32>;
    default: return nullptr;
    }
}
