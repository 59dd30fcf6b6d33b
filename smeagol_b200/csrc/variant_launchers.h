// Dispatch table of the batched SNV scoring kernel instantiations (dual-strand padded widths 2..16).
#pragma once
#include <cuda_runtime.h>

namespace smg { struct VariantParams; }
typedef void (*smg_variant_launch_fn)(const smg::VariantParams&, int n_blocks, int n_groups, int nm, cudaStream_t);
smg_variant_launch_fn smg_get_variant_launcher(int wb);
