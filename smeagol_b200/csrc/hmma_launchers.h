// Dispatch table of the tensor-core prefilter instantiations (KS k16-steps = ceil(W/4), NB 16-column blocks per warp).
#pragma once
#include <cuda_runtime.h>

namespace smg { struct HmmaParams; }
typedef void (*smg_hmma_launch_fn)(const smg::HmmaParams&, int n_chunks, int n_colgroups, cudaStream_t);
smg_hmma_launch_fn smg_get_hmma_launcher(int ks, int nb);
