// Launcher of the tcgen05 / TMEM prefilter (scan_tc5.cuh), compiled in its own translation unit (tc5_inst.cu).
#pragma once
#include <cuda_runtime.h>

namespace smg { struct Tc5Params; }
// grid_x persistent CTAs per column group (grid_x * n_colgroups <= number of SMs), max_image_bytes = largest weight image
cudaError_t smg_launch_tc5(const smg::Tc5Params& hp, int grid_x, int n_colgroups, int max_image_bytes, cudaStream_t stream);
int smg_tc5_max_image_bytes();
int smg_tc5_max_blocks();
