// The tcgen05 / TMEM prefilter kernel and its launcher.
#include "scan_tc5.cuh"
#include "tc5_launchers.h"

cudaError_t smg_launch_tc5(const smg::Tc5Params& hp, int grid_x, int n_colgroups, int max_image_bytes, cudaStream_t stream)
{
    const size_t smem = smg::tc5_smem_bytes(max_image_bytes);
    cudaError_t e = cudaFuncSetAttribute(smg::scan_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)grid_x, (unsigned)n_colgroups, 1);
    smg::scan_tc5_kernel<<<grid, smg::kTc5Threads, smem, stream>>>(hp);
    return cudaGetLastError();
}
int smg_tc5_max_image_bytes() { return smg::kTc5MaxImageBytes; }
int smg_tc5_max_blocks() { return smg::kTc5MaxBlk; }
