// Dispatch table of the register-table scan kernel instantiations: padded width WB (2..32) x PWMs per lane NM (1, 2).
#pragma once
#include <cuda_runtime.h>

struct ScanParams;
typedef void (*smg_launch_fn)(const ScanParams&, int n_chunks, int n_groups, cudaStream_t);
smg_launch_fn smg_get_launcher(int wb, int nm);
smg_launch_fn smg_get_launcher_part0(int wb, int nm);
smg_launch_fn smg_get_launcher_part1(int wb, int nm);
smg_launch_fn smg_get_launcher_part2(int wb, int nm);
smg_launch_fn smg_get_launcher_part3(int wb, int nm);
smg_launch_fn smg_get_launcher_part4(int wb, int nm);
smg_launch_fn smg_get_launcher_part5(int wb, int nm);
smg_launch_fn smg_get_launcher_part6(int wb, int nm);
smg_launch_fn smg_get_launcher_part7(int wb, int nm);
