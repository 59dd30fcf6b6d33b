// Dispatch table of the register-table scan kernel instantiations (one per padded width / strand packing).
#pragma once
#include <cuda_runtime.h>

struct ScanParams;
typedef void (*smg_launch_fn)(const ScanParams&, int n_chunks, int n_groups, cudaStream_t);
smg_launch_fn smg_get_launcher(int wb, bool dual);
smg_launch_fn smg_get_launcher_part0(int wb, bool dual);
smg_launch_fn smg_get_launcher_part1(int wb, bool dual);
smg_launch_fn smg_get_launcher_part2(int wb, bool dual);
smg_launch_fn smg_get_launcher_part3(int wb, bool dual);
smg_launch_fn smg_get_launcher_part4(int wb, bool dual);
smg_launch_fn smg_get_launcher_part5(int wb, bool dual);
