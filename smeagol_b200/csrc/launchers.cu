// Launcher lookup across the instantiation translation units.
#include "launchers.h"

smg_launch_fn smg_get_launcher(int wb, bool dual)
{
    smg_launch_fn f = nullptr;
    if ((f = smg_get_launcher_part0(wb, dual))) return f;
    if ((f = smg_get_launcher_part1(wb, dual))) return f;
    if ((f = smg_get_launcher_part2(wb, dual))) return f;
    if ((f = smg_get_launcher_part3(wb, dual))) return f;
    if ((f = smg_get_launcher_part4(wb, dual))) return f;
    if ((f = smg_get_launcher_part5(wb, dual))) return f;
    return nullptr;
}
