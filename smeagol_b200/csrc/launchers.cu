// Launcher lookup across the instantiation translation units.
#include "launchers.h"

smg_launch_fn smg_get_launcher(int wb, int nm)
{
    smg_launch_fn f = nullptr;
    if ((f = smg_get_launcher_part0(wb, nm))) return f;
    if ((f = smg_get_launcher_part1(wb, nm))) return f;
    if ((f = smg_get_launcher_part2(wb, nm))) return f;
    if ((f = smg_get_launcher_part3(wb, nm))) return f;
    if ((f = smg_get_launcher_part4(wb, nm))) return f;
    if ((f = smg_get_launcher_part5(wb, nm))) return f;
    if ((f = smg_get_launcher_part6(wb, nm))) return f;
    if ((f = smg_get_launcher_part7(wb, nm))) return f;
    return nullptr;
}
