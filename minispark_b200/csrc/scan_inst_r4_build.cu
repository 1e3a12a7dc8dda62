// Explicit instantiation of the fused scan kernel: 4 rows per lane, MODE_BUILD (see scan_kernel.cuh).
#include "scan_kernel.cuh"

template int mscan::launch_scan<4, mscan::MODE_BUILD>(msc_ctx*, mscan::LaunchPlan*);
