// Explicit instantiation of the fused scan kernel: 4 rows per lane, MODE_PROJECT (see scan_kernel.cuh).
#include "scan_kernel.cuh"

template int mscan::launch_scan<4, mscan::MODE_PROJECT>(msc_ctx*, mscan::LaunchPlan*);
