// Explicit instantiation of the fused scan kernel: 8 rows per lane, MODE_HASH (see scan_kernel.cuh).
#include "scan_kernel.cuh"

template int mscan::launch_scan<8, mscan::MODE_HASH>(msc_ctx*, mscan::LaunchPlan*);
