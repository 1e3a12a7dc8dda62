// Explicit instantiation of the fused scan kernel: 8 rows per lane, MODE_DENSE (see scan_kernel.cuh).
#include "scan_kernel.cuh"

template int mscan::launch_scan<8, mscan::MODE_DENSE>(msc_ctx*, mscan::LaunchPlan*);
