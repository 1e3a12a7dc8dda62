// Explicit instantiation of the fused scan kernel: 8 rows per lane, MODE_RUNS (see scan_kernel.cuh).
#include "scan_kernel.cuh"

template int mscan::launch_scan<8, mscan::MODE_RUNS>(msc_ctx*, mscan::LaunchPlan*);
