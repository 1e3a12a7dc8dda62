// scan_regvm.h -- host-visible declarations of the regvm dense aggregate kernel (scan_regvm.cu).
#pragma once
#include "regvm_handlers.h"
#include "scan_kernel.cuh"

#define MSC_RV_MAX_CODE 128

namespace mscan {

struct RegvmProgram {
  uint32_t n;
  uint32_t code[MSC_RV_MAX_CODE];
};

int launch_regvm_dense(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);

}  // namespace mscan
