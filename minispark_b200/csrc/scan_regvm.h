// scan_regvm.h -- host-visible declarations of the regvm dense aggregate kernels (scan_regvm_ng<N>.cu).
#pragma once
#include "regvm_handlers.h"
#include "scan_kernel.cuh"

#define MSC_RV_MAX_CODE 128

namespace mscan {

struct RegvmProgram {
  uint32_t n;
  uint32_t code[MSC_RV_MAX_CODE];
};

// variant 0: any number of groups; variants 1..MSC_RV_MAX_NG: exactly that many groups, SUM_F / COUNT only
int launch_regvm_dense_ng0(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);
int launch_regvm_dense_ng1(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);
int launch_regvm_dense_ng2(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);
int launch_regvm_dense_ng3(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);
int launch_regvm_dense_ng4(msc_ctx* ctx, LaunchPlan* lp, const RegvmProgram* prog);

}  // namespace mscan
