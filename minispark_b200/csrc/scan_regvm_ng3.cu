// scan_regvm_ng3.cu -- regvm dense aggregate kernel, variant NG = 3 (see scan_regvm_impl.cuh, gen_regvm.py).
#define MSC_RV_NG 3
#define MSC_RV_MIN_CTAS 3
#define MSC_RV_PTX_INC "regvm_ptx_ng3.inc"
#include "scan_regvm_impl.cuh"
