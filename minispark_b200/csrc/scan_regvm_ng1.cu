// scan_regvm_ng1.cu -- regvm dense aggregate kernel, variant NG = 1 (see scan_regvm_impl.cuh, gen_regvm.py).
#define MSC_RV_NG 1
#define MSC_RV_MIN_CTAS 4
#define MSC_RV_PTX_INC "regvm_ptx_ng1.inc"
#include "scan_regvm_impl.cuh"
