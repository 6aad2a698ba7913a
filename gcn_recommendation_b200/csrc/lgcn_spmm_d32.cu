// lgcn_spmm_d32.cu -- the SpMM kernels of lgcn_spmm_impl.cuh for 32-float table rows.
#include "lgcn_spmm_impl.cuh"

LGCN_SPMM_DEFINE_WIDTH(32)
