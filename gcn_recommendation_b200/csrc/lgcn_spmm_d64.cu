// lgcn_spmm_d64.cu -- the SpMM kernels of lgcn_spmm_impl.cuh for 64-float table rows.
#include "lgcn_spmm_impl.cuh"

LGCN_SPMM_DEFINE_WIDTH(64)
