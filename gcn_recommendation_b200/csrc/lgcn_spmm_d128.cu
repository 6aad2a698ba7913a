// lgcn_spmm_d128.cu -- the SpMM kernels of lgcn_spmm_impl.cuh for 128-float table rows.
#include "lgcn_spmm_impl.cuh"

LGCN_SPMM_DEFINE_WIDTH(128)
