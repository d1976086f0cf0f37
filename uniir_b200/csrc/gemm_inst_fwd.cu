// explicit instantiations of the tcgen05 GEMM kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace uniir {
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_STORE32)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_STORE32)
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_ACT)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_ACT)
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_ACT_AUX)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_ACT_AUX)
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_RES32)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_RES32)
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_RES32_RS)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_RES32_RS)
}  // namespace uniir
