// explicit instantiations of the CTA-pair (cta_group::2) tcgen05 GEMM kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace uniir {
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_STORE32)
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_ACT)
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_ACT_AUX)
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_RES32)
UNIIR_INSTANTIATE_GEMM_2CTA(false, false, EPI_RES32_RS)
UNIIR_INSTANTIATE_GEMM_2CTA(false, true, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM_2CTA(false, true, EPI_AGS)
UNIIR_INSTANTIATE_GEMM_2CTA(true, true, EPI_ACC32)
}  // namespace uniir
