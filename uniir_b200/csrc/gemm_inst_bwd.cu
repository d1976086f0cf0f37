// explicit instantiations of the tcgen05 GEMM kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace uniir {
UNIIR_INSTANTIATE_GEMM(128, false, true, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM(256, false, true, EPI_STORE16)
UNIIR_INSTANTIATE_GEMM(128, false, true, EPI_STORE32)
UNIIR_INSTANTIATE_GEMM(256, false, true, EPI_STORE32)
UNIIR_INSTANTIATE_GEMM(128, false, true, EPI_AGS)
UNIIR_INSTANTIATE_GEMM(256, false, true, EPI_AGS)
UNIIR_INSTANTIATE_GEMM(128, true, true, EPI_ACC32)
UNIIR_INSTANTIATE_GEMM(256, true, true, EPI_ACC32)
UNIIR_INSTANTIATE_GEMM(128, true, true, EPI_ATOMIC)
UNIIR_INSTANTIATE_GEMM(256, true, true, EPI_ATOMIC)
}  // namespace uniir
