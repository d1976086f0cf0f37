// explicit instantiations of the tcgen05 GEMM kernel (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace uniir {
UNIIR_INSTANTIATE_GEMM(128, false, false, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(256, false, false, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(128, false, true, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(256, false, true, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(128, true, false, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(256, true, false, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(128, true, true, EPI_DYN)
UNIIR_INSTANTIATE_GEMM(256, true, true, EPI_DYN)
}  // namespace uniir
