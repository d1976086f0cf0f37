// Error reporting, device queries and TMA tensor-map encoding for libuniir_b200.so.
#include <stdarg.h>
#include <atomic>

#include "host_util.cuh"

namespace uniir {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* gptr, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return UNIIR_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(gptr) & 15) != 0 || (row_stride_elems * 2) % 16 != 0) {
    set_error("TMA operand must be 16B aligned with a 16B-multiple row stride (ptr=%p ld=%llu)",
              gptr, (unsigned long long)row_stride_elems);
    return UNIIR_ERR_INVALID_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype == UNIIR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                            : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)",
              (int)r, (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)row_stride_elems);
    return UNIIR_ERR_CUDA;
  }
  return UNIIR_OK;
}

}  // namespace uniir

extern "C" {
int uniir_abi_version(void) { return UNIIR_ABI_VERSION; }
const char* uniir_last_error(void) { return uniir::g_err; }
int64_t uniir_launch_count(void) { return uniir::g_launches.load(); }
}
