// Host-side helpers shared by the C-ABI translation units: error reporting, device
// properties, TMA tensor-map encoding through the driver entry point (no -lcuda link).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/uniir_b200.h"

namespace uniir {

void set_error(const char* fmt, ...);  // api.cu
int num_sms();                         // api.cu (cached, current device)
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// true the first time it is called with this mask on the current device: function attributes
// (cudaFuncSetAttribute) are per device, and one process may drive several GPUs (single-process retrieval)
inline bool first_on_device(std::atomic<uint64_t>& mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  const uint64_t bit = 1ull << dev;
  return (mask.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
}

#define UNIIR_CHECK_ARG(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      ::uniir::set_error(__VA_ARGS__);          \
      return UNIIR_ERR_INVALID_ARG;             \
    }                                           \
  } while (0)

#define UNIIR_CUDA_OK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::uniir::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                      \
      return UNIIR_ERR_CUDA;                                                             \
    }                                                                                    \
  } while (0)

// 2-D tensor map over a row-major [outer, inner] matrix of 2-byte elements with a
// (box_inner x box_outer) box, 128B swizzle, zero fill out of bounds.
// dtype: 0 = bf16, 1 = fp16.   Returns 0 on success.
int make_tmap_2d(CUtensorMap* out, const void* gptr, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer);

}  // namespace uniir
