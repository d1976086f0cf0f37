// Small device helpers shared by the tcgen05 attention kernels (attention_tc.cu forward,
// attention_bwd_tc.cu backward).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace uniir {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

// byte offset of 16-byte chunk `ch` of row `r` in a [rows x 64] bf16 tile with the TMA 128B swizzle
__device__ __forceinline__ uint32_t sw128(int r, int ch) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4));
}

// bit 0: tcgen05 forward, bit 1: tcgen05 backward (uniir_attn_set_mode / UNIIR_ATTN_TC)
int attn_tc_mode_bits();

}  // namespace uniir
