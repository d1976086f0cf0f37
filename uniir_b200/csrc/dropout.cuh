// Counter-based dropout masks shared by dropout.cu and the attention kernels.
//
// The reference trains its BERT (med.py:97,217,244,325: hidden_dropout_prob / attention_probs_dropout_prob = 0.1),
// its ViT-L (vit.py:151-152, timm DropPath, rate 0.1) and its T5 fusion stack (dropout_rate 0.1) with torch's
// stateful Philox masks.  Here a mask is a pure function of (step seed, site, element index of the WHOLE batch), so
//   * nothing is stored: the gradient-cache re-forward and the backward regenerate the same bits,
//   * the micro-batch tiling does not matter (phase A and phase C tile the batch differently),
//   * CUDA graphs replay with a new seed: seed and first item index live in device memory (`ctx`).
// keep(idx) = mix32(idx ^ key(seed, site)) >= floor(p * 2^32); mix32 is a three-round 32-bit finaliser.
#pragma once
#include <stdint.h>

namespace uniir {

// Three-round multiply-xorshift finaliser.  A two-round one (0x7feb352d / 0x846ca68b) left a measurable correlation
// between elements 768 apart -- the same feature of neighbouring tokens at width 768 (6 sigma over 2^21 draws,
// tests/test_host_logic.py); with these constants every lag, site and seed pair tested is within chance.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 17; x *= 0xed5ad4bbu;
  x ^= x >> 11; x *= 0xac4c1b51u;
  x ^= x >> 15; x *= 0x31848babu;
  x ^= x >> 14;
  return x;
}
// ctx[0] = seed of this step (and encoder), ctx[1] = index of the micro-batch's first item within the batch
__device__ __forceinline__ uint32_t drop_key(const uint32_t* ctx, uint32_t site) {
  return mix32(__ldg(ctx) + site * 0x9E3779B1u);
}
__device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t idx, uint32_t thresh) {
  return mix32(idx ^ key) >= thresh;
}
inline uint32_t drop_thresh(float p) {
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : static_cast<uint32_t>(t));
}

}  // namespace uniir
