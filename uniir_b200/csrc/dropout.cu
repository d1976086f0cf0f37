// Dropout / stochastic depth for the training arms that have them in the reference (BLIP BERT: med.py:97,244,325;
// BLIP ViT-L DropPath: vit.py:151-152 + blip.py:253; CLIP_FF T5 stack: dropout_rate 0.1).  Masks are counter-based
// (dropout.cuh); the attention-probability dropout (med.py:217) lives in the attention kernels (attention.cu).
#include <cuda_bf16.h>

#include "dropout.cuh"
#include "host_util.cuh"

namespace uniir {

template <typename T> __device__ __forceinline__ float ld1(const T* p, long long i);
template <> __device__ __forceinline__ float ld1<float>(const float* p, long long i) { return p[i]; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void st1(T* p, long long i, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, long long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// out[i] = (keep(i) ? x[i] / (1 - p) : 0) + res[i];  element i of this micro-batch is element
// ctx[1] * per_item + i of the batch.  4 elements per thread (the widths in use are multiples of 8).
template <typename TI, typename TO>
__global__ void dropout_kernel(const TI* __restrict__ x, const float* __restrict__ res, TO* __restrict__ out, long long n,
                               const uint32_t* __restrict__ ctx, uint32_t site, uint32_t per_item, uint32_t thresh,
                               float scale) {
  const uint32_t key = drop_key(ctx, site);
  const uint32_t base = __ldg(ctx + 1) * per_item;
  const long long stride = 4LL * gridDim.x * blockDim.x;
  for (long long i0 = 4LL * (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x); i0 < n; i0 += stride) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = i0 + j;
      if (i < n) {
        float v = drop_keep(key, base + static_cast<uint32_t>(i), thresh) ? __fmul_rn(ld1(x, i), scale) : 0.f;   // no FMA with the residual
        if (res) v = __fadd_rn(v, res[i]);
        st1(out, i, v);
      }
    }
  }
}

__global__ void dropout_mask_kernel(const uint32_t* __restrict__ ctx, uint32_t site, uint32_t per_item, long long n,
                                    uint32_t thresh, uint8_t* __restrict__ keep) {
  const uint32_t key = drop_key(ctx, site);
  const uint32_t base = __ldg(ctx + 1) * per_item;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = drop_keep(key, base + static_cast<uint32_t>(i), thresh) ? 1 : 0;
}

// scales[s * mb + b] = keep ? 1 / (1 - p[s]) : 0 for DropPath site s and item b (timm DropPath, scale_by_keep)
__global__ void droppath_scales_kernel(const uint32_t* __restrict__ ctx, uint32_t site_base, int n_sites, int mb,
                                       const float* __restrict__ p_sites, float* __restrict__ scales) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sites * mb) return;
  const int s = i / mb, b = i % mb;
  const float p = p_sites[s];
  const double t = static_cast<double>(p) * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : static_cast<uint32_t>(t));
  const bool keep = drop_keep(drop_key(ctx, site_base + s), __ldg(ctx + 1) + b, thresh);
  scales[i] = keep ? 1.f / (1.f - p) : 0.f;
}

// x[r, c] *= scales[r / rows_per_scale]   (bf16, in place): the branch gradient under DropPath
__global__ void row_scale_kernel(__nv_bfloat16* __restrict__ x, long long rows, int cols, long long ld,
                                 const float* __restrict__ scales, int rows_per_scale) {
  const int cpr = cols >> 3;   // 16 B chunks per row
  const long long total = rows * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cpr;
    const int c = static_cast<int>(i - r * cpr) << 3;
    const float s = __ldg(scales + r / rows_per_scale);
    if (s == 1.f) continue;
    uint4* p = reinterpret_cast<uint4*>(x + r * ld + c);
    uint4 v = *p;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      h[j] = __floats2bfloat162_rn(f.x * s, f.y * s);
    }
    *p = v;
  }
}

}  // namespace uniir

using namespace uniir;
#define ST(s) static_cast<cudaStream_t>(s)

static unsigned blocks_for(long long threads) {
  const long long cap = 148LL * 16;
  long long b = (threads + 255) / 256;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int uniir_dropout(const void* x, int32_t x_dtype, const float* residual, void* out, int32_t out_dtype,
                             int64_t n, const uint32_t* ctx, uint32_t site, int64_t per_item, float p, void* stream) {
  UNIIR_CHECK_ARG(x && out && ctx && n >= 0 && per_item > 0 && p >= 0.f && p < 1.f, "uniir_dropout: bad args");
  UNIIR_CHECK_ARG((x_dtype == UNIIR_F32 || x_dtype == UNIIR_BF16) && (out_dtype == UNIIR_F32 || out_dtype == UNIIR_BF16),
                  "uniir_dropout: x / out must be fp32 or bf16");
  if (n == 0) return UNIIR_OK;
  const uint32_t th = drop_thresh(p);
  const float sc = 1.f / (1.f - p);
  const unsigned g = blocks_for((n + 3) / 4);
  const uint32_t pi = static_cast<uint32_t>(per_item);
#define GO(TI, TO)                                                                                                     \
  dropout_kernel<TI, TO><<<g, 256, 0, ST(stream)>>>(static_cast<const TI*>(x), residual, static_cast<TO*>(out), n, ctx, \
                                                    site, pi, th, sc)
  if (x_dtype == UNIIR_F32 && out_dtype == UNIIR_F32) GO(float, float);
  else if (x_dtype == UNIIR_F32) GO(float, __nv_bfloat16);
  else if (out_dtype == UNIIR_F32) GO(__nv_bfloat16, float);
  else GO(__nv_bfloat16, __nv_bfloat16);
#undef GO
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_dropout_mask(const uint32_t* ctx, uint32_t site, int64_t per_item, int64_t n, float p, uint8_t* keep,
                                  void* stream) {
  UNIIR_CHECK_ARG(ctx && keep && n >= 0 && per_item > 0 && p >= 0.f && p < 1.f, "uniir_dropout_mask: bad args");
  if (n == 0) return UNIIR_OK;
  dropout_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, ST(stream)>>>(
      ctx, site, static_cast<uint32_t>(per_item), n, drop_thresh(p), keep);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_droppath_scales(const uint32_t* ctx, uint32_t site_base, int32_t n_sites, int32_t mb,
                                     const float* p_sites, float* scales, void* stream) {
  UNIIR_CHECK_ARG(ctx && p_sites && scales && n_sites > 0 && mb > 0, "uniir_droppath_scales: bad args");
  const int n = n_sites * mb;
  droppath_scales_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(ctx, site_base, n_sites, mb, p_sites, scales);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_row_scale(void* x_bf16, int64_t rows, int32_t cols, int64_t ld, const float* scales,
                               int32_t rows_per_scale, void* stream) {
  UNIIR_CHECK_ARG(x_bf16 && scales && rows >= 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && rows_per_scale > 0,
                  "uniir_row_scale: bad args (cols and ld must be multiples of 8)");
  if (rows == 0) return UNIIR_OK;
  row_scale_kernel<<<blocks_for(rows * (cols / 8)), 256, 0, ST(stream)>>>(static_cast<__nv_bfloat16*>(x_bf16), rows, cols, ld,
                                                                          scales, rows_per_scale);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
