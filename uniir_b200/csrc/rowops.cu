// HBM-bound row / elementwise kernels around the tensor-core GEMMs: LayerNorm fwd/bwd,
// patch im2col, vision/text embedding assembly (+ backward), score fusion, bias-gradient
// column sums, pad/cast helpers and the fused AdamW step.  All use 128-bit coalesced
// accesses; none of them has data reuse beyond a row, so no shared-memory tiling except
// the block reductions.
#include <cuda_bf16.h>

#include "host_util.cuh"

namespace uniir {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over a block of up to 1024 threads; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* sh /*>=33 floats*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
__device__ __forceinline__ float4 unpack4_bf16(uint2 u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                     __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}

// ---------------------------------------------------------------------------------
// LayerNorm forward: one CTA per row, D <= 4 * blockDim * 2.  x fp32 (row r at
// x + rowmap(r)*ldx), y bf16 or fp32.  Two-pass (mean, then centred variance) in fp32 like
// torch.nn.LayerNorm / upstream CLIP's fp32 LayerNorm subclass.
// ---------------------------------------------------------------------------------
template <int VEC_PER_THREAD>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long ldx,
                                     const int* __restrict__ row_index, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, int D,
                                     void* __restrict__ y, long long ldy, int y_f32,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  __shared__ float sh[33];
  const long long r = blockIdx.x;
  const long long src = row_index ? row_index[r] : r;
  const float* xr = x + src * ldx;
  float4 v[VEC_PER_THREAD];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int c = (threadIdx.x + i * blockDim.x) * 4;
    v[i] = (c < D) ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0, 0, 0, 0);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = block_sum(s, sh) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int c = (threadIdx.x + i * blockDim.x) * 4;
    if (c < D) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(block_sum(q, sh) / D + eps);
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[r] = mean;
    if (rstd_out) rstd_out[r] = rstd;
  }
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int c = (threadIdx.x + i * blockDim.x) * 4;
    if (c < D) {
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 b = *reinterpret_cast<const float4*>(beta + c);
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (y_f32)
        *reinterpret_cast<float4*>(static_cast<float*>(y) + r * ldy + c) = make_float4(o0, o1, o2, o3);
      else
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + r * ldy + c) = pack4_bf16(o0, o1, o2, o3);
    }
  }
}

// LayerNorm backward.  dy (bf16 or fp32) [rows, D]; x fp32 (same row addressing as fwd);
//   dx[src row] (+)= rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat))      (fp32)
//   dgamma += sum_r dy*xhat ; dbeta += sum_r dy   (per-CTA partial over a strip of rows,
//   then one atomicAdd per column per CTA)
template <int VEC_PER_THREAD>
__global__ void layernorm_bwd_kernel(const void* __restrict__ dy, long long lddy, int dy_f32,
                                     const float* __restrict__ x, long long ldx,
                                     const int* __restrict__ row_index,
                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ gamma, int D, long long rows,
                                     int rows_per_cta, float* __restrict__ dx, long long lddx,
                                     int dx_accumulate, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dx16) {
  __shared__ float sh[33];
  float4 dg[VEC_PER_THREAD], db[VEC_PER_THREAD], g[VEC_PER_THREAD];
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int c = (threadIdx.x + i * blockDim.x) * 4;
    dg[i] = db[i] = make_float4(0, 0, 0, 0);
    g[i] = (c < D) ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0, 0, 0, 0);
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (long long r = r0; r < r1; ++r) {
    const long long src = row_index ? row_index[r] : r;
    const float mu = mean[r], rs = rstd[r];
    float4 xh[VEC_PER_THREAD], d[VEC_PER_THREAD];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
      const int c = (threadIdx.x + i * blockDim.x) * 4;
      if (c < D) {
        const float4 xv = *reinterpret_cast<const float4*>(x + src * ldx + c);
        d[i] = dy_f32 ? *reinterpret_cast<const float4*>(static_cast<const float*>(dy) + r * lddy + c)
                      : unpack4_bf16(*reinterpret_cast<const uint2*>(
                            static_cast<const __nv_bfloat16*>(dy) + r * lddy + c));
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        const float a0 = d[i].x * g[i].x, a1 = d[i].y * g[i].y, a2 = d[i].z * g[i].z, a3 = d[i].w * g[i].w;
        s1 += a0 + a1 + a2 + a3;
        s2 += a0 * xh[i].x + a1 * xh[i].y + a2 * xh[i].z + a3 * xh[i].w;
        dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y;
        dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
        db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      } else {
        xh[i] = d[i] = make_float4(0, 0, 0, 0);
      }
    }
    const float m1 = block_sum(s1, sh) / D;
    const float m2 = block_sum(s2, sh) / D;
#pragma unroll
    for (int i = 0; i < VEC_PER_THREAD; ++i) {
      const int c = (threadIdx.x + i * blockDim.x) * 4;
      if (c < D) {
        float4 o;
        o.x = rs * (d[i].x * g[i].x - m1 - xh[i].x * m2);
        o.y = rs * (d[i].y * g[i].y - m1 - xh[i].y * m2);
        o.z = rs * (d[i].z * g[i].z - m1 - xh[i].z * m2);
        o.w = rs * (d[i].w * g[i].w - m1 - xh[i].w * m2);
        float* dp = dx + src * lddx + c;
        if (dx_accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(dp);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        *reinterpret_cast<float4*>(dp) = o;
        // optional bf16 copy of the updated gradient: the next GEMMs' operand (saves a cast pass)
        if (dx16) *reinterpret_cast<uint2*>(dx16 + src * lddx + c) = pack4_bf16(o.x, o.y, o.z, o.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC_PER_THREAD; ++i) {
    const int c = (threadIdx.x + i * blockDim.x) * 4;
    if (c < D && r0 < r1) {
      atomicAdd(dgamma + c + 0, dg[i].x); atomicAdd(dgamma + c + 1, dg[i].y);
      atomicAdd(dgamma + c + 2, dg[i].z); atomicAdd(dgamma + c + 3, dg[i].w);
      atomicAdd(dbeta + c + 0, db[i].x); atomicAdd(dbeta + c + 1, db[i].y);
      atomicAdd(dbeta + c + 2, db[i].z); atomicAdd(dbeta + c + 3, db[i].w);
    }
  }
}

// ---------------------------------------------------------------------------------
// Warp-per-row LayerNorm (D <= 1024): the row lives in the registers of one warp (NV float4
// per lane, lane-interleaved so every access is a full 512 B line group), statistics by warp
// shuffles only -- no __syncthreads, no shared memory; several rows per CTA and ~10 KB per
// warp in flight keep HBM busy.  Same arithmetic as the CTA-per-row kernels above (two-pass
// centred variance in fp32), which remain the path for wider rows.
// ---------------------------------------------------------------------------------
constexpr int kLnWarpsFwd = 4;
template <int NV>
__global__ void __launch_bounds__(kLnWarpsFwd * 32)
layernorm_fwd_warp_kernel(const float* __restrict__ x, long long ldx, const int* __restrict__ row_index,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                          int D, long long rows, void* __restrict__ y, long long ldy, int y_f32,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * kLnWarpsFwd + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long src = row_index ? row_index[r] : r;
  const float* xr = x + src * ldx;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    v[i] = (c < D) ? __ldcs(reinterpret_cast<const float4*>(xr + c)) : make_float4(0, 0, 0, 0);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    if (c < D) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[r] = mean;
    if (rstd_out) rstd_out[r] = rstd;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    if (c < D) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float o0 = (v[i].x - mean) * rstd * g.x + b.x, o1 = (v[i].y - mean) * rstd * g.y + b.y;
      const float o2 = (v[i].z - mean) * rstd * g.z + b.z, o3 = (v[i].w - mean) * rstd * g.w + b.w;
      if (y_f32)
        *reinterpret_cast<float4*>(static_cast<float*>(y) + r * ldy + c) = make_float4(o0, o1, o2, o3);
      else
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + r * ldy + c) = pack4_bf16(o0, o1, o2, o3);
    }
  }
}

// T5LayerNorm: y = x * rsqrt(mean(x^2) + eps) * gamma.  Warp per row like the LayerNorm forward above.
template <int NV>
__global__ void __launch_bounds__(kLnWarpsFwd * 32)
rmsnorm_fwd_warp_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, float eps, int D,
                        long long rows, void* __restrict__ y, long long ldy, int y_f32) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * kLnWarpsFwd + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + r * ldx;
  float4 v[NV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    v[i] = (c < D) ? __ldcs(reinterpret_cast<const float4*>(xr + c)) : make_float4(0, 0, 0, 0);
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + i * 32) * 4;
    if (c < D) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float o0 = v[i].x * rstd * g.x, o1 = v[i].y * rstd * g.y, o2 = v[i].z * rstd * g.z, o3 = v[i].w * rstd * g.w;
      if (y_f32)
        *reinterpret_cast<float4*>(static_cast<float*>(y) + r * ldy + c) = make_float4(o0, o1, o2, o3);
      else
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + r * ldy + c) = pack4_bf16(o0, o1, o2, o3);
    }
  }
}

// out[b, c..c+3] = mean_s x[b, s, c..c+3]; one thread per float4 column of one item, consecutive threads on
// consecutive columns (coalesced rows)
__global__ void token_mean_kernel(const float* __restrict__ x, int S, int D, float* __restrict__ out) {
  const int b = blockIdx.y, c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const float* xb = x + static_cast<long long>(b) * S * D + c;
  float4 a = make_float4(0, 0, 0, 0);
#pragma unroll 4
  for (int s = 0; s < S; ++s) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(xb + static_cast<long long>(s) * D));
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  const float inv = 1.f / S;
  *reinterpret_cast<float4*>(out + static_cast<long long>(b) * D + c) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

// Backward, warp per row, 8 warps per CTA, each warp walking rows r0+w, r0+w+8, ... of the
// CTA's strip.  dgamma / dbeta (and, when `dcolsum` is given, the column sums of the UPDATED dx
// -- the bias gradient of the linear layer whose output gradient this dx is, which saves a
// separate pass over dx) accumulate per lane in registers, are reduced over the CTA's warps
// through shared memory and leave with one atomicAdd per column per CTA.
constexpr int kLnWarpsBwd = 8;
template <int NV, bool COLSUM, bool DY_F32>
__global__ void __launch_bounds__(kLnWarpsBwd * 32)
layernorm_bwd_warp_kernel(const void* __restrict__ dy, long long lddy,
                          const float* __restrict__ x, long long ldx, const int* __restrict__ row_index,
                          const float* __restrict__ mean, const float* __restrict__ rstd,
                          const float* __restrict__ gamma, int D, long long rows, int rows_per_cta,
                          float* __restrict__ dx, long long lddx, int dx_accumulate,
                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                          __nv_bfloat16* __restrict__ dx16, float* __restrict__ dcolsum) {
  __shared__ float4 red[kLnWarpsBwd][NV * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 dg[NV], db[NV], dc[COLSUM ? NV : 1];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = db[i] = make_float4(0, 0, 0, 0);
    if (COLSUM) dc[i] = make_float4(0, 0, 0, 0);
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (long long r = r0 + w; r < r1; r += kLnWarpsBwd) {
    const long long src = row_index ? row_index[r] : r;
    const float mu = mean[r], rs = rstd[r];
    float4 xh[NV], d[NV], old[NV];
    uint2 raw[DY_F32 ? 1 : NV];
    // Issue EVERY load of the row before the first use: nothing in this loop may consume a loaded
    // value (ncu showed the bf16 unpack placed here serialising the row into NV memory round trips).
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      if (c < D) {
        xh[i] = __ldcs(reinterpret_cast<const float4*>(x + src * ldx + c));
        if (DY_F32) d[i] = __ldcs(reinterpret_cast<const float4*>(static_cast<const float*>(dy) + r * lddy + c));
        else raw[DY_F32 ? 0 : i] = __ldcs(reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy) + r * lddy + c));
        if (dx_accumulate) old[i] = *reinterpret_cast<const float4*>(dx + src * lddx + c);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      if (c < D) {
        if (!DY_F32) d[i] = unpack4_bf16(raw[DY_F32 ? 0 : i]);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));   // 4 KB, L1-resident
        xh[i] = make_float4((xh[i].x - mu) * rs, (xh[i].y - mu) * rs, (xh[i].z - mu) * rs, (xh[i].w - mu) * rs);
        dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y;
        dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
        db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
        d[i] = make_float4(d[i].x * g.x, d[i].y * g.y, d[i].z * g.z, d[i].w * g.w);   // from here on: g * dy
        s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
        s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // the two reductions share the shuffle latency
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / D, m2 = s2 / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      if (c < D) {
        float4 o;
        o.x = rs * (d[i].x - m1 - xh[i].x * m2);
        o.y = rs * (d[i].y - m1 - xh[i].y * m2);
        o.z = rs * (d[i].z - m1 - xh[i].z * m2);
        o.w = rs * (d[i].w - m1 - xh[i].w * m2);
        if (dx_accumulate) { o.x += old[i].x; o.y += old[i].y; o.z += old[i].z; o.w += old[i].w; }
        *reinterpret_cast<float4*>(dx + src * lddx + c) = o;
        if (dx16) *reinterpret_cast<uint2*>(dx16 + src * lddx + c) = pack4_bf16(o.x, o.y, o.z, o.w);
        if (COLSUM) { dc[i].x += o.x; dc[i].y += o.y; dc[i].z += o.z; dc[i].w += o.w; }
      }
    }
  }
  // CTA reduction of the column sums: [warp][column-vector] in shared memory, one output at a time
  float* outs[3] = {dgamma, dbeta, dcolsum};
#pragma unroll
  for (int which = 0; which < (COLSUM ? 3 : 2); ++which) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) red[w][lane + i * 32] = which == 0 ? dg[i] : (which == 1 ? db[i] : dc[COLSUM ? i : 0]);
    __syncthreads();
    const float* rf = reinterpret_cast<const float*>(&red[0][0]);
    for (int col = threadIdx.x; col < D; col += kLnWarpsBwd * 32) {
      // column `col` sits in vector (col/4) at lane-interleaved slot: c = (lane + 32 i)*4 -> slot = c/4
      float t = 0.f;
#pragma unroll
      for (int ww = 0; ww < kLnWarpsBwd; ++ww) t += rf[(ww * NV * 32 + (col >> 2)) * 4 + (col & 3)];
      atomicAdd(outs[which] + col, t);
    }
  }
}

// T5LayerNorm backward (y = x * rstd * gamma, rstd = rsqrt(mean(x^2) + eps), recomputed from x -- nothing is saved):
//   g = dy o gamma,  x^ = x * rstd,  dx (+)= rstd * (g - x^ * mean(g o x^)),  dgamma += sum_rows dy o x^.
// Same structure as the LayerNorm backward above (warp per row, per-CTA column reduction, one atomic per column).
template <int NV>
__global__ void __launch_bounds__(kLnWarpsBwd * 32)
rmsnorm_bwd_warp_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                        float eps, int D, long long rows, int rows_per_cta, float* __restrict__ dx, int dx_accumulate,
                        float* __restrict__ dgamma) {
  __shared__ float4 red[kLnWarpsBwd][NV * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 dg[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = make_float4(0, 0, 0, 0);
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (long long r = r0 + w; r < r1; r += kLnWarpsBwd) {
    float4 xh[NV], d[NV], old[NV];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      xh[i] = d[i] = old[i] = make_float4(0, 0, 0, 0);
      if (c < D) {
        xh[i] = __ldcs(reinterpret_cast<const float4*>(x + r * D + c));
        d[i] = __ldcs(reinterpret_cast<const float4*>(dy + r * D + c));
        if (dx_accumulate) old[i] = *reinterpret_cast<const float4*>(dx + r * D + c);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) q += (xh[i].x * xh[i].x + xh[i].y * xh[i].y) + (xh[i].z * xh[i].z + xh[i].w * xh[i].w);
    const float rs = rsqrtf(warp_sum(q) / D + eps);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      if (c < D) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
        xh[i] = make_float4(xh[i].x * rs, xh[i].y * rs, xh[i].z * rs, xh[i].w * rs);
        dg[i].x += d[i].x * xh[i].x; dg[i].y += d[i].y * xh[i].y;
        dg[i].z += d[i].z * xh[i].z; dg[i].w += d[i].w * xh[i].w;
        d[i] = make_float4(d[i].x * g.x, d[i].y * g.y, d[i].z * g.z, d[i].w * g.w);   // from here on: gamma o dy
        s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
      }
    }
    const float m2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + i * 32) * 4;
      if (c < D) {
        float4 o;
        o.x = rs * (d[i].x - xh[i].x * m2) + old[i].x;
        o.y = rs * (d[i].y - xh[i].y * m2) + old[i].y;
        o.z = rs * (d[i].z - xh[i].z * m2) + old[i].z;
        o.w = rs * (d[i].w - xh[i].w * m2) + old[i].w;
        *reinterpret_cast<float4*>(dx + r * D + c) = o;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) red[w][lane + i * 32] = dg[i];
  __syncthreads();
  const float* rf = reinterpret_cast<const float*>(&red[0][0]);
  for (int col = threadIdx.x; col < D; col += kLnWarpsBwd * 32) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < kLnWarpsBwd; ++ww) t += rf[(ww * NV * 32 + (col >> 2)) * 4 + (col & 3)];
    atomicAdd(dgamma + col, t);
  }
}

// dx[b, s, c..c+3] = d[b, c..c+3] / S: backward of token_mean_kernel
__global__ void token_mean_bwd_kernel(const float* __restrict__ d, int S, int D, float* __restrict__ dx) {
  const int b = blockIdx.y, c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const float inv = 1.f / S;
  float4 v = *reinterpret_cast<const float4*>(d + static_cast<long long>(b) * D + c);
  v = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  float* o = dx + static_cast<long long>(b) * S * D + c;
  for (int s = 0; s < S; ++s) *reinterpret_cast<float4*>(o + static_cast<long long>(s) * D) = v;
}

// ---------------------------------------------------------------------------------
// im2col for the stride=patch conv: img fp32 [B,3,H,W] -> bf16 [B*g*g, Kpad],
// k = c*P*P + py*P + px (conv weight [out, c, py, px] flattened), zero padded to Kpad.
// ---------------------------------------------------------------------------------
__global__ void im2col_kernel(const float* __restrict__ img, int B, int H, int W, int P, int g,
                              int Kpad, __nv_bfloat16* __restrict__ out) {
  const long long row = blockIdx.x;  // b*g*g + gy*g + gx
  const int b = row / (g * g), rem = row % (g * g), gy = rem / g, gx = rem % g;
  const int K = 3 * P * P;
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    float v = 0.f;
    if (k < K) {
      const int c = k / (P * P), r2 = k % (P * P), py = r2 / P, px = r2 % P;
      v = img[((static_cast<long long>(b) * 3 + c) * H + gy * P + py) * W + gx * P + px];
    }
    out[row * Kpad + k] = __float2bfloat16(v);
  }
}

// x0[b, s, :] = (s == 0 ? cls : patch[b*(S-1) + s-1, :]) + pos[s, :]      (fp32 out)
__global__ void vision_embed_fwd_kernel(const __nv_bfloat16* __restrict__ patch, const float* __restrict__ cls,
                                        const float* __restrict__ pos, int S, int D, float* __restrict__ x0) {
  const long long row = blockIdx.x;  // b*S + s
  const int s = row % S;
  const long long b = row / S;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 pv = *reinterpret_cast<const float4*>(pos + static_cast<long long>(s) * D + c);
    float4 v;
    if (s == 0) v = *reinterpret_cast<const float4*>(cls + c);
    else v = unpack4_bf16(*reinterpret_cast<const uint2*>(patch + (b * (S - 1) + s - 1) * D + c));
    *reinterpret_cast<float4*>(x0 + row * D + c) = make_float4(v.x + pv.x, v.y + pv.y, v.z + pv.z, v.w + pv.w);
  }
}
// backward: d_patch (bf16) = d_x0[:,1:], d_cls += sum_b d_x0[:,0], d_pos[s] += sum_b d_x0[b,s]
// grid = S CTAs (one per token position), each loops over the batch.
__global__ void vision_embed_bwd_kernel(const float* __restrict__ dx0, int B, int S, int D,
                                        __nv_bfloat16* __restrict__ d_patch, float* __restrict__ d_cls,
                                        float* __restrict__ d_pos) {
  const int s = blockIdx.x;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int b = 0; b < B; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dx0 + (static_cast<long long>(b) * S + s) * D + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (s > 0)
        *reinterpret_cast<uint2*>(d_patch + (static_cast<long long>(b) * (S - 1) + s - 1) * D + c) =
            pack4_bf16(v.x, v.y, v.z, v.w);
    }
    float* dp = d_pos + static_cast<long long>(s) * D + c;
    dp[0] += acc.x; dp[1] += acc.y; dp[2] += acc.z; dp[3] += acc.w;
    if (s == 0) { d_cls[c] += acc.x; d_cls[c + 1] += acc.y; d_cls[c + 2] += acc.z; d_cls[c + 3] += acc.w; }
  }
}

// text: x[b,s,:] = tok_emb[ids[b,s], :] + pos[s, :];  also eot[b] = argmax_s ids[b,s] (first max)
__global__ void text_embed_fwd_kernel(const int* __restrict__ ids, const float* __restrict__ tok,
                                      const float* __restrict__ pos, int S, int D, float* __restrict__ x,
                                      int* __restrict__ eot_row) {
  const long long row = blockIdx.x;
  const int s = row % S;
  const long long b = row / S;
  const int id = ids[row];
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    const float4 t = *reinterpret_cast<const float4*>(tok + static_cast<long long>(id) * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + static_cast<long long>(s) * D + c);
    *reinterpret_cast<float4*>(x + row * D + c) = make_float4(t.x + p.x, t.y + p.y, t.z + p.z, t.w + p.w);
  }
  if (s == 0 && threadIdx.x == 0 && eot_row) {
    int best = ids[b * S], bi = 0;
    for (int j = 1; j < S; ++j) {
      const int v = ids[b * S + j];
      if (v > best) { best = v; bi = j; }
    }
    eot_row[b] = static_cast<int>(b * S + bi);
  }
}
// backward: d_tok[ids[b,s]] += dx[b,s] (atomic), d_pos[s] += sum_b dx[b,s]; grid = S CTAs
__global__ void text_embed_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ ids, int B,
                                      int S, int D, float* __restrict__ d_tok, float* __restrict__ d_pos) {
  const int s = blockIdx.x;
  for (int c = threadIdx.x * 4; c < D; c += blockDim.x * 4) {
    float4 acc = make_float4(0, 0, 0, 0);
    for (int b = 0; b < B; ++b) {
      const long long row = static_cast<long long>(b) * S + s;
      const float4 v = *reinterpret_cast<const float4*>(dx + row * D + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      float* tp = d_tok + static_cast<long long>(ids[row]) * D + c;
      atomicAdd(tp, v.x); atomicAdd(tp + 1, v.y); atomicAdd(tp + 2, v.z); atomicAdd(tp + 3, v.w);
    }
    float* dp = d_pos + static_cast<long long>(s) * D + c;
    dp[0] += acc.x; dp[1] += acc.y; dp[2] += acc.z; dp[3] += acc.w;
  }
}

// score fusion (reference clip_sf.py:61-63): out = t*mt + v*mv  (masks int64 0/1)
__global__ void fuse_fwd_kernel(const float* __restrict__ t, const float* __restrict__ v,
                                const long long* __restrict__ mt, const long long* __restrict__ mv,
                                long long n, int E, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / E;
  out[i] = t[i] * static_cast<float>(mt[r]) + v[i] * static_cast<float>(mv[r]);
}
// backward: dt = d*mt, dv = d*mv as bf16 (inputs of the projection dgrad/wgrad GEMMs)
__global__ void fuse_bwd_kernel(const float* __restrict__ d, const long long* __restrict__ mt,
                                const long long* __restrict__ mv, long long n, int E,
                                __nv_bfloat16* __restrict__ dt, __nv_bfloat16* __restrict__ dv) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / E;
  dt[i] = __float2bfloat16(d[i] * static_cast<float>(mt[r]));
  dv[i] = __float2bfloat16(d[i] * static_cast<float>(mv[r]));
}

// Score fusion over COMPACTED rows (masked towers are skipped: a text-only item never enters the
// image tower).  `rows[r]` is the batch row that compact row r came from:
//   scatter:  out[rows[r], :] += mask[rows[r]] * emb[r, :]            (clip_sf.py:61-63, one tower's term)
//   gather :  d[r, :] = bf16(mask[rows[r]] * d_out[rows[r], :])  for r < n, 0 for n <= r < n_pad
__global__ void fuse_scatter_kernel(const float* __restrict__ emb, const int* __restrict__ rows,
                                    const long long* __restrict__ mask, int n, int E, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n) * E) return;
  const int r = static_cast<int>(i / E), c = static_cast<int>(i % E);
  const long long dst = rows[r];
  out[dst * E + c] += static_cast<float>(mask[dst]) * emb[i];
}
__global__ void fuse_gather_bwd_kernel(const float* __restrict__ d_out, const int* __restrict__ rows,
                                       const long long* __restrict__ mask, int n, int n_pad, int E,
                                       __nv_bfloat16* __restrict__ d) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(n_pad) * E) return;
  const int r = static_cast<int>(i / E), c = static_cast<int>(i % E);
  float v = 0.f;
  if (r < n) {
    const long long src = rows[r];
    v = static_cast<float>(mask[src]) * d_out[src * E + c];
  }
  d[i] = __float2bfloat16(v);
}

// bias gradient: out[n] += sum_m dy[m, n]; CTA = 64 columns x a strip of rows
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ dy, long long M, int N, long long ld,
                              int rows_per_cta, float* __restrict__ out) {
  __shared__ float sh[8][64];
  const int col = blockIdx.x * 64 + (threadIdx.x & 63);
  const int sub = threadIdx.x >> 6;  // 0..3 (256 threads)
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(M, r0 + rows_per_cta);
  float acc = 0.f;
  if (col < N)
    for (long long r = r0 + sub; r < r1; r += 4) acc += __bfloat162float(dy[r * ld + col]);
  sh[sub][threadIdx.x & 63] = acc;
  __syncthreads();
  if (sub == 0 && col < N) atomicAdd(out + col, sh[0][threadIdx.x] + sh[1][threadIdx.x] + sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}

// 128-bit version (N, ld multiples of 8, 16 B-aligned base): a warp reads 512 contiguous bytes of a
// row (256 columns, 8 per lane), the CTA's 8 warps take rows r0+w, r0+w+8, ... of the strip four at
// a time, partial sums meet in shared memory, one atomicAdd per column per CTA.
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const __nv_bfloat16* __restrict__ dy, long long M, int N, long long ld, int rows_per_cta,
                  float* __restrict__ out) {
  __shared__ float sh[8][256 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 256 + lane * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(M, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (col0 < N) {
    const __nv_bfloat16* base = dy + col0;
    long long r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldcs(reinterpret_cast<const uint4*>(base + (r + 8 * k) * ld));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 lo = unpack4_bf16(make_uint2(u[k].x, u[k].y)), hi = unpack4_bf16(make_uint2(u[k].z, u[k].w));
        acc[0] += lo.x; acc[1] += lo.y; acc[2] += lo.z; acc[3] += lo.w;
        acc[4] += hi.x; acc[5] += hi.y; acc[6] += hi.z; acc[7] += hi.w;
      }
    }
    for (; r < r1; r += 8) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(base + r * ld));
      const float4 lo = unpack4_bf16(make_uint2(u.x, u.y)), hi = unpack4_bf16(make_uint2(u.z, u.w));
      acc[0] += lo.x; acc[1] += lo.y; acc[2] += lo.z; acc[3] += lo.w;
      acc[4] += hi.x; acc[5] += hi.y; acc[6] += hi.z; acc[7] += hi.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[w][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < N) {
    float t = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) t += sh[ww][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

// fp32 [rows, K] -> bf16 [rows, Kpad] (zero pad); and the reverse accumulate for its gradient
__global__ void pad_cast_kernel(const float* __restrict__ src, long long rows, int K, int Kpad,
                                __nv_bfloat16* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * Kpad) return;
  const long long r = i / Kpad;
  const int k = i % Kpad;
  dst[i] = __float2bfloat16(k < K ? src[r * K + k] : 0.f);
}
__global__ void unpad_accum_kernel(const float* __restrict__ src, long long rows, int K, int Kpad,
                                   float* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const long long r = i / K;
  const int k = i % K;
  dst[i] += src[r * Kpad + k];
}

// row gather/scatter of fp32 rows (EOT / CLS pooling and their backward)
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ src, long long n, __nv_bfloat16* __restrict__ dst) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = pack4_bf16(v.x, v.y, v.z, v.w);
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}

// ---------------------------------------------------------------------------------
// Fused AdamW over a flat parameter segment (torch.optim.AdamW semantics, reference
// clip_scorefusion/train.py:52-61): decoupled weight decay, bias correction, writes the
// bf16 shadow copy used by the GEMMs and zeroes the gradient.  16 B/param read, 14 written.
// ---------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, __nv_bfloat16* __restrict__ p16, long long n,
                             float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                             float grad_scale) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gp[j] * grad_scale;
      pp[j] *= (1.f - lr * wd);
      mp[j] = b1 * mp[j] + (1.f - b1) * gr;
      vp[j] = b2 * vp[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vp[j]) / sqrtf(bc2) + eps;
      pp[j] -= (lr / bc1) * mp[j] / denom;
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    *reinterpret_cast<float4*>(g + i) = make_float4(0, 0, 0, 0);
    if (p16) *reinterpret_cast<uint2*>(p16 + i) = pack4_bf16(pv.x, pv.y, pv.z, pv.w);
  } else {
    for (long long j = i; j < n; ++j) {
      const float gr = g[j] * grad_scale;
      float pj = p[j] * (1.f - lr * wd);
      m[j] = b1 * m[j] + (1.f - b1) * gr;
      v[j] = b2 * v[j] + (1.f - b2) * gr * gr;
      pj -= (lr / bc1) * m[j] / (sqrtf(v[j]) / sqrtf(bc2) + eps);
      p[j] = pj;
      g[j] = 0.f;
      if (p16) p16[j] = __float2bfloat16(pj);
    }
  }
}

// momentum encoder EMA (reference blip_ff.py:289-292): p_m = m*p_m + (1-m)*p over a flat segment,
// also refreshing the momentum encoder's bf16 shadow
__global__ void ema_kernel(float* __restrict__ pm, const float* __restrict__ p, long long n, float m,
                           __nv_bfloat16* __restrict__ pm16) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 3 < n) {
    float4 a = *reinterpret_cast<float4*>(pm + i);
    const float4 b = *reinterpret_cast<const float4*>(p + i);
    a.x = a.x * m + b.x * (1.f - m); a.y = a.y * m + b.y * (1.f - m);
    a.z = a.z * m + b.z * (1.f - m); a.w = a.w * m + b.w * (1.f - m);
    *reinterpret_cast<float4*>(pm + i) = a;
    if (pm16) *reinterpret_cast<uint2*>(pm16 + i) = pack4_bf16(a.x, a.y, a.z, a.w);
  } else {
    for (long long j = i; j < n; ++j) {
      const float a = pm[j] * m + p[j] * (1.f - m);
      pm[j] = a;
      if (pm16) pm16[j] = __float2bfloat16(a);
    }
  }
}

// dz = dy * (1 - y*y) (tanh backward; BERT pooler), bf16 out for the following GEMMs
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n,
                                __nv_bfloat16* __restrict__ dz) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dz[i] = __float2bfloat16(dy[i] * (1.f - y[i] * y[i]));
}

}  // namespace uniir

using namespace uniir;

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int uniir_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index,
                                   const float* gamma, const float* beta, float eps, int64_t rows,
                                   int32_t D, void* y, int64_t ldy, int32_t y_dtype, float* mean,
                                   float* rstd, void* stream) {
  UNIIR_CHECK_ARG(x && gamma && beta && y && rows >= 0 && D > 0 && D % 4 == 0 && D <= 4096,
                  "uniir_layernorm_fwd: bad args (D=%d must be a multiple of 4, <= 4096)", D);
  UNIIR_CHECK_ARG(y_dtype == UNIIR_BF16 || y_dtype == UNIIR_F32, "uniir_layernorm_fwd: y dtype");
  if (rows == 0) return UNIIR_OK;
  const int f32 = y_dtype == UNIIR_F32;
  if (D <= 1024) {
    // warp per row
    const int nv = (D + 127) / 128;
    const unsigned grid = static_cast<unsigned>((rows + kLnWarpsFwd - 1) / kLnWarpsFwd);
#define LN_FWD_W(V) layernorm_fwd_warp_kernel<V><<<grid, kLnWarpsFwd * 32, 0, ST(stream)>>>(x, ldx, row_index, gamma, beta, eps, D, rows, y, ldy, f32, mean, rstd)
    if (nv <= 1) LN_FWD_W(1); else if (nv <= 2) LN_FWD_W(2); else if (nv <= 4) LN_FWD_W(4);
    else if (nv <= 6) LN_FWD_W(6); else LN_FWD_W(8);
#undef LN_FWD_W
  } else {
    const int threads = 256;
    const int vecs = (D / 4 + threads - 1) / threads;
#define LN_FWD(V) layernorm_fwd_kernel<V><<<(unsigned)rows, threads, 0, ST(stream)>>>(x, ldx, row_index, gamma, beta, eps, D, y, ldy, f32, mean, rstd)
    if (vecs <= 2) LN_FWD(2); else LN_FWD(4);
#undef LN_FWD
  }
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_rmsnorm_fwd(const float* x, int64_t ldx, const float* gamma, float eps, int64_t rows, int32_t D,
                                 void* y, int64_t ldy, int32_t y_dtype, void* stream) {
  UNIIR_CHECK_ARG(x && gamma && y && D > 0 && D % 4 == 0 && D <= 1024 && ldx % 4 == 0 && ldy % 4 == 0,
                  "uniir_rmsnorm_fwd: bad args (D must be a multiple of 4, at most 1024)");
  UNIIR_CHECK_ARG(y_dtype == UNIIR_F32 || y_dtype == UNIIR_BF16, "uniir_rmsnorm_fwd: y must be fp32 or bf16");
  if (rows == 0) return UNIIR_OK;
  const int f32 = y_dtype == UNIIR_F32;
  const int nv = (D + 127) / 128;
  const unsigned grid = static_cast<unsigned>((rows + kLnWarpsFwd - 1) / kLnWarpsFwd);
#define RMS_FWD_W(V) rmsnorm_fwd_warp_kernel<V><<<grid, kLnWarpsFwd * 32, 0, ST(stream)>>>(x, ldx, gamma, eps, D, rows, y, ldy, f32)
  if (nv <= 1) RMS_FWD_W(1); else if (nv <= 2) RMS_FWD_W(2); else if (nv <= 4) RMS_FWD_W(4);
  else if (nv <= 6) RMS_FWD_W(6); else RMS_FWD_W(8);
#undef RMS_FWD_W
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_rmsnorm_bwd(const float* dy, const float* x, const float* gamma, float eps, int64_t rows, int32_t D,
                                 float* dx, int32_t dx_accumulate, float* dgamma, void* stream) {
  UNIIR_CHECK_ARG(dy && x && gamma && dx && dgamma && D > 0 && D % 4 == 0 && D <= 1024,
                  "uniir_rmsnorm_bwd: bad args (D must be a multiple of 4, at most 1024)");
  if (rows == 0) return UNIIR_OK;
  long long ctas = 2LL * num_sms();
  long long rpc = (rows + ctas - 1) / ctas;
  rpc = (rpc + kLnWarpsBwd - 1) / kLnWarpsBwd * kLnWarpsBwd;
  const unsigned grid = static_cast<unsigned>((rows + rpc - 1) / rpc);
  const int nv = (D + 127) / 128;
#define RMS_BWD_W(V) rmsnorm_bwd_warp_kernel<V><<<grid, kLnWarpsBwd * 32, 0, ST(stream)>>>(dy, x, gamma, eps, D, rows, (int)rpc, dx, dx_accumulate, dgamma)
  if (nv <= 1) RMS_BWD_W(1); else if (nv <= 2) RMS_BWD_W(2); else if (nv <= 4) RMS_BWD_W(4);
  else if (nv <= 6) RMS_BWD_W(6); else RMS_BWD_W(8);
#undef RMS_BWD_W
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_token_mean_bwd(const float* d_out, int32_t B, int32_t S, int32_t D, float* dx, void* stream) {
  UNIIR_CHECK_ARG(d_out && dx && B > 0 && S > 0 && D > 0 && D % 4 == 0 && B <= 65535, "uniir_token_mean_bwd: bad args");
  const int threads = 128;
  token_mean_bwd_kernel<<<dim3((D / 4 + threads - 1) / threads, B), threads, 0, ST(stream)>>>(d_out, S, D, dx);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_token_mean(const float* x, int32_t B, int32_t S, int32_t D, float* out, void* stream) {
  UNIIR_CHECK_ARG(x && out && B > 0 && S > 0 && D > 0 && D % 4 == 0 && B <= 65535, "uniir_token_mean: bad args");
  const int threads = 128;
  token_mean_kernel<<<dim3((D / 4 + threads - 1) / threads, B), threads, 0, ST(stream)>>>(x, S, D, out);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_dtype, const float* x,
                                   int64_t ldx, const int32_t* row_index, const float* mean,
                                   const float* rstd, const float* gamma, int64_t rows, int32_t D,
                                   float* dx, int64_t lddx, int32_t dx_accumulate, float* dgamma,
                                   float* dbeta, void* dx_bf16, float* dcolsum, void* stream) {
  UNIIR_CHECK_ARG(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && D > 0 && D % 4 == 0 && D <= 4096,
                  "uniir_layernorm_bwd: bad args");
  if (rows == 0) return UNIIR_OK;
  const int f32 = dy_dtype == UNIIR_F32;
  __nv_bfloat16* dx16 = static_cast<__nv_bfloat16*>(dx_bf16);
  if (D <= 1024) {
    // warp per row; 2 CTAs of 8 warps per SM-worth of strips keeps the atomics at ~300 per column
    long long ctas = 2LL * num_sms();
    long long rpc = (rows + ctas - 1) / ctas;
    rpc = (rpc + kLnWarpsBwd - 1) / kLnWarpsBwd * kLnWarpsBwd;
    const unsigned grid = static_cast<unsigned>((rows + rpc - 1) / rpc);
    const int nv = (D + 127) / 128;
#define LN_BWD_L(V, C, F)                                                                                    \
  layernorm_bwd_warp_kernel<V, C, F><<<grid, kLnWarpsBwd * 32, 0, ST(stream)>>>(                             \
      dy, lddy, x, ldx, row_index, mean, rstd, gamma, D, rows, (int)rpc, dx, lddx, dx_accumulate, dgamma,    \
      dbeta, dx16, dcolsum)
#define LN_BWD_W(V)                                                                                          \
  do {                                                                                                       \
    if (dcolsum) { if (f32) LN_BWD_L(V, true, true); else LN_BWD_L(V, true, false); }                        \
    else         { if (f32) LN_BWD_L(V, false, true); else LN_BWD_L(V, false, false); }                      \
  } while (0)
    if (nv <= 1) LN_BWD_W(1); else if (nv <= 2) LN_BWD_W(2); else if (nv <= 4) LN_BWD_W(4);
    else if (nv <= 6) LN_BWD_W(6); else LN_BWD_W(8);
#undef LN_BWD_W
#undef LN_BWD_L
    count_launch();
  } else {
    const int threads = 256;
    const int vecs = (D / 4 + threads - 1) / threads;
    // ~4 CTAs per SM keeps the dgamma/dbeta atomics to a few thousand per column
    long long ctas = 4LL * num_sms();
    int rpc = static_cast<int>((rows + ctas - 1) / ctas);
    if (rpc < 1) rpc = 1;
    const unsigned grid = static_cast<unsigned>((rows + rpc - 1) / rpc);
#define LN_BWD(V) layernorm_bwd_kernel<V><<<grid, threads, 0, ST(stream)>>>(dy, lddy, f32, x, ldx, row_index, mean, rstd, gamma, D, rows, rpc, dx, lddx, dx_accumulate, dgamma, dbeta, dx16)
    if (vecs <= 2) LN_BWD(2); else LN_BWD(4);
#undef LN_BWD
    count_launch();
    if (dcolsum) {
      // wide rows: the column sums of the updated dx take the separate pass
      UNIIR_CHECK_ARG(row_index == nullptr && dx16, "uniir_layernorm_bwd: dcolsum with D > 1024 needs dx_bf16 and no row_index");
      const int r = uniir_colsum_accum(dx16, rows, D, lddx, dcolsum, stream);
      if (r != UNIIR_OK) return r;
    }
  }
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_patch_im2col(const float* img, int32_t B, int32_t H, int32_t W, int32_t P,
                                  int32_t Kpad, void* out, void* stream) {
  UNIIR_CHECK_ARG(img && out && B > 0 && H % P == 0 && W % P == 0 && H == W && Kpad >= 3 * P * P && Kpad % 8 == 0,
                  "uniir_patch_im2col: bad args");
  const int g = H / P;
  im2col_kernel<<<(unsigned)(B * g * g), 256, 0, ST(stream)>>>(img, B, H, W, P, g, Kpad,
                                                                static_cast<__nv_bfloat16*>(out));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_vision_embed_fwd(const void* patch, const float* cls, const float* pos, int32_t B,
                                      int32_t S, int32_t D, float* x0, void* stream) {
  UNIIR_CHECK_ARG(patch && cls && pos && x0 && B > 0 && S > 1 && D % 4 == 0, "uniir_vision_embed_fwd: bad args");
  vision_embed_fwd_kernel<<<(unsigned)(B * S), 256, 0, ST(stream)>>>(
      static_cast<const __nv_bfloat16*>(patch), cls, pos, S, D, x0);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_vision_embed_bwd(const float* dx0, int32_t B, int32_t S, int32_t D, void* d_patch,
                                      float* d_cls, float* d_pos, void* stream) {
  UNIIR_CHECK_ARG(dx0 && d_patch && d_cls && d_pos && D % 4 == 0, "uniir_vision_embed_bwd: bad args");
  vision_embed_bwd_kernel<<<(unsigned)S, 256, 0, ST(stream)>>>(dx0, B, S, D, static_cast<__nv_bfloat16*>(d_patch), d_cls, d_pos);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_text_embed_fwd(const int32_t* ids, const float* tok, const float* pos, int32_t B,
                                    int32_t S, int32_t D, float* x, int32_t* eot_row, void* stream) {
  UNIIR_CHECK_ARG(ids && tok && pos && x && B > 0 && S > 0 && D % 4 == 0, "uniir_text_embed_fwd: bad args");
  text_embed_fwd_kernel<<<(unsigned)(B * S), 192, 0, ST(stream)>>>(ids, tok, pos, S, D, x, eot_row);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_text_embed_bwd(const float* dx, const int32_t* ids, int32_t B, int32_t S, int32_t D,
                                    float* d_tok, float* d_pos, void* stream) {
  UNIIR_CHECK_ARG(dx && ids && d_tok && d_pos && D % 4 == 0, "uniir_text_embed_bwd: bad args");
  text_embed_bwd_kernel<<<(unsigned)S, 192, 0, ST(stream)>>>(dx, ids, B, S, D, d_tok, d_pos);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_score_fuse_fwd(const float* t, const float* v, const int64_t* mt, const int64_t* mv,
                                    int64_t B, int32_t E, float* out, void* stream) {
  UNIIR_CHECK_ARG(t && v && mt && mv && out, "uniir_score_fuse_fwd: null");
  const long long n = B * E;
  if (n == 0) return UNIIR_OK;
  fuse_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(
      t, v, reinterpret_cast<const long long*>(mt), reinterpret_cast<const long long*>(mv), n, E, out);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_score_fuse_bwd(const float* d, const int64_t* mt, const int64_t* mv, int64_t B,
                                    int32_t E, void* dt, void* dv, void* stream) {
  UNIIR_CHECK_ARG(d && mt && mv && dt && dv, "uniir_score_fuse_bwd: null");
  const long long n = B * E;
  if (n == 0) return UNIIR_OK;
  fuse_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(
      d, reinterpret_cast<const long long*>(mt), reinterpret_cast<const long long*>(mv), n, E,
      static_cast<__nv_bfloat16*>(dt), static_cast<__nv_bfloat16*>(dv));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_score_fuse_scatter(const float* emb, const int32_t* rows, const int64_t* mask, int32_t n,
                                        int32_t E, float* out, void* stream) {
  UNIIR_CHECK_ARG(emb && rows && mask && out && n >= 0 && E > 0, "uniir_score_fuse_scatter: bad args");
  if (n == 0) return UNIIR_OK;
  const long long t = static_cast<long long>(n) * E;
  fuse_scatter_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(emb, rows, reinterpret_cast<const long long*>(mask), n, E, out);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_score_fuse_gather_bwd(const float* d_out, const int32_t* rows, const int64_t* mask, int32_t n,
                                           int32_t n_pad, int32_t E, void* d_bf16, void* stream) {
  UNIIR_CHECK_ARG(d_out && rows && mask && d_bf16 && n >= 0 && n_pad >= n && E > 0, "uniir_score_fuse_gather_bwd: bad args");
  if (n_pad == 0) return UNIIR_OK;
  const long long t = static_cast<long long>(n_pad) * E;
  fuse_gather_bwd_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(
      d_out, rows, reinterpret_cast<const long long*>(mask), n, n_pad, E, static_cast<__nv_bfloat16*>(d_bf16));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_colsum_accum(const void* dy, int64_t M, int32_t N, int64_t ld, float* out, void* stream) {
  UNIIR_CHECK_ARG(dy && out && M >= 0 && N > 0, "uniir_colsum_accum: bad args");
  if (M == 0) return UNIIR_OK;
  const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(dy);
  if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const int nbx = (N + 255) / 256;
    long long strips = (8LL * num_sms() + nbx - 1) / nbx;
    long long rpc = (M + strips - 1) / strips;
    rpc = (rpc + 31) / 32 * 32;
    dim3 grid(nbx, static_cast<unsigned>((M + rpc - 1) / rpc));
    colsum_vec_kernel<<<grid, 256, 0, ST(stream)>>>(p, M, N, ld, static_cast<int>(rpc), out);
  } else {
    const int rpc = 512;
    dim3 grid((N + 63) / 64, static_cast<unsigned>((M + rpc - 1) / rpc));
    colsum_kernel<<<grid, 256, 0, ST(stream)>>>(p, M, N, ld, rpc, out);
  }
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_pad_cast_bf16(const float* src, int64_t rows, int32_t K, int32_t Kpad, void* dst, void* stream) {
  UNIIR_CHECK_ARG(src && dst && Kpad >= K, "uniir_pad_cast_bf16: bad args");
  const long long n = rows * Kpad;
  if (n == 0) return UNIIR_OK;
  pad_cast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(src, rows, K, Kpad, static_cast<__nv_bfloat16*>(dst));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_unpad_accum(const float* src, int64_t rows, int32_t K, int32_t Kpad, float* dst, void* stream) {
  UNIIR_CHECK_ARG(src && dst && Kpad >= K, "uniir_unpad_accum: bad args");
  const long long n = rows * K;
  if (n == 0) return UNIIR_OK;
  unpad_accum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(src, rows, K, Kpad, dst);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_cast_bf16(const float* src, int64_t n, void* dst, void* stream) {
  UNIIR_CHECK_ARG(src && dst && n >= 0, "uniir_cast_bf16: bad args");
  if (n == 0) return UNIIR_OK;
  const long long t = (n + 3) / 4;
  cast_f32_to_bf16_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(src, n, static_cast<__nv_bfloat16*>(dst));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                float grad_scale, void* stream) {
  UNIIR_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "uniir_adamw_step: bad args");
  if (n == 0) return UNIIR_OK;
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const long long t = (n + 3) / 4;
  adamw_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(
      p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_ema_update(float* p_m, const float* p, int64_t n, float momentum, void* p_m_bf16, void* stream) {
  UNIIR_CHECK_ARG(p_m && p && n >= 0, "uniir_ema_update: bad args");
  if (n == 0) return UNIIR_OK;
  const long long t = (n + 3) / 4;
  ema_kernel<<<(unsigned)((t + 255) / 256), 256, 0, ST(stream)>>>(p_m, p, n, momentum, static_cast<__nv_bfloat16*>(p_m_bf16));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
extern "C" int uniir_tanh_bwd(const float* dy, const float* y, int64_t n, void* dz_bf16, void* stream) {
  UNIIR_CHECK_ARG(dy && y && dz_bf16 && n >= 0, "uniir_tanh_bwd: bad args");
  if (n == 0) return UNIIR_OK;
  tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(dy, y, n, static_cast<__nv_bfloat16*>(dz_bf16));
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
