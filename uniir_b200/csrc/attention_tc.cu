// tcgen05 attention forward for the short sequences of the UniIR towers (S = 257 / 197 / 100 / 77,
// head_dim 64, bf16).  Replaces the forward of nn.MultiheadAttention's softmax(QK^T/8 [+mask])V
// (upstream CLIP resblocks, BLIP ViT / BERT self- and cross-attention) when Skv <= 304.
//
// One persistent CTA per SM walks (batch, head) items.  Because a whole score row (<= 384 keys) fits in
// TMEM there is no online-softmax rescaling: per 128-query tile
//     S[128, Skv]  = Q K^T        tcgen05.mma, fp32 in TMEM columns [0, skv_pad)
//     m = rowmax(S), P = exp2(S*c - m*c), l = rowsum(P)      two passes over TMEM, 256 threads:
//                                                            row = TMEM lane, two threads per row
//     O[128, 64]  += P_c V_c      per 64-key chunk c; P_c staged bf16 in shared memory (K-major,
//                                 128B swizzle, double buffered) so the MMAs overlap the exponentials
//     out = O / l, lse = m*c' + log(l)
// Warp roles (320 threads): warp 0 = TMA producer (K/V of the NEXT head double-buffered against the
// current one, Q tiles through a 2-deep ring), warp 1 = MMA issuer + TMEM owner, warps 2-9 = softmax /
// epilogue.  Query rows that do not fill a 128-row MMA tile and are few (257 = 2 x 128 + 1: the class
// token) are computed on CUDA cores from the K/V already in shared memory instead of a third, 99 %
// empty tile.  Masks: keys >= kv_len[b] (BERT padding) and key > query (CLIP text tower, causal).
#include <cuda_bf16.h>

#include "gemm_sm100.cuh"
#include "host_util.cuh"

namespace uniir {

constexpr int kTcThreads = 320;
constexpr int kTcSoftmaxThreads = 256;
constexpr int kTcMaxKv = 304;        // K and V of two heads (double buffer) must fit shared memory: 4 x 38 KB
constexpr int kTcOCol = 320;         // TMEM: scores in columns [0, skv_pad <= 304), O in [320, 384)
constexpr int kTcTailMax = 4;        // leftover query rows computed on CUDA cores
constexpr float kTcLog2e = 1.4426950408889634f;

struct AttnTcParams {
  CUtensorMap tmQ, tmK, tmV;         // box 64 x 128 rows
  CUtensorMap tmK16, tmV16;          // box 64 x 16 rows (tail of the key range)
  const __nv_bfloat16* q;            // tail rows are read directly
  long long ldq, ldo;
  __nv_bfloat16* out;
  float* lse;                        // [B, H, Sq] or null
  const int* kv_len;                 // [B] or null
  int B, H, Sq, Skv, causal;
  int n_qt;                          // 128-row MMA tiles per head
  int tail_rows;                     // Sq - 128 * n_qt when that is <= kTcTailMax (else 0 and n_qt is ceil)
  int skv_pad;                       // Skv rounded up to 16
  int kv_bytes;                      // skv_pad * 128 rounded up to 1024
  float scale;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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

struct TcSmem {
  uint8_t *k0, *v0, *q0, *p0;   // K[2], V[2] (kv_bytes each), Q[2], P[2] (16 KB each)
  int kv_bytes;
  float *xmax, *xsum;      // [2][128]
  float *tail_s;           // [kTcMaxKv]
  float *tail_q;           // [64]
  float *tail_o;           // [4][64]
  float *red;              // [16]
  uint64_t *kv_full, *kv_empty, *q_full, *q_empty, *p_full, *p_empty;   // [2] each
  uint64_t *s_full, *s_free, *o_full, *o_free;
  uint32_t* tmem_ptr;
  __device__ __forceinline__ uint8_t* k(int i) const { return k0 + i * kv_bytes; }
  __device__ __forceinline__ uint8_t* v(int i) const { return v0 + i * kv_bytes; }
  __device__ __forceinline__ uint8_t* q(int i) const { return q0 + i * 16384; }
  __device__ __forceinline__ uint8_t* p(int i) const { return p0 + i * 16384; }
};

__device__ __forceinline__ TcSmem carve(uint8_t* base, int kv_bytes) {
  TcSmem s;
  uint8_t* c = base;
  s.kv_bytes = kv_bytes;
  s.k0 = c; c += 2 * kv_bytes;
  s.v0 = c; c += 2 * kv_bytes;
  s.q0 = c; c += 2 * 16384;
  s.p0 = c; c += 2 * 16384;
  s.xmax = reinterpret_cast<float*>(c); c += 2 * 128 * 4;
  s.xsum = reinterpret_cast<float*>(c); c += 2 * 128 * 4;
  s.tail_s = reinterpret_cast<float*>(c); c += kTcMaxKv * 4;
  s.tail_q = reinterpret_cast<float*>(c); c += 64 * 4;
  s.tail_o = reinterpret_cast<float*>(c); c += 4 * 64 * 4;
  s.red = reinterpret_cast<float*>(c); c += 16 * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(c);
  s.kv_full = b; s.kv_empty = b + 2; s.q_full = b + 4; s.q_empty = b + 6; s.p_full = b + 8; s.p_empty = b + 10;
  s.s_full = b + 12; s.s_free = b + 13; s.o_full = b + 14; s.o_free = b + 15;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b + 16);
  return s;
}
static int attn_tc_smem_bytes(int kv_bytes) {
  return 1024 /*alignment slack*/ + 4 * kv_bytes + 4 * 16384 + 2 * 2 * 128 * 4 + kTcMaxKv * 4 + 64 * 4 + 4 * 64 * 4 + 16 * 4 +
         16 * 8 + 16;
}

// reduce over the 256 softmax threads (warps 2..9) through shared memory; IS_MAX ? max : sum
template <bool IS_MAX>
__device__ __forceinline__ float softmax_group_reduce(float v, float* red, int sw /*0..7*/, int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = IS_MAX ? fmaxf(v, t) : v + t;
  }
  named_bar_sync(1, kTcSoftmaxThreads);     // previous users of `red` are done
  if (lane == 0) red[sw] = v;
  named_bar_sync(1, kTcSoftmaxThreads);
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = IS_MAX ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __launch_bounds__(kTcThreads, 1) attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem sm = carve(align_smem_1024(smem_raw), p.kv_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_heads = p.B * p.H;
  const int n_chunks = (p.skv_pad + 63) >> 6;          // 64-key P chunks per tile

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmQ); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmK16); tma_prefetch_desc(&p.tmV16);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.kv_full[i], 1);
      mbar_init(&sm.kv_empty[i], 1 + kTcSoftmaxThreads);   // MMA commit + every softmax thread (tail reads K/V)
      mbar_init(&sm.q_full[i], 1);
      mbar_init(&sm.q_empty[i], 1);
      mbar_init(&sm.p_full[i], kTcSoftmaxThreads);
      mbar_init(&sm.p_empty[i], 1);
    }
    mbar_init(sm.s_full, 1);
    mbar_init(sm.s_free, kTcSoftmaxThreads);
    mbar_init(sm.o_full, 1);
    mbar_init(sm.o_free, kTcSoftmaxThreads);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(sm.tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      int kv_it = 0, q_it = 0;
      for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++kv_it) {
        const int b = head / p.H, h = head % p.H;
        const int kb = kv_it & 1;
        if (kv_it >= 2) mbar_wait(&sm.kv_empty[kb], ((kv_it >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&sm.kv_full[kb], 2u * p.skv_pad * 128u);
        int r0 = 0;
        for (; r0 + 128 <= p.skv_pad; r0 += 128) {
          tma_load_2d(sm.k(kb) + r0 * 128, &p.tmK, &sm.kv_full[kb], h * 64, b * p.Skv + r0);
          tma_load_2d(sm.v(kb) + r0 * 128, &p.tmV, &sm.kv_full[kb], h * 64, b * p.Skv + r0);
        }
        for (; r0 < p.skv_pad; r0 += 16) {
          tma_load_2d(sm.k(kb) + r0 * 128, &p.tmK16, &sm.kv_full[kb], h * 64, b * p.Skv + r0);
          tma_load_2d(sm.v(kb) + r0 * 128, &p.tmV16, &sm.kv_full[kb], h * 64, b * p.Skv + r0);
        }
        for (int t = 0; t < p.n_qt; ++t, ++q_it) {
          const int qb = q_it & 1;
          if (q_it >= 2) mbar_wait(&sm.q_empty[qb], ((q_it >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&sm.q_full[qb], 16384u);
          tma_load_2d(sm.q(qb), &p.tmQ, &sm.q_full[qb], h * 64, b * p.Sq + t * 128);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      const uint32_t idesc_pv = make_idesc_f16(true, false, true, 128, 64);
      int kv_it = 0, q_it = 0, tile_it = 0, chunk_it = 0;
      for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++kv_it) {
        const int kb = kv_it & 1;
        mbar_wait(&sm.kv_full[kb], (kv_it >> 1) & 1);
        const uint32_t sK = smem_u32(sm.k(kb)), sV = smem_u32(sm.v(kb));
        for (int t = 0; t < p.n_qt; ++t, ++q_it, ++tile_it) {
          const int qb = q_it & 1;
          mbar_wait(&sm.q_full[qb], (q_it >> 1) & 1);
          if (tile_it >= 1) mbar_wait(sm.s_free, (tile_it - 1) & 1);   // previous tile's scores are consumed
          tc_fence_after();
          const uint32_t sQ = smem_u32(sm.q(qb));
          for (int n0 = 0; n0 < p.skv_pad; n0 += 256) {
            const int nn = min(256, p.skv_pad - n0);
            const uint32_t idesc = make_idesc_f16(true, false, false, 128, static_cast<uint32_t>(nn));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + n0, make_smem_desc(sQ + k * 32, 16, 1024),
                       make_smem_desc(sK + n0 * 128 + k * 32, 16, 1024), idesc, k > 0 ? 1u : 0u);
          }
          umma_commit(&sm.q_empty[qb]);
          umma_commit(sm.s_full);
          if (tile_it >= 1) mbar_wait(sm.o_free, (tile_it - 1) & 1);   // previous tile's O has been read
          for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
            const int pb = chunk_it & 1;
            mbar_wait(&sm.p_full[pb], (chunk_it >> 1) & 1);
            tc_fence_after();
            const uint32_t sP = smem_u32(sm.p(pb));
            const int ks = min(4, (p.skv_pad - 64 * c) >> 4);
            for (int k = 0; k < ks; ++k)
              umma_f16(tmem_base + kTcOCol, make_smem_desc(sP + k * 32, 16, 1024),
                       make_smem_desc(sV + (64 * c + 16 * k) * 128, 64 * 128, 1024), idesc_pv,
                       (c > 0 || k > 0) ? 1u : 0u);
            umma_commit(&sm.p_empty[pb]);
          }
          umma_commit(sm.o_full);
        }
        umma_commit(&sm.kv_empty[kb]);     // every MMA that reads this head's K/V has retired
      }
    }
  } else {
    // =============================== softmax / epilogue ===============================
    const int sw = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lanes [32 q, 32 q + 32)
    const int half = sw >> 2;                // which 32 columns of each 64-key chunk / which half of O
    const int rit = quarter * 32 + lane;     // row in tile
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = p.scale * kTcLog2e;
    const int n32 = (p.skv_pad + 31) >> 5;
    const int stid = threadIdx.x - 64;       // 0..255
    int kv_it = 0, tile_it = 0, chunk_it = 0;
    for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++kv_it) {
      const int b = head / p.H, h = head % p.H;
      const int kb = kv_it & 1;
      const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
      __nv_bfloat16* ob = p.out + static_cast<long long>(b) * p.Sq * p.ldo + h * 64;
      float* lp = p.lse ? p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq : nullptr;
      for (int t = 0; t < p.n_qt; ++t, ++tile_it) {
        const int row = t * 128 + rit;
        const bool warp_valid = t * 128 + quarter * 32 < p.Sq;     // warp-uniform
        const int klim = p.causal ? min(kvl, row + 1) : kvl;       // keys [0, klim) are visible to this row
        mbar_wait(sm.s_full, tile_it & 1);
        tc_fence_after();
        // ---- pass 1: row maximum (this thread: every other 32-column group)
        float mx = -INFINITY;
        if (warp_valid) {
          for (int j = half; j < n32; j += 2) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tlane + j * 32, v);
            tmem_ld_wait();
            const int k0 = j * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = (k0 + i < klim) ? fmaxf(mx, __uint_as_float(v[i])) : mx;
          }
        }
        sm.xmax[half * 128 + rit] = mx;
        named_bar_sync(1, kTcSoftmaxThreads);
        mx = fmaxf(sm.xmax[rit], sm.xmax[128 + rit]);
        const float base = (mx == -INFINITY) ? 0.f : mx * sc;
        // ---- pass 2: P = exp2(S*sc - base) chunk by chunk into shared memory, row sums
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
          const int pb = chunk_it & 1;
          if (chunk_it >= 2) mbar_wait(&sm.p_empty[pb], ((chunk_it >> 1) - 1) & 1);
          const int k0 = c * 64 + half * 32;
          if (warp_valid && k0 < p.skv_pad) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tlane + k0, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float e0 = (k0 + i < klim) ? ex2_approx(fmaf(__uint_as_float(v[i]), sc, -base)) : 0.f;
              const float e1 = (k0 + i + 1 < klim) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -base)) : 0.f;
              sum += e0 + e1;
              pk[i >> 1] = pack2_bf16(e0, e1);
            }
            uint8_t* prow = sm.p(pb);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              *reinterpret_cast<uint4*>(prow + sw128(rit, half * 4 + g)) =
                  make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
          fence_proxy_async();             // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(&sm.p_full[pb]);
        }
        tc_fence_before();
        mbar_arrive(sm.s_free);            // score columns may be overwritten by the next tile's QK^T
        sm.xsum[half * 128 + rit] = sum;
        named_bar_sync(1, kTcSoftmaxThreads);
        const float total = sm.xsum[rit] + sm.xsum[128 + rit];
        const float inv = total > 0.f ? 1.f / total : 0.f;
        // ---- epilogue: O / l
        mbar_wait(sm.o_full, tile_it & 1);
        tc_fence_after();
        if (warp_valid) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tlane + kTcOCol + half * 32, v);
          tmem_ld_wait();
          if (row < p.Sq) {
            uint4* dst = reinterpret_cast<uint4*>(ob + static_cast<long long>(row) * p.ldo + half * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w[e] = pack2_bf16(__uint_as_float(v[8 * g + 2 * e]) * inv, __uint_as_float(v[8 * g + 2 * e + 1]) * inv);
              dst[g] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (lp && half == 0) lp[row] = (mx * sc + log2f(total)) / kTcLog2e;
          }
        }
        tc_fence_before();
        mbar_arrive(sm.o_free);
      }
      // ---- leftover query rows on CUDA cores (K/V of this head are in shared memory) ----
      if (p.tail_rows > 0) {
        mbar_wait(&sm.kv_full[kb], (kv_it >> 1) & 1);
        const uint8_t* sK = sm.k(kb);
        const uint8_t* sV = sm.v(kb);
        for (int tr = 0; tr < p.tail_rows; ++tr) {
          const int row = p.n_qt * 128 + tr;
          const int klim = p.causal ? min(kvl, row + 1) : kvl;
          named_bar_sync(1, kTcSoftmaxThreads);       // tail_q / tail_s / tail_o of the previous row are free
          if (stid < 32) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(
                p.q + (static_cast<long long>(b) * p.Sq + row) * p.ldq + h * 64 + stid * 2);
            sm.tail_q[2 * stid] = bf16_lo(u);
            sm.tail_q[2 * stid + 1] = bf16_hi(u);
          }
          named_bar_sync(1, kTcSoftmaxThreads);
          float lmax = -INFINITY;
          for (int j = stid; j < p.skv_pad; j += kTcSoftmaxThreads) {
            float acc = 0.f;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              const uint4 kk = *reinterpret_cast<const uint4*>(sK + sw128(j, ch));
              const float* qv = sm.tail_q + ch * 8;
              acc += bf16_lo(kk.x) * qv[0] + bf16_hi(kk.x) * qv[1] + bf16_lo(kk.y) * qv[2] + bf16_hi(kk.y) * qv[3] +
                     bf16_lo(kk.z) * qv[4] + bf16_hi(kk.z) * qv[5] + bf16_lo(kk.w) * qv[6] + bf16_hi(kk.w) * qv[7];
            }
            const float s = (j < klim) ? acc * sc : -INFINITY;
            sm.tail_s[j] = s;
            lmax = fmaxf(lmax, s);
          }
          const float mx = softmax_group_reduce<true>(lmax, sm.red, sw, lane);
          const float base = (mx == -INFINITY) ? 0.f : mx;
          float lsum = 0.f;
          for (int j = stid; j < p.skv_pad; j += kTcSoftmaxThreads) {
            const float e = (j < klim) ? ex2_approx(sm.tail_s[j] - base) : 0.f;
            sm.tail_s[j] = e;
            lsum += e;
          }
          const float total = softmax_group_reduce<false>(lsum, sm.red, sw, lane);   // its barriers publish tail_s
          const int d = stid & 63, part = stid >> 6;
          float acc = 0.f;
          for (int j = part; j < klim; j += 4) {
            const uint16_t vv = *reinterpret_cast<const uint16_t*>(sV + sw128(j, d >> 3) + (d & 7) * 2);
            acc = fmaf(sm.tail_s[j], __uint_as_float(static_cast<uint32_t>(vv) << 16), acc);
          }
          sm.tail_o[part * 64 + d] = acc;
          named_bar_sync(1, kTcSoftmaxThreads);
          if (stid < 64) {
            const float inv = total > 0.f ? 1.f / total : 0.f;
            const float o = (sm.tail_o[stid] + sm.tail_o[64 + stid] + sm.tail_o[128 + stid] + sm.tail_o[192 + stid]) * inv;
            ob[static_cast<long long>(row) * p.ldo + stid] = __float2bfloat16(o);
            if (lp && stid == 0) lp[row] = (mx + log2f(total)) / kTcLog2e;
          }
        }
      }
      mbar_arrive(&sm.kv_empty[kb]);       // this thread no longer reads this head's K/V
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace uniir

using namespace uniir;

// mode: 0 = mma.sync head-resident kernels (attention.cu), 1 = tcgen05 forward where it applies.
static int g_attn_tc_mode = -1;
static int attn_tc_mode() {
  if (g_attn_tc_mode < 0) {
    const char* e = getenv("UNIIR_ATTN_TC");
    g_attn_tc_mode = e ? (atoi(e) != 0) : 0;
  }
  return g_attn_tc_mode;
}
extern "C" int uniir_attn_set_mode(int32_t mode) {
  const int prev = attn_tc_mode();
  g_attn_tc_mode = mode != 0;
  return prev;
}

namespace uniir {
// Called by attention.cu's forward entry points.  Returns UNIIR_OK after launching, a negative error,
// or +1 when the tcgen05 path does not apply (caller falls through to the mma.sync kernel).
int attn_fwd_tc_try(const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                    long long ldkv, int B, int Sq, int Skv, int H, int causal, const int* kv_len,
                    __nv_bfloat16* out, long long ldo, float* lse, float scale, void* stream) {
  if (!attn_tc_mode()) return 1;
  if (Skv > kTcMaxKv || ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return 1;
  AttnTcParams p{};
  p.q = q; p.ldq = ldq; p.ldo = ldo; p.out = out; p.lse = lse; p.kv_len = kv_len;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.causal = causal; p.scale = scale;
  const int full = Sq / 128, rem = Sq % 128;
  if (full >= 1 && rem > 0 && rem <= kTcTailMax) { p.n_qt = full; p.tail_rows = rem; }
  else { p.n_qt = (Sq + 127) / 128; p.tail_rows = 0; }
  p.skv_pad = (Skv + 15) / 16 * 16;
  p.kv_bytes = (p.skv_pad * 128 + 1023) / 1024 * 1024;
  int rc;
  const uint64_t cols = static_cast<uint64_t>(H) * 64;
  if ((rc = make_tmap_2d(&p.tmQ, q, UNIIR_BF16, cols, static_cast<uint64_t>(B) * Sq, ldq, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&p.tmK, k, UNIIR_BF16, cols, static_cast<uint64_t>(B) * Skv, ldkv, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&p.tmV, v, UNIIR_BF16, cols, static_cast<uint64_t>(B) * Skv, ldkv, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&p.tmK16, k, UNIIR_BF16, cols, static_cast<uint64_t>(B) * Skv, ldkv, 64, 16))) return rc;
  if ((rc = make_tmap_2d(&p.tmV16, v, UNIIR_BF16, cols, static_cast<uint64_t>(B) * Skv, ldkv, 64, 16))) return rc;
  const int smem = attn_tc_smem_bytes(p.kv_bytes);
  static int smem_set = 0;
  if (smem_set < smem) {
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    smem_set = 227 * 1024;
  }
  const int heads = B * H;
  const int grid = heads < num_sms() ? heads : num_sms();
  attn_fwd_tc_kernel<<<grid, kTcThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
}  // namespace uniir
