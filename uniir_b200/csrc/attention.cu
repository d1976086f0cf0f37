// Multi-head self-attention core for short sequences (S = 257 / 197 / 77 causal / 100),
// head_dim 64, bf16 in / fp32 softmax + accumulate.  Replaces the softmax(QK^T/8 [+mask])V
// inside nn.MultiheadAttention (upstream CLIP resblocks; causal mask = upstream
// build_attention_mask) and its autograd backward.
//
// Flash-style: a CTA owns a 64-row query (or key) block, streams the other operand through
// shared memory in 64-row blocks, keeps scores in registers and never writes the S x S matrix.
// Tensor-core path is mma.sync m16n8k16 (HMMA): attention is ~4 % of the tower FLOPs and the
// whole K/V of a head fits in one SM's shared memory, so the tcgen05/TMEM machinery (128-row
// UMMA tiles, single-thread issue) buys nothing at S <= 257; the GEMMs that carry 96 % of the
// FLOPs are tcgen05 (gemm_sm100.cuh).
//   fwd     : grid (H, B)   K,V resident; warps stride over 16-row query tiles        O, LSE
//   bwd dQ  : grid (H, B)   K,V resident; per-warp Q / dO tiles
//   bwd dKV : grid (H, B)   Q,dO,LSE,delta resident; per-warp K / V tiles
// S is recomputed in both backward kernels (7 instead of 5 GEMM units; no atomics, no dQ
// round trip through HBM).
#include <cuda_bf16.h>

#include "dropout.cuh"
#include "host_util.cuh"

namespace uniir {

constexpr int kHD = 64;          // head dim
constexpr int kBlk = 64;         // rows per block
constexpr int kTileBytes = kBlk * kHD * 2;  // 8 KB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t s_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// 64x64 bf16 tile, 128 B rows, 16 B chunks XOR-swizzled by (row & 7).
// Rows >= valid_rows are zero-filled.  128 threads.
__device__ __forceinline__ void load_tile(uint32_t tile, const __nv_bfloat16* g, long long ld, int row0,
                                          int valid_rows /*absolute row limit*/) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = threadIdx.x + i * 128;
    const int r = id >> 3, c = id & 7;
    const bool ok = (row0 + r) < valid_rows;
    const __nv_bfloat16* src = g + static_cast<long long>(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(tile + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
  }
}
// A fragment (16 rows x 16 k) of a tile: rows [rbase, rbase+16), k-step kk (16 dims)
__device__ __forceinline__ void frag_a(uint32_t tile, int rbase, int kk, uint32_t (&a)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  const int r = rbase + (mi & 1) * 8 + (l & 7);
  const int ch = kk * 2 + (mi >> 1);
  ldsm_x4(tile + r * 128 + ((ch ^ (r & 7)) << 4), a[0], a[1], a[2], a[3]);
}
// B fragments, operand stored [n][k] (k contiguous): n-tiles j, j+1 (8 rows each), k-step kk.
// b[0],b[1] -> n-tile j ; b[2],b[3] -> n-tile j+1
__device__ __forceinline__ void frag_b_nk(uint32_t tile, int j, int kk, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  const int r = (j + (mi >> 1)) * 8 + (l & 7);
  const int ch = kk * 2 + (mi & 1);
  ldsm_x4(tile + r * 128 + ((ch ^ (r & 7)) << 4), b[0], b[1], b[2], b[3]);
}
// B fragments, operand stored [k][n] (n contiguous): k-step t (16 rows), n-tiles jd, jd+1.
__device__ __forceinline__ void frag_b_kn(uint32_t tile, int t, int jd, uint32_t (&b)[4]) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  const int r = t * 16 + (mi & 1) * 8 + (l & 7);
  const int ch = jd + (mi >> 1);
  ldsm_x4_t(tile + r * 128 + ((ch ^ (r & 7)) << 4), b[0], b[1], b[2], b[3]);
}

struct AttnParams {
  const __nv_bfloat16* qkv;  // [B*S, 3*D]   (packed self-attention; backward kernels)
  const __nv_bfloat16 *q, *k, *v;  // forward: q rows [B*Sq] stride ldq; k/v rows [B*Skv] stride ldkv
  long long ldq, ldkv, ldo;
  int Sq, Skv;
  const int* kv_len;         // optional [B]: keys >= kv_len[b] are masked (BERT padding mask)
  const float* bias;         // optional fp32 [H, Sq, Skv] added to scale * QK^T (T5 relative-position bias)
  float* dbias;              // backward with bias: fp32 [H, Sq, Skv] += sum over the batch of dS (atomics)
  __nv_bfloat16 *dq, *dk, *dv;     // backward outputs: dq rows stride lddq; dk/dv rows stride lddkv
  long long lddq, lddkv, lddo;     // lddo: row stride of dout
  __nv_bfloat16* out;        // [B*S, D]
  float* lse;                // [B, H, S]
  const __nv_bfloat16* dout; // [B*S, D]
  float* delta;              // [B, H, S]
  __nv_bfloat16* dqkv;       // [B*S, 3*D]
  int B, S, H, D, causal;
  float scale;               // 1/sqrt(64)
  // attention-probability dropout (DROP instantiations only; med.py:217, T5Attention): element (b, h, q, key) of the
  // batch-wide [items, H, Sq, Skv] probability tensor, counter-based mask of dropout.cuh
  const uint32_t* drop_ctx;
  uint32_t drop_site, drop_thresh;
  float drop_scale;          // 1 / (1 - p)
};

// first element index of batch row b, head h in the batch-wide probability tensor (mod 2^32)
__device__ __forceinline__ uint32_t drop_row_base(const AttnParams& p, int b, int h) {
  return ((__ldg(p.drop_ctx + 1) + static_cast<uint32_t>(b)) * static_cast<uint32_t>(p.H) + static_cast<uint32_t>(h)) *
         static_cast<uint32_t>(p.Sq);
}

// =================================================================================
// Head-resident kernels: one CTA per (batch, head).  The whole K and V (forward, dQ) or
// Q and dO (dK/dV) of the head live in shared memory (S <= 768), loaded once; each warp
// strides over 16-row tiles of the other operand held in a private 2 KB buffer.  The
// streamed dimension is walked in 64-row blocks with a 16-row-granular tail, so S = 257
// costs 17 x 272 score tiles instead of the 320 x 320 of plain 64 x 64 blocking, and
// there is no block-wide barrier inside the loops.
// =================================================================================
constexpr int kAttnWarps = 6;
constexpr int kAttnThreads = kAttnWarps * 32;

// rows [0, spad) of a [S, 64] bf16 matrix -> swizzled smem (zero fill beyond S); all threads
__device__ __forceinline__ void load_rows(uint32_t dst, const __nv_bfloat16* g, long long ld, int S, int spad) {
  for (int id = threadIdx.x; id < spad * 8; id += blockDim.x) {
    const int r = id >> 3, c = id & 7;
    const bool ok = r < S;
    cp_async16(dst + r * 128 + ((c ^ (r & 7)) << 4), g + static_cast<long long>(ok ? r : 0) * ld + c * 8, ok);
  }
}
// 16 rows starting at row0 -> a warp-private swizzled tile; one warp
__device__ __forceinline__ void load_rows16_warp(uint32_t dst, const __nv_bfloat16* g, long long ld, int row0, int S) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + i * 32, r = id >> 3, c = id & 7;
    const bool ok = row0 + r < S;
    cp_async16(dst + r * 128 + ((c ^ (r & 7)) << 4), g + static_cast<long long>(ok ? row0 + r : 0) * ld + c * 8, ok);
  }
}

// BIAS = false is the tower kernel; BIAS = true adds p.bias[h, q, key] (in score units: divided by the scale
// that rides in the exp2 FFMA) -- a separate instantiation, so the tower path compiles exactly as before.
// DROP = true (separate instantiations again): the un-normalised weights are masked and scaled by 1/(1-p) AFTER the
// row sum, so O = dropout(softmax(S)) V as in the reference; the log-sum-exp is that of the undropped softmax.
template <bool BIAS, bool DROP = false>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(const AttnParams p, int spad) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sK = s_u32(sm), sV = sK + spad * 128, sQ0 = sV + spad * 128;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Sq * p.ldq + h * kHD;
  const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;     // valid keys of this batch row
  load_rows(sK, p.k + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD, p.ldkv, p.Skv, spad);
  load_rows(sV, p.v + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD, p.ldkv, p.Skv, spad);
  cp_async_wait_all();
  __syncthreads();
  const uint32_t sQ = sQ0 + warp * 2048;
  const float sc = p.scale * kLog2e;
  __nv_bfloat16* ob = p.out + static_cast<long long>(b) * p.Sq * p.ldo + h * kHD;
  float* lp = p.lse ? p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq : nullptr;
  const int n_qt = (p.Sq + 15) >> 4;
  for (int qt = warp; qt < n_qt; qt += kAttnWarps) {
    const int q0 = qt * 16;
    __syncwarp();
    load_rows16_warp(sQ, qb, p.ldq, q0, p.Sq);
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) frag_a(sQ, 0, kk, qa[kk]);
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
    const int kv_end = p.causal ? min(spad, q0 + 16) : spad;
    for (int k0 = 0; k0 < kv_end; k0 += kBlk) {
      const int nt = min(kBlk, kv_end - k0) >> 3;  // n-tiles of 8 keys in this block (even)
      const uint32_t bK = sK + k0 * 128, bV = sV + k0 * 128;
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          if (j < nt) {
            uint32_t bb[4];
            frag_b_nk(bK, j, kk, bb);
            mma_bf16(s[j], qa[kk], bb[0], bb[1]);
            mma_bf16(s[j + 1], qa[kk], bb[2], bb[3]);
          }
        }
      }
      if constexpr (BIAS) {
        const float inv_scale = 1.f / p.scale;
        const float* bh = p.bias + static_cast<long long>(h) * p.Sq * p.Skv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int qr = (e < 2) ? r0 : r1;
            if (j < nt && key < p.Skv && qr < p.Sq)
              s[j][e] = fmaf(__ldg(bh + static_cast<long long>(qr) * p.Skv + key), inv_scale, s[j][e]);
          }
        }
      }
      // masking only where a block touches the sequence end or the causal diagonal (warp-uniform)
      const bool need_mask = nt < 8 || k0 + kBlk > kvl || (p.causal && k0 + kBlk - 1 > q0);
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int qr = (e < 2) ? r0 : r1;
            if (j >= nt || key >= kvl || (p.causal && key > qr)) s[j][e] = -INFINITY;
          }
        }
      }
      // running max in raw score units; the softmax scale rides in the FFMA that feeds exp2
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
        bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
      const float base0 = (nm0 == -INFINITY) ? 0.f : nm0 * sc, base1 = (nm1 == -INFINITY) ? 0.f : nm1 * sc;
      const float c0 = exp2f(fmaf(m0, sc, -base0)), c1 = exp2f(fmaf(m1, sc, -base1));
      m0 = nm0; m1 = nm1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(fmaf(s[j][0], sc, -base0)); s[j][1] = exp2f(fmaf(s[j][1], sc, -base0));
        s[j][2] = exp2f(fmaf(s[j][2], sc, -base1)); s[j][3] = exp2f(fmaf(s[j][3], sc, -base1));
        rs0 += s[j][0] + s[j][1];
        rs1 += s[j][2] + s[j][3];
        o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
      if constexpr (DROP) {
        const uint32_t dkey = drop_key(p.drop_ctx, p.drop_site), dbase = drop_row_base(p, b, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const uint32_t qr = (e < 2) ? r0 : r1;
            s[j][e] = drop_keep(dkey, (dbase + qr) * static_cast<uint32_t>(p.Skv) + key, p.drop_thresh) ? s[j][e] * p.drop_scale : 0.f;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (2 * t < nt) {
          uint32_t pa[4];
          pa[0] = pack_bf16(s[2 * t][0], s[2 * t][1]);
          pa[1] = pack_bf16(s[2 * t][2], s[2 * t][3]);
          pa[2] = pack_bf16(s[2 * t + 1][0], s[2 * t + 1][1]);
          pa[3] = pack_bf16(s[2 * t + 1][2], s[2 * t + 1][3]);
#pragma unroll
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t bb[4];
            frag_b_kn(bV, t, jd, bb);
            mma_bf16(o[jd], pa, bb[0], bb[1]);
            mma_bf16(o[jd + 1], pa, bb[2], bb[3]);
          }
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + (lane & 3) * 2;
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * p.ldo + col) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * p.ldo + col) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
    }
    if (lp && (lane & 3) == 0) {
      if (r0 < p.Sq) lp[r0] = (m0 * sc + log2f(l0)) / kLog2e;
      if (r1 < p.Sq) lp[r1] = (m1 * sc + log2f(l1)) / kLog2e;
    }
  }
}

// delta[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d]; one warp per (b,s,h)
__global__ void attn_delta_kernel(const AttnParams p) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(p.B) * p.Sq * p.H;
  if (w >= total) return;
  const int lane = threadIdx.x & 31;
  const int h = w % p.H;
  const long long bs = w / p.H;
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(p.dout + bs * p.lddo + h * kHD + lane * 2);
  const __nv_bfloat162 c = *reinterpret_cast<const __nv_bfloat162*>(p.out + bs * p.ldo + h * kHD + lane * 2);
  float v = __bfloat162float(a.x) * __bfloat162float(c.x) + __bfloat162float(a.y) * __bfloat162float(c.y);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) {
    const long long b = bs / p.Sq, s_ = bs % p.Sq;
    p.delta[(b * p.H + h) * p.Sq + s_] = v;
  }
}

// backward dQ: K, V resident; per warp 16 query rows (Q and dO tiles private).
// BIAS (separate instantiation, the tower path compiles as before): scores get p.bias[h, q, key] / scale before the
// exponential, and dS = P o (dP - delta) -- the gradient of the bias -- is accumulated over the batch into p.dbias.
// DROP: dP = mask / (1-p) o (dO V^T); delta = rowsum(dO o O) already holds the dropped output.
template <bool BIAS, bool DROP = false>
__global__ void __launch_bounds__(kAttnThreads, 2) attn_bwd_dq_kernel(const AttnParams p, int spad) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sK = s_u32(sm), sV = sK + spad * 128, sW = sV + spad * 128;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Sq * p.ldq + h * kHD;
  const __nv_bfloat16* dob = p.dout + static_cast<long long>(b) * p.Sq * p.lddo + h * kHD;
  const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
  load_rows(sK, p.k + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD, p.ldkv, p.Skv, spad);
  load_rows(sV, p.v + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD, p.ldkv, p.Skv, spad);
  cp_async_wait_all();
  __syncthreads();
  const uint32_t sQ = sW + warp * 4096, sDO = sQ + 2048;
  const float* lp = p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq;
  const float* dp = p.delta + (static_cast<long long>(b) * p.H + h) * p.Sq;
  const float sc = p.scale * kLog2e;
  __nv_bfloat16* ob = p.dq + static_cast<long long>(b) * p.Sq * p.lddq + h * kHD;
  const int n_qt = (p.Sq + 15) >> 4;
  for (int qt = warp; qt < n_qt; qt += kAttnWarps) {
    const int q0 = qt * 16;
    __syncwarp();
    load_rows16_warp(sQ, qb, p.ldq, q0, p.Sq);
    load_rows16_warp(sDO, dob, p.lddo, q0, p.Sq);
    cp_async_wait_all();
    __syncwarp();
    uint32_t qa[4][4], da[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { frag_a(sQ, 0, kk, qa[kk]); frag_a(sDO, 0, kk, da[kk]); }
    const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
    const float lse0 = (r0 < p.Sq) ? lp[r0] * kLog2e : 0.f, lse1 = (r1 < p.Sq) ? lp[r1] * kLog2e : 0.f;
    const float dl0 = (r0 < p.Sq) ? dp[r0] : 0.f, dl1 = (r1 < p.Sq) ? dp[r1] : 0.f;
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
    const int kv_end = p.causal ? min(spad, q0 + 16) : spad;
    for (int k0 = 0; k0 < kv_end; k0 += kBlk) {
      const int nt = min(kBlk, kv_end - k0) >> 3;
      const uint32_t bK = sK + k0 * 128, bV = sV + k0 * 128;
      float s[8][4], dpv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        dpv[j][0] = dpv[j][1] = dpv[j][2] = dpv[j][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          if (j < nt) {
            uint32_t bb[4];
            frag_b_nk(bK, j, kk, bb);
            mma_bf16(s[j], qa[kk], bb[0], bb[1]);
            mma_bf16(s[j + 1], qa[kk], bb[2], bb[3]);
            frag_b_nk(bV, j, kk, bb);
            mma_bf16(dpv[j], da[kk], bb[0], bb[1]);
            mma_bf16(dpv[j + 1], da[kk], bb[2], bb[3]);
          }
        }
      }
      if constexpr (BIAS) {
        const float inv_scale = 1.f / p.scale;
        const float* bh = p.bias + static_cast<long long>(h) * p.Sq * p.Skv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int qr = (e < 2) ? r0 : r1;
            if (j < nt && key < p.Skv && qr < p.Sq)
              s[j][e] = fmaf(__ldg(bh + static_cast<long long>(qr) * p.Skv + key), inv_scale, s[j][e]);
          }
        }
      }
      const bool need_mask = nt < 8 || k0 + kBlk > kvl || q0 + 16 > p.Sq || (p.causal && k0 + kBlk - 1 > q0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] = exp2f(fmaf(s[j][0], sc, -lse0)); s[j][1] = exp2f(fmaf(s[j][1], sc, -lse0));
        s[j][2] = exp2f(fmaf(s[j][2], sc, -lse1)); s[j][3] = exp2f(fmaf(s[j][3], sc, -lse1));
      }
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int qr = (e < 2) ? r0 : r1;
            if (j >= nt || key >= kvl || qr >= p.Sq || (p.causal && key > qr)) s[j][e] = 0.f;
          }
        }
      }
      if constexpr (DROP) {
        const uint32_t dkey = drop_key(p.drop_ctx, p.drop_site), dbase = drop_row_base(p, b, h);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const uint32_t qr = (e < 2) ? r0 : r1;
            dpv[j][e] = drop_keep(dkey, (dbase + qr) * static_cast<uint32_t>(p.Skv) + key, p.drop_thresh) ? dpv[j][e] * p.drop_scale : 0.f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j][0] *= dpv[j][0] - dl0; s[j][1] *= dpv[j][1] - dl0;
        s[j][2] *= dpv[j][2] - dl1; s[j][3] *= dpv[j][3] - dl1;
      }
      if constexpr (BIAS) {
        float* gh = p.dbias + static_cast<long long>(h) * p.Sq * p.Skv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = k0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int qr = (e < 2) ? r0 : r1;
            if (j < nt && key < p.Skv && qr < p.Sq) atomicAdd(gh + static_cast<long long>(qr) * p.Skv + key, s[j][e]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (2 * t < nt) {
          uint32_t pa[4];
          pa[0] = pack_bf16(s[2 * t][0], s[2 * t][1]);
          pa[1] = pack_bf16(s[2 * t][2], s[2 * t][3]);
          pa[2] = pack_bf16(s[2 * t + 1][0], s[2 * t + 1][1]);
          pa[3] = pack_bf16(s[2 * t + 1][2], s[2 * t + 1][3]);
#pragma unroll
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t bb[4];
            frag_b_kn(bK, t, jd, bb);
            mma_bf16(dq[jd], pa, bb[0], bb[1]);
            mma_bf16(dq[jd + 1], pa, bb[2], bb[3]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + (lane & 3) * 2;
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r0) * p.lddq + col) = pack_bf16(dq[j][0] * p.scale, dq[j][1] * p.scale);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(r1) * p.lddq + col) = pack_bf16(dq[j][2] * p.scale, dq[j][3] * p.scale);
    }
  }
}

// backward dK/dV: Q, dO, LSE, delta resident; per warp 16 keys (K and V tiles private).
// BIAS: as in the dQ kernel (scores only; the bias gradient is accumulated there).
// DROP: dV takes the dropped weights, dS = P o (mask / (1-p) o dP - delta).
template <bool BIAS, bool DROP = false>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_dkv_kernel(const AttnParams p, int spad) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t sQ = s_u32(sm), sDO = sQ + spad * 128, sW = sDO + spad * 128;
  float* s_lse = reinterpret_cast<float*>(sm + 2 * spad * 128 + kAttnWarps * 4096);
  float* s_delta = s_lse + spad;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Sq * p.ldq + h * kHD;
  const __nv_bfloat16* kb_ = p.k + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD;
  const __nv_bfloat16* vb_ = p.v + static_cast<long long>(b) * p.Skv * p.ldkv + h * kHD;
  const __nv_bfloat16* dob = p.dout + static_cast<long long>(b) * p.Sq * p.lddo + h * kHD;
  const float* lp = p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq;
  const float* dp = p.delta + (static_cast<long long>(b) * p.H + h) * p.Sq;
  const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
  load_rows(sQ, qb, p.ldq, p.Sq, spad);          // spad = Sq rounded up to 16
  load_rows(sDO, dob, p.lddo, p.Sq, spad);
  for (int i = threadIdx.x; i < spad; i += blockDim.x) {
    s_lse[i] = (i < p.Sq) ? lp[i] * kLog2e : 0.f;
    s_delta[i] = (i < p.Sq) ? dp[i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  const uint32_t sKw = sW + warp * 4096, sVw = sKw + 2048;
  const float sc = p.scale * kLog2e;
  __nv_bfloat16* okb = p.dk + static_cast<long long>(b) * p.Skv * p.lddkv + h * kHD;
  __nv_bfloat16* ovb = p.dv + static_cast<long long>(b) * p.Skv * p.lddkv + h * kHD;
  const int n_kt = (p.Skv + 15) >> 4;
  for (int kt = warp; kt < n_kt; kt += kAttnWarps) {
    const int kbase = kt * 16;
    __syncwarp();
    load_rows16_warp(sKw, kb_, p.ldkv, kbase, p.Skv);
    load_rows16_warp(sVw, vb_, p.ldkv, kbase, p.Skv);
    cp_async_wait_all();
    __syncwarp();
    uint32_t ka[4][4], va[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { frag_a(sKw, 0, kk, ka[kk]); frag_a(sVw, 0, kk, va[kk]); }
    const int key0 = kbase + (lane >> 2), key1 = key0 + 8;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
    const int q_begin = p.causal ? (kbase / kBlk) * kBlk : 0;  // earlier query blocks see none of these keys
    for (int q0 = q_begin; q0 < spad; q0 += kBlk) {
      const int nt = min(kBlk, spad - q0) >> 3;
      const uint32_t bQ = sQ + q0 * 128, bDO = sDO + q0 * 128;
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
        dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          if (j < nt) {
            uint32_t bb[4];
            frag_b_nk(bQ, j, kk, bb);
            mma_bf16(st[j], ka[kk], bb[0], bb[1]);
            mma_bf16(st[j + 1], ka[kk], bb[2], bb[3]);
            frag_b_nk(bDO, j, kk, bb);
            mma_bf16(dpt[j], va[kk], bb[0], bb[1]);
            mma_bf16(dpt[j + 1], va[kk], bb[2], bb[3]);
          }
        }
      }
      if constexpr (BIAS) {
        const float inv_scale = 1.f / p.scale;
        const float* bh = p.bias + static_cast<long long>(h) * p.Sq * p.Skv;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int q = q0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int key = (e < 2) ? key0 : key1;
            if (j < nt && key < p.Skv && q < p.Sq)
              st[j][e] = fmaf(__ldg(bh + static_cast<long long>(q) * p.Skv + key), inv_scale, st[j][e]);
          }
        }
      }
      const bool need_mask = nt < 8 || q0 + kBlk > p.Sq || kbase + 16 > kvl || (p.causal && kbase + 15 > q0);
      uint32_t pt[4][4], dst[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float pe[4], de[4];
        const int ql = min(q0 + j * 8 + (lane & 3) * 2, spad - 2);   // this thread's two query columns
        const float2 ls = *reinterpret_cast<const float2*>(s_lse + ql);
        const float2 dl = *reinterpret_cast<const float2*>(s_delta + ql);
        pe[0] = exp2f(fmaf(st[j][0], sc, -ls.x)); pe[1] = exp2f(fmaf(st[j][1], sc, -ls.y));
        pe[2] = exp2f(fmaf(st[j][2], sc, -ls.x)); pe[3] = exp2f(fmaf(st[j][3], sc, -ls.y));
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int q = q0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const int key = (e < 2) ? key0 : key1;
            if (j >= nt || key >= kvl || q >= p.Sq || (p.causal && key > q)) pe[e] = 0.f;
          }
        }
        if constexpr (DROP) {
          const uint32_t dkey = drop_key(p.drop_ctx, p.drop_site), dbase = drop_row_base(p, b, h);
          float ms[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t q = q0 + j * 8 + (lane & 3) * 2 + (e & 1);
            const uint32_t key = (e < 2) ? key0 : key1;
            ms[e] = drop_keep(dkey, (dbase + q) * static_cast<uint32_t>(p.Skv) + key, p.drop_thresh) ? p.drop_scale : 0.f;
          }
          de[0] = pe[0] * (dpt[j][0] * ms[0] - dl.x); de[1] = pe[1] * (dpt[j][1] * ms[1] - dl.y);
          de[2] = pe[2] * (dpt[j][2] * ms[2] - dl.x); de[3] = pe[3] * (dpt[j][3] * ms[3] - dl.y);
#pragma unroll
          for (int e = 0; e < 4; ++e) pe[e] *= ms[e];
        } else {
          de[0] = pe[0] * (dpt[j][0] - dl.x); de[1] = pe[1] * (dpt[j][1] - dl.y);
          de[2] = pe[2] * (dpt[j][2] - dl.x); de[3] = pe[3] * (dpt[j][3] - dl.y);
        }
        const int t = j >> 1, hi = (j & 1) * 2;
        pt[t][hi] = pack_bf16(pe[0], pe[1]);
        pt[t][hi + 1] = pack_bf16(pe[2], pe[3]);
        dst[t][hi] = pack_bf16(de[0], de[1]);
        dst[t][hi + 1] = pack_bf16(de[2], de[3]);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        if (2 * t < nt) {
#pragma unroll
          for (int jd = 0; jd < 8; jd += 2) {
            uint32_t bb[4];
            frag_b_kn(bDO, t, jd, bb);
            mma_bf16(dv[jd], pt[t], bb[0], bb[1]);
            mma_bf16(dv[jd + 1], pt[t], bb[2], bb[3]);
            frag_b_kn(bQ, t, jd, bb);
            mma_bf16(dk[jd], dst[t], bb[0], bb[1]);
            mma_bf16(dk[jd + 1], dst[t], bb[2], bb[3]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j * 8 + (lane & 3) * 2;
      if (key0 < p.Skv) {
        *reinterpret_cast<uint32_t*>(okb + static_cast<long long>(key0) * p.lddkv + col) = pack_bf16(dk[j][0] * p.scale, dk[j][1] * p.scale);
        *reinterpret_cast<uint32_t*>(ovb + static_cast<long long>(key0) * p.lddkv + col) = pack_bf16(dv[j][0], dv[j][1]);
      }
      if (key1 < p.Skv) {
        *reinterpret_cast<uint32_t*>(okb + static_cast<long long>(key1) * p.lddkv + col) = pack_bf16(dk[j][2] * p.scale, dk[j][3] * p.scale);
        *reinterpret_cast<uint32_t*>(ovb + static_cast<long long>(key1) * p.lddkv + col) = pack_bf16(dv[j][2], dv[j][3]);
      }
    }
  }
}

}  // namespace uniir

using namespace uniir;
#define ST(s) static_cast<cudaStream_t>(s)

static int attn_smem(int spad, bool dkv) {
  return 2 * spad * 128 + kAttnWarps * (dkv ? 4096 : 2048) + (dkv ? 2 * spad * 4 : 0);
}
static int attn_set_smem() {
  static bool done = false;
  if (!done) {
    const int mx = 227 * 1024;
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    done = true;
  }
  return UNIIR_OK;
}

namespace uniir {
int attn_fwd_tc_try(const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                    long long ldkv, int B, int Sq, int Skv, int H, int causal, const int* kv_len,
                    __nv_bfloat16* out, long long ldo, float* lse, float scale, void* stream);   // attention_tc.cu
}

static int attn_fwd_launch(AttnParams& p, void* stream) {
  // tcgen05 forward (attention_tc.cu) when enabled and the key range fits; > 0 means "not applicable"
  if (!p.bias && !p.drop_ctx) {     // the tcgen05 kernel takes no score bias and no dropout
    const int tc = attn_fwd_tc_try(p.q, p.ldq, p.k, p.v, p.ldkv, p.B, p.Sq, p.Skv, p.H, p.causal, p.kv_len, p.out,
                                   p.ldo, p.lse, p.scale, stream);
    if (tc <= 0) return tc;
  }
  UNIIR_CHECK_ARG(p.Skv <= 768, "uniir_attn_fwd: Skv=%d exceeds the head-resident kernel's limit of 768", p.Skv);
  if (int rc = attn_set_smem()) return rc;
  const int spad = (p.Skv + 15) / 16 * 16;
  const dim3 grid(p.H, p.B);
  const int smem = attn_smem(spad, false);
  if (p.drop_ctx) {
    if (p.bias) attn_fwd_kernel<true, true><<<grid, kAttnThreads, smem, ST(stream)>>>(p, spad);
    else attn_fwd_kernel<false, true><<<grid, kAttnThreads, smem, ST(stream)>>>(p, spad);
  } else if (p.bias) attn_fwd_kernel<true><<<grid, kAttnThreads, smem, ST(stream)>>>(p, spad);
  else attn_fwd_kernel<false><<<grid, kAttnThreads, smem, ST(stream)>>>(p, spad);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, int32_t causal, void* out,
                              float* lse, void* stream) {
  UNIIR_CHECK_ARG(qkv && out && B > 0 && S > 0 && H > 0, "uniir_attn_fwd: bad args");
  AttnParams p{};
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.D = H * kHD;
  p.q = p.qkv; p.k = p.qkv + p.D; p.v = p.qkv + 2 * p.D;
  p.ldq = p.ldkv = 3LL * p.D; p.ldo = p.D;
  p.Sq = p.Skv = S;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.B = B; p.S = S; p.H = H; p.causal = causal; p.scale = 0.125f;
  return attn_fwd_launch(p, stream);
}

extern "C" int uniir_attn_fwd_ex(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                 int32_t B, int32_t Sq, int32_t Skv, int32_t H, int32_t causal,
                                 const int32_t* kv_len, void* out, int64_t ldo, float* lse, void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && B > 0 && Sq > 0 && Skv > 0 && H > 0, "uniir_attn_fwd_ex: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0, "uniir_attn_fwd_ex: strides must keep 16 B row alignment");
  UNIIR_CHECK_ARG(!causal || Sq == Skv, "uniir_attn_fwd_ex: causal needs Sq == Skv");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv; p.ldo = ldo;
  p.Sq = Sq; p.Skv = Skv; p.kv_len = kv_len;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.B = B; p.S = Sq; p.H = H; p.D = H * kHD; p.causal = causal; p.scale = 0.125f;
  return attn_fwd_launch(p, stream);
}

extern "C" int uniir_attn_fwd_bias(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int32_t B,
                                   int32_t Sq, int32_t Skv, int32_t H, float scale, const float* bias, void* out,
                                   int64_t ldo, float* lse, void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && bias && B > 0 && Sq > 0 && Skv > 0 && H > 0 && scale > 0.f, "uniir_attn_fwd_bias: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0, "uniir_attn_fwd_bias: strides must keep 16 B row alignment");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv; p.ldo = ldo;
  p.Sq = Sq; p.Skv = Skv; p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.B = B; p.S = Sq; p.H = H; p.D = H * kHD; p.causal = 0; p.scale = scale;
  return attn_fwd_launch(p, stream);
}

namespace uniir {
long long attn_bwd_ws_floats(int B, int Sq, int Skv, int H);   // attention_bwd_tc.cu
int attn_bwd_tc_try(const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                    long long ldkv, const __nv_bfloat16* out, long long ldo, const __nv_bfloat16* dout,
                    long long lddo, const float* lse, int B, int Sq, int Skv, int H, int causal, const int* kv_len,
                    float* ws, __nv_bfloat16* dq, long long lddq, __nv_bfloat16* dk, __nv_bfloat16* dv,
                    long long lddkv, float scale, void* stream);
}

extern "C" int64_t uniir_attn_bwd_workspace_floats(int32_t B, int32_t Sq, int32_t Skv, int32_t H) {
  return attn_bwd_ws_floats(B, Sq, Skv, H);
}

static int attn_bwd_launch(AttnParams& p, void* stream) {
  // tcgen05 backward (attention_bwd_tc.cu) when enabled and the shape fits; > 0 means "not applicable"
  if (!p.bias && !p.drop_ctx) {     // the tcgen05 kernel takes no score bias and no dropout
    const int tc = attn_bwd_tc_try(p.q, p.ldq, p.k, p.v, p.ldkv, p.out, p.ldo, p.dout, p.lddo, p.lse, p.B, p.Sq, p.Skv,
                                   p.H, p.causal, p.kv_len, p.delta, p.dq, p.lddq, p.dk, p.dv, p.lddkv, p.scale, stream);
    if (tc <= 0) return tc;
  }
  UNIIR_CHECK_ARG(p.Sq <= 768 && p.Skv <= 768, "uniir_attn_bwd: sequence length exceeds the head-resident limit of 768");
  if (int rc = attn_set_smem()) return rc;
  const int spad_kv = (p.Skv + 15) / 16 * 16, spad_q = (p.Sq + 15) / 16 * 16;
  const long long warps = 1LL * p.B * p.Sq * p.H;
  attn_delta_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, ST(stream)>>>(p);
  if (p.drop_ctx) {
    if (p.bias) {
      attn_bwd_dq_kernel<true, true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_kv, true), ST(stream)>>>(p, spad_kv);
      attn_bwd_dkv_kernel<true, true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_q, true), ST(stream)>>>(p, spad_q);
    } else {
      attn_bwd_dq_kernel<false, true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_kv, true), ST(stream)>>>(p, spad_kv);
      attn_bwd_dkv_kernel<false, true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_q, true), ST(stream)>>>(p, spad_q);
    }
  } else if (p.bias) {
    attn_bwd_dq_kernel<true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_kv, true), ST(stream)>>>(p, spad_kv);
    attn_bwd_dkv_kernel<true><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_q, true), ST(stream)>>>(p, spad_q);
  } else {
    attn_bwd_dq_kernel<false><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_kv, true), ST(stream)>>>(p, spad_kv);
    attn_bwd_dkv_kernel<false><<<dim3(p.H, p.B), kAttnThreads, attn_smem(spad_q, true), ST(stream)>>>(p, spad_q);
  }
  count_launch(3);
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                              int32_t B, int32_t S, int32_t H, int32_t causal, float* delta_ws,
                              void* dqkv, void* stream) {
  UNIIR_CHECK_ARG(qkv && out && dout && lse && delta_ws && dqkv && B > 0 && S > 0 && H > 0,
                  "uniir_attn_bwd: bad args");
  AttnParams p{};
  p.D = H * kHD;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.q = p.qkv; p.k = p.qkv + p.D; p.v = p.qkv + 2 * p.D;
  p.ldq = p.ldkv = 3LL * p.D;
  p.out = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(out));
  p.ldo = p.D;
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.lddo = p.D;
  p.lse = const_cast<float*>(lse);
  p.delta = delta_ws;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.dq = p.dqkv; p.dk = p.dqkv + p.D; p.dv = p.dqkv + 2 * p.D;
  p.lddq = p.lddkv = 3LL * p.D;
  p.B = B; p.S = S; p.Sq = p.Skv = S; p.H = H; p.causal = causal; p.scale = 0.125f;
  return attn_bwd_launch(p, stream);
}

extern "C" int uniir_attn_bwd_ex(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                 const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                 int32_t B, int32_t Sq, int32_t Skv, int32_t H, int32_t causal,
                                 const int32_t* kv_len, float* delta_ws, void* dq, int64_t lddq, void* dk,
                                 void* dv, int64_t lddkv, void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && dout && lse && delta_ws && dq && dk && dv && B > 0 && Sq > 0 && Skv > 0 && H > 0,
                  "uniir_attn_bwd_ex: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddq % 2 == 0 && lddkv % 2 == 0,
                  "uniir_attn_bwd_ex: strides must keep row alignment");
  UNIIR_CHECK_ARG(!causal || Sq == Skv, "uniir_attn_bwd_ex: causal needs Sq == Skv");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv;
  p.out = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(out));
  p.ldo = ldo;
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.lddo = lddo;
  p.lse = const_cast<float*>(lse);
  p.delta = delta_ws;
  p.kv_len = kv_len;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddkv = lddkv;
  p.B = B; p.S = Sq; p.Sq = Sq; p.Skv = Skv; p.H = H; p.D = H * kHD; p.causal = causal; p.scale = 0.125f;
  return attn_bwd_launch(p, stream);
}


extern "C" int uniir_attn_bwd_bias(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* out,
                                   int64_t ldo, const void* dout, int64_t lddo, const float* lse, int32_t B, int32_t Sq,
                                   int32_t Skv, int32_t H, float scale, const float* bias, float* dbias, float* delta_ws,
                                   void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv, void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && dout && lse && delta_ws && dq && dk && dv && bias && dbias && B > 0 && Sq > 0 &&
                  Skv > 0 && H > 0 && scale > 0.f, "uniir_attn_bwd_bias: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddq % 2 == 0 && lddkv % 2 == 0,
                  "uniir_attn_bwd_bias: strides must keep row alignment");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv;
  p.out = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(out));
  p.ldo = ldo;
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.lddo = lddo;
  p.lse = const_cast<float*>(lse);
  p.delta = delta_ws;
  p.bias = bias; p.dbias = dbias;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddkv = lddkv;
  p.B = B; p.S = Sq; p.Sq = Sq; p.Skv = Skv; p.H = H; p.D = H * kHD; p.causal = 0; p.scale = scale;
  return attn_bwd_launch(p, stream);
}

// ---- attention with dropout on the probabilities (training of the BLIP BERT stack and the CLIP_FF T5 stack).
// General addressing (as the _ex / _bias entry points), optional key-padding mask, optional score bias, caller's scale.
static void set_drop(AttnParams& p, const uint32_t* ctx, uint32_t site, float pd) {
  p.drop_ctx = ctx; p.drop_site = site; p.drop_thresh = drop_thresh(pd); p.drop_scale = 1.f / (1.f - pd);
}

extern "C" int uniir_attn_fwd_drop(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int32_t B,
                                   int32_t Sq, int32_t Skv, int32_t H, int32_t causal, const int32_t* kv_len, float scale,
                                   const float* bias, void* out, int64_t ldo, float* lse, const uint32_t* drop_ctx,
                                   uint32_t site, float p_drop, void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && drop_ctx && B > 0 && Sq > 0 && Skv > 0 && H > 0 && scale > 0.f && p_drop >= 0.f &&
                  p_drop < 1.f, "uniir_attn_fwd_drop: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0, "uniir_attn_fwd_drop: strides must keep 16 B row alignment");
  UNIIR_CHECK_ARG(!causal || Sq == Skv, "uniir_attn_fwd_drop: causal needs Sq == Skv");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv; p.ldo = ldo;
  p.Sq = Sq; p.Skv = Skv; p.kv_len = kv_len; p.bias = bias;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  p.B = B; p.S = Sq; p.H = H; p.D = H * kHD; p.causal = causal; p.scale = scale;
  set_drop(p, drop_ctx, site, p_drop);
  return attn_fwd_launch(p, stream);
}

extern "C" int uniir_attn_bwd_drop(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* out,
                                   int64_t ldo, const void* dout, int64_t lddo, const float* lse, int32_t B, int32_t Sq,
                                   int32_t Skv, int32_t H, int32_t causal, const int32_t* kv_len, float scale,
                                   const float* bias, float* dbias, float* delta_ws, void* dq, int64_t lddq, void* dk,
                                   void* dv, int64_t lddkv, const uint32_t* drop_ctx, uint32_t site, float p_drop,
                                   void* stream) {
  UNIIR_CHECK_ARG(q && k && v && out && dout && lse && delta_ws && dq && dk && dv && drop_ctx && B > 0 && Sq > 0 && Skv > 0 &&
                  H > 0 && scale > 0.f && p_drop >= 0.f && p_drop < 1.f && (!bias == !dbias), "uniir_attn_bwd_drop: bad args");
  UNIIR_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddq % 2 == 0 && lddkv % 2 == 0,
                  "uniir_attn_bwd_drop: strides must keep row alignment");
  UNIIR_CHECK_ARG(!causal || Sq == Skv, "uniir_attn_bwd_drop: causal needs Sq == Skv");
  AttnParams p{};
  p.q = static_cast<const __nv_bfloat16*>(q);
  p.k = static_cast<const __nv_bfloat16*>(k);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ldq = ldq; p.ldkv = ldkv;
  p.out = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(out));
  p.ldo = ldo;
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.lddo = lddo;
  p.lse = const_cast<float*>(lse);
  p.delta = delta_ws;
  p.kv_len = kv_len;
  p.bias = bias; p.dbias = dbias;
  p.dq = static_cast<__nv_bfloat16*>(dq); p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.lddq = lddq; p.lddkv = lddkv;
  p.B = B; p.S = Sq; p.Sq = Sq; p.Skv = Skv; p.H = H; p.D = H * kHD; p.causal = causal; p.scale = scale;
  set_drop(p, drop_ctx, site, p_drop);
  return attn_bwd_launch(p, stream);
}
