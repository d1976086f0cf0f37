// tcgen05 attention forward for the short sequences of the UniIR towers (S = 257 / 197 / 100 / 77,
// head_dim 64, bf16).  Replaces the forward of nn.MultiheadAttention's softmax(QK^T/8 [+mask])V
// (upstream CLIP resblocks, BLIP ViT / BERT self- and cross-attention) when Skv <= 272.
//
// One persistent CTA per SM walks (batch, head) items.  Because a whole score row (<= 384 keys) fits in
// TMEM there is no online-softmax rescaling: per 128-query tile
//     S[128, Skv]  = Q K^T        tcgen05.mma, fp32 in TMEM columns [0, skv_pad)
//     m = rowmax(S), P = exp2(S*c - m*c), l = rowsum(P)      two passes over TMEM, 256 threads:
//                                                            row = TMEM lane, two threads per row
//     O[128, 64]  += P_c V_c      per 64-key chunk c; P_c staged bf16 in shared memory (K-major,
//                                 128B swizzle, double buffered) so the MMAs overlap the exponentials
//     out = O / l, lse = m*c' + log(l)
// Warp roles (352 threads): warp 0 = TMA producer (K/V of the NEXT head double-buffered against the
// current one, Q tiles through a 2-deep ring), warp 1 = MMA issuer + TMEM owner, warps 2-9 = softmax /
// epilogue, warp 10 = tail rows.  Query rows that do not fill a 128-row MMA tile and are few (257 = 2 x
// 128 + 1: the class token) are computed by warp 10 on CUDA cores from the K/V already in shared memory,
// concurrently with the tiles, instead of a third, 99 % empty tile.  Masks: keys >= kv_len[b] (BERT
// padding) and key > query (CLIP text tower, causal), applied only in the 32-key groups they touch.
// ncu on the first version (one TMEM load at a time, per-element masks, tail on the softmax warps) showed
// the kernel latency-bound at 43 % issue utilisation and 20 % XU: hence two TMEM loads in flight in pass 1,
// alternate chunk ownership of the two warps per SM sub-partition in pass 2, a 3-deep P ring.
#include <cuda_bf16.h>

#include "attn_tc_common.cuh"
#include "gemm_sm100.cuh"
#include "host_util.cuh"

namespace uniir {

constexpr int kTcThreads = 352;
constexpr int kTcPBufs = 3;          // P chunk ring
constexpr int kTcSoftmaxThreads = 256;
constexpr int kTcMaxKv = 288;        // K and V of two heads (double buffer) + Q ring + P ring must fit 227 KB
constexpr int kTcOCol = 320;         // TMEM: scores in columns [0, skv_pad <= 288), O in [320, 384)
constexpr int kTcTailMax = 4;        // leftover query rows computed on CUDA cores
constexpr float kTcLog2e = 1.4426950408889634f;
constexpr float kTcLn2 = 0.6931471805599453f;

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
  long long* trace;                  // diagnostic: SM-clock timestamps of CTA 0's phases (null = off)
};

// diagnostic timeline: role r (0 softmax warp 2, 1 MMA thread, 2 tail warp), slot i < 256
#ifdef UNIIR_ATTN_TRACE   // build with -DUNIIR_ATTN_TRACE for scripts/trace_attn.py; off by default (register pressure)
#define TC_TRACE(r, i) do { if (p.trace && blockIdx.x == 0 && (i) < 256) p.trace[(r) * 256 + (i)] = clock64(); } while (0)
#else
#define TC_TRACE(r, i) do { } while (0)
#endif

struct TcSmem {
  uint8_t *k0, *v0, *q0, *p0;   // K[2], V[2] (kv_bytes each), Q[2], P[kTcPBufs] (16 KB each)
  int kv_bytes;
  float *xmax, *xsum;      // [2][128]
  float *tail_s;           // [kTcMaxKv]  (tail warp's probabilities)
  uint64_t *kv_full, *kv_empty, *q_full, *q_empty;   // [2] each
  uint64_t *p_full, *p_empty;                        // [kTcPBufs] each
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
  s.p0 = c; c += kTcPBufs * 16384;
  s.xmax = reinterpret_cast<float*>(c); c += 2 * 128 * 4;
  s.xsum = reinterpret_cast<float*>(c); c += 2 * 128 * 4;
  s.tail_s = reinterpret_cast<float*>(c); c += kTcMaxKv * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(c);
  s.kv_full = b; s.kv_empty = b + 2; s.q_full = b + 4; s.q_empty = b + 6; s.p_full = b + 8; s.p_empty = b + 12;
  s.s_full = b + 16; s.s_free = b + 17; s.o_full = b + 18; s.o_free = b + 19;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b + 20);
  return s;
}
static int attn_tc_smem_bytes(int kv_bytes) {
  return 1024 /*alignment slack*/ + 4 * kv_bytes + (2 + kTcPBufs) * 16384 + 2 * 2 * 128 * 4 + kTcMaxKv * 4 + 20 * 8 + 16;
}

// running maximum of 32 score columns starting at key k0; MASK: only keys < klim count
template <bool MASK>
__device__ __forceinline__ float max32(const uint32_t (&v)[32], int k0, int klim, float mx) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float x = __uint_as_float(v[i]);
    mx = (!MASK || k0 + i < klim) ? fmaxf(mx, x) : mx;
  }
  return mx;
}
// P = exp2(S*sc - base) for 32 keys of one row -> bf16, 64 bytes of the row's 128-byte line in the P chunk
// (16-byte pieces first_piece .. +3, 128B-swizzled); returns the fp32 row-sum contribution
template <bool MASK>
__device__ __forceinline__ float exp32_store(const uint32_t (&v)[32], int k0, int klim, float sc, float base,
                                             uint8_t* pchunk, int rit, int first_piece) {
  float sum = 0.f;
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float e0 = ex2_approx(fmaf(__uint_as_float(v[i]), sc, -base));
    float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), sc, -base));
    if (MASK) {
      e0 = (k0 + i < klim) ? e0 : 0.f;
      e1 = (k0 + i + 1 < klim) ? e1 : 0.f;
    }
    sum += e0 + e1;
    pk[i >> 1] = pack2_bf16(e0, e1);
  }
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(pchunk + sw128(rit, first_piece + g)) =
        make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
  return sum;
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
      mbar_init(&sm.kv_empty[i], p.tail_rows > 0 ? 2 : 1);   // MMA commit (+ the tail warp, which reads K/V)
      mbar_init(&sm.q_full[i], 1);
      mbar_init(&sm.q_empty[i], 1);
    }
    for (int i = 0; i < kTcPBufs; ++i) {
      mbar_init(&sm.p_full[i], kTcSoftmaxThreads / 2);   // the 128 threads of the half that owns the chunk
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
          TC_TRACE(1, tile_it * 12 + 0);
          mbar_wait(&sm.q_full[qb], (q_it >> 1) & 1);
          TC_TRACE(1, tile_it * 12 + 1);
          if (tile_it >= 1) mbar_wait(sm.s_free, (tile_it - 1) & 1);   // previous tile's scores are consumed
          tc_fence_after();
          TC_TRACE(1, tile_it * 12 + 2);
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
          TC_TRACE(1, tile_it * 12 + 3);
          if (tile_it >= 1) mbar_wait(sm.o_free, (tile_it - 1) & 1);   // previous tile's O has been read
          TC_TRACE(1, tile_it * 12 + 4);
          for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
            const int pb = chunk_it % kTcPBufs;
            mbar_wait(&sm.p_full[pb], (chunk_it / kTcPBufs) & 1);
            tc_fence_after();
            TC_TRACE(1, tile_it * 12 + 5 + c);
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
  } else if (warp < 10) {
    // =============================== softmax / epilogue ===============================
    const int sw = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lanes [32 q, 32 q + 32)
    const int half = sw >> 2;                // which 32 columns of each 64-key chunk / which half of O
    const int rit = quarter * 32 + lane;     // row in tile
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = p.scale * kTcLog2e;
    const int n32 = (p.skv_pad + 31) >> 5;
    int tile_it = 0, chunk_it = 0;
    for (int head = blockIdx.x; head < n_heads; head += gridDim.x) {
      const int b = head / p.H, h = head % p.H;
      const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
      for (int t = 0; t < p.n_qt; ++t, ++tile_it) {
        const int wrow0 = t * 128 + quarter * 32;                  // first row of this warp
        const int row = wrow0 + lane;
        const bool warp_valid = wrow0 < p.Sq;                      // warp-uniform
        const int klim = p.causal ? min(kvl, row + 1) : kvl;       // keys [0, klim) are visible to this row
        // a 32-key group starting at k0 needs per-element masking iff it crosses kvl or the causal diagonal
        auto needs_mask = [&](int k0) { return k0 + 32 > kvl || (p.causal && k0 + 31 > wrow0); };
        const bool tr_on = warp == 2 && lane == 0;
        if (tr_on) TC_TRACE(0, tile_it * 12 + 0);
        mbar_wait(sm.s_full, tile_it & 1);
        tc_fence_after();
        if (tr_on) TC_TRACE(0, tile_it * 12 + 1);
        // ---- pass 1: row maximum (this thread: every other 32-column group), two TMEM loads in flight
        float mx = -INFINITY;
        if (warp_valid) {
          uint32_t va[32], vb[32];
          int j = half;
          for (; j + 2 < n32; j += 4) {
            tmem_ld_32x32b_x32(tlane + j * 32, va);
            tmem_ld_32x32b_x32(tlane + (j + 2) * 32, vb);
            tmem_ld_wait();
            mx = needs_mask(j * 32) ? max32<true>(va, j * 32, klim, mx) : max32<false>(va, j * 32, klim, mx);
            mx = needs_mask((j + 2) * 32) ? max32<true>(vb, (j + 2) * 32, klim, mx) : max32<false>(vb, (j + 2) * 32, klim, mx);
          }
          if (j < n32) {
            tmem_ld_32x32b_x32(tlane + j * 32, va);
            tmem_ld_wait();
            mx = needs_mask(j * 32) ? max32<true>(va, j * 32, klim, mx) : max32<false>(va, j * 32, klim, mx);
          }
        }
        sm.xmax[half * 128 + rit] = mx;
        if (tr_on) TC_TRACE(0, tile_it * 12 + 2);
        named_bar_sync(1, kTcSoftmaxThreads);
        if (tr_on) TC_TRACE(0, tile_it * 12 + 3);
        mx = fmaxf(sm.xmax[rit], sm.xmax[128 + rit]);
        const float base = (mx == -INFINITY) ? 0.f : mx * sc;
        // ---- pass 2: P = exp2(S*sc - base), 64-key chunks into the shared-memory ring, row sums.  The two
        // warps that share an SM sub-partition (same TMEM lane quarter, half 0 / 1) own ALTERNATE chunks, so
        // one is in its MUFU-bound exponentials while the other loads TMEM / waits for a free P buffer /
        // fences and signals (ncu + SM-clock trace: with both halves on the same chunk those ~650 cycles of
        // per-chunk overhead were exposed on top of ~500 cycles of exponentials).
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c, ++chunk_it) {
          if ((c & 1) != half) continue;                       // the other half's chunk
          const int pb = chunk_it % kTcPBufs, use = chunk_it / kTcPBufs;
          const int k0 = c * 64;
          const bool second = k0 + 32 < p.skv_pad;             // keys [k0+32, k0+64) exist
          uint32_t va[32];                   // one 32-key group at a time: the sibling warp hides the latency
          if (warp_valid) tmem_ld_32x32b_x32(tlane + k0, va);
          if (use >= 1) mbar_wait(&sm.p_empty[pb], (use - 1) & 1);
          if (warp_valid) {
            tmem_ld_wait();
            sum += needs_mask(k0) ? exp32_store<true>(va, k0, klim, sc, base, sm.p(pb), rit, 0)
                                  : exp32_store<false>(va, k0, klim, sc, base, sm.p(pb), rit, 0);
            if (second) {
              tmem_ld_32x32b_x32(tlane + k0 + 32, va);
              tmem_ld_wait();
              sum += needs_mask(k0 + 32) ? exp32_store<true>(va, k0 + 32, klim, sc, base, sm.p(pb), rit, 4)
                                         : exp32_store<false>(va, k0 + 32, klim, sc, base, sm.p(pb), rit, 4);
            }
          }
          fence_proxy_async();             // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(&sm.p_full[pb]);
          if (tr_on) TC_TRACE(0, tile_it * 12 + 4 + c);
        }
        tc_fence_before();
        mbar_arrive(sm.s_free);            // score columns may be overwritten by the next tile's QK^T
        sm.xsum[half * 128 + rit] = sum;
        named_bar_sync(1, kTcSoftmaxThreads);
        if (tr_on) TC_TRACE(0, tile_it * 12 + 9);
        const float total = sm.xsum[rit] + sm.xsum[128 + rit];
        const float inv = total > 0.f ? 1.f / total : 0.f;
        // ---- epilogue: O / l
        mbar_wait(sm.o_full, tile_it & 1);
        tc_fence_after();
        if (tr_on) TC_TRACE(0, tile_it * 12 + 10);
        if (warp_valid) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tlane + kTcOCol + half * 32, v);
          tmem_ld_wait();
          if (row < p.Sq) {
            uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * 64 + half * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                w[e] = pack2_bf16(__uint_as_float(v[8 * g + 2 * e]) * inv, __uint_as_float(v[8 * g + 2 * e + 1]) * inv);
              dst[g] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (p.lse && half == 0)
              p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + row] = (mx * sc + lg2_approx(total)) * kTcLn2;
          }
        }
        tc_fence_before();
        mbar_arrive(sm.o_free);
        if (tr_on) TC_TRACE(0, tile_it * 12 + 11);
      }
    }
  } else {
    // =============================== tail rows (warp 10) ===============================
    // The few query rows beyond the last full 128-row tile, against the K/V in shared memory: one warp,
    // shuffles only, concurrent with the tiles.  Lane l scores keys l, l+32, ...; probabilities go through a
    // warp-private vector; lane l then accumulates output dims 2l, 2l+1.
    if (p.tail_rows > 0) {
      const float sc = p.scale * kTcLog2e;
      int kv_it = 0;
      for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++kv_it) {
        const int b = head / p.H, h = head % p.H;
        const int kb = kv_it & 1;
        const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
        if (lane == 0) TC_TRACE(2, kv_it * 3 + 0);
        mbar_wait(&sm.kv_full[kb], (kv_it >> 1) & 1);
        if (lane == 0) TC_TRACE(2, kv_it * 3 + 1);
        const uint8_t* sK = sm.k(kb);
        const uint8_t* sV = sm.v(kb);
        for (int tr = 0; tr < p.tail_rows; ++tr) {
          const int row = p.n_qt * 128 + tr;
          const int klim = p.causal ? min(kvl, row + 1) : kvl;
          const uint4* qg = reinterpret_cast<const uint4*>(p.q + (static_cast<long long>(b) * p.Sq + row) * p.ldq + h * 64);
          uint4 qp[8];                                   // the query row, packed bf16 (unpacked on the fly)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) qp[ch] = __ldg(qg + ch);
          // scores of keys lane, lane+32, ... -> tail_s (each lane re-reads only what it wrote itself)
          __syncwarp();                                  // previous row's readers of tail_s are done
          float mx = -INFINITY;
#pragma unroll 2
          for (int j = lane; j < p.skv_pad; j += 32) {
            float acc = 0.f;
            if (j < klim) {
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                const uint4 kk = *reinterpret_cast<const uint4*>(sK + sw128(j, ch));
                const uint4 qq = qp[ch];
                acc += bf16_lo(kk.x) * bf16_lo(qq.x) + bf16_hi(kk.x) * bf16_hi(qq.x) + bf16_lo(kk.y) * bf16_lo(qq.y) +
                       bf16_hi(kk.y) * bf16_hi(qq.y) + bf16_lo(kk.z) * bf16_lo(qq.z) + bf16_hi(kk.z) * bf16_hi(qq.z) +
                       bf16_lo(kk.w) * bf16_lo(qq.w) + bf16_hi(kk.w) * bf16_hi(qq.w);
              }
            }
            const float sj = (j < klim) ? acc * sc : -INFINITY;
            sm.tail_s[j] = sj;
            mx = fmaxf(mx, sj);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          const float base = (mx == -INFINITY) ? 0.f : mx;
          float total = 0.f;
          for (int j = lane; j < p.skv_pad; j += 32) {
            const float e = (j < klim) ? ex2_approx(sm.tail_s[j] - base) : 0.f;
            sm.tail_s[j] = e;
            total += e;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
          __syncwarp();                                  // probabilities visible to every lane
          // O: lane = (key group g = lane / 8, dims 8 (lane % 8) .. +7); 4 independent key streams, then a
          // butterfly over the groups; lanes 0..7 write the 128-byte output row
          const int g = lane >> 3, d8 = lane & 7;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 4
          for (int j = g; j < klim; j += 4) {
            const uint4 vv = *reinterpret_cast<const uint4*>(sV + sw128(j, d8));
            const float pj = sm.tail_s[j];
            acc[0] = fmaf(pj, bf16_lo(vv.x), acc[0]); acc[1] = fmaf(pj, bf16_hi(vv.x), acc[1]);
            acc[2] = fmaf(pj, bf16_lo(vv.y), acc[2]); acc[3] = fmaf(pj, bf16_hi(vv.y), acc[3]);
            acc[4] = fmaf(pj, bf16_lo(vv.z), acc[4]); acc[5] = fmaf(pj, bf16_hi(vv.z), acc[5]);
            acc[6] = fmaf(pj, bf16_lo(vv.w), acc[6]); acc[7] = fmaf(pj, bf16_hi(vv.w), acc[7]);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
          }
          const float inv = total > 0.f ? 1.f / total : 0.f;
          if (lane < 8)
            *reinterpret_cast<uint4*>(p.out + (static_cast<long long>(b) * p.Sq + row) * p.ldo + h * 64 + 8 * lane) =
                make_uint4(pack2_bf16(acc[0] * inv, acc[1] * inv), pack2_bf16(acc[2] * inv, acc[3] * inv),
                           pack2_bf16(acc[4] * inv, acc[5] * inv), pack2_bf16(acc[6] * inv, acc[7] * inv));
          if (p.lse && lane == 0)
            p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + row] = (mx + lg2_approx(total)) * kTcLn2;
        }
        __syncwarp();
        if (lane == 0) TC_TRACE(2, kv_it * 3 + 2);
        if (lane == 0) mbar_arrive(&sm.kv_empty[kb]);    // the tail warp no longer reads this head's K/V
      }
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

// mode bits: 1 = tcgen05 forward, 2 = tcgen05 backward (attention_bwd_tc.cu), wherever they apply and pay off;
// 4 = also where they apply but measured slower (the causal 77-token forward); 0 = the mma.sync head-resident
// kernels (attention.cu) everywhere.  Default 3; environment UNIIR_ATTN_TC overrides.
static int g_attn_tc_mode = -1;
static int attn_tc_mode() {
  if (g_attn_tc_mode < 0) {
    const char* e = getenv("UNIIR_ATTN_TC");
    g_attn_tc_mode = e ? (atoi(e) & 7) : 3;
  }
  return g_attn_tc_mode;
}
namespace uniir {
int attn_tc_mode_bits() { return attn_tc_mode(); }
}
static long long* g_attn_tc_trace = nullptr;
extern "C" int uniir_attn_set_trace(void* device_buf_3x256_int64) {
  g_attn_tc_trace = static_cast<long long*>(device_buf_3x256_int64);
  return UNIIR_OK;
}
extern "C" int uniir_attn_set_mode(int32_t mode) {
  const int prev = attn_tc_mode();
  g_attn_tc_mode = mode & 7;
  return prev;
}

namespace uniir {
// Called by attention.cu's forward entry points.  Returns UNIIR_OK after launching, a negative error,
// or +1 when the tcgen05 path does not apply (caller falls through to the mma.sync kernel).
int attn_fwd_tc_try(const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                    long long ldkv, int B, int Sq, int Skv, int H, int causal, const int* kv_len,
                    __nv_bfloat16* out, long long ldo, float* lse, float scale, void* stream) {
  if (!(attn_tc_mode() & 1)) return 1;
  if (causal && !(attn_tc_mode() & 4)) return 1;   // CLIP text tower (77 tokens): 33.6 us vs 32.2 us for mma.sync
  if (Skv > kTcMaxKv || ldo % 8 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return 1;
  AttnTcParams p{};
  p.q = q; p.ldq = ldq; p.ldo = ldo; p.out = out; p.lse = lse; p.kv_len = kv_len;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.causal = causal; p.scale = scale;
  p.trace = g_attn_tc_trace;
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
  if (smem > 227 * 1024) return 1;     // K/V double buffer does not fit: the mma.sync kernel takes it
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
