// tcgen05 attention backward for the short sequences of the UniIR towers (Sq <= 260, head_dim 64, bf16).
// Replaces the autograd backward of nn.MultiheadAttention's softmax(QK^T/8 [+mask])V (upstream CLIP
// resblocks; BLIP vit.py:75-101 Attention, med.py:97-222 BertSelfAttention incl. cross-attention).
//
//   P  = exp(S/8 - LSE)           S = Q K^T (recomputed)
//   dV = P^T dO                   dP = dO V^T
//   dS = P o (dP - delta) / 8     delta = rowsum(dO o O)
//   dQ = dS K                     dK = dS^T Q
//
// Two launches:
//  1. attn_bwd_prep_kernel (CUDA cores, one CTA per head): delta, the base-2 LSE, both padded to a multiple
//     of 16 rows (+inf LSE in the padding so that padded query columns produce P = 0 with no branch), and the
//     "+1" of 257 = 2 x 128 + 1: the few key columns / query rows beyond the last full 128-row MMA tile.
//  2. attn_bwd_tc_kernel (persistent, one CTA per SM, work item = head x 128-key tile j x 64-query chunk i).
//     Keys are the MMA M dimension (= TMEM lanes), queries the N dimension:
//        S^T_ji  [128 keys x 64 q] = K_j  Q_i^T          dP^T_ji = V_j dO_i^T         (TMEM, double buffered)
//        256 threads (row = key): P^T, dS^T -> bf16 -> shared memory (128B-swizzled rows of 64 queries)
//        dV_j [128 x 64] += P^T_ji  dO_i     dK_j += dS^T_ji Q_i      (A = staged tile, K-major; B MN-major)
//        dQ_T [128 q x 64] += dS_Tj K_j      (A = a PAIR of staged dS^T tiles read MN-major, B MN-major)
//     TMEM columns: S^T 2 x 64 | dP^T 2 x 64 | dK_j 64 | dV_j 64 | dQ 2 x 64 = 512.  P and S are computed
//     once (the mma.sync kernels recompute them in a dQ and a dK/dV kernel).
//     warp 0 = TMA producer (K_j / V_j tiles in 2 slots; Q_i, dO_i, LSE, delta chunks in a 4-stage ring that
//     runs across tile and head boundaries), warp 1 = MMA issuer (the S^T/dP^T MMAs of chunk g+1 are issued
//     before the dV/dK/dQ MMAs of chunk g, so they overlap the exponentials of chunk g), warps 2-9 =
//     elementwise, warps 10-13 = epilogues (dK_j/dV_j after each key tile, dQ after each head; ncu on the
//     first version, where the elementwise warps also ran the epilogues: 28 % of their samples waiting for the
//     accumulators or for the epilogue's global loads).
#include <cuda_bf16.h>
#include <math.h>

#include "attn_tc_common.cuh"
#include "gemm_sm100.cuh"
#include "host_util.cuh"

namespace uniir {

// warp 0 TMA, warp 1 MMA, EW elementwise warps (8: two 16-query units of a chunk per thread; 16: one), then 4
// epilogue warps
constexpr int bw_threads(int ew_warps, bool fold = false) { return 64 + 32 * ew_warps + 128 + (fold ? 64 : 0); }
constexpr int kBwEpThreads = 128;
constexpr int kBwStages = 4;         // Q_i / dO_i / stats ring
constexpr int kBwTailMax = 4;        // leftover key columns / query rows handled by the prep kernel
constexpr int kBwMaxQTiles = 2;      // dQ accumulators in TMEM
constexpr float kBwLog2e = 1.4426950408889634f;
// TMEM columns
constexpr int kColS = 0, kColDP = 128, kColDK = 256, kColDV = 320, kColDQ = 384;
static_assert(kColDV == kColDK + 64, "the dK/dV epilogue walks both accumulators as one 128-column range");

struct AttnBwdTcParams {
  CUtensorMap tmQ64, tmQ16, tmDO64, tmDO16;   // box 64 cols x 64 / 16 rows
  CUtensorMap tmK, tmV;                       // box 64 cols x 128 rows
  const float* lse2;        // [B*H, sq_pad]  LSE * log2(e), +inf in the padding
  const float* delta_s;     // [B*H, sq_pad]  delta * scale
  const float* ds_col;      // [B*H, tail_k, sq_pad]  dS of the tail key columns (fp32), for the dQ epilogue
  float* fold_ws;           // FOLD: per CTA 2 x sq_pad (dS of the tail key column, by head parity) + 192 floats
  const __nv_bfloat16* k;   // tail key rows are read directly in the dQ epilogue
  long long ldkv;
  const __nv_bfloat16 *v_g, *q_g, *do_g;   // FOLD: tail rows of V, Q, dO
  long long ldq_g, lddo_g;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddkv;
  const int* kv_len;
  int B, H, Sq, Skv, causal;
  int sq_pad;               // Sq rounded up to 16
  int n_chunks;             // 64-query chunks per key tile
  int n_kt, tail_k;         // 128-key MMA tiles; leftover key columns (prep kernel)
  int n_qt, tail_q;         // 128-query dQ tiles; leftover query rows (prep kernel)
  float scale;
};

struct BwSmem {
  uint8_t *q0, *do0;        // [kBwStages] x 8 KB
  uint8_t *k0, *v0;         // [2] x 16 KB
  uint8_t *p0;              // [2] x 16 KB    P^T chunks
  uint8_t *ds0;             // [4] x 16 KB    dS^T chunks: 2 pairs, the halves of a pair adjacent
  float* stats;             // [kBwStages][128]: lse2[64] | delta_s[64]
  uint64_t *st_full, *st_empty;     // [kBwStages]
  uint64_t *kv_full, *kv_empty;     // [2]
  uint64_t *s_full;                 // [2]  S^T / dP^T of chunk g are in TMEM buffer g & 1
  uint64_t *chunk_done;             // [2]  the 256 elementwise threads have read TMEM buffer g & 1 and staged P^T / dS^T
  uint64_t *bufs_free;              // [2]  the dV / dK / dQ MMAs of chunk g (and every earlier MMA) have retired
  uint64_t *akv_full, *akv_free, *dq_full, *dq_free;
  uint64_t *tail_done;               // [2] FOLD: the tail warps have published dS of the tail key column (by head parity)
  uint32_t* tmem_ptr;
};
__device__ __forceinline__ BwSmem bw_carve(uint8_t* base) {
  BwSmem s;
  uint8_t* c = base;
  s.q0 = c; c += kBwStages * 8192;
  s.do0 = c; c += kBwStages * 8192;
  s.k0 = c; c += 2 * 16384;
  s.v0 = c; c += 2 * 16384;
  s.p0 = c; c += 2 * 16384;
  s.ds0 = c; c += 4 * 16384;
  s.stats = reinterpret_cast<float*>(c); c += kBwStages * 128 * 4;
  uint64_t* b = reinterpret_cast<uint64_t*>(c);
  s.st_full = b; b += kBwStages;
  s.st_empty = b; b += kBwStages;
  s.kv_full = b; b += 2;
  s.kv_empty = b; b += 2;
  s.s_full = b; b += 2;
  s.chunk_done = b; b += 2;
  s.bufs_free = b; b += 2;
  s.akv_full = b++; s.akv_free = b++; s.dq_full = b++; s.dq_free = b++;
  s.tail_done = b; b += 2;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(b);
  return s;
}
// the dynamic shared memory array is declared 1024-byte aligned; 512 bytes of slack in case the driver places
// it at a smaller alignment (checked in the kernel: a larger shift traps instead of overrunning)
constexpr int kBwSmemSlack = 512;
constexpr int kBwSmemBytes = kBwSmemSlack + 2 * kBwStages * 8192 + 10 * 16384 + kBwStages * 512 +
                             (2 * kBwStages + 16) * 8 + 16;
static_assert(kBwSmemBytes <= 227 * 1024, "shared memory budget");

// ---------------------------------------------------------------------------------------------------
// prep kernel
// ---------------------------------------------------------------------------------------------------
struct AttnBwdPrepParams {
  const __nv_bfloat16 *q, *k, *v, *out, *dout;
  long long ldq, ldkv, ldo, lddo;
  const float* lse;         // [B, H, Sq]
  const int* kv_len;
  float *lse2, *delta_s, *ds_col;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddkv;
  int B, H, Sq, Skv, causal, sq_pad, n_kt, tail_k, n_qt, tail_q;
  float scale;
};

__device__ __forceinline__ float bw_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float2 bw_ld_bf2(const __nv_bfloat16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}

// 64-wide dot product of two bf16 rows held as 8 x uint4
__device__ __forceinline__ float dot64(const uint4 (&a)[8], const uint4 (&b)[8]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint32_t x[4] = {a[c].x, a[c].y, a[c].z, a[c].w}, y[4] = {b[c].x, b[c].y, b[c].z, b[c].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(bf16_lo(x[e]), bf16_lo(y[e]), s0);
      s1 = fmaf(bf16_hi(x[e]), bf16_hi(y[e]), s1);
    }
  }
  return s0 + s1;
}
__device__ __forceinline__ void ld_row(uint4 (&r)[8], const __nv_bfloat16* p) {
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int c = 0; c < 8; ++c) r[c] = __ldg(p4 + c);
}
// row held in registers . row in global memory (the same row for every thread: broadcast loads, L1 hits)
__device__ __forceinline__ float dot64_g(const uint4 (&a)[8], const __nv_bfloat16* g) {
  const uint4* g4 = reinterpret_cast<const uint4*>(g);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 y4 = __ldg(g4 + c);
    const uint32_t x[4] = {a[c].x, a[c].y, a[c].z, a[c].w}, y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(bf16_lo(x[e]), bf16_lo(y[e]), s0);
      s1 = fmaf(bf16_hi(x[e]), bf16_hi(y[e]), s1);
    }
  }
  return s0 + s1;
}
// acc[0..8) += w * (8 bf16 of a row chunk)
__device__ __forceinline__ void axpy8(float (&acc)[8], float w, const uint4 x) {
  acc[0] = fmaf(w, bf16_lo(x.x), acc[0]); acc[1] = fmaf(w, bf16_hi(x.x), acc[1]);
  acc[2] = fmaf(w, bf16_lo(x.y), acc[2]); acc[3] = fmaf(w, bf16_hi(x.y), acc[3]);
  acc[4] = fmaf(w, bf16_lo(x.z), acc[4]); acc[5] = fmaf(w, bf16_hi(x.z), acc[5]);
  acc[6] = fmaf(w, bf16_lo(x.w), acc[6]); acc[7] = fmaf(w, bf16_hi(x.w), acc[7]);
}

constexpr int kPrepThreads = 288;                  // one query / key row per thread
constexpr int kPrepRowGroups = kPrepThreads / 8;   // reduction pass: 8 threads x 16 bytes per row, 36 rows at a time

// grid (H, B).  One query (and key) row per thread, every dot product thread-local, all row loads of a pass in
// flight at once; then one pass of 16-byte loads for the three weighted column sums (dV, dK of the tail key, dQ
// of the tail query).  History: v1 walked rows warp by warp with shuffle reductions and was latency-bound (259 us
// for the ViT-L/14 shape, against 305 us for the tcgen05 kernel itself); v2 (thread per row, but 2-byte column
// sums and a second trip for the 257th row) 158 us.
__global__ void __launch_bounds__(kPrepThreads, 2) attn_bwd_prep_kernel(const AttnBwdPrepParams p) {
  __shared__ float w_p[kPrepThreads], w_ds[kPrepThreads], w_dq[kPrepThreads];
  __shared__ float red[kPrepRowGroups][3][64];
  const int h = blockIdx.x, b = blockIdx.y, bh = b * p.H + h;
  const int tid = threadIdx.x, r = threadIdx.x;
  float* lse2 = p.lse2 + static_cast<long long>(bh) * p.sq_pad;
  float* dls = p.delta_s + static_cast<long long>(bh) * p.sq_pad;
  const float* lse = p.lse + static_cast<long long>(bh) * p.Sq;
  const float sc = p.scale * kBwLog2e;
  const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
  const __nv_bfloat16* qb = p.q + static_cast<long long>(b) * p.Sq * p.ldq + h * 64;
  const __nv_bfloat16* dob = p.dout + static_cast<long long>(b) * p.Sq * p.lddo + h * 64;
  const __nv_bfloat16* ob = p.out + static_cast<long long>(b) * p.Sq * p.ldo + h * 64;
  const __nv_bfloat16* kb = p.k + static_cast<long long>(b) * p.Skv * p.ldkv + h * 64;
  const __nv_bfloat16* vb = p.v + static_cast<long long>(b) * p.Skv * p.ldkv + h * 64;
  const int n_pass = max(1, max(p.tail_k, p.tail_q));

  for (int t = 0; t < n_pass; ++t) {
    const bool do_k = t < p.tail_k, do_q = t < p.tail_q;
    const int kt = p.n_kt * 128 + t, qt = p.n_qt * 128 + t;
    // rows beyond one thread per row only exist without tails (delta only): strided loop
    for (int q = r; q < p.sq_pad; q += kPrepThreads) {
      float pr = 0.f, ds = 0.f;
      if (q < p.Sq && (t == 0 || do_k)) {
        uint4 dd[8];
        ld_row(dd, dob + q * p.lddo);
        float dl, l2;
        if (t == 0) {   // delta * scale and LSE * log2(e)
          uint4 o[8];
          ld_row(o, ob + q * p.ldo);
          dl = dot64(o, dd) * p.scale;
          l2 = lse[q] * kBwLog2e;
          dls[q] = dl;
          lse2[q] = l2;
        } else {
          dl = dls[q];
          l2 = lse2[q];
        }
        if (do_k) {     // tail key column kt against this query
          uint4 qq[8];
          ld_row(qq, qb + q * p.ldq);
          const float s = dot64_g(qq, kb + kt * p.ldkv), dp = dot64_g(dd, vb + kt * p.ldkv);
          pr = (kt < kvl && (!p.causal || kt <= q)) ? ex2_approx(fmaf(s, sc, -l2)) : 0.f;
          ds = pr * fmaf(dp, p.scale, -dl);
          if (q < p.n_qt * 128) p.ds_col[(static_cast<long long>(bh) * p.tail_k + t) * p.sq_pad + q] = ds;
        }
      } else if (t == 0) {
        dls[q] = 0.f;
        lse2[q] = INFINITY;
      }
      if (q < kPrepThreads) { w_p[q] = pr; w_ds[q] = ds; }
    }
    if (!do_k && !do_q) return;   // no tails (uniform)
    {
      float dsq = 0.f;            // tail query row qt against key r
      const int klim = p.causal ? min(kvl, qt + 1) : kvl;
      if (do_q && r < klim) {
        uint4 kk[8], vv[8];
        ld_row(kk, kb + r * p.ldkv);
        ld_row(vv, vb + r * p.ldkv);
        uint4 dq_[8];
        ld_row(dq_, dob + qt * p.lddo);
        const float dlq = dot64_g(dq_, ob + qt * p.ldo) * p.scale, l2q = lse[qt] * kBwLog2e;
        const float s = dot64_g(kk, qb + qt * p.ldq), dp = dot64(vv, dq_);
        dsq = ex2_approx(fmaf(s, sc, -l2q)) * fmaf(dp, p.scale, -dlq);
      }
      w_dq[r] = dsq;
    }
    __syncthreads();
    // ---- weighted column sums: dV[kt] = sum_q P dO[q], dK[kt] = sum_q dS Q[q], dQ[qt] = sum_k dS K[k]
    {
      const int dg = tid & 7, rg = tid >> 3;
      float av[8], ak[8], aq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) av[e] = ak[e] = aq[e] = 0.f;
      const int nrow = max(p.Sq, p.Skv);
#pragma unroll 2
      for (int rr = rg; rr < nrow; rr += kPrepRowGroups) {
        if (do_k && rr < p.Sq) {
          axpy8(av, w_p[rr], __ldg(reinterpret_cast<const uint4*>(dob + rr * p.lddo) + dg));
          axpy8(ak, w_ds[rr], __ldg(reinterpret_cast<const uint4*>(qb + rr * p.ldq) + dg));
        }
        if (do_q && rr < p.Skv) axpy8(aq, w_dq[rr], __ldg(reinterpret_cast<const uint4*>(kb + rr * p.ldkv) + dg));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[rg][0][8 * dg + e] = av[e];
        red[rg][1][8 * dg + e] = ak[e];
        red[rg][2][8 * dg + e] = aq[e];
      }
    }
    __syncthreads();
    if (tid < 192) {
      const int which = tid >> 6, d = tid & 63;
      float s = 0.f;
#pragma unroll 4
      for (int g = 0; g < kPrepRowGroups; ++g) s += red[g][which][d];
      if (which < 2 && do_k)
        (which == 0 ? p.dv : p.dk)[(static_cast<long long>(b) * p.Skv + kt) * p.lddkv + h * 64 + d] = __float2bfloat16(s);
      if (which == 2 && do_q)
        p.dq[(static_cast<long long>(b) * p.Sq + qt) * p.lddq + h * 64 + d] = __float2bfloat16(s);
    }
    __syncthreads();   // w_* / red are rewritten by the next pass
  }
}

// ---------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------
// One 16-query unit of one key row: P^T and dS^T -> two 16-byte pieces each (pieces 2u, 2u+1 of the row).
// MASK: this thread's key may be padding (key_ok) and / or the chunk crosses the causal diagonal (key > query
// -> 0); the unmasked instantiation is the common case (warp-uniform choice by the caller).
template <bool MASK>
__device__ __forceinline__ void bw_unit(const uint32_t (&s)[16], const uint32_t (&dp)[16], const float* st_lse,
                                        const float* st_dl, int u, float sc, float scale, bool key_ok, int qmin,
                                        uint8_t* pbuf, uint8_t* dsbuf, int row) {
  // qmin: first query this key may attend to, relative to the unit's first column (<= 0: all of them)
  uint32_t pp[8], dd[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 l4 = *reinterpret_cast<const float4*>(st_lse + 16 * u + 4 * g);
    const float4 d4 = *reinterpret_cast<const float4*>(st_dl + 16 * u + 4 * g);
    const float l[4] = {l4.x, l4.y, l4.z, l4.w};
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    float pr[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * g + e;
      float x = ex2_approx(fmaf(__uint_as_float(s[i]), sc, -l[e]));
      if (MASK) x = (key_ok && i >= qmin) ? x : 0.f;
      pr[e] = x;
      ds[e] = x * fmaf(__uint_as_float(dp[i]), scale, -d[e]);
    }
    pp[2 * g] = pack2_bf16(pr[0], pr[1]); pp[2 * g + 1] = pack2_bf16(pr[2], pr[3]);
    dd[2 * g] = pack2_bf16(ds[0], ds[1]); dd[2 * g + 1] = pack2_bf16(ds[2], ds[3]);
  }
  *reinterpret_cast<uint4*>(pbuf + sw128(row, 2 * u)) = make_uint4(pp[0], pp[1], pp[2], pp[3]);
  *reinterpret_cast<uint4*>(pbuf + sw128(row, 2 * u + 1)) = make_uint4(pp[4], pp[5], pp[6], pp[7]);
  *reinterpret_cast<uint4*>(dsbuf + sw128(row, 2 * u)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
  *reinterpret_cast<uint4*>(dsbuf + sw128(row, 2 * u + 1)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
}

__device__ __forceinline__ void bw_store32(__nv_bfloat16* dst, const uint32_t (&v)[32]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int g = 0; g < 4; ++g)
    d4[g] = make_uint4(pack2_bf16(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1])),
                       pack2_bf16(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                       pack2_bf16(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                       pack2_bf16(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
}

// FOLD (opt-in diagnostic, UNIIR_ATTN_BWD_FOLD=1; run on a B200 in round 2: correct on every shape of
// scripts/check_attn_bwd.py but 2.2x SLOWER than the default <8, false> + prep kernel -- 849 vs 382 us on the
// ViT-L/14 shape -- so it is not the default): two more warps compute the tail key column and the tail query
// row (257 = 2 x 128 + 1) from the operands already in shared memory, in registers, instead of the prep kernel's
// second pass over Q / K / V / dO in HBM.  Only tail_k <= 1 and tail_q <= 1.
template <int EW, bool FOLD = false>
__global__ void __launch_bounds__(bw_threads(EW, FOLD), 1) attn_bwd_tc_kernel(const __grid_constant__ AttnBwdTcParams p) {
  constexpr int kBwEwThreads = 32 * EW;
  constexpr int kUnits = 16 / EW;            // 16-query units per thread and chunk (4 units x 4 lane quarters / EW warps)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_al = align_smem_1024(smem_raw);
  if (smem_al - smem_raw > kBwSmemSlack) __trap();
  const BwSmem sm = bw_carve(smem_al);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_heads = p.B * p.H;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmQ64); tma_prefetch_desc(&p.tmQ16); tma_prefetch_desc(&p.tmDO64);
    tma_prefetch_desc(&p.tmDO16); tma_prefetch_desc(&p.tmK); tma_prefetch_desc(&p.tmV);
    for (int i = 0; i < kBwStages; ++i) { mbar_init(&sm.st_full[i], 1); mbar_init(&sm.st_empty[i], FOLD ? 3 : 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.kv_full[i], 1); mbar_init(&sm.kv_empty[i], FOLD ? 3 : 1);
      mbar_init(&sm.tail_done[i], 2);
      mbar_init(&sm.s_full[i], 1); mbar_init(&sm.chunk_done[i], kBwEwThreads); mbar_init(&sm.bufs_free[i], 1);
    }
    mbar_init(sm.akv_full, 1); mbar_init(sm.akv_free, kBwEpThreads);
    mbar_init(sm.dq_full, 1); mbar_init(sm.dq_free, kBwEpThreads);
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
      int g = 0, kvt = 0;
      for (int head = blockIdx.x; head < n_heads; head += gridDim.x) {
        const int b = head / p.H, h = head % p.H;
        const float* lse2 = p.lse2 + static_cast<long long>(head) * p.sq_pad;
        const float* dls = p.delta_s + static_cast<long long>(head) * p.sq_pad;
        for (int j = 0; j < p.n_kt; ++j, ++kvt) {
          const int slot = kvt & 1;
          if (kvt >= 2) mbar_wait(&sm.kv_empty[slot], ((kvt >> 1) - 1) & 1);
          mbar_arrive_expect_tx(&sm.kv_full[slot], 32768u);
          tma_load_2d(sm.k0 + slot * 16384, &p.tmK, &sm.kv_full[slot], h * 64, b * p.Skv + j * 128);
          tma_load_2d(sm.v0 + slot * 16384, &p.tmV, &sm.kv_full[slot], h * 64, b * p.Skv + j * 128);
          for (int i = 0; i < p.n_chunks; ++i, ++g) {
            const int st = g % kBwStages, use = g / kBwStages;
            if (use >= 1) mbar_wait(&sm.st_empty[st], (use - 1) & 1);
            const int nq = min(64, p.sq_pad - 64 * i);
            mbar_arrive_expect_tx(&sm.st_full[st], static_cast<uint32_t>(nq) * (256u + 8u));
            uint8_t* sq = sm.q0 + st * 8192;
            uint8_t* sdo = sm.do0 + st * 8192;
            const int row0 = b * p.Sq + 64 * i;
            if (nq == 64) {
              tma_load_2d(sq, &p.tmQ64, &sm.st_full[st], h * 64, row0);
              tma_load_2d(sdo, &p.tmDO64, &sm.st_full[st], h * 64, row0);
            } else {
              for (int r = 0; r < nq; r += 16) {
                tma_load_2d(sq + r * 128, &p.tmQ16, &sm.st_full[st], h * 64, row0 + r);
                tma_load_2d(sdo + r * 128, &p.tmDO16, &sm.st_full[st], h * 64, row0 + r);
              }
            }
            bulk_load_1d(sm.stats + st * 128, lse2 + 64 * i, static_cast<uint32_t>(nq) * 4u, &sm.st_full[st]);
            bulk_load_1d(sm.stats + st * 128 + 64, dls + 64 * i, static_cast<uint32_t>(nq) * 4u, &sm.st_full[st]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      const uint32_t idesc_acc = make_idesc_f16(true, false, true, 128, 64);   // dV / dK: A K-major, B MN-major
      const uint32_t idesc_dq = make_idesc_f16(true, true, true, 128, 64);     // dQ: A MN-major, B MN-major
      // Chunk g's S^T / dP^T MMAs are issued as soon as the elementwise threads have released TMEM buffer g & 1
      // (chunk g-2 done), AHEAD of chunk g-2's dV / dK / dQ MMAs, so that they are complete long before the
      // elementwise threads finish chunk g-1 (ncu on the first version, which issued them after: 27 % of the
      // elementwise warps' samples waiting on s_full).  Two chunks are therefore pending at any time.
      struct Pending { int valid, g, i, j, nq, st, slot, kvt, head_it, pairctr; };
      Pending pdA{}, pdB{};      // chunk g-2, chunk g-1
      // dV / dK (/ dQ) MMAs of chunk c; its chunk_done phase has been observed unless `flush`
      auto back = [&](const Pending& c, bool flush) {
        const int pb = c.g & 1, ps = c.pairctr & 1, hb = c.i & 1;
        const bool last = c.i == p.n_chunks - 1;
        if (flush) mbar_wait(&sm.chunk_done[pb], (c.g >> 1) & 1);
        if (c.i == 0 && c.kvt >= 1) mbar_wait(sm.akv_free, (c.kvt - 1) & 1);
        tc_fence_after();
        const uint32_t sP = smem_u32(sm.p0 + pb * 16384), sDS = smem_u32(sm.ds0 + (2 * ps + hb) * 16384);
        const uint32_t sQ = smem_u32(sm.q0 + c.st * 8192), sDO = smem_u32(sm.do0 + c.st * 8192);
        const int ks = c.nq >> 4;
        for (int k = 0; k < ks; ++k)
          umma_f16(tmem_base + kColDV, make_smem_desc(sP + k * 32, 16, 1024),
                   make_smem_desc(sDO + k * 2048, 8192, 1024), idesc_acc, (c.i > 0 || k > 0) ? 1u : 0u);
        for (int k = 0; k < ks; ++k)
          umma_f16(tmem_base + kColDK, make_smem_desc(sDS + k * 32, 16, 1024),
                   make_smem_desc(sQ + k * 2048, 8192, 1024), idesc_acc, (c.i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&sm.st_empty[c.st]);
        if (hb == 1 || last) {
          const int T = c.i >> 1;
          if (T < p.n_qt) {
            if (c.j == 0 && T == 0 && c.head_it >= 1) {
              mbar_wait(sm.dq_free, (c.head_it - 1) & 1);
              tc_fence_after();
            }
            const uint32_t sPair = smem_u32(sm.ds0 + 2 * ps * 16384), sK = smem_u32(sm.k0 + c.slot * 16384);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(tmem_base + kColDQ + 64 * T, make_smem_desc(sPair + k * 2048, 16384, 1024),
                       make_smem_desc(sK + k * 2048, 8192, 1024), idesc_dq, (c.j > 0 || k > 0) ? 1u : 0u);
          }
        }
        // every MMA issued so far has retired when this fires: P^T buffer g & 1 and the dS^T buffers of the
        // pair before last (whose dQ MMAs were issued at least one chunk ago) may be rewritten
        umma_commit(&sm.bufs_free[pb]);
        if (last) {
          umma_commit(sm.akv_full);
          umma_commit(&sm.kv_empty[c.slot]);
          if (c.j == p.n_kt - 1) umma_commit(sm.dq_full);
        }
      };
      int g = 0, kvt = 0, head_it = 0, pairctr = 0;
      for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++head_it) {
        for (int j = 0; j < p.n_kt; ++j, ++kvt) {
          const int slot = kvt & 1;
          // single-chunk tiles: the last chunk of tile kvt-2 is still pending, and the producer refills this
          // K/V slot only after that chunk's MMAs are committed
          if (p.n_chunks == 1 && pdA.valid) { back(pdA, true); pdA.valid = 0; }
          mbar_wait(&sm.kv_full[slot], (kvt >> 1) & 1);
          const uint32_t sK = smem_u32(sm.k0 + slot * 16384), sV = smem_u32(sm.v0 + slot * 16384);
          for (int i = 0; i < p.n_chunks; ++i, ++g) {
            const int st = g % kBwStages, tb = g & 1;
            const int nq = min(64, p.sq_pad - 64 * i);
            mbar_wait(&sm.st_full[st], (g / kBwStages) & 1);
            if (g >= 2) mbar_wait(&sm.chunk_done[tb], ((g >> 1) - 1) & 1);   // chunk g-2: TMEM buffer free, P^T / dS^T staged
            tc_fence_after();
            const uint32_t sQ = smem_u32(sm.q0 + st * 8192), sDO = smem_u32(sm.do0 + st * 8192);
            const uint32_t idesc_s = make_idesc_f16(true, false, false, 128, static_cast<uint32_t>(nq));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + kColS + 64 * tb, make_smem_desc(sK + k * 32, 16, 1024),
                       make_smem_desc(sQ + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_base + kColDP + 64 * tb, make_smem_desc(sV + k * 32, 16, 1024),
                       make_smem_desc(sDO + k * 32, 16, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&sm.s_full[tb]);
            if (pdA.valid) back(pdA, false);
            pdA = pdB;
            pdB.valid = 1; pdB.g = g; pdB.i = i; pdB.j = j; pdB.nq = nq; pdB.st = st; pdB.slot = slot; pdB.kvt = kvt;
            pdB.head_it = head_it; pdB.pairctr = pairctr;
            if ((i & 1) || i == p.n_chunks - 1) ++pairctr;
          }
        }
      }
      if (pdA.valid) back(pdA, true);
      if (pdB.valid) back(pdB, true);
    }
  } else if (warp < 2 + EW) {
    // =============================== elementwise ===============================
    const int ew = warp - 2;                 // 0..EW-1
    const int quarter = warp & 3;            // TMEM lanes [32 quarter, +32)
    const int part = ew >> 2;                // which 64 / (EW / 4) of the chunk's 64 query columns
    const int rit = quarter * 32 + lane;     // key row in the tile
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float sc = p.scale * kBwLog2e;
    int g = 0, pairctr = 0;
    for (int head = blockIdx.x; head < n_heads; head += gridDim.x) {
      const int b = head / p.H;
      const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
      for (int j = 0; j < p.n_kt; ++j) {
        const int key = j * 128 + rit;
        const bool key_ok = key < kvl;
        const int wkey0 = j * 128 + quarter * 32;          // first key of this warp
        for (int i = 0; i < p.n_chunks; ++i, ++g) {
          const int st = g % kBwStages, tb = g & 1, pb = g & 1, ps = pairctr & 1, hb = i & 1;
          const int nq = min(64, p.sq_pad - 64 * i);
          const bool last = i == p.n_chunks - 1;
          mbar_wait(&sm.st_full[st], (g / kBwStages) & 1);        // lse2 / delta_s of the chunk
          mbar_wait(&sm.s_full[tb], (g >> 1) & 1);
          tc_fence_after();
          // P^T buffer g & 1 (chunk g-2) and the dS^T buffers of the pair before last are free once chunk g-2's
          // MMAs have retired (tcgen05.commit covers every MMA issued before it)
          if (g >= 2) mbar_wait(&sm.bufs_free[pb], ((g >> 1) - 1) & 1);
          const float* st_lse = sm.stats + st * 128;
          const float* st_dl = st_lse + 64;
          uint8_t* pbuf = sm.p0 + pb * 16384;
          uint8_t* dsbuf = sm.ds0 + (2 * ps + hb) * 16384;
          // masks only where they can bite (warp-uniform): padded keys in this warp, or (causal) some key of the
          // warp beyond some query of the chunk
          const bool masked = (wkey0 + 31 >= kvl) || (p.causal && (wkey0 + 31 > 64 * i));
          {
            // this thread's 16-query units; all of its TMEM loads in flight before the first exponential
            uint32_t s[kUnits][16], dp[kUnits][16];
#pragma unroll
            for (int uu = 0; uu < kUnits; ++uu) {
              const int u = kUnits * part + uu;
              if (16 * u < nq) {
                tmem_ld_32x32b_x16(tlane + kColS + 64 * tb + 16 * u, s[uu]);
                tmem_ld_32x32b_x16(tlane + kColDP + 64 * tb + 16 * u, dp[uu]);
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int uu = 0; uu < kUnits; ++uu) {
              const int u = kUnits * part + uu;
              if (16 * u < nq) {
                // causal: key may attend to queries >= key, i.e. unit columns >= key - (64 i + 16 u)
                const int qm = p.causal ? key - (64 * i + 16 * u) : 0;
                if (masked) bw_unit<true>(s[uu], dp[uu], st_lse, st_dl, u, sc, p.scale, key_ok, qm, pbuf, dsbuf, rit);
                else bw_unit<false>(s[uu], dp[uu], st_lse, st_dl, u, sc, p.scale, key_ok, qm, pbuf, dsbuf, rit);
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();             // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(&sm.chunk_done[tb]);
          if (hb == 1 || last) ++pairctr;
        }
      }
    }
  } else if (warp < 2 + EW + 4) {
    // =============================== epilogue (4 warps) ===============================
    // dK_j / dV_j after each key tile, dQ after each head; off the elementwise warps' critical path: the MMA
    // issuer needs the accumulators back only one (dK/dV) or two (dQ) chunks into the next tile / head.
    const int quarter = warp & 3;
    const int rit = quarter * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int kvt = 0, head_it = 0;
    for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++head_it) {
      const int b = head / p.H, h = head % p.H;
      for (int j = 0; j < p.n_kt; ++j, ++kvt) {
        const int key = j * 128 + rit;
        mbar_wait(sm.akv_full, kvt & 1);
        tc_fence_after();
        const long long off = (static_cast<long long>(b) * p.Skv + key) * p.lddkv + h * 64;
#pragma unroll
        for (int part = 0; part < 4; ++part) {     // dK[0:32], dK[32:64], dV[0:32], dV[32:64]: one TMEM load at a time
          uint32_t a[32];
          tmem_ld_32x32b_x32(tlane + kColDK + 32 * part, a);      // kColDV == kColDK + 64
          tmem_ld_wait();
          if (key < p.Skv) bw_store32((part < 2 ? p.dk : p.dv) + off + 32 * (part & 1), a);
        }
        tc_fence_before();
        mbar_arrive(sm.akv_free);
      }
      // ---- dQ
      // tail key columns: this row's dS (prep kernel), requested before the accumulator is complete
      float dsc[kBwMaxQTiles][kBwTailMax];
#pragma unroll
      for (int T = 0; T < kBwMaxQTiles; ++T)
#pragma unroll
        for (int t = 0; t < kBwTailMax; ++t)
          dsc[T][t] = (!FOLD && T < p.n_qt && t < p.tail_k && T * 128 + rit < p.Sq)
                          ? __ldg(p.ds_col + (static_cast<long long>(head) * p.tail_k + t) * p.sq_pad + T * 128 + rit)
                          : 0.f;
      if constexpr (FOLD) {
        if (p.tail_k > 0) {      // the tail warps wrote this head's column into the CTA's scratch (plain loads: same SM)
          mbar_wait(&sm.tail_done[head_it & 1], (head_it >> 1) & 1);
          const float* col = p.fold_ws + (static_cast<long long>(blockIdx.x) * 2 + (head_it & 1)) * p.sq_pad;
#pragma unroll
          for (int T = 0; T < kBwMaxQTiles; ++T)
            if (T < p.n_qt && T * 128 + rit < p.Sq) dsc[T][0] = *reinterpret_cast<const volatile float*>(col + T * 128 + rit);
        }
      }
      mbar_wait(sm.dq_full, head_it & 1);
      tc_fence_after();
#pragma unroll
      for (int T = 0; T < kBwMaxQTiles; ++T) {
        if (T >= p.n_qt) break;
        const int q = T * 128 + rit;
        const bool q_ok = q < p.Sq;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t a[32];
          tmem_ld_32x32b_x32(tlane + kColDQ + 64 * T + 32 * hf, a);
          tmem_ld_wait();
          if (q_ok) {
            for (int t = 0; t < p.tail_k; ++t) {   // + dS[q, tail key] * K[tail key]
              const float ds = dsc[T][t];
              const uint4* kr = reinterpret_cast<const uint4*>(
                  p.k + (static_cast<long long>(b) * p.Skv + p.n_kt * 128 + t) * p.ldkv + h * 64 + 32 * hf);
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                const uint4 kk = __ldg(kr + gq);
                const uint32_t w[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  a[8 * gq + 2 * e] = __float_as_uint(fmaf(ds, bf16_lo(w[e]), __uint_as_float(a[8 * gq + 2 * e])));
                  a[8 * gq + 2 * e + 1] = __float_as_uint(fmaf(ds, bf16_hi(w[e]), __uint_as_float(a[8 * gq + 2 * e + 1])));
                }
              }
            }
            bw_store32(p.dq + (static_cast<long long>(b) * p.Sq + q) * p.lddq + h * 64 + 32 * hf, a);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(sm.dq_free);
    }
  }

  else if constexpr (FOLD) {
    // =============================== tail warps (2) ===============================
    const int tw = warp - (2 + EW + 4);                  // 0 / 1
    const float sc = p.scale * kBwLog2e;
    const int kt = p.n_kt * 128, qt = p.n_qt * 128;      // the tail key / the tail query (tail_k, tail_q <= 1)
    float* ws_cta = p.fold_ws + static_cast<long long>(gridDim.x) * 2 * p.sq_pad + static_cast<long long>(blockIdx.x) * 192;
    int g = 0, kvt = 0, head_it = 0;
    for (int head = blockIdx.x; head < n_heads; head += gridDim.x, ++head_it) {
      const int b = head / p.H, h = head % p.H;
      const int kvl = p.kv_len ? min(p.kv_len[b], p.Skv) : p.Skv;
      // lane l owns head dims 2l, 2l+1 of the tail key's K / V rows and of the tail query's Q / dO rows
      float2 Kt = make_float2(0.f, 0.f), Vt = Kt, Qt = Kt, dOt = Kt;
      float l2t = INFINITY, dlt = 0.f;
      if (p.tail_k > 0) {
        Kt = bw_ld_bf2(p.k + (static_cast<long long>(b) * p.Skv + kt) * p.ldkv + h * 64 + 2 * lane);
        Vt = bw_ld_bf2(p.v_g + (static_cast<long long>(b) * p.Skv + kt) * p.ldkv + h * 64 + 2 * lane);
      }
      if (p.tail_q > 0) {
        Qt = bw_ld_bf2(p.q_g + (static_cast<long long>(b) * p.Sq + qt) * p.ldq_g + h * 64 + 2 * lane);
        dOt = bw_ld_bf2(p.do_g + (static_cast<long long>(b) * p.Sq + qt) * p.lddo_g + h * 64 + 2 * lane);
        l2t = p.lse2[static_cast<long long>(head) * p.sq_pad + qt];
        dlt = p.delta_s[static_cast<long long>(head) * p.sq_pad + qt];
      }
      const bool kt_ok = kt < kvl;
      float2 accV = make_float2(0.f, 0.f), accK = accV, accQ = accV;
      float* col = p.fold_ws + (static_cast<long long>(blockIdx.x) * 2 + (head_it & 1)) * p.sq_pad;
      for (int j = 0; j < p.n_kt; ++j, ++kvt) {
        const int slot = kvt & 1;
        for (int i = 0; i < p.n_chunks; ++i, ++g) {
          const int st = g % kBwStages;
          const int nq = min(64, p.sq_pad - 64 * i);
          mbar_wait(&sm.st_full[st], (g / kBwStages) & 1);
          if (j == 0 && p.tail_k > 0) {
            // tail key column against this chunk's queries: this warp's 32 rows, one at a time, all lanes on one row
            const uint8_t* sq = sm.q0 + st * 8192;
            const uint8_t* sdo = sm.do0 + st * 8192;
            const float* st_lse = sm.stats + st * 128;
            const float* st_dl = st_lse + 64;
            float my_ds = 0.f;
            for (int rr = 0; rr < 32; ++rr) {
              const int r = 32 * tw + rr;
              if (r >= nq) break;                                    // warp-uniform
              const uint32_t off = sw128(r, lane >> 2) + (lane & 3) * 4;
              const float2 qv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sq + off));
              const float2 dv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sdo + off));
              const float s = bw_warp_sum(qv.x * Kt.x + qv.y * Kt.y), dp = bw_warp_sum(dv.x * Vt.x + dv.y * Vt.y);
              const int q = 64 * i + r;
              const float pr = (kt_ok && (!p.causal || kt <= q)) ? ex2_approx(fmaf(s, sc, -st_lse[r])) : 0.f;
              const float ds = pr * fmaf(dp, p.scale, -st_dl[r]);
              accV.x = fmaf(pr, dv.x, accV.x); accV.y = fmaf(pr, dv.y, accV.y);
              accK.x = fmaf(ds, qv.x, accK.x); accK.y = fmaf(ds, qv.y, accK.y);
              if (lane == rr) my_ds = ds;
            }
            if (32 * tw + lane < nq) col[64 * i + 32 * tw + lane] = my_ds;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.st_empty[st]);
        }
        if (p.tail_q > 0) {
          // tail query row against this tile's keys: this warp's 64 keys
          mbar_wait(&sm.kv_full[slot], (kvt >> 1) & 1);
          const uint8_t* sk = sm.k0 + slot * 16384;
          const uint8_t* sv = sm.v0 + slot * 16384;
          const int klim = p.causal ? min(kvl, qt + 1) : kvl;
          for (int kk = 0; kk < 64; ++kk) {
            const int r = 64 * tw + kk, key = 128 * j + r;
            if (key >= klim) break;                                  // warp-uniform (keys ascend)
            const uint32_t off = sw128(r, lane >> 2) + (lane & 3) * 4;
            const float2 kv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sk + off));
            const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sv + off));
            const float s = bw_warp_sum(kv.x * Qt.x + kv.y * Qt.y), dp = bw_warp_sum(vv.x * dOt.x + vv.y * dOt.y);
            const float ds = ex2_approx(fmaf(s, sc, -l2t)) * fmaf(dp, p.scale, -dlt);
            accQ.x = fmaf(ds, kv.x, accQ.x); accQ.y = fmaf(ds, kv.y, accQ.y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.kv_empty[slot]);
        if (j == 0 && lane == 0) mbar_arrive(&sm.tail_done[head_it & 1]);   // release: the column is published
      }
      // corner (tail query x tail key) for dQ[qt]: recomputed by warp 0
      if (tw == 0 && p.tail_q > 0 && p.tail_k > 0) {
        const float s = bw_warp_sum(Qt.x * Kt.x + Qt.y * Kt.y), dp = bw_warp_sum(dOt.x * Vt.x + dOt.y * Vt.y);
        const float pr = (kt_ok && (!p.causal || kt <= qt)) ? ex2_approx(fmaf(s, sc, -l2t)) : 0.f;
        const float ds = pr * fmaf(dp, p.scale, -dlt);
        accQ.x = fmaf(ds, Kt.x, accQ.x); accQ.y = fmaf(ds, Kt.y, accQ.y);
      }
      // the two warps' partial sums meet in the CTA's global scratch; warp 0 stores the three 64-wide rows
      if (tw == 1) {
        ws_cta[2 * lane] = accV.x; ws_cta[2 * lane + 1] = accV.y;
        ws_cta[64 + 2 * lane] = accK.x; ws_cta[64 + 2 * lane + 1] = accK.y;
        ws_cta[128 + 2 * lane] = accQ.x; ws_cta[128 + 2 * lane + 1] = accQ.y;
      }
      named_bar_sync(3, 64);
      if (tw == 0) {
        const volatile float* wv = ws_cta;
        if (p.tail_k > 0) {
          const long long o = (static_cast<long long>(b) * p.Skv + kt) * p.lddkv + h * 64 + 2 * lane;
          *reinterpret_cast<uint32_t*>(p.dv + o) = pack2_bf16(accV.x + wv[2 * lane], accV.y + wv[2 * lane + 1]);
          *reinterpret_cast<uint32_t*>(p.dk + o) = pack2_bf16(accK.x + wv[64 + 2 * lane], accK.y + wv[64 + 2 * lane + 1]);
        }
        if (p.tail_q > 0)
          *reinterpret_cast<uint32_t*>(p.dq + (static_cast<long long>(b) * p.Sq + qt) * p.lddq + h * 64 + 2 * lane) =
              pack2_bf16(accQ.x + wv[128 + 2 * lane], accQ.y + wv[128 + 2 * lane + 1]);
      }
      named_bar_sync(3, 64);            // the scratch is rewritten by warp 1 for the next head
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Returns UNIIR_OK after launching both kernels, a negative error, or +1 when the tcgen05 path does not
// apply (the caller falls through to the mma.sync kernels).  ws: attn_bwd_ws_floats(...) floats.
long long attn_bwd_ws_floats(int B, int Sq, int Skv, int H) {
  const long long sq_pad = (Sq + 15) / 16 * 16;
  const int rem = Skv % 128;
  const int tail_k = (Skv > 128 && rem > 0 && rem <= kBwTailMax) ? rem : 0;
  // + the per-CTA scratch of the FOLD variant (2 x sq_pad + 192 floats per SM; small, always included)
  return static_cast<long long>(B) * H * sq_pad * (2 + tail_k) + static_cast<long long>(num_sms()) * (2 * sq_pad + 192);
}

int attn_bwd_tc_try(const __nv_bfloat16* q, long long ldq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                    long long ldkv, const __nv_bfloat16* out, long long ldo, const __nv_bfloat16* dout,
                    long long lddo, const float* lse, int B, int Sq, int Skv, int H, int causal, const int* kv_len,
                    float* ws, __nv_bfloat16* dq, long long lddq, __nv_bfloat16* dk, __nv_bfloat16* dv,
                    long long lddkv, float scale, void* stream) {
  if (!(attn_tc_mode_bits() & 2)) return 1;
  auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  if (lddq % 8 != 0 || lddkv % 8 != 0 || ldkv % 8 != 0 || ldo % 8 != 0 || !al16(dq) || !al16(dk) || !al16(dv) ||
      !al16(k) || !al16(out) || !al16(ws))
    return 1;
  AttnBwdPrepParams pp{};
  AttnBwdTcParams p{};
  const int sq_pad = (Sq + 15) / 16 * 16;
  int n_qt, tail_q, n_kt, tail_k;
  {
    const int rq = Sq % 128, rk = Skv % 128;
    if (Sq > 128 && rq > 0 && rq <= kBwTailMax) { n_qt = Sq / 128; tail_q = rq; } else { n_qt = (Sq + 127) / 128; tail_q = 0; }
    if (Skv > 128 && rk > 0 && rk <= kBwTailMax) { n_kt = Skv / 128; tail_k = rk; } else { n_kt = (Skv + 127) / 128; tail_k = 0; }
  }
  if (n_qt > kBwMaxQTiles) return 1;
  if ((tail_k > 0 || tail_q > 0) && (Skv > kPrepThreads || Sq > kPrepThreads)) return 1;   // prep: a thread per row
  const long long bh = static_cast<long long>(B) * H;
  float* lse2 = ws;
  float* delta_s = ws + bh * sq_pad;
  float* ds_col = ws + 2 * bh * sq_pad;
  float* fold_ws = ds_col + bh * tail_k * sq_pad;
  // opt-in: tail column / row inside the main kernel (see the FOLD note above the kernel); one tail at most
  static int fold_env = -1;
  if (fold_env < 0) {
    const char* e = getenv("UNIIR_ATTN_BWD_FOLD");
    fold_env = (e && atoi(e) != 0);
  }
  const bool fold = fold_env && (tail_k > 0 || tail_q > 0) && tail_k <= 1 && tail_q <= 1;

  pp.q = q; pp.k = k; pp.v = v; pp.out = out; pp.dout = dout;
  pp.ldq = ldq; pp.ldkv = ldkv; pp.ldo = ldo; pp.lddo = lddo;
  pp.lse = lse; pp.kv_len = kv_len; pp.lse2 = lse2; pp.delta_s = delta_s; pp.ds_col = ds_col;
  pp.dq = dq; pp.dk = dk; pp.dv = dv; pp.lddq = lddq; pp.lddkv = lddkv;
  pp.B = B; pp.H = H; pp.Sq = Sq; pp.Skv = Skv; pp.causal = causal; pp.sq_pad = sq_pad;
  pp.n_kt = n_kt; pp.tail_k = fold ? 0 : tail_k; pp.n_qt = n_qt; pp.tail_q = fold ? 0 : tail_q; pp.scale = scale;
  // (fold: the prep kernel only writes delta and the base-2 LSE)

  p.lse2 = lse2; p.delta_s = delta_s; p.ds_col = ds_col; p.k = k; p.ldkv = ldkv;
  p.fold_ws = fold_ws; p.v_g = v; p.q_g = q; p.do_g = dout; p.ldq_g = ldq; p.lddo_g = lddo;
  p.dq = dq; p.dk = dk; p.dv = dv; p.lddq = lddq; p.lddkv = lddkv; p.kv_len = kv_len;
  p.B = B; p.H = H; p.Sq = Sq; p.Skv = Skv; p.causal = causal; p.sq_pad = sq_pad;
  p.n_chunks = (sq_pad + 63) / 64; p.n_kt = n_kt; p.tail_k = tail_k; p.n_qt = n_qt; p.tail_q = tail_q;
  p.scale = scale;
  int rc;
  const uint64_t cols = static_cast<uint64_t>(H) * 64;
  const uint64_t rows_q = static_cast<uint64_t>(B) * Sq, rows_k = static_cast<uint64_t>(B) * Skv;
  if ((rc = make_tmap_2d(&p.tmQ64, q, UNIIR_BF16, cols, rows_q, ldq, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&p.tmQ16, q, UNIIR_BF16, cols, rows_q, ldq, 64, 16))) return rc;
  if ((rc = make_tmap_2d(&p.tmDO64, dout, UNIIR_BF16, cols, rows_q, lddo, 64, 64))) return rc;
  if ((rc = make_tmap_2d(&p.tmDO16, dout, UNIIR_BF16, cols, rows_q, lddo, 64, 16))) return rc;
  if ((rc = make_tmap_2d(&p.tmK, k, UNIIR_BF16, cols, rows_k, ldkv, 64, 128))) return rc;
  if ((rc = make_tmap_2d(&p.tmV, v, UNIIR_BF16, cols, rows_k, ldkv, 64, 128))) return rc;
  // 8 elementwise warps (two 16-query units per thread) by default; UNIIR_ATTN_BWD_EW=16 selects the 16-warp
  // variant (one unit per thread), measured 3 % slower on B200 (ViT-L/14 shape: 390 us vs 379 us with prep)
  static int ew16 = -1;
  if (ew16 < 0) {
    const char* e = getenv("UNIIR_ATTN_BWD_EW");
    ew16 = (e && atoi(e) == 16);
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmemBytes));
    UNIIR_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmemBytes));
    UNIIR_CUDA_OK((cudaFuncSetAttribute(attn_bwd_tc_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmemBytes)));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  attn_bwd_prep_kernel<<<dim3(H, B), kPrepThreads, 0, st>>>(pp);
  const int heads = B * H;
  const int grid = heads < num_sms() ? heads : num_sms();
  if (fold) attn_bwd_tc_kernel<8, true><<<grid, bw_threads(8, true), kBwSmemBytes, st>>>(p);
  else if (ew16) attn_bwd_tc_kernel<16><<<grid, bw_threads(16), kBwSmemBytes, st>>>(p);
  else attn_bwd_tc_kernel<8><<<grid, bw_threads(8), kBwSmemBytes, st>>>(p);
  count_launch(2);
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

}  // namespace uniir
