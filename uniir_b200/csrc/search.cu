// Dense retrieval: fused fp16 tcgen05 GEMM + per-row top-k (replaces FAISS GpuIndexFlat
// inner-product search, reference common/mbeir_retriever.py:231) plus the row-norm and
// k-way merge kernels around it.
//
// Work decomposition: item = (query tile of 128 rows) x (one CHUNK of ~16 k candidates).
// Items are ordered chunk-major, so at any moment all 148 CTAs work on different query tiles
// of the SAME ~25 MB candidate chunk: it is fetched from HBM once and served to everyone else
// from L2.  (A first version gave each CTA a whole candidate range; ncu showed 10.9 TB of DRAM
// reads for the 190k x 5.6M problem -- CTAs drift apart after the first wave and every query
// tile re-streamed its range from HBM.)  The per-row running top-k (threshold, count, buffer)
// lives in global memory and is handed from the CTA that processed (tile, chunk c) to the one
// that processes (tile, chunk c+1) through a release/acquire flag per query tile; the static
// chunk-major order guarantees the predecessor is always scheduled earlier, so the wait cannot
// deadlock on a persistent grid.  Small query sets additionally split the candidates into
// independent ranges (own state, merged afterwards) to keep every SM busy.  Scores never leave the SM: the epilogue reads the fp32 accumulators from
// TMEM, scales by 1/||c|| and rejects >99.9 % of them with one max-reduce + compare against
// the row's current k-th best; survivors are appended to a small per-row buffer that is
// compacted (warp-cooperative bitonic sort) when it fills.
#include <cuda_fp16.h>
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "gemm_sm100.cuh"
#include "host_util.cuh"

namespace uniir {

constexpr int kSearchBN = 256;
constexpr int kCap = 128;      // per-row candidate buffer (entries) of the streaming kernels
constexpr int kCapMax = 256;   // ... of the deferred-compaction experiment; the workspace is sized for it
constexpr int kMaxK = 64;
constexpr float kNegInf = -INFINITY;
constexpr float kFaissNeg = -3.4028234663852886e38f;

constexpr int kDefaultTilesPerChunk = 64;   // 64 x 256 candidates x 1536 B = 25 MB: L2-resident per wave
constexpr int kDefaultTilesPerChunkTs = 128; // 16 k candidates (25 MB) per chunk in 128-candidate tiles: L2-resident across the ~10 query rounds; same-box sweeps r2c9 / r2c10

struct SearchParams {
  long long Q, N;
  int k, splits, n_qtiles, tiles_per_split, n_tiles, kb_total, n_chunks;
  const float* q_inv;
  const float* c_inv;
  float* part_scores;
  int* part_idx;
  float* ws_scores;  // [splits][n_qtiles*128][kCap]   per-row candidate buffers
  int* ws_idx;
  float* g_thr;      // [splits][n_qtiles*128]         running k-th best per row
  int* g_cnt;        // [splits][n_qtiles*128]         entries in the row buffer
  int* done;         // [splits][n_qtiles]             chunks completed (release/acquire hand-off)
};

struct SearchSched {
  int n_units, splits, tiles_per_split, n_tiles, kb_total, n_chunks, tpc;   // unit = query tile (or tile pair); tpc = tiles per chunk
  int n_workers;   // CTAs (or CTA pairs) of the persistent grid; 0 = every worker walks a chunk in the same order
  __device__ __forceinline__ int num_items() const { return n_chunks * splits * n_units; }
  __device__ __forceinline__ bool get(int item, WorkItem& w) const {
    const int slots = splits * n_units;
    const int chunk = item / slots;
    const int slot = item - chunk * slots;        // (split, unit)
    const int split = slot / n_units;
    w.m_blk = slot - split * n_units;
    const int split_begin = split * tiles_per_split;
    const int split_end = min(n_tiles, split_begin + tiles_per_split);
    w.n_begin = split_begin + chunk * tpc;
    w.n_end = min(split_end, w.n_begin + tpc);
    w.kb_begin = 0;
    w.kb_end = kb_total;
    w.tag = slot;
    // aux: chunk index, bit 30 = last chunk of this split
    const int last_chunk = (split_end - split_begin + tpc - 1) / tpc - 1;
    w.aux = chunk | (chunk >= last_chunk ? (1 << 30) : 0);
    // Stagger the workers over the chunk: worker j starts at tile j * len / n_workers and wraps around.  All
    // workers sweep the same chunk at the same time; started at the SAME tile they would all miss on the same
    // L2 lines at once, and this L2 does not merge misses to a line that is already in flight -- measured with
    // one query tile pair per CTA pair and the queries L2-resident: 249 GB of DRAM reads for 8.6 GB of candidates
    // (profiles/search_l2_diagnosis_r02.txt).  Staggered, each tile is fetched by one worker and hit by the rest.
    const int len = w.n_end - w.n_begin;
    w.rot = (n_workers > 0 && len > 1) ? static_cast<int>((static_cast<long long>(item % n_workers) * len) / n_workers) : 0;
    return w.n_begin < w.n_end || chunk == 0;     // an empty split still emits its (padded) result
  }
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// strict total order "a ranks before b": higher score first, lower index on ties
__device__ __forceinline__ bool ranks_before(float ka, long long ia, float kb, long long ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

// Bitonic sort (descending by ranks_before) of 32*R elements held as element e = r*32 + lane.
template <int R, typename IdxT>
__device__ __forceinline__ void warp_bitonic_desc(float (&key)[R], IdxT (&idx)[R]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k2 = 2; k2 <= 32 * R; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & jr) == 0) {
            const int r2 = r | jr;
            const int e = r * 32 + lane;
            const bool desc = (e & k2) == 0;
            const bool first_wins = ranks_before(key[r], (long long)idx[r], key[r2], (long long)idx[r2]);
            if (first_wins != desc) {
              const float tk = key[r]; key[r] = key[r2]; key[r2] = tk;
              const IdxT ti = idx[r]; idx[r] = idx[r2]; idx[r2] = ti;
            }
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          const float ok = __shfl_xor_sync(0xffffffffu, key[r], j);
          const IdxT oi = __shfl_xor_sync(0xffffffffu, idx[r], j);
          const bool lower = (lane & j) == 0;
          const bool desc = (e & k2) == 0;
          const bool mine_first = ranks_before(key[r], (long long)idx[r], ok, (long long)oi);
          // the lower position keeps the element that ranks first iff this run is descending
          const bool keep_mine = (lower == desc) ? mine_first : !mine_first;
          const bool same = (key[r] == ok) && (idx[r] == oi);
          if (!keep_mine && !same) { key[r] = ok; idx[r] = oi; }
        }
      }
    }
  }
}

template <int BN, int CAP = kCap, bool kSelect = false>
struct SearchEpilogueT {
  const SearchParams& p;
  float* invc_s;  // [2][BN] in shared memory
  float thr;
  int cnt;
  int tcount;
  __device__ SearchEpilogueT(const SearchParams& p_, float* s) : p(p_), invc_s(s), thr(kNegInf), cnt(0), tcount(0) {}

  long long slot_row0;  // first state row of the current (split, query tile) slot

  __device__ __forceinline__ long long ws_base(int row) const { return (slot_row0 + row) * CAP; }

  // Pick up the running top-k state of this query tile from whoever processed the previous chunk.
  __device__ __forceinline__ void item_begin(const WorkItem& w, int row) {
    slot_row0 = static_cast<long long>(w.tag) * kBM;
    const int chunk = w.aux & 0x3fffffff;
    if (chunk == 0) {
      thr = kNegInf;
      cnt = 0;
      return;
    }
    if ((threadIdx.x & 31) == 0) {
      unsigned spins = 0;
      while (ld_acquire_gpu(p.done + w.tag) < chunk) {
        __nanosleep(200);
        if (++spins > (1u << 24)) {
          printf("uniir: retrieval hand-off timeout (slot %d chunk %d)\n", w.tag, chunk);
          __trap();
        }
      }
    }
    __syncwarp();
    thr = __ldcg(p.g_thr + slot_row0 + row);
    cnt = __ldcg(p.g_cnt + slot_row0 + row);
  }

  // Compact the buffers of the rows selected by `mask` (warp-uniform) to their top-k.
  // final != 0: also emit the sorted partial result for (query row, split).
  __device__ __forceinline__ void compact(uint32_t mask, int row, bool final, const WorkItem& w) {
    const int lane = threadIdx.x & 31;
    const int row0 = row - lane;
    __syncwarp();
    while (mask) {
      const int L = __ffs(mask) - 1;
      mask &= mask - 1;
      const long long base = ws_base(row0 + L);
      const int cntL = __shfl_sync(0xffffffffu, cnt, L);
      constexpr int R = CAP / 32;
      float key[R];
      int idx[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int e = r * 32 + lane;
        const bool ok = e < cntL;
        // L2 loads: earlier chunks of this row were written by other SMs (L1 is not coherent)
        key[r] = ok ? __ldcg(p.ws_scores + base + e) : kNegInf;
        idx[r] = ok ? __ldcg(p.ws_idx + base + e) : 0x7fffffff;
      }
      if constexpr (kSelect) {
        // Mid-scan compaction by SELECTION: the exact k-th best value by an MSB-first radix select over the
        // order-preserving integer image of the scores (32 rounds of R ballots), then every entry >= it is kept, in
        // any order (the final compaction sorts).  ~3x cheaper than the bitonic sort of the same buffer and
        // independent of its size, which is what makes a 256-entry buffer (5.7 instead of 13 compactions per row at
        // k = 50) pay.  Ties beyond k stay (a superset of the exact top-k); a buffer that is mostly ties falls back
        // to the sort so that it always shrinks.
        if (!final && cntL >= p.k) {
          uint32_t u[R];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const uint32_t b = __float_as_uint(key[r]);
            u[r] = (r * 32 + lane < cntL) ? ((b & 0x80000000u) ? ~b : (b | 0x80000000u)) : 0u;
          }
          uint32_t prefix = 0;
          int remaining = p.k;
#pragma unroll 1
          for (int bit = 31; bit >= 0; --bit) {
            const uint32_t trial = prefix | (1u << bit);
            const uint32_t hi = 0xffffffffu << bit;
            int c = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) c += __popc(__ballot_sync(0xffffffffu, (u[r] & hi) == trial));
            if (c >= remaining) prefix = trial; else remaining -= c;
          }
          int ge = 0;
#pragma unroll
          for (int r = 0; r < R; ++r) ge += __popc(__ballot_sync(0xffffffffu, u[r] >= prefix));
          if (ge <= CAP / 2) {
            int at = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const bool keep_it = u[r] >= prefix;
              const uint32_t m = __ballot_sync(0xffffffffu, keep_it);
              if (keep_it) {
                const int pos = at + __popc(m & ((1u << lane) - 1u));
                p.ws_scores[base + pos] = key[r];
                p.ws_idx[base + pos] = idx[r];
              }
              at += __popc(m);
            }
            if (lane == L) {
              cnt = ge;
              thr = __uint_as_float((prefix & 0x80000000u) ? (prefix ^ 0x80000000u) : ~prefix);
            }
            continue;
          }
        }
      }
      warp_bitonic_desc<R, int>(key, idx);
      const int keep = min(cntL, p.k);
#pragma unroll
      for (int r = 0; r < (kMaxK + 31) / 32; ++r) {
        const int e = r * 32 + lane;
        if (e < keep) {
          p.ws_scores[base + e] = key[r];
          p.ws_idx[base + e] = idx[r];
        }
      }
      // k-th best (element k-1) becomes the new admission threshold
      const int kr = (p.k - 1) >> 5, kl = (p.k - 1) & 31;
      float kth = kNegInf;
#pragma unroll
      for (int r = 0; r < (kMaxK + 31) / 32; ++r) {
        const float v = __shfl_sync(0xffffffffu, key[r], kl);
        if (r == kr) kth = v;
      }
      if (lane == L) {
        cnt = keep;
        thr = (cntL >= p.k) ? kth : kNegInf;
      }
      if (final) {
        const long long q = static_cast<long long>(w.m_blk) * kBM + row0 + L;
        if (q < p.Q) {
          const int split = w.tag / p.n_qtiles;
          const long long ob = (q * p.splits + split) * p.k;
          const float qi = p.q_inv[q];
#pragma unroll
          for (int r = 0; r < (kMaxK + 31) / 32; ++r) {
            const int e = r * 32 + lane;
            if (e < p.k) {
              p.part_scores[ob + e] = (e < keep) ? key[r] * qi : kFaissNeg;
              p.part_idx[ob + e] = (e < keep) ? idx[r] : -1;
            }
          }
        }
      }
    }
    __syncwarp();
  }

  // `staged`: the tile's BN inverse candidate norms already sit in shared memory (the caller prefetched them a
  // tile ahead); otherwise they are fetched here and published to the four epilogue warps with a named barrier.
  // kDefer (experiment, off): the row buffer has room for a whole tile (cnt <= CAP - BN on entry), nothing is
  // compacted inside the tile; the caller releases the accumulator first and then calls compact_deferred().
  template <bool kPipe = true, bool kDefer = false>
  __device__ __forceinline__ void tile(const WorkItem& w, int n_blk, uint32_t tmem_acc, int row, int /*col_half*/,
                                       const float* staged = nullptr) {
    const float* invc = staged;
    const long long col_base = static_cast<long long>(n_blk) * BN;
    if (staged == nullptr) {
      float* buf = invc_s + (tcount & 1) * BN;
      ++tcount;
#pragma unroll
      for (int i = row; i < BN; i += kBM) {
        const long long c = col_base + i;
        buf[i] = (c < p.N) ? p.c_inv[c] : 0.f;
      }
      named_bar_sync(1, 128);
      invc = buf;
    }
    const bool tail = col_base + BN > p.N;
    const long long wsb = ws_base(row);
    // one 32-column chunk of this row: scale, reject against the row's k-th best, keep the rare survivors
    auto chunk = [&](int c, const uint32_t (&r)[32]) {
      float s[32];
      float m = kNegInf;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 iv = *reinterpret_cast<const float4*>(invc + c * 32 + j4 * 4);
        s[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) * iv.x;
        s[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) * iv.y;
        s[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) * iv.z;
        s[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) * iv.w;
      }
      if (tail) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col_base + c * 32 + j >= p.N) s[j] = kNegInf;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) m = fmaxf(m, s[j]);
      if (m > thr) {
        // rare path (~1e-4 of the scores): locate the survivors by 8-wide group so that a
        // single hit costs 4 group tests + 8 predicated appends instead of 32
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float mg = s[g * 8];
#pragma unroll
          for (int j = 1; j < 8; ++j) mg = fmaxf(mg, s[g * 8 + j]);
          if (mg > thr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (s[g * 8 + j] > thr) {
                p.ws_scores[wsb + cnt] = s[g * 8 + j];
                p.ws_idx[wsb + cnt] = static_cast<int>(col_base + c * 32 + g * 8 + j);
                ++cnt;
              }
            }
          }
        }
      }
      if constexpr (!kDefer) {
        const uint32_t need = __ballot_sync(0xffffffffu, cnt > CAP - 32);
        if (need) compact(need, row, false, w);
      }
    };
    // two register buffers: the TMEM load of the next chunk is in flight while this one is processed
    static_assert((BN / 32) % 2 == 0, "chunks are processed in pairs");
    if constexpr (kPipe) {
      uint32_t ra[32], rb[32];
      tmem_ld_32x32b_x32(tmem_acc, ra);
      tmem_ld_wait();
#pragma unroll 1
      for (int c = 0; c < BN / 32; c += 2) {
        tmem_ld_32x32b_x32(tmem_acc + (c + 1) * 32, rb);
        chunk(c, ra);
        tmem_ld_wait();
        if (c + 2 < BN / 32) tmem_ld_32x32b_x32(tmem_acc + (c + 2) * 32, ra);
        chunk(c + 1, rb);
        tmem_ld_wait();
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_acc + c * 32, r);
        tmem_ld_wait();
        chunk(c, r);
      }
    }
  }

  __device__ __forceinline__ void compact_deferred(const WorkItem& w, int row) {
    const uint32_t need = __ballot_sync(0xffffffffu, cnt > CAP - BN);
    if (need) compact(need, row, false, w);
  }

  __device__ __forceinline__ void item_end(const WorkItem& w, int row) {
    if (w.aux & (1 << 30)) {           // last chunk of the split: final sort + emit the partial list
      compact(0xffffffffu, row, true, w);
      return;
    }
    // hand the state over to the CTA that takes (this tile, next chunk)
    p.g_thr[slot_row0 + row] = thr;
    p.g_cnt[slot_row0 + row] = cnt;
    __threadfence();
    named_bar_sync(2, 128);
    if (row == 0) st_release_gpu(p.done + w.tag, (w.aux & 0x3fffffff) + 1);
  }
};

using SearchEpilogue = SearchEpilogueT<kSearchBN>;

__global__ void __launch_bounds__(num_threads(4), 1)
ip_topk_kernel(const __grid_constant__ GemmTensorMaps maps, const SearchParams p,
               const SearchSched sched, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  using SM = GemmSmem<kSearchBN>;
  float* invc_s = reinterpret_cast<float*>(smem + SM::kRingBytes + SM::kBarBytes);
  SearchEpilogue epi(p, invc_s);
  gemm_body<kSearchBN, false, false, 4>(&maps.a, &maps.b, idesc, sched, epi, smem);
}

// CTA-pair variant: two CTAs (one cluster) take two adjacent query tiles and share every candidate
// tile through one 256-row cta_group::2 UMMA; each CTA stages half of the candidate tile.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads(4), 1)
ip_topk_kernel_2cta(const __grid_constant__ GemmTensorMaps maps, const SearchParams p,
                    const SearchSched sched, const uint32_t idesc, const int hint_a, const int hint_b) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  using SM = GemmSmem2<kSearchBN>;
  float* invc_s = reinterpret_cast<float*>(smem + SM::kRingBytes + SM::kBarBytes);
  SearchEpilogue epi(p, invc_s);
  gemm_body_2cta<kSearchBN, false, false, 4>(&maps.a, &maps.b, idesc, sched, epi, smem, hint_a, hint_b);
}


// ---------------------------------------------------------------------------------------------------------
// Queries resident on the SM (default).  ncu on the kernels above (profiles/search_l2_diagnosis_r02.txt): they
// stream BOTH operands from L2 for every 256 x 256 tile, 64 B/clk/SM at full tensor rate = 8.5 TB/s chip-wide
// measured -- the L2 -> SM throughput cap -- and 1.0 PFLOP/s whatever the SM clock.  Here a CTA pair keeps its 256
// query rows (128 per CTA) on chip for a whole work item and streams only the candidates:
//   * k-blocks 0-7 of the queries (K = 512) sit in TENSOR MEMORY -- 256 columns, two fp16 per 32-bit column,
//     lane = row -- and enter `tcgen05.mma` as the A operand from TMEM;
//   * k-blocks 8-11 (K = 256, 64 KB per CTA) sit in SHARED MEMORY, loaded once per item by TMA, and enter as an
//     ordinary shared-memory A operand;
//   * the other 256 TMEM columns are two 128-column accumulators: the tile is 256 queries x 128 candidates,
//     double-buffered against the epilogue.  (All 384 query columns in TMEM leave room for 64-candidate tiles
//     only; that version was correct but issue-bound: a 256 x 64 x 16 MMA lasts 32 clocks and the single issuing
//     thread needs ~40 for the descriptor arithmetic -- 52 % tensor pipe, ncu r2c5.)
// L2 -> SM bytes per FLOP halve (256 FLOP/B); the ring holds nothing but candidates (9 stages x 16 KB per CTA).
//   warp 0   : TMA producer (both CTAs): per item the shared-memory part of its queries, then its 64 rows of each
//              128-candidate tile, 2 k-blocks (16 KB) per stage
//   warp 1   : MMA issuer (leader CTA): per k-block 4 x tcgen05.mma.cta_group::2 (M 256, N 128, K 16)
//   warps 2-5: per item, copy k-blocks 0-7 of this CTA's 128 query rows global -> registers -> TMEM
//              (tcgen05.st); then the top-k epilogue on 128-column tiles
// ---------------------------------------------------------------------------------------------------------
constexpr int kTsBN = 128;                // candidates per tile
constexpr int kTsStageKb = 2;             // k-blocks per smem stage
constexpr int kTsStages = 9;
constexpr int kTsKbBytes = (kTsBN / 2) * kBK * 2;              // 8 KB: this CTA's 64 rows of one k-block
constexpr int kTsStageBytes = kTsStageKb * kTsKbBytes;         // 16 KB
constexpr int kTsRingBytes = kTsStages * kTsStageBytes;        // 144 KB
constexpr int kTsTmemKb = 8;              // query k-blocks held in TMEM (8 x 32 = 256 columns)
constexpr int kTsSmemKb = 4;              // query k-blocks held in shared memory
constexpr int kTsMaxKb = kTsTmemKb + kTsSmemKb;                // 12 x 64 = 768
constexpr int kTsAKbBytes = kBM * kBK * 2;                     // 16 KB: this CTA's 128 rows of one k-block
constexpr int kTsASmemBytes = kTsSmemKb * kTsAKbBytes;         // 64 KB
constexpr int kTsACols = kTsTmemKb * (kBK / 2);                // 256
constexpr int kTsBarBytes = ((2 * kTsStages + 6) * 8 + 8 + 127) / 128 * 128;
constexpr int kTsInvBytes = 4 * 2 * kTsBN * static_cast<int>(sizeof(float));   // per epilogue warp: two tiles of 1/||c||
constexpr int kTsSmemBytes = kTsRingBytes + kTsASmemBytes + kTsBarBytes + kTsInvBytes + 1024;
static_assert(kTsACols + 2 * kTsBN == 512, "TMEM budget: queries + two accumulators");
static_assert(kTsSmemBytes <= 232448, "shared memory budget");
// Row buffers and their compaction (k = 50; same launch with k = 1: 1173 / 159 ms at 5.6 M / 0.7 M rows):
//   128 entries, bitonic sort in the tile (round-2a)                         1259 / 193 ms   (same box, r2c16)
//   256 entries, exact radix SELECT in the tile (default)                    1241 / 185 ms
//   256 entries, bitonic sort AFTER the accumulator release (kDefer)         slower than 128-in-tile (r2c14): a sort
//                                                                            of 256 costs 2.6x one of 128 and the
//                                                                            double-buffered accumulator does not hide it
constexpr bool kTsDeferCompaction = false;
// Row buffers of the queries-on-chip kernel: 256 entries, compacted in the tile by selection (see compact()).
using SearchEpilogueTs = SearchEpilogueT<kTsBN, kCapMax, true>;
using SearchEpilogueTsSort = SearchEpilogueT<kTsBN, kCap, false>;     // UNIIR_SEARCH_SELECT=0: 128 entries, bitonic sort

// One 128-candidate tile's MMAs.  kFull: d = 768 (12 k-blocks, 6 full stages): everything but the stage slot is a
// compile-time constant, so each MMA costs the issuing thread two adds.
template <bool kFull>
__device__ __forceinline__ void ts_issue_tile(uint32_t tmem_d, uint32_t tmem_a, uint64_t adesc0, uint64_t bdesc0,
                                              uint32_t idesc, int kb_total, uint64_t* full_bar, uint64_t* empty_bar,
                                              int& stage, uint32_t& phase) {
  if constexpr (kFull) {
#pragma unroll
    for (int s6 = 0; s6 < kTsMaxKb / kTsStageKb; ++s6) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (kTsStageBytes >> 4));
#pragma unroll
      for (int j = 0; j < kTsStageKb; ++j) {
        const int kb = s6 * kTsStageKb + j;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t db = bd + static_cast<uint64_t>((j * kTsKbBytes + k * 32) >> 4);
          if (kb < kTsTmemKb)
            umma_f16_ts_2sm(tmem_d, tmem_a + kb * (kBK / 2) + k * (kUmmaK / 2), db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          else
            umma_f16_2sm(tmem_d, adesc0 + static_cast<uint64_t>(((kb - kTsTmemKb) * kTsAKbBytes + k * 32) >> 4), db, idesc, 1u);
        }
      }
      umma_commit_2sm(&empty_bar[stage], 0x3);
      if (++stage == kTsStages) { stage = 0; phase ^= 1; }
    }
  } else {
    for (int kb0 = 0; kb0 < kb_total; kb0 += kTsStageKb) {
      const int nkb = min(kTsStageKb, kb_total - kb0);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * (kTsStageBytes >> 4));
      for (int j = 0; j < nkb; ++j) {
        const int kb = kb0 + j;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t db = bd + static_cast<uint64_t>((j * kTsKbBytes + k * 32) >> 4);
          if (kb < kTsTmemKb)
            umma_f16_ts_2sm(tmem_d, tmem_a + kb * (kBK / 2) + k * (kUmmaK / 2), db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          else
            umma_f16_2sm(tmem_d, adesc0 + static_cast<uint64_t>(((kb - kTsTmemKb) * kTsAKbBytes + k * 32) >> 4), db, idesc, 1u);
        }
      }
      umma_commit_2sm(&empty_bar[stage], 0x3);
      if (++stage == kTsStages) { stage = 0; phase ^= 1; }
    }
  }
}

template <bool kPipe, bool kSel = true>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads(4), 1)
ip_topk_kernel_ts(const __grid_constant__ GemmTensorMaps maps, const __half* __restrict__ q, const int d,
                  const SearchParams p, const SearchSched sched, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* ring = smem;
  uint8_t* a_smem = smem + kTsRingBytes;                      // query k-blocks 8..11, K-major 128B-swizzled
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kTsRingBytes + kTsASmemBytes);
  uint64_t* empty_bar = full_bar + kTsStages;
  uint64_t* tfull_bar = empty_bar + kTsStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* afull_bar = tempty_bar + 2;
  uint64_t* aempty_bar = afull_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aempty_bar + 1);
  float* invc_s = reinterpret_cast<float*>(smem + kTsRingBytes + kTsASmemBytes + kTsBarBytes);

  const int warp_idx = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    for (int i = 0; i < kTsStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * 4);     // one arrival per epilogue warp of both CTAs
    }
    mbar_init(afull_bar, 2 * 128 + 1);      // TMEM part: every epilogue thread of both CTAs; smem part: the leader's expect_tx
    mbar_init(aempty_bar, 1);               // the item's last MMA has retired: the shared-memory queries may be replaced
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc_2sm<512>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_a = tmem_base;                  // columns [0, 256)
  const uint32_t tmem_d0 = tmem_base + kTsACols;      // two 128-column accumulators

  const int n_items = sched.num_items();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int kb_total = sched.kb_total;
  const int kb_tmem = min(kb_total, kTsTmemKb), kb_smem = kb_total - kb_tmem;

  if (warp_idx == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      const uint32_t af = mapa_cluster(smem_u32(afull_bar), 0);
      WorkItem w;
      for (int item = pair; item < n_items; item += n_pairs) {
        if (!sched.get(item, w)) continue;
        if (w.n_begin >= w.n_end) continue;
        // The candidate ring does not depend on the queries: the first ring-full of this item's stages is issued
        // while the previous item is still draining; only then wait until its MMAs have let go of the
        // shared-memory query buffer and load this item's query k-blocks 8..11.
        const int a_row = (w.m_blk * 2 + static_cast<int>(rank)) * kBM;
        const int stages_per_tile = (kb_total + kTsStageKb - 1) / kTsStageKb;
        const int n_stage_loads = (w.n_end - w.n_begin) * stages_per_tile;
        const int a_after = min(kTsStages, n_stage_loads);      // query load goes in after this many stage loads
        auto load_queries = [&]() {
          if (kb_smem > 0) mbar_wait(aempty_bar, a_phase ^ 1);
          a_phase ^= 1;
          if (leader) mbar_arrive_expect_tx(afull_bar, 2 * kb_smem * kTsAKbBytes);
          for (int j = 0; j < kb_smem; ++j) tma_load_2d_2sm(a_smem + j * kTsAKbBytes, &maps.a, af, (kTsTmemKb + j) * kBK, a_row);
        };
        int issued = 0;
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          const int b_row = w.tile_at(ti) * kTsBN + static_cast<int>(rank) * (kTsBN / 2);
          for (int kb0 = 0; kb0 < kb_total; kb0 += kTsStageKb) {
            if (issued == a_after) load_queries();
            ++issued;
            const int nkb = min(kTsStageKb, kb_total - kb0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * nkb * kTsKbBytes);
            uint8_t* sb = ring + stage * kTsStageBytes;
            for (int j = 0; j < nkb; ++j) tma_load_2d_2sm(sb + j * kTsKbBytes, &maps.b, fb, (kb0 + j) * kBK, b_row);
            if (++stage == kTsStages) { stage = 0; phase ^= 1; }
          }
        }
        if (issued <= a_after) load_queries();
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer (leader only) ============================
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint64_t bdesc0 = make_smem_desc(smem_u32(ring), 16, 1024);
      const uint64_t adesc0 = make_smem_desc(smem_u32(a_smem), 16, 1024);
      const bool full_k = kb_total == kTsMaxKb;
      WorkItem w;
      for (int item = pair; item < n_items; item += n_pairs) {
        if (!sched.get(item, w)) continue;
        if (w.n_begin >= w.n_end) continue;   // an empty split: nothing to multiply (the epilogue pads its result)
        mbar_wait(afull_bar, a_phase);        // this item's queries are on chip in both CTAs
        a_phase ^= 1;
        tc_fence_after();
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_d0 + acc * kTsBN;
          if (full_k) ts_issue_tile<true>(tmem_d, tmem_a, adesc0, bdesc0, idesc, kb_total, full_bar, empty_bar, stage, phase);
          else ts_issue_tile<false>(tmem_d, tmem_a, adesc0, bdesc0, idesc, kb_total, full_bar, empty_bar, stage, phase);
          umma_commit_2sm(&tfull_bar[acc], 0x3);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        if (kb_smem > 0) umma_commit_2sm(aempty_bar, 0x3);
      }
    }
  } else {
    // ============================ query loader + epilogue (both CTAs) ================================
    typename std::conditional<kSel, SearchEpilogueTs, SearchEpilogueTsSort>::type epi(p, invc_s);
    const int quarter = warp_idx & 3;
    const int row = quarter * 32 + lane_id();
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t te0 = mapa_cluster(smem_u32(&tempty_bar[0]), 0), te1 = mapa_cluster(smem_u32(&tempty_bar[1]), 0);
    const uint32_t af = mapa_cluster(smem_u32(afull_bar), 0);
    float* winv_base = invc_s + (warp_idx - 2) * 2 * kTsBN;
    int inv_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int item = pair; item < n_items; item += n_pairs) {
      if (!sched.get(item, w)) continue;
      w.m_blk = w.m_blk * 2 + static_cast<int>(rank);
      w.tag = w.tag * 2 + static_cast<int>(rank);
      // Every MMA that read the previous item's queries has retired: this thread consumed that item's last
      // accumulator, and the commit that published it covers all earlier MMAs of the issuer.
      if (w.n_begin < w.n_end) {
        const long long qrow = static_cast<long long>(w.m_blk) * kBM + row;
        const uint4* src = reinterpret_cast<const uint4*>(q + qrow * d);
        const bool live = qrow < p.Q;
        for (int c = 0; c < kb_tmem; ++c) {        // 64 fp16 = 128 B = 32 TMEM columns per step
          uint32_t r[32];
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int e = c * kBK + v * 8;          // first element of this 16-byte piece
            uint4 u = make_uint4(0u, 0u, 0u, 0u);   // rows beyond Q and columns beyond d are zero
            if (live && e < d) u = __ldg(src + c * 8 + v);
            r[v * 4 + 0] = u.x; r[v * 4 + 1] = u.y; r[v * 4 + 2] = u.z; r[v * 4 + 3] = u.w;
          }
          tmem_st_32x32b_x32(tmem_a + lane_base + c * (kBK / 2), r);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_cluster(af);
      }
      epi.item_begin(w, row);
      // 1/||c|| of a tile: fetched one tile ahead into registers, parked in this warp's own slice of shared memory
      // before the accumulator is waited for, so neither a global-load latency nor a cross-warp barrier sits
      // between "accumulator ready" and "accumulator released"
      const int n_tiles_item = w.n_end - w.n_begin;
      float nx[kTsBN / 32];
      auto fetch_inv = [&](int n_blk) {
#pragma unroll
        for (int t = 0; t < kTsBN / 32; ++t) {
          const long long c0 = static_cast<long long>(n_blk) * kTsBN + t * 32 + lane_id();
          nx[t] = (c0 < p.N) ? __ldg(p.c_inv + c0) : 0.f;
        }
      };
      if (n_tiles_item > 0) fetch_inv(w.tile_at(0));
      for (int ti = 0; ti < n_tiles_item; ++ti) {
        const int n = w.tile_at(ti);
        if (ti == n_tiles_item - 16 || (n_tiles_item < 16 && ti == 0)) {
          // shortly before the item ends: pull the NEXT item's query row into L2, so that the row-per-thread loads
          // that refill tensor memory at the item boundary (nothing else runs then) hit L2 instead of DRAM
          WorkItem w2;
          const int nxt = item + n_pairs;
          if (nxt < n_items && sched.get(nxt, w2)) {
            const long long qrow2 = static_cast<long long>(w2.m_blk * 2 + static_cast<int>(rank)) * kBM + row;
            if (qrow2 < p.Q) {
              const char* base = reinterpret_cast<const char*>(q + qrow2 * d);
              for (int b = 0; b < d * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + b));
            }
          }
        }
        float* winv = winv_base + (inv_buf & 1) * kTsBN;
        ++inv_buf;
#pragma unroll
        for (int t = 0; t < kTsBN / 32; ++t) winv[t * 32 + lane_id()] = nx[t];
        __syncwarp();
        if (ti + 1 < n_tiles_item) fetch_inv(w.tile_at(ti + 1));
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        epi.template tile<kPipe, kTsDeferCompaction>(w, n, tmem_d0 + lane_base + acc * kTsBN, row, 0, winv);
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive_remote(acc ? te1 : te0);
        if constexpr (kTsDeferCompaction) epi.compact_deferred(w, row);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      epi.item_end(w, row);
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------
// inv-norm: one warp per row
// ---------------------------------------------------------------------------------
__global__ void row_inv_norm_kernel(const uint16_t* __restrict__ x, int dtype, long long rows, int d,
                                    long long ld, float* __restrict__ inv) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint16_t* xr = x + row * ld;
  float acc = 0.f;
  for (int i = lane * 8; i < d; i += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float a, b;
      if (dtype == UNIIR_BF16) {
        a = __uint_as_float(wv[t] << 16);
        b = __uint_as_float(wv[t] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wv[t]));
        a = f.x; b = f.y;
      }
      acc = fmaf(a, a, acc);
      acc = fmaf(b, b, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) inv[row] = acc > 0.f ? 1.f / sqrtf(acc) : 0.f;
}

// ---------------------------------------------------------------------------------
// k-way merge: one warp per query, bitonic over L*k <= 32*R entries
// ---------------------------------------------------------------------------------
template <int R>
__global__ void topk_merge_kernel(const float* __restrict__ in_s, const void* __restrict__ in_i,
                                  int idx64, long long Q, int L, int k, long long q_stride,
                                  long long l_stride, const long long* __restrict__ list_offset,
                                  const long long* __restrict__ id_table, float* __restrict__ out_s,
                                  long long* __restrict__ out_i) {
  const long long q = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const int lane = threadIdx.x & 31;
  constexpr long long kNone = 0x7fffffffffffffffLL;
  float key[R];
  long long idx[R];
  const int total = L * k;
  // Streaming merge: the first k slots hold the running best, the remaining 32R-k slots are
  // refilled with unseen entries and the whole buffer re-sorted until the input is consumed.
  int next = 0;      // next unseen input entry
  int fill_from = 0; // first slot to refill
  do {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int e = r * 32 + lane;
      if (e >= fill_from) {
        key[r] = kNegInf;
        idx[r] = kNone;
        const int src = next + (e - fill_from);
        if (src < total) {
          const int l = src / k, j = src - l * k;
          const long long off = q * q_stride + l * l_stride + j;
          const long long raw = idx64 ? static_cast<const long long*>(in_i)[off]
                                      : static_cast<long long>(static_cast<const int*>(in_i)[off]);
          if (raw >= 0) {
            key[r] = in_s[off];
            idx[r] = raw + (list_offset ? list_offset[l] : 0);
          }
        }
      }
    }
    next += 32 * R - fill_from;
    fill_from = k;
    warp_bitonic_desc<R, long long>(key, idx);
  } while (next < total);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    if (e < k) {
      const bool ok = idx[r] != kNone;
      out_s[q * k + e] = ok ? key[r] : kFaissNeg;
      out_i[q * k + e] = ok ? (id_table ? id_table[idx[r]] : idx[r]) : -1;
    }
  }
}

// Kernel family: 2 = queries in TMEM (CTA pairs, 64-candidate tiles; needs d <= 768), 1 = CTA pairs streaming both
// operands (256-candidate tiles), 0 = single CTAs.  UNIIR_SEARCH_MODE overrides; UNIIR_SEARCH_2CTA=0 is the round-1
// spelling of mode 0.
static int search_mode(int d) {
  int m = 2;
  if (const char* e = getenv("UNIIR_SEARCH_MODE")) m = atoi(e);
  if (const char* e = getenv("UNIIR_SEARCH_2CTA")) { if (atoi(e) == 0) m = 0; }
  if (m == 2 && d > kTsMaxKb * kBK) m = 1;
  return m < 0 ? 0 : (m > 2 ? 2 : m);
}
static bool search_pairs() {
  const char* e = getenv("UNIIR_SEARCH_2CTA");
  if (e && atoi(e) == 0) return false;
  if (const char* m = getenv("UNIIR_SEARCH_MODE")) return atoi(m) != 0;
  return true;
}

// n_qtiles is the number of 128-row state slots per split (padded to an even count in pair mode).  The number of
// splits does not depend on the kernel family (the plan call does not know d): it is sized on 256-candidate tiles;
// tiles_per_split / n_tiles are in tiles of `bn` candidates.
static void plan_splits(long long Q, long long N, int bn, int* splits, int* tiles_per_split, int* n_tiles,
                        int* n_qtiles) {
  const int sms = num_sms();
  const bool pairs = search_pairs();
  int nq = static_cast<int>((Q + kBM - 1) / kBM);
  if (pairs) nq = (nq + 1) / 2 * 2;
  const int units = pairs ? nq / 2 : nq;          // schedulable units per chunk and split
  const int workers = pairs ? sms / 2 : sms;
  const int nt256 = static_cast<int>((N + kSearchBN - 1) / kSearchBN);
  // Enough independent (split, unit) slots to keep every SM busy; one split when the query set
  // alone provides them.  Each split needs a few tiles to be worth its ramp-up.
  int s = 1;
  if (units < workers) {
    s = (workers + units - 1) / units;
    if (s > 32) s = 32;
    while (s > 1 && (nt256 + s - 1) / s < 8) --s;
  }
  const int nt = static_cast<int>((N + bn - 1) / bn);
  *splits = s;
  *tiles_per_split = ((nt256 + s - 1) / s) * (kSearchBN / bn);   // split boundaries do not move with the tile width
  *n_tiles = nt;
  *n_qtiles = nq;
}

static size_t search_ws_bytes(int splits, int nq) {
  const size_t rows = static_cast<size_t>(splits) * nq * kBM;
  return rows * kCapMax * (sizeof(float) + sizeof(int)) + rows * (sizeof(float) + sizeof(int)) +
         static_cast<size_t>(splits) * nq * sizeof(int) + 256;
}

}  // namespace uniir

using namespace uniir;

extern "C" int uniir_row_inv_norm(const void* x, int32_t dtype, int64_t rows, int32_t d, int64_t ld,
                                  float* inv, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UNIIR_CHECK_ARG(x && inv && rows >= 0 && d > 0, "uniir_row_inv_norm: bad args");
  UNIIR_CHECK_ARG(d % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                  "uniir_row_inv_norm: d, ld must be multiples of 8 and x 16B aligned");
  UNIIR_CHECK_ARG(dtype == UNIIR_BF16 || dtype == UNIIR_F16, "uniir_row_inv_norm: dtype");
  if (rows == 0) return UNIIR_OK;
  const int threads = 256;
  const long long blocks = (rows * 32 + threads - 1) / threads;
  row_inv_norm_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      static_cast<const uint16_t*>(x), dtype, rows, d, ld, inv);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_ip_topk_plan(int64_t Q, int64_t N, int32_t k, int32_t* splits,
                                  size_t* workspace_bytes) {
  UNIIR_CHECK_ARG(Q > 0 && N > 0 && k > 0 && k <= kMaxK && splits && workspace_bytes,
                  "uniir_ip_topk_plan: bad args (k must be in [1,%d])", kMaxK);
  int s, tps, nt, nq;
  plan_splits(Q, N, kSearchBN, &s, &tps, &nt, &nq);
  *splits = s;
  *workspace_bytes = search_ws_bytes(s, nq);
  return UNIIR_OK;
}

extern "C" int uniir_ip_topk(const void* q, const float* q_inv, const void* c, const float* c_inv,
                             int64_t Q, int64_t N, int32_t d, int32_t k, int32_t splits,
                             float* part_scores, int32_t* part_idx, void* workspace,
                             size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UNIIR_CHECK_ARG(q && q_inv && c && c_inv && part_scores && part_idx && workspace,
                  "uniir_ip_topk: null pointer");
  UNIIR_CHECK_ARG(Q > 0 && N > 0 && N < (1LL << 31) && d > 0 && d % 8 == 0,
                  "uniir_ip_topk: bad shape Q=%lld N=%lld d=%d", (long long)Q, (long long)N, d);
  UNIIR_CHECK_ARG(k > 0 && k <= kMaxK, "uniir_ip_topk: k must be in [1,%d]", kMaxK);
  const int mode = search_mode(d);
  const int bn = mode == 2 ? kTsBN : kSearchBN;
  int s, tps, nt, nq;
  plan_splits(Q, N, bn, &s, &tps, &nt, &nq);
  UNIIR_CHECK_ARG(splits == s, "uniir_ip_topk: splits=%d does not match uniir_ip_topk_plan (%d)",
                  splits, s);
  const int sms = num_sms();
  const size_t need = search_ws_bytes(s, nq);
  UNIIR_CHECK_ARG(workspace_bytes >= need, "uniir_ip_topk: workspace too small (%zu < %zu)",
                  workspace_bytes, need);

  const bool pairs = search_pairs();
  GemmTensorMaps maps;
  int rc = make_tmap_2d(&maps.a, q, UNIIR_F16, d, Q, d, kBK, kBM);
  if (rc) return rc;
  rc = make_tmap_2d(&maps.b, c, UNIIR_F16, d, N, d, kBK, mode == 2 ? kTsBN / 2 : (pairs ? kSearchBN / 2 : kSearchBN));
  if (rc) return rc;

  SearchParams p;
  p.Q = Q; p.N = N; p.k = k; p.splits = s; p.n_qtiles = nq; p.tiles_per_split = tps;
  p.n_tiles = nt; p.kb_total = (d + kBK - 1) / kBK;
  p.q_inv = q_inv; p.c_inv = c_inv;
  p.part_scores = part_scores; p.part_idx = part_idx;
  const size_t rows = static_cast<size_t>(s) * nq * kBM;
  int tpc = mode == 2 ? kDefaultTilesPerChunkTs : kDefaultTilesPerChunk;
  if (const char* e = getenv("UNIIR_SEARCH_TPC")) { const int v = atoi(e); if (v >= 1) tpc = v; }
  p.n_chunks = (tps + tpc - 1) / tpc;
  p.ws_scores = static_cast<float*>(workspace);
  p.ws_idx = reinterpret_cast<int*>(p.ws_scores + rows * kCapMax);
  p.g_thr = reinterpret_cast<float*>(p.ws_idx + rows * kCapMax);
  p.g_cnt = reinterpret_cast<int*>(p.g_thr + rows);
  p.done = p.g_cnt + rows;
  UNIIR_CUDA_OK(cudaMemsetAsync(p.done, 0, static_cast<size_t>(s) * nq * sizeof(int), stream));
  const int units = pairs ? nq / 2 : nq;
  SearchSched sched{units, s, tps, nt, p.kb_total, p.n_chunks, tpc, 0};
  bool stagger = true;
  if (const char* e = getenv("UNIIR_SEARCH_ROT")) stagger = atoi(e) != 0;
  const long long items = 1LL * units * s * p.n_chunks;
  UNIIR_CHECK_ARG(items < (1LL << 31), "uniir_ip_topk: too many work items");

  if (mode == 2) {
    static std::atomic<uint64_t> attr_ts{0};
    if (first_on_device(attr_ts)) {
      UNIIR_CUDA_OK(cudaFuncSetAttribute(ip_topk_kernel_ts<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes));
      UNIIR_CUDA_OK(cudaFuncSetAttribute(ip_topk_kernel_ts<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes));
      UNIIR_CUDA_OK(cudaFuncSetAttribute(ip_topk_kernel_ts<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmemBytes));
    }
    bool select = true;    // UNIIR_SEARCH_SELECT=0: the round-2a epilogue (128-entry buffers compacted by a bitonic sort)
    if (const char* e = getenv("UNIIR_SEARCH_SELECT")) select = atoi(e) != 0;
    // Epilogue with double-buffered TMEM loads: measured SLOWER on the same box (1107 vs 1177 TFLOP/s at 512-tile
    // chunks, gpurun_out/r2c8_modes_full.txt; 183 vs 138 registers, twice the epilogue code) -> opt-in only.
    bool pipe = false;
    if (const char* e = getenv("UNIIR_SEARCH_EPIPIPE")) pipe = atoi(e) != 0;
    int npairs = sms / 2;
    if (items < npairs) npairs = static_cast<int>(items);
    sched.n_workers = stagger ? npairs : 0;
    const uint32_t idesc = make_idesc_f16(false, false, false, 2 * kBM, kTsBN);
    if (!select) ip_topk_kernel_ts<false, false><<<2 * npairs, num_threads(4), kTsSmemBytes, stream>>>(maps, static_cast<const __half*>(q), d, p, sched, idesc);
    else if (pipe) ip_topk_kernel_ts<true, true><<<2 * npairs, num_threads(4), kTsSmemBytes, stream>>>(maps, static_cast<const __half*>(q), d, p, sched, idesc);
    else ip_topk_kernel_ts<false, true><<<2 * npairs, num_threads(4), kTsSmemBytes, stream>>>(maps, static_cast<const __half*>(q), d, p, sched, idesc);
  } else if (pairs) {
    using SM = GemmSmem2<kSearchBN>;
    const int smem_bytes = SM::kRingBytes + SM::kBarBytes + 2 * kSearchBN * sizeof(float) + 1024;
    static std::atomic<uint64_t> attr2{0};
    if (first_on_device(attr2))
      UNIIR_CUDA_OK(cudaFuncSetAttribute(ip_topk_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int npairs = sms / 2;
    if (items < npairs) npairs = static_cast<int>(items);
    const uint32_t idesc = make_idesc_f16(false, false, false, 2 * kBM, kSearchBN);
    sched.n_workers = stagger ? npairs : 0;
    int hint_a = kL2None, hint_b = kL2None;   // diagnostics: UNIIR_SEARCH_HINT=<a><b>, digits 0 none 1 normal 2 evict_first 3 evict_last
    if (const char* e = getenv("UNIIR_SEARCH_HINT")) { const int v = atoi(e); hint_a = (v / 10) % 10; hint_b = v % 10; }
    ip_topk_kernel_2cta<<<2 * npairs, num_threads(4), smem_bytes, stream>>>(maps, p, sched, idesc, hint_a, hint_b);
  } else {
    using SM = GemmSmem<kSearchBN>;
    const int smem_bytes = SM::kRingBytes + SM::kBarBytes + 2 * kSearchBN * sizeof(float) + 1024;
    static std::atomic<uint64_t> attr_set{0};
    if (first_on_device(attr_set))
      UNIIR_CUDA_OK(cudaFuncSetAttribute(ip_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int grid = items < sms ? static_cast<int>(items) : sms;
    sched.n_workers = stagger ? grid : 0;
    const uint32_t idesc = make_idesc_f16(false, false, false, kBM, kSearchBN);
    ip_topk_kernel<<<grid, num_threads(4), smem_bytes, stream>>>(maps, p, sched, idesc);
  }
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_topk_merge(const float* in_scores, const void* in_idx, int32_t idx_is_64,
                                int64_t Q, int32_t L, int32_t k, int64_t q_stride, int64_t l_stride,
                                const int64_t* list_offset, const int64_t* id_table,
                                float* out_scores, int64_t* out_ids, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UNIIR_CHECK_ARG(in_scores && in_idx && out_scores && out_ids, "uniir_topk_merge: null pointer");
  UNIIR_CHECK_ARG(Q >= 0 && L > 0 && k > 0 && k <= kMaxK,
                  "uniir_topk_merge: need k in [1,%d] (L=%d k=%d)", kMaxK, L, k);
  if (Q == 0) return UNIIR_OK;
  const int total = L * k;
  const int threads = 128;
  const long long blocks = (Q * 32 + threads - 1) / threads;
  const long long* lo = reinterpret_cast<const long long*>(list_offset);
  const long long* idt = reinterpret_cast<const long long*>(id_table);
  long long* oi = reinterpret_cast<long long*>(out_ids);
#define UNIIR_MERGE(R_)                                                                        \
  topk_merge_kernel<R_><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(                \
      in_scores, in_idx, idx_is_64, Q, L, k, q_stride, l_stride, lo, idt, out_scores, oi)
  if (total <= 64) UNIIR_MERGE(2);
  else if (total <= 128) UNIIR_MERGE(4);
  else if (total <= 256) UNIIR_MERGE(8);
  else UNIIR_MERGE(16);
#undef UNIIR_MERGE
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
