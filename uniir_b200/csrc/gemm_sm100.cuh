// Warp-specialised persistent tcgen05 GEMM mainloop for sm_100a.
//
//   D[M,N] = A · Bᵀ  with A = [M,K] and B = [N,K] as *logical* operands, fp16/bf16 in,
//   fp32 accumulate in TMEM.  Each operand may be stored K-major (row-major [rows,K]:
//   the contraction index is contiguous) or MN-major (row-major [K,rows]: the non-
//   contracted index is contiguous) -- the latter is what dgrad/wgrad need without
//   materialising transposes.
//
// Roles (192 threads, 1 CTA/SM):
//   warp 0   : TMA producer  (one elected lane) -- fills a kStages-deep smem ring
//   warp 1   : MMA issuer    (one elected lane) -- tcgen05.mma into one of 2 TMEM
//              accumulator buffers; tcgen05.commit releases smem slots / publishes tiles
//   warps 2-5: epilogue      (TMEM lane quarter = warp_idx % 4) -- user policy `Epi`
// Three pipelines: smem full/empty, TMEM full/empty, and the persistent work loop.
//
// Work is described by a scheduler policy that maps a linear item index to
//   {m_blk, n_begin, n_end, kb_begin, kb_end}: the CTA computes output tiles
//   (m_blk, n) for n in [n_begin, n_end), each contracting k-blocks [kb_begin, kb_end).
// Plain GEMM: one tile per item.  Retrieval: one query tile x a whole candidate range.
#pragma once
#include "ptx.cuh"

namespace uniir {

constexpr int kBM = 128;   // UMMA M (cta_group::1)
constexpr int kBK = 64;    // 64 x 2B = one 128B swizzle row
constexpr int kUmmaK = 16;
// threads = 2 control warps (TMA, MMA) + NE epilogue warps (4: one per TMEM lane quarter;
// 8: two per quarter, each taking half of the tile's columns)
constexpr int num_threads(int ne) { return 64 + 32 * ne; }

struct WorkItem {
  int m_blk, n_begin, n_end, kb_begin, kb_end;
  int tag, aux;  // scheduler-defined (retrieval: state slot, chunk index / last-chunk flag)
  int rot;       // the item's tiles are visited in the cyclic order n_begin + (i + rot) % (n_end - n_begin)
  // i-th tile of the item in visiting order (0 <= rot < n_end - n_begin)
  __device__ __forceinline__ int tile_at(int i) const {
    const int n = n_begin + i + rot;
    return n >= n_end ? n - (n_end - n_begin) : n;
  }
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kRingBytes = kStages * kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then tmem ptr
  static constexpr int kBarBytes = ((2 * kStages + 4) * 8 + 8 + 127) / 128 * 128;  // keeps what follows 128B aligned
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers
};

struct GemmTensorMaps {
  CUtensorMap a, b;
};

// ---------------------------------------------------------------------------------
// The kernel body.  `Sched` : bool get(int item, WorkItem&) const;  int num_items() const
//                   `Epi`   : see gemm.cu / search.cu.  Runs on the 128 epilogue threads:
//       void item_begin(const WorkItem&, int row_in_tile)
//       void tile(const WorkItem&, int n_blk, uint32_t tmem_acc /*lane-adjusted*/, int row_in_tile,
//                 int col_half /*0..NE/4-1: which slice of the BN columns this warp owns*/)
//       void item_end(const WorkItem&, int row_in_tile)
// `extra_smem` bytes after the ring/barriers are handed to Epi via its ctor.
// ---------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int NE, class Sched, class Epi>
__device__ __forceinline__ void gemm_body(const CUtensorMap* tmA, const CUtensorMap* tmB,
                                          uint32_t idesc, const Sched& sched, Epi& epi,
                                          uint8_t* smem /*1024-aligned*/) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kRingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(tmA);
    tma_prefetch_desc(tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 32 * NE);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_items = sched.num_items();

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      WorkItem w;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (!sched.get(item, w)) continue;
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          const int n = w.tile_at(ti);
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = ring + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
            if constexpr (!A_MN) {
              tma_load_2d(sa, tmA, &full_bar[stage], kb * kBK, w.m_blk * kBM);
            } else {
#pragma unroll
              for (int g = 0; g < kBM / 64; ++g)
                tma_load_2d(sa + g * (kBK * 128), tmA, &full_bar[stage], w.m_blk * kBM + g * 64,
                            kb * kBK);
            }
            if constexpr (!B_MN) {
              tma_load_2d(sb, tmB, &full_bar[stage], kb * kBK, n * BN);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 64; ++g)
                tma_load_2d(sb + g * (kBK * 128), tmB, &full_bar[stage], n * BN + g * 64, kb * kBK);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ==============================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkItem w;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        if (!sched.get(item, w)) continue;
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * S::kStageBytes);
            const uint32_t sb = sa + S::kABytes;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // K-major : 8-row groups 1024B apart (SBO); +32B per UMMA_K inside the swizzle row
              // MN-major: 64-wide MN groups kBK*128 B apart (LBO); 8-k-row groups 1024B apart
              //           (SBO); +16 k-rows = 2048B per UMMA_K
              const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, kBK * 128, 1024)
                                       : make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, kBK * 128, 1024)
                                       : make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16(tmem_d, da, db, idesc, (kb > w.kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ============================ epilogue ================================
    const int quarter = warp_idx & 3;  // TMEM lanes [32q, 32q+32) are visible to this warp
    const int col_half = (warp_idx - 2) >> 2;
    const int row = quarter * 32 + lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      if (!sched.get(item, w)) continue;
      epi.item_begin(w, row);
      for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
        const int n = w.tile_at(ti);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        epi.tile(w, n, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN, row, col_half);
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      epi.item_end(w, row);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}


// ---------------------------------------------------------------------------------
// CTA-pair variant: a cluster of two CTAs (two SMs of one TPC) computes a 256 x 256 tile with
// one `tcgen05.mma.cta_group::2` (M = 256).  Each CTA stages only ITS half of both operands
// (A rows [r*128, +128), B rows [r*128, +128) of the tile), so shared-memory fill traffic per
// FLOP halves relative to the 128 x 256 single-CTA tile, and holds its own 128 accumulator rows
// in its own TMEM.  Protocol (CUTLASS-style):
//   full[s]   : lives in the LEADER (rank 0); the leader's producer arms it for BOTH CTAs'
//               bytes, both CTAs' TMA loads credit it (2-SM TMA form);
//   empty[s]  : one per CTA; released by the leader's multicast tcgen05.commit;
//   tfull[b]  : one per CTA; published by the same multicast commit;
//   tempty[b] : lives in the leader; the epilogue threads of BOTH CTAs arrive on it.
// Work items address 256-row blocks; the epilogue sees m_blk in 128-row units (2*blk + rank)
// so the same epilogue policies serve both bodies.
// ---------------------------------------------------------------------------------
template <int BN>
struct GemmSmem2 {
  static_assert(BN == 256, "the pair tile is 256 x 256");
  static constexpr int kStages = 6;
  static constexpr int kABytes = kBM * kBK * 2;          // this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kBK * 2;     // this CTA's 128 rows of B
  static constexpr int kStageBytes = kABytes + kBBytes;  // 32 KB
  static constexpr int kRingBytes = kStages * kStageBytes;
  static constexpr int kBarBytes = ((2 * kStages + 4) * 8 + 8 + 127) / 128 * 128;
  static constexpr int kTmemCols = 2 * BN;
};

template <int BN, bool A_MN, bool B_MN, int NE, class Sched, class Epi>
__device__ __forceinline__ void gemm_body_2cta(const CUtensorMap* tmA, const CUtensorMap* tmB,
                                               uint32_t idesc, const Sched& sched, Epi& epi,
                                               uint8_t* smem /*1024-aligned*/,
                                               int l2_hint_a = kL2None, int l2_hint_b = kL2None) {
  using S = GemmSmem2<BN>;
  constexpr int kStages = S::kStages;
  uint8_t* ring = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kRingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(tmA);
    tma_prefetch_desc(tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * 32 * NE);  // epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) tmem_alloc_2sm<S::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anyone signals remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int n_items = sched.num_items();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp_idx == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      WorkItem w;
      const uint64_t pol_a = make_l2_policy(l2_hint_a), pol_b = make_l2_policy(l2_hint_b);
      for (int item = pair; item < n_items; item += n_pairs) {
        if (!sched.get(item, w)) continue;
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          const int n = w.tile_at(ti);
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = ring + stage * S::kStageBytes;
            uint8_t* sb = sa + S::kABytes;
            const uint32_t fb = mapa_cluster(smem_u32(&full_bar[stage]), 0);  // the leader's barrier
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            const int a_row = w.m_blk * 2 * kBM + rank * kBM;
            const int b_row = n * BN + rank * (BN / 2);
            if constexpr (!A_MN) {
              if (l2_hint_a) tma_load_2d_2sm_hint(sa, tmA, fb, kb * kBK, a_row, pol_a);
              else tma_load_2d_2sm(sa, tmA, fb, kb * kBK, a_row);
            } else {
#pragma unroll
              for (int g = 0; g < kBM / 64; ++g) tma_load_2d_2sm(sa + g * (kBK * 128), tmA, fb, a_row + g * 64, kb * kBK);
            }
            if constexpr (!B_MN) {
              if (l2_hint_b) tma_load_2d_2sm_hint(sb, tmB, fb, kb * kBK, b_row, pol_b);
              else tma_load_2d_2sm(sb, tmB, fb, kb * kBK, b_row);
            } else {
#pragma unroll
              for (int g = 0; g < BN / 128; ++g) tma_load_2d_2sm(sb + g * (kBK * 128), tmB, fb, b_row + g * 64, kb * kBK);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer (leader only) ============================
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      WorkItem w;
      for (int item = pair; item < n_items; item += n_pairs) {
        if (!sched.get(item, w)) continue;
        for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BN;
          for (int kb = w.kb_begin; kb < w.kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * S::kStageBytes);
            const uint32_t sb = sa + S::kABytes;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, kBK * 128, 1024)
                                       : make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, kBK * 128, 1024)
                                       : make_smem_desc(sb + k * 32, 16, 1024);
              umma_f16_2sm(tmem_d, da, db, idesc, (kb > w.kb_begin || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage], 0x3);   // frees the slot in BOTH CTAs
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_2sm(&tfull_bar[acc], 0x3);        // accumulator ready in BOTH CTAs
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ============================ epilogue (both CTAs) ================================
    const int quarter = warp_idx & 3;
    const int col_half = (warp_idx - 2) >> 2;
    const int row = quarter * 32 + lane_id();
    const uint32_t te0 = mapa_cluster(smem_u32(&tempty_bar[0]), 0), te1 = mapa_cluster(smem_u32(&tempty_bar[1]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    WorkItem w;
    for (int item = pair; item < n_items; item += n_pairs) {
      if (!sched.get(item, w)) continue;
      w.m_blk = w.m_blk * 2 + static_cast<int>(rank);   // this CTA's 128-row block
      w.tag = w.tag * 2 + static_cast<int>(rank);
      epi.item_begin(w, row);
      for (int ti = 0; ti < w.n_end - w.n_begin; ++ti) {
        const int n = w.tile_at(ti);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        epi.tile(w, n, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN, row, col_half);
        tc_fence_before();
        mbar_arrive_remote(acc ? te1 : te0);   // accumulator release: no global-memory ordering involved
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      epi.item_end(w, row);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody tears down while the peer may still signal / read
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<S::kTmemCols>(tmem_base);
  }
}

__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* p) {
  const uint32_t a = smem_u32(p);
  return p + ((1024u - (a & 1023u)) & 1023u);
}

}  // namespace uniir
