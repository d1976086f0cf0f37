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
        for (int n = w.n_begin; n < w.n_end; ++n) {
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
        for (int n = w.n_begin; n < w.n_end; ++n) {
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
      for (int n = w.n_begin; n < w.n_end; ++n) {
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

__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* p) {
  const uint32_t a = smem_u32(p);
  return p + ((1024u - (a & 1023u)) & 1023u);
}

}  // namespace uniir
