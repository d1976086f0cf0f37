// tcgen05 GEMM kernel + compile-time specialised epilogues.  Included by the gemm_*.cu
// translation units (explicit instantiations, built in parallel) and by gemm.cu (dispatch).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "gemm_sm100.cuh"
#include "host_util.cuh"

namespace uniir {

struct GemmEpiParams {
  int M, N;
  void* out; long long ldo; int out_dtype;
  const float* bias;
  int act;
  const void* residual; long long ldr; int res_dtype;
  void* aux_out; long long ld_aux;
  const void* ags; long long ld_ags;
  int in_dtype;
  float alpha;
  int accumulate;
  int atomic;
  // NB: this struct is a by-value kernel parameter of every GEMM instantiation, and its SIZE steers nvcc's register
  // allocation of all of them (measured: +16 bytes took the CTA-pair residual kernel from 162 to 140 registers and the
  // training step from 574 to 506 pairs/s).  EPI_RES32_RS therefore borrows two fields it never uses otherwise:
  // ags = row_scale (fp32 [ceil(M / rows)]), ld_ags = rows per scale.
};

struct TileSched {
  int m_blocks, n_blocks, kb_total, split_k, group_m;
  __device__ __forceinline__ int num_items() const { return m_blocks * n_blocks * split_k; }
  __device__ __forceinline__ bool get(int item, WorkItem& w) const {
    const int ks = item % split_k;
    const int t = item / split_k;
    // grouped order: walk n within bands of `group_m` row-blocks so that the tiles in flight
    // at any moment form a compact rectangle (operand reuse in L2)
    const int band = group_m * n_blocks;
    const int g = t / band;
    const int first_m = g * group_m;
    const int rows = min(group_m, m_blocks - first_m);
    const int r = t - g * band;
    w.m_blk = first_m + r % rows;
    w.n_begin = r / rows;
    w.n_end = w.n_begin + 1;
    w.rot = 0;
    const int per = (kb_total + split_k - 1) / split_k;
    w.kb_begin = ks * per;
    w.kb_end = min(kb_total, w.kb_begin + per);
    return w.kb_begin < w.kb_end;
  }
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid(1.702 x) = 0.5 + 0.5 tanh(0.851 x): one MUFU op instead of exp + full-precision divide
__device__ __forceinline__ float qg_sigmoid(float x) { return fmaf(0.5f, tanh_approx(0.851f * x), 0.5f); }
__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == UNIIR_ACT_QUICKGELU) return x * qg_sigmoid(x);
  if (act == UNIIR_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (act == UNIIR_ACT_TANH) return tanhf(x);
  if (act == UNIIR_ACT_RELU) return fmaxf(x, 0.f);
  return x;
}
__device__ __forceinline__ float act_bwd(float x, int act) {
  if (act == UNIIR_ACT_QUICKGELU) {
    const float s = qg_sigmoid(x);
    return s * fmaf(1.702f * x, 1.f - s, 1.f);
  }
  if (act == UNIIR_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  }
  if (act == UNIIR_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  return 1.f;
}

// 8 consecutive 16-bit elements <-> 8 floats
__device__ __forceinline__ void load8_16b(const void* p, int dtype, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == UNIIR_BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void store8_16b(void* p, int dtype, const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (dtype == UNIIR_BF16) {
      const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&b);
    } else {
      const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kGemmNE = 8;  // epilogue warps of the plain GEMM kernel

enum : int {
  EPI_DYN = 0,   // all features selected at run time
  EPI_STORE16,   // (+bias) -> 16-bit store                      qkv fwd, plain dgrad
  EPI_STORE32,   // (+bias) -> fp32 store                        pooled projections
  EPI_ACT,       // +bias, activation -> 16-bit                  c_fc forward (inference)
  EPI_ACT_AUX,   // +bias, save pre-activation, activation       c_fc forward (training)
  EPI_AGS,       // * act'(saved pre-activation) -> 16-bit       c_proj dgrad
  EPI_RES32,     // +bias, + fp32 residual -> fp32               out_proj / c_proj forward
  EPI_ACC32,     // out += value (fp32)                          wgrad
  EPI_ATOMIC,    // atomicAdd(out, value) (fp32, split-K)        wgrad, narrow outputs
  EPI_RES32_RS,  // (+bias) * row_scale[row / rs_rows] + fp32 residual -> fp32   out_proj / c_proj forward under DropPath
};

template <int BN, int EPI>
struct GemmEpilogue {
  static constexpr bool DYN = EPI == EPI_DYN;
  __device__ __forceinline__ bool do_aux() const { return DYN ? p.aux_out != nullptr : EPI == EPI_ACT_AUX; }
  __device__ __forceinline__ bool do_act() const { return DYN ? (p.act != UNIIR_ACT_NONE && !p.ags) : (EPI == EPI_ACT || EPI == EPI_ACT_AUX); }
  __device__ __forceinline__ bool do_ags() const { return DYN ? p.ags != nullptr : EPI == EPI_AGS; }
  __device__ __forceinline__ bool res32() const { return DYN ? (p.residual && p.res_dtype == UNIIR_F32) : (EPI == EPI_RES32 || EPI == EPI_RES32_RS); }
  __device__ __forceinline__ bool row_scaled() const { return EPI == EPI_RES32_RS; }   // static variant only
  __device__ __forceinline__ bool res16() const { return DYN ? (p.residual && p.res_dtype != UNIIR_F32) : false; }
  __device__ __forceinline__ bool out32() const { return DYN ? p.out_dtype == UNIIR_F32 : (EPI == EPI_STORE32 || EPI == EPI_RES32 || EPI == EPI_RES32_RS || EPI == EPI_ACC32 || EPI == EPI_ATOMIC); }
  __device__ __forceinline__ bool atomic() const { return DYN ? p.atomic != 0 : EPI == EPI_ATOMIC; }
  __device__ __forceinline__ bool acc() const { return DYN ? (p.accumulate && !p.atomic) : EPI == EPI_ACC32; }

  const GemmEpiParams& p;
  float* stage_base;  // kGemmNE x 4 KB warp-private staging tiles in shared memory
  __device__ GemmEpilogue(const GemmEpiParams& p_, float* st) : p(p_), stage_base(st) {}
  __device__ __forceinline__ void item_begin(const WorkItem&, int) {}
  __device__ __forceinline__ void item_end(const WorkItem&, int) {}

  // Epilogue inputs of one unit (8 consecutive output columns of one row), fetched early so
  // that their global-load latency overlaps the TMEM->smem transposition of the chunk.
  struct UnitIn {
    uint4 a, b;  // fp32 residual / accumulate target (a,b = 8 floats)  OR  a = 16-bit residual, b = ags
  };
  __device__ __forceinline__ bool f32_side_input() const { return res32() || (acc() && !res16()); }
  __device__ __forceinline__ void prefetch(UnitIn& u, long long grow, int col) const {
    if (grow >= p.M || col >= p.N) return;
    if (f32_side_input()) {
      const float* rp = res32() ? static_cast<const float*>(p.residual) + grow * p.ldr + col
                                : static_cast<const float*>(p.out) + grow * p.ldo + col;
      u.a = *reinterpret_cast<const uint4*>(rp);
      u.b = *reinterpret_cast<const uint4*>(rp + 4);
    } else {
      if (res16())
        u.a = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col);
      if (do_ags())
        u.b = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.ags) + grow * p.ld_ags + col);
    }
  }
  __device__ __forceinline__ static void unpack8(const uint4& u, int dtype, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dtype == UNIIR_BF16) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
      } else {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
      }
    }
  }
  __device__ __forceinline__ static void add8(float (&v)[8], const UnitIn& in) {
    v[0] += __uint_as_float(in.a.x); v[1] += __uint_as_float(in.a.y);
    v[2] += __uint_as_float(in.a.z); v[3] += __uint_as_float(in.a.w);
    v[4] += __uint_as_float(in.b.x); v[5] += __uint_as_float(in.b.y);
    v[6] += __uint_as_float(in.b.z); v[7] += __uint_as_float(in.b.w);
  }

  // scale / bias / (aux copy) / act / act-grad / residual / accumulate, then store
  __device__ __forceinline__ void unit(float (&v)[8], const UnitIn& in, long long grow, int col) {
    if (grow >= p.M || col >= p.N) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
    if (p.bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
      const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (do_aux()) store8_16b(static_cast<uint16_t*>(p.aux_out) + grow * p.ld_aux + col, p.in_dtype, v);
    if (do_act()) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_fwd(v[j], p.act);
    }
    if (do_ags()) {
      float s[8];
      if (f32_side_input())  // rare combination (EPI_DYN only): prefetch slots hold the fp32 input
        unpack8(*reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.ags) + grow * p.ld_ags + col), p.in_dtype, s);
      else
        unpack8(in.b, p.in_dtype, s);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= act_bwd(s[j], p.act);
    }
    if (row_scaled()) {
      const float rs = __ldg(static_cast<const float*>(p.ags) + static_cast<int>(grow) / static_cast<int>(p.ld_ags));
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= rs;
    }
    if (res32()) add8(v, in);
    if (res16()) {
      float s[8];
      unpack8(in.a, p.res_dtype, s);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += s[j];
    }
    if (out32()) {
      float* op = static_cast<float*>(p.out) + grow * p.ldo + col;
      if (atomic()) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(op + j, v[j]);
      } else {
        if (acc()) {
          if (!res32() && !res16()) {
            add8(v, in);
          } else {
            const float4 a = *reinterpret_cast<const float4*>(op);
            const float4 b = *reinterpret_cast<const float4*>(op + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
          }
        }
        *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      store8_16b(static_cast<uint16_t*>(p.out) + grow * p.ldo + col, p.out_dtype, v);
    }
  }

  // One 32-row x 32-column accumulator chunk.  TMEM hands each lane one ROW (32 columns);
  // issuing global I/O in that shape touches 32 cache lines per instruction.  The chunk is
  // therefore transposed through a 4 KB warp-private staging tile (16 B chunks XOR-swizzled
  // by row, conflict-free both ways) so that 4 adjacent lanes cover one row's 32 columns and
  // every global load/store instruction covers 8 rows x one contiguous segment.
  __device__ __forceinline__ void chunk(uint32_t tmem_addr, float* stage, long long grow0, int gcol0) {
    const int lane = threadIdx.x & 31;
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem_addr, r);
    // side inputs (residual / act-grad source / accumulate target) are requested while the TMEM
    // load is in flight, in the coalesced (row = u/4, 8-column group = u%4) mapping
    UnitIn in[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = i * 32 + lane;
      prefetch(in[i], grow0 + (u >> 2), gcol0 + (u & 3) * 8);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<uint4*>(stage + lane * 32 + ((c ^ (lane & 7)) << 2)) =
          make_uint4(r[c * 4], r[c * 4 + 1], r[c * 4 + 2], r[c * 4 + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int u = i * 32 + lane, row = u >> 2, g = u & 3;
      const float4 lo = *reinterpret_cast<const float4*>(stage + row * 32 + (((2 * g) ^ (row & 7)) << 2));
      const float4 hi = *reinterpret_cast<const float4*>(stage + row * 32 + (((2 * g + 1) ^ (row & 7)) << 2));
      float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      unit(v, in[i], grow0 + row, gcol0 + g * 8);
    }
    __syncwarp();
  }

  // This warp owns columns [col_half*W, (col_half+1)*W) of the tile, W = BN / (NE/4); the two
  // epilogue warps that share an SMSP overlap each other's TMEM / global-load latencies.
  __device__ __forceinline__ void tile(const WorkItem& w, int n_blk, uint32_t tmem_acc, int row,
                                       int col_half) {
    constexpr int W = BN / (kGemmNE / 4);
    constexpr int NC = W / 32;
    const int lane = threadIdx.x & 31;
    const long long grow0 = static_cast<long long>(w.m_blk) * kBM + (row - lane);
    float* stage = stage_base + ((threadIdx.x >> 5) - 2) * 1024;
    const int c0 = col_half * W;
#pragma unroll 1
    for (int c = 0; c < NC; ++c)
      chunk(tmem_acc + c0 + c * 32, stage, grow0, n_blk * BN + c0 + c * 32);
  }
};

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm(const GemmTensorMaps& maps, const GemmEpiParams& p, const TileSched& s, bool bf16,
                cudaStream_t stream);

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(num_threads(kGemmNE), 1)
gemm_epi_kernel(const __grid_constant__ GemmTensorMaps maps, const GemmEpiParams p,
                const TileSched sched, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  using SM = GemmSmem<BN>;
  GemmEpilogue<BN, EPI> epi(p, reinterpret_cast<float*>(smem + SM::kRingBytes + SM::kBarBytes));
  gemm_body<BN, A_MN, B_MN, kGemmNE>(&maps.a, &maps.b, idesc, sched, epi, smem);
}

template <int BN, bool A_MN, bool B_MN, int EPI>
int launch_gemm_impl(const GemmTensorMaps& maps, const GemmEpiParams& p, const TileSched& s,
                     bool bf16, cudaStream_t stream) {
  using SM = GemmSmem<BN>;
  const int smem_bytes = SM::kRingBytes + SM::kBarBytes + kGemmNE * 4096 + 1024;
  auto kern = gemm_epi_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    UNIIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int items = s.m_blocks * s.n_blocks * s.split_k;
  const int grid = items < num_sms() ? items : num_sms();
  const uint32_t idesc = make_idesc_f16(bf16, A_MN, B_MN, kBM, BN);
  kern<<<grid, num_threads(kGemmNE), smem_bytes, stream>>>(maps, p, s, idesc);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

#define UNIIR_INSTANTIATE_GEMM(BN_, A_, B_, EPI_)                                                  \
  template <>                                                                                      \
  int launch_gemm<BN_, A_, B_, EPI_>(const GemmTensorMaps& m, const GemmEpiParams& p,              \
                                     const TileSched& s, bool bf16, cudaStream_t st) {             \
    return launch_gemm_impl<BN_, A_, B_, EPI_>(m, p, s, bf16, st);                                 \
  }

// Static epilogue variants: the runtime-flag epilogue (EPI_DYN) compiles every feature into a
// 4x-unrolled unit and the kernel grows to ~80 KB of SASS, which thrashes the instruction
// cache (ncu: stall_no_inst on the epilogue FMULs).  The variants below cover the patterns the
// CLIP towers use; anything else takes EPI_DYN.

// ---- CTA-pair kernel (256 x 256 tiles, cta_group::2) -------------------------------------
template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(num_threads(kGemmNE), 1)
gemm_epi_kernel_2cta(const __grid_constant__ GemmTensorMaps maps, const GemmEpiParams p,
                     const TileSched sched, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  using SM = GemmSmem2<256>;
  GemmEpilogue<256, EPI> epi(p, reinterpret_cast<float*>(smem + SM::kRingBytes + SM::kBarBytes));
  gemm_body_2cta<256, A_MN, B_MN, kGemmNE>(&maps.a, &maps.b, idesc, sched, epi, smem);
}

template <bool A_MN, bool B_MN, int EPI>
int launch_gemm_2cta(const GemmTensorMaps& maps, const GemmEpiParams& p, const TileSched& s, bool bf16,
                     cudaStream_t stream);

template <bool A_MN, bool B_MN, int EPI>
int launch_gemm_2cta_impl(const GemmTensorMaps& maps, const GemmEpiParams& p, const TileSched& s, bool bf16,
                          cudaStream_t stream) {
  using SM = GemmSmem2<256>;
  const int smem_bytes = SM::kRingBytes + SM::kBarBytes + kGemmNE * 4096 + 1024;
  auto kern = gemm_epi_kernel_2cta<A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    UNIIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int items = s.m_blocks * s.n_blocks * s.split_k;
  int pairs = num_sms() / 2;
  if (items < pairs) pairs = items;
  const uint32_t idesc = make_idesc_f16(bf16, A_MN, B_MN, 2 * kBM, 256);
  kern<<<2 * pairs, num_threads(kGemmNE), smem_bytes, stream>>>(maps, p, s, idesc);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

#define UNIIR_INSTANTIATE_GEMM_2CTA(A_, B_, EPI_)                                                   \
  template <>                                                                                       \
  int launch_gemm_2cta<A_, B_, EPI_>(const GemmTensorMaps& m, const GemmEpiParams& p,               \
                                     const TileSched& s, bool bf16, cudaStream_t st) {              \
    return launch_gemm_2cta_impl<A_, B_, EPI_>(m, p, s, bf16, st);                                  \
  }

}  // namespace uniir
