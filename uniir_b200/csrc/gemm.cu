// uniir_gemm_epi: tcgen05 GEMM with fused bias / activation / activation-grad / residual /
// accumulate epilogue.  See gemm_sm100.cuh for the mainloop and include/uniir_b200.h for
// the contract.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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
    const int per = (kb_total + split_k - 1) / split_k;
    w.kb_begin = ks * per;
    w.kb_end = min(kb_total, w.kb_begin + per);
    return w.kb_begin < w.kb_end;
  }
};

__device__ __forceinline__ float act_fwd(float x, int act) {
  if (act == UNIIR_ACT_QUICKGELU) return x / (1.f + __expf(-1.702f * x));
  if (act == UNIIR_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  return x;
}
__device__ __forceinline__ float act_bwd(float x, int act) {
  if (act == UNIIR_ACT_QUICKGELU) {
    const float s = 1.f / (1.f + __expf(-1.702f * x));
    return s * (1.f + 1.702f * x * (1.f - s));
  }
  if (act == UNIIR_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
    return cdf + x * 0.3989422804014327f * __expf(-0.5f * x * x);
  }
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

template <int BN>
struct GemmEpilogue {
  const GemmEpiParams& p;
  __device__ explicit GemmEpilogue(const GemmEpiParams& p_) : p(p_) {}
  __device__ __forceinline__ void item_begin(const WorkItem&, int) {}
  __device__ __forceinline__ void item_end(const WorkItem&, int) {}

  __device__ __forceinline__ void tile(const WorkItem& w, int n_blk, uint32_t tmem_acc, int row) {
    const long long grow = static_cast<long long>(w.m_blk) * kBM + row;
    const bool row_ok = grow < p.M;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_acc + c * 32, r);
      tmem_ld_wait();
      const int gcol0 = n_blk * BN + c * 32;
      if (!row_ok || gcol0 >= p.N) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = gcol0 + g * 8;
        if (col >= p.N) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]) * p.alpha;
        if (p.bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
          const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (p.aux_out)
          store8_16b(static_cast<uint16_t*>(p.aux_out) + grow * p.ld_aux + col, p.in_dtype, v);
        if (p.act != UNIIR_ACT_NONE && !p.ags) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = act_fwd(v[j], p.act);
        }
        if (p.ags) {
          float s[8];
          load8_16b(static_cast<const uint16_t*>(p.ags) + grow * p.ld_ags + col, p.in_dtype, s);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= act_bwd(s[j], p.act);
        }
        if (p.residual) {
          if (p.res_dtype == UNIIR_F32) {
            const float* rp = static_cast<const float*>(p.residual) + grow * p.ldr + col;
            const float4 a = *reinterpret_cast<const float4*>(rp);
            const float4 b = *reinterpret_cast<const float4*>(rp + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
          } else {
            float s[8];
            load8_16b(static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col, p.res_dtype, s);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += s[j];
          }
        }
        if (p.out_dtype == UNIIR_F32) {
          float* op = static_cast<float*>(p.out) + grow * p.ldo + col;
          if (p.atomic) {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(op + j, v[j]);
          } else {
            if (p.accumulate) {
              const float4 a = *reinterpret_cast<const float4*>(op);
              const float4 b = *reinterpret_cast<const float4*>(op + 4);
              v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
              v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
            }
            *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        } else {
          store8_16b(static_cast<uint16_t*>(p.out) + grow * p.ldo + col, p.out_dtype, v);
        }
      }
    }
  }
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_epi_kernel(const __grid_constant__ GemmTensorMaps maps, const GemmEpiParams p,
                const TileSched sched, const uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  GemmEpilogue<BN> epi(p);
  gemm_body<BN, A_MN, B_MN>(&maps.a, &maps.b, idesc, sched, epi, smem);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const GemmTensorMaps& maps, const GemmEpiParams& p, const TileSched& s,
                       bool bf16, cudaStream_t stream) {
  using SM = GemmSmem<BN>;
  const int smem_bytes = SM::kRingBytes + SM::kBarBytes + 1024;
  auto kern = gemm_epi_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    UNIIR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  const int items = s.m_blocks * s.n_blocks * s.split_k;
  const int grid = items < num_sms() ? items : num_sms();
  const uint32_t idesc = make_idesc_f16(bf16, A_MN, B_MN, kBM, BN);
  kern<<<grid, kNumThreads, smem_bytes, stream>>>(maps, p, s, idesc);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

}  // namespace uniir

using namespace uniir;

extern "C" int uniir_gemm_epi(const uniir_gemm_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UNIIR_CHECK_ARG(d && d->A && d->B && d->out, "uniir_gemm_epi: null operand");
  UNIIR_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "uniir_gemm_epi: bad shape %d %d %d", d->M,
                  d->N, d->K);
  UNIIR_CHECK_ARG(d->in_dtype == UNIIR_BF16 || d->in_dtype == UNIIR_F16,
                  "uniir_gemm_epi: in_dtype must be bf16/fp16");
  UNIIR_CHECK_ARG(d->N % 8 == 0 && d->K % 8 == 0 && (d->a_mn || d->b_mn ? d->M % 8 == 0 || !d->a_mn : true),
                  "uniir_gemm_epi: N, K must be multiples of 8 (N=%d K=%d)", d->N, d->K);
  UNIIR_CHECK_ARG(d->ldo % 4 == 0, "uniir_gemm_epi: ldo must be a multiple of 4");
  const bool atomic = d->split_k > 1;
  UNIIR_CHECK_ARG(!(atomic || d->accumulate) || d->out_dtype == UNIIR_F32,
                  "uniir_gemm_epi: accumulate/split_k need an fp32 output");
  UNIIR_CHECK_ARG(!atomic || (!d->bias && !d->residual && !d->aux_out && !d->act_grad_src &&
                              d->act == UNIIR_ACT_NONE),
                  "uniir_gemm_epi: split_k supports only the plain accumulate epilogue");

  // Tile width: 256 when it does not cost a wave, else 128.
  const int m_blocks = (d->M + kBM - 1) / kBM;
  const int sms = num_sms();
  int bn = 256;
  {
    const long long t256 = 1LL * m_blocks * ((d->N + 255) / 256) * (atomic ? d->split_k : 1);
    const long long t128 = 1LL * m_blocks * ((d->N + 127) / 128) * (atomic ? d->split_k : 1);
    const long long w256 = (t256 + sms - 1) / sms * 2;  // cost in 128-wide units
    const long long w128 = (t128 + sms - 1) / sms;
    if (d->N <= 128 || w128 < w256) bn = 128;
  }

  GemmTensorMaps maps;
  int rc;
  if (!d->a_mn)
    rc = make_tmap_2d(&maps.a, d->A, d->in_dtype, d->K, d->M, d->lda, kBK, kBM);
  else
    rc = make_tmap_2d(&maps.a, d->A, d->in_dtype, d->M, d->K, d->lda, 64, kBK);
  if (rc) return rc;
  if (!d->b_mn)
    rc = make_tmap_2d(&maps.b, d->B, d->in_dtype, d->K, d->N, d->ldb, kBK, bn);
  else
    rc = make_tmap_2d(&maps.b, d->B, d->in_dtype, d->N, d->K, d->ldb, 64, kBK);
  if (rc) return rc;

  GemmEpiParams p;
  p.M = d->M; p.N = d->N;
  p.out = d->out; p.ldo = d->ldo; p.out_dtype = d->out_dtype;
  p.bias = d->bias; p.act = d->act;
  p.residual = d->residual; p.ldr = d->ldr; p.res_dtype = d->res_dtype;
  p.aux_out = d->aux_out; p.ld_aux = d->ld_aux;
  p.ags = d->act_grad_src; p.ld_ags = d->ld_ags;
  p.in_dtype = d->in_dtype;
  p.alpha = d->alpha;
  p.accumulate = d->accumulate;
  p.atomic = atomic ? 1 : 0;

  TileSched s;
  s.m_blocks = m_blocks;
  s.n_blocks = (d->N + bn - 1) / bn;
  s.kb_total = (d->K + kBK - 1) / kBK;
  s.split_k = atomic ? d->split_k : 1;
  if (s.split_k > s.kb_total) s.split_k = s.kb_total;
  s.group_m = 16;
  const bool bf16 = d->in_dtype == UNIIR_BF16;

#define UNIIR_DISPATCH(BN_)                                                               \
  if (!d->a_mn && !d->b_mn) return launch_gemm<BN_, false, false>(maps, p, s, bf16, stream); \
  if (!d->a_mn && d->b_mn) return launch_gemm<BN_, false, true>(maps, p, s, bf16, stream);   \
  if (d->a_mn && !d->b_mn) return launch_gemm<BN_, true, false>(maps, p, s, bf16, stream);   \
  return launch_gemm<BN_, true, true>(maps, p, s, bf16, stream);
  if (bn == 256) { UNIIR_DISPATCH(256) }
  UNIIR_DISPATCH(128)
#undef UNIIR_DISPATCH
}
