// uniir_gemm_epi: tcgen05 GEMM with fused bias / activation / activation-grad / residual /
// accumulate epilogue -- host-side validation, tile/epilogue selection and dispatch.  Kernels are
// in gemm_kernel.cuh, instantiated in gemm_inst_*.cu.  Contract: include/uniir_b200.h.
#include "gemm_kernel.cuh"

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
  UNIIR_CHECK_ARG(!d->row_scale || (d->row_scale_rows > 0 && d->split_k <= 1),
                  "uniir_gemm_epi: row_scale needs row_scale_rows > 0 and no split_k");
  UNIIR_CHECK_ARG(!(atomic || d->accumulate) || d->out_dtype == UNIIR_F32,
                  "uniir_gemm_epi: accumulate/split_k need an fp32 output");
  UNIIR_CHECK_ARG(!atomic || (!d->bias && !d->residual && !d->aux_out && !d->act_grad_src &&
                              d->act == UNIIR_ACT_NONE),
                  "uniir_gemm_epi: split_k supports only the plain accumulate epilogue");

  // Tile width.  A 128x128 tile reads 8 KB of operands per 64-cycle UMMA = 128 B/clk, the whole
  // shared-memory bandwidth of the SM, so 128x256 (96 B/clk) is preferred unless the problem is
  // too narrow or 256-wide tiles would leave most of the last wave empty.
  const int m_blocks = (d->M + kBM - 1) / kBM;
  const int sms = num_sms();
  int bn = 256;
  {
    const long long t256 = 1LL * m_blocks * ((d->N + 255) / 256) * (atomic ? d->split_k : 1);
    const long long t128 = 1LL * m_blocks * ((d->N + 127) / 128) * (atomic ? d->split_k : 1);
    const double c256 = static_cast<double>((t256 + sms - 1) / sms) * 1.0;
    const double c128 = static_cast<double>((t128 + sms - 1) / sms) * 0.65;
    if (d->N <= 128 || c128 < c256) bn = 128;
    if (const char* e = getenv("UNIIR_GEMM_BN")) { const int v = atoi(e); if (v == 128 || (v == 256 && d->N > 128)) bn = v; }
  }

  // CTA-pair path (256 x 256 tiles, cta_group::2): large, non-split problems whose epilogue has a pair
  // instantiation.  UNIIR_GEMM_2CTA=0 disables, =1 forces where legal.
  bool pair = false;
  {
    const char* e = getenv("UNIIR_GEMM_2CTA");
    const int mode = e ? atoi(e) : 1;
    const long long tiles2 = 1LL * ((d->M + 255) / 256) * ((d->N + 255) / 256);
    pair = mode != 0 && bn == 256 && !atomic && d->M >= 256 && !(d->a_mn && !d->b_mn) &&
           (mode == 1 ? tiles2 >= sms / 2 : true);
  }

  GemmTensorMaps maps;
  int rc;
  const int b_box = pair ? 128 : bn;
  if (!d->a_mn)
    rc = make_tmap_2d(&maps.a, d->A, d->in_dtype, d->K, d->M, d->lda, kBK, kBM);
  else
    rc = make_tmap_2d(&maps.a, d->A, d->in_dtype, d->M, d->K, d->lda, 64, kBK);
  if (rc) return rc;
  if (!d->b_mn)
    rc = make_tmap_2d(&maps.b, d->B, d->in_dtype, d->K, d->N, d->ldb, kBK, b_box);
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
  if (d->row_scale) { p.ags = d->row_scale; p.ld_ags = d->row_scale_rows; }   // EPI_RES32_RS borrows these (gemm_kernel.cuh)

  TileSched s;
  s.m_blocks = m_blocks;
  s.n_blocks = (d->N + bn - 1) / bn;
  s.kb_total = (d->K + kBK - 1) / kBK;
  s.split_k = atomic ? d->split_k : 1;
  if (s.split_k > s.kb_total) s.split_k = s.kb_total;
  s.group_m = 16;
  const bool bf16 = d->in_dtype == UNIIR_BF16;

  // ---- pick a statically specialised epilogue when the descriptor matches one -------------
  const bool has_res = d->residual != nullptr, has_aux = d->aux_out != nullptr, has_ags = d->act_grad_src != nullptr;
  const bool has_act = d->act != UNIIR_ACT_NONE;
  const bool out32 = d->out_dtype == UNIIR_F32;
  const bool has_rs = d->row_scale != nullptr;
  int epi = EPI_DYN;
  if (has_rs) {
    UNIIR_CHECK_ARG(!d->a_mn && !d->b_mn && has_res && d->res_dtype == UNIIR_F32 && out32 && !has_aux && !has_ags && !has_act &&
                    !d->accumulate && !atomic,
                    "uniir_gemm_epi: row_scale is implemented for the fp32 residual epilogue only (A, B K-major, fp32 out)");
    epi = EPI_RES32_RS;
  } else if (!d->a_mn && !has_res && !has_aux && !has_ags && !has_act && !d->accumulate && !atomic)
    epi = out32 ? EPI_STORE32 : EPI_STORE16;
  else if (!d->a_mn && !d->b_mn && !has_res && !has_ags && has_act && !out32 && !d->accumulate && !atomic)
    epi = has_aux ? EPI_ACT_AUX : EPI_ACT;
  else if (!d->a_mn && d->b_mn && has_ags && has_act && !has_res && !has_aux && !out32 && !d->bias)
    epi = EPI_AGS;
  else if (!d->a_mn && !d->b_mn && has_res && d->res_dtype == UNIIR_F32 && out32 && !has_aux && !has_ags &&
           !has_act && !d->accumulate && !atomic)
    epi = EPI_RES32;
  else if (d->a_mn && d->b_mn && out32 && !has_res && !has_aux && !has_ags && !has_act && !d->bias &&
           (atomic || d->accumulate))
    epi = atomic ? EPI_ATOMIC : EPI_ACC32;
  if (getenv("UNIIR_GEMM_DYN") && !has_rs) epi = EPI_DYN;

  if (pair) {
    TileSched s2 = s;
    s2.m_blocks = (d->M + 255) / 256;
    s2.group_m = 8;
#define UNIIR_GO2(A_, B_, E_) return launch_gemm_2cta<A_, B_, E_>(maps, p, s2, bf16, stream)
    switch (epi) {
      case EPI_STORE16: if (d->b_mn) UNIIR_GO2(false, true, EPI_STORE16); else UNIIR_GO2(false, false, EPI_STORE16);
      case EPI_STORE32: if (!d->b_mn) UNIIR_GO2(false, false, EPI_STORE32); break;
      case EPI_ACT: UNIIR_GO2(false, false, EPI_ACT);
      case EPI_ACT_AUX: UNIIR_GO2(false, false, EPI_ACT_AUX);
      case EPI_AGS: UNIIR_GO2(false, true, EPI_AGS);
      case EPI_RES32: UNIIR_GO2(false, false, EPI_RES32);
      case EPI_RES32_RS: UNIIR_GO2(false, false, EPI_RES32_RS);
      case EPI_ACC32: UNIIR_GO2(true, true, EPI_ACC32);
      default: break;   // no pair instantiation: fall through to the single-CTA kernels
    }
#undef UNIIR_GO2
    // rebuild the B map for the single-CTA box
    if (!d->b_mn) {
      rc = make_tmap_2d(&maps.b, d->B, d->in_dtype, d->K, d->N, d->ldb, kBK, bn);
      if (rc) return rc;
    }
  }
#define UNIIR_GO(BN_, A_, B_, E_) return launch_gemm<BN_, A_, B_, E_>(maps, p, s, bf16, stream)
#define UNIIR_BY_BN(A_, B_, E_) do { if (bn == 256) UNIIR_GO(256, A_, B_, E_); else UNIIR_GO(128, A_, B_, E_); } while (0)
  switch (epi) {
    case EPI_STORE16: if (d->b_mn) UNIIR_BY_BN(false, true, EPI_STORE16); else UNIIR_BY_BN(false, false, EPI_STORE16);
    case EPI_STORE32: if (d->b_mn) UNIIR_BY_BN(false, true, EPI_STORE32); else UNIIR_BY_BN(false, false, EPI_STORE32);
    case EPI_ACT: UNIIR_BY_BN(false, false, EPI_ACT);
    case EPI_ACT_AUX: UNIIR_BY_BN(false, false, EPI_ACT_AUX);
    case EPI_AGS: UNIIR_BY_BN(false, true, EPI_AGS);
    case EPI_RES32: UNIIR_BY_BN(false, false, EPI_RES32);
    case EPI_RES32_RS: UNIIR_BY_BN(false, false, EPI_RES32_RS);
    case EPI_ACC32: UNIIR_BY_BN(true, true, EPI_ACC32);
    case EPI_ATOMIC: UNIIR_BY_BN(true, true, EPI_ATOMIC);
    default: break;
  }
  if (!d->a_mn && !d->b_mn) UNIIR_BY_BN(false, false, EPI_DYN);
  if (!d->a_mn && d->b_mn) UNIIR_BY_BN(false, true, EPI_DYN);
  if (d->a_mn && !d->b_mn) UNIIR_BY_BN(true, false, EPI_DYN);
  UNIIR_BY_BN(true, true, EPI_DYN);
#undef UNIIR_GO
#undef UNIIR_BY_BN
}
