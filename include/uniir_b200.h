/* uniir_b200.h -- C ABI of the B200-native UniIR hot path (libuniir_b200.so).
 *
 * The reference (TIGER-AI-Lab/UniIR) is 100 % Python and has no FFI of its own: every
 * device op on its hot path is a PyTorch / cuBLAS / FAISS library call.  Each entry point
 * below therefore cites the reference call site(s) whose library call it replaces
 * (paths relative to the reference root; "upstream CLIP" = the un-vendored openai/CLIP
 * dependency, restated in the reference at clip_featurefusion/clip_ff.py).
 *
 * Conventions:
 *   - plain `extern "C"`, no torch types; every pointer is a DEVICE pointer owned by the
 *     caller unless a comment says "host";
 *   - `stream` is a cudaStream_t passed as void*; the call only enqueues work on it
 *     (no hidden synchronisation, no hidden allocation -- scratch is passed in);
 *   - return value: 0 = ok, negative = error (UNIIR_ERR_*); `uniir_last_error()` returns
 *     a thread-local human-readable message for the last failing call;
 *   - row-major everywhere, leading dimensions in ELEMENTS.
 */
#ifndef UNIIR_B200_H_
#define UNIIR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNIIR_ABI_VERSION 1

#define UNIIR_OK 0
#define UNIIR_ERR_INVALID_ARG (-1)
#define UNIIR_ERR_CUDA (-2)
#define UNIIR_ERR_UNSUPPORTED (-3)

#define UNIIR_BF16 0
#define UNIIR_F16 1
#define UNIIR_F32 2

#define UNIIR_ACT_NONE 0
#define UNIIR_ACT_QUICKGELU 1 /* x*sigmoid(1.702x): upstream CLIP QuickGELU */
#define UNIIR_ACT_GELU 2      /* exact erf GELU: BLIP vit.py:39 / med.py:308 */
#define UNIIR_ACT_TANH 3      /* BERT pooler, med.py:505-511 (forward only) */
#define UNIIR_ACT_RELU 4      /* T5 DenseReluDense of the CLIP_FF fusion stack (clip_ff.py:82-98) */

int uniir_abi_version(void);
const char* uniir_last_error(void);
/* number of this library's kernel launches since process start (bench.py "gpu_launches") */
int64_t uniir_launch_count(void);

/* ------------------------------------------------------------------------------------
 * GEMM with fused epilogue on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *   out[M,N] (+)= epilogue( alpha * A . B^T )     A logical [M,K], B logical [N,K]
 * a_mn / b_mn = 0: operand stored [rows, K] (K contiguous);  1: stored [K, rows].
 * Epilogue order: *alpha, +bias[N], (aux_out := value), act, *act'(act_grad_src),
 *                 +residual, store / accumulate / atomic-accumulate (split_k > 1).
 * Replaces: nn.MultiheadAttention in_proj/out_proj, mlp.c_fc/c_proj, visual.conv1 (as
 * im2col GEMM), visual.proj, text_projection (upstream CLIP; restated clip_ff.py:35-59,
 * 148-156) and their autograd backward (dgrad = b_mn=1, wgrad = a_mn=b_mn=1).
 * Constraints: K % 8 == 0 (16 B rows), lda/ldb % 8 == 0, N % 8 == 0, 16 B aligned bases.
 * ---------------------------------------------------------------------------------- */
typedef struct uniir_gemm_desc {
  int32_t M, N, K;
  const void* A; int64_t lda; int32_t a_mn;
  const void* B; int64_t ldb; int32_t b_mn;
  int32_t in_dtype;                 /* UNIIR_BF16 | UNIIR_F16 */
  void* out; int64_t ldo; int32_t out_dtype; /* BF16 | F16 | F32 */
  const float* bias;                /* [N] or NULL */
  int32_t act;                      /* UNIIR_ACT_* */
  const void* residual; int64_t ldr; int32_t res_dtype; /* added last; NULL = none */
  void* aux_out; int64_t ld_aux;    /* optional copy of the pre-activation, in_dtype */
  const void* act_grad_src; int64_t ld_ags; /* optional: value *= act'(src), in_dtype */
  float alpha;
  int32_t accumulate;               /* out += value (out_dtype must be F32) */
  int32_t split_k;                  /* >1: split contraction, atomic fp32 accumulate */
  const float* row_scale;           /* optional fp32 [ceil(M / row_scale_rows)]: (alpha*AB + bias) is multiplied by   */
  int32_t row_scale_rows;           /* row_scale[row / row_scale_rows] before the residual is added: per-item DropPath */
                                    /* of a residual branch (vit.py:151-152, timm).  Only with the fp32-residual       */
                                    /* epilogue: A, B K-major, fp32 residual, fp32 out, no act / aux / accumulate.     */
} uniir_gemm_desc;
int uniir_gemm_epi(const uniir_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------
 * Dense retrieval: exact cosine top-k (FAISS "IDMap,Flat" + METRIC_INNER_PRODUCT over
 * L2-normalised rows).  Replaces faiss.normalize_L2 + GpuIndexFlat.search:
 * common/mbeir_retriever.py:69,76 (index build), :190,:194,:231 (search).
 * Embeddings stay as the stored fp16 values; 1/||row|| is applied in fp32 in the epilogue.
 * ---------------------------------------------------------------------------------- */
/* inv[r] = 1/||x[r,:]||_2 in fp32 (0 for an all-zero row, matching normalize_L2). */
int uniir_row_inv_norm(const void* x, int32_t dtype, int64_t rows, int32_t d, int64_t ld,
                       float* inv, void* stream);
/* Number of candidate-range splits the search kernel will use for (Q, N) and the scratch
 * it needs.  Partial outputs are [Q, splits, k]. */
int uniir_ip_topk_plan(int64_t Q, int64_t N, int32_t k, int32_t* splits, size_t* workspace_bytes);
/* Fused fp16 GEMM + per-row top-k.  part_scores fp32 [Q, splits, k] (descending, cosine),
 * part_idx int32 [Q, splits, k] (candidate row in [0,N), -1 padding).  k <= 64. */
int uniir_ip_topk(const void* q, const float* q_inv, const void* c, const float* c_inv, int64_t Q,
                  int64_t N, int32_t d, int32_t k, int32_t splits, float* part_scores,
                  int32_t* part_idx, void* workspace, size_t workspace_bytes, void* stream);
/* k-way merge of L sorted partial lists per query (splits of one GPU, or shards of W GPUs):
 * in_scores/in_idx element (q, l, j) at  q*q_stride + l*l_stride + j; in_idx is int32
 * (idx_is_64 = 0) or int64 (idx_is_64 = 1); negative idx = padding.
 * list_offset[l] (int64, may be NULL) is added to idx (shard row offset);
 * id_table (int64, may be NULL) maps the resulting row to the user id (add_with_ids,
 * mbeir_retriever.py:103).  Output: fp32 [Q,k] descending + int64 [Q,k]; missing = -1 with
 * score -3.4028235e38 (FAISS convention).  k <= 64, any L.  Replaces the host-side
 * IndexShards merge behind mbeir_retriever.py:206,231. */
int uniir_topk_merge(const float* in_scores, const void* in_idx, int32_t idx_is_64, int64_t Q,
                     int32_t L, int32_t k, int64_t q_stride, int64_t l_stride,
                     const int64_t* list_offset, const int64_t* id_table, float* out_scores,
                     int64_t* out_ids, void* stream);

/* ------------------------------------------------------------------------------------
 * Row-wise kernels around the GEMMs (HBM-bound, 128-bit coalesced).
 * ---------------------------------------------------------------------------------- */
/* y[r,:] = LN(x[row_index ? row_index[r] : r, :]) * gamma + beta, fp32 statistics (two-pass).
 * Replaces upstream CLIP `LayerNorm` (fp32 subclass; ln_pre/ln_1/ln_2/ln_post/ln_final,
 * restated clip_ff.py:49,53,58,154).  y_dtype BF16 (GEMM input) or F32 (in-place ln_pre).
 * mean/rstd [rows] optional (saved for backward). */
int uniir_layernorm_fwd(const float* x, int64_t ldx, const int32_t* row_index, const float* gamma,
                        const float* beta, float eps, int64_t rows, int32_t D, void* y, int64_t ldy,
                        int32_t y_dtype, float* mean, float* rstd, void* stream);
/* dx[src row] (+)= LN backward; dgamma/dbeta += column sums (atomic fp32); dx_bf16 (optional,
 * same row addressing as dx) receives a bf16 copy of the updated dx (next GEMM's operand);
 * dcolsum [D] (optional) += column sums of the updated dx over the processed rows = the bias
 * gradient of the linear layer this dx flows into next (autograd's sum over tokens), fused here
 * so that no separate pass reads dx again. */
int uniir_layernorm_bwd(const void* dy, int64_t lddy, int32_t dy_dtype, const float* x, int64_t ldx,
                        const int32_t* row_index, const float* mean, const float* rstd,
                        const float* gamma, int64_t rows, int32_t D, float* dx, int64_t lddx,
                        int32_t dx_accumulate, float* dgamma, float* dbeta, void* dx_bf16,
                        float* dcolsum, void* stream);
/* T5LayerNorm (RMS norm, no mean subtraction, no bias): y = x * rsqrt(mean(x^2) + eps) * gamma; x fp32 rows at stride
 * ldx, y bf16 or fp32 rows at stride ldy.  transformers T5LayerNorm as used by the CLIP_FF fusion stack. */
int uniir_rmsnorm_fwd(const float* x, int64_t ldx, const float* gamma, float eps, int64_t rows, int32_t D, void* y,
                      int64_t ldy, int32_t y_dtype, void* stream);
/* out[b, :] = mean over the S token rows of x[b, :, :] (x fp32 [B, S, D] contiguous): the mean pooling that ends
 * CLIPFeatureFusion.encode_multimodal_input (clip_ff.py:185-192). */
int uniir_token_mean(const float* x, int32_t B, int32_t S, int32_t D, float* out, void* stream);
/* Backward of uniir_rmsnorm_fwd (rstd recomputed from x; x, dy, dx fp32 [rows, D] contiguous):
 * dx (+)= rstd * (gamma o dy - x^ * mean(gamma o dy o x^)), dgamma += sum over rows of dy o x^. */
int uniir_rmsnorm_bwd(const float* dy, const float* x, const float* gamma, float eps, int64_t rows, int32_t D, float* dx,
                      int32_t dx_accumulate, float* dgamma, void* stream);
/* Backward of uniir_token_mean: dx[b, s, :] = d_out[b, :] / S. */
int uniir_token_mean_bwd(const float* d_out, int32_t B, int32_t S, int32_t D, float* dx, void* stream);
/* visual.conv1 (stride = kernel = patch, no bias; clip_ff.py:36) as im2col:
 * img fp32 [B,3,H,W] -> bf16 [B*(H/P)^2, Kpad], k = c*P*P + py*P + px, zero padded. */
int uniir_patch_im2col(const float* img, int32_t B, int32_t H, int32_t W, int32_t P, int32_t Kpad,
                       void* out, void* stream);
/* x0[b,s,:] = (s==0 ? class_embedding : patch[b,s-1,:]) + positional_embedding[s,:]
 * (clip_ff.py:37-48).  patch bf16 [B*(S-1), D], x0 fp32 [B,S,D]. */
int uniir_vision_embed_fwd(const void* patch, const float* cls, const float* pos, int32_t B, int32_t S,
                           int32_t D, float* x0, void* stream);
int uniir_vision_embed_bwd(const float* dx0, int32_t B, int32_t S, int32_t D, void* d_patch,
                           float* d_cls, float* d_pos, void* stream);
/* x[b,s,:] = token_embedding[ids[b,s]] + positional_embedding[s] (upstream CLIP.encode_text;
 * clip_ff.py:149); eot_row[b] = b*S + argmax_s ids[b,s] (EOT pooling row). */
int uniir_text_embed_fwd(const int32_t* ids, const float* tok, const float* pos, int32_t B, int32_t S,
                         int32_t D, float* x, int32_t* eot_row, void* stream);
int uniir_text_embed_bwd(const float* dx, const int32_t* ids, int32_t B, int32_t S, int32_t D,
                         float* d_tok, float* d_pos, void* stream);
/* Score fusion, clip_sf.py:61-63: out = t*txt_mask[:,None] + v*img_mask[:,None]. */
int uniir_score_fuse_fwd(const float* t, const float* v, const int64_t* mt, const int64_t* mv, int64_t B,
                         int32_t E, float* out, void* stream);
int uniir_score_fuse_bwd(const float* d, const int64_t* mt, const int64_t* mv, int64_t B, int32_t E,
                         void* dt_bf16, void* dv_bf16, void* stream);
/* Score fusion over compacted rows -- the reference encodes "" / a black image for a missing
 * modality and multiplies the result by mask 0 (clip_sf.py:53-63, mbeir_dataset.py:503-504); here a
 * tower only sees the rows whose mask is non-zero and its term is scattered back:
 *   scatter   : out[rows[r],:] += mask[rows[r]] * emb[r,:]                         r < n
 *   gather_bwd: d[r,:] = bf16(mask[rows[r]] * d_out[rows[r],:]) for r < n, 0 for n <= r < n_pad
 * rows int32 [n] (batch row of each compact row), mask int64 [B], emb/out/d_out fp32 [*,E]. */
int uniir_score_fuse_scatter(const float* emb, const int32_t* rows, const int64_t* mask, int32_t n,
                             int32_t E, float* out, void* stream);
int uniir_score_fuse_gather_bwd(const float* d_out, const int32_t* rows, const int64_t* mask, int32_t n,
                                int32_t n_pad, int32_t E, void* d_bf16, void* stream);
/* out[n] += sum_m dy[m,n]  (bias gradients; dy bf16). */
int uniir_colsum_accum(const void* dy, int64_t M, int32_t N, int64_t ld, float* out, void* stream);
int uniir_pad_cast_bf16(const float* src, int64_t rows, int32_t K, int32_t Kpad, void* dst, void* stream);
int uniir_unpad_accum(const float* src, int64_t rows, int32_t K, int32_t Kpad, float* dst, void* stream);
int uniir_cast_bf16(const float* src, int64_t n, void* dst, void* stream);
/* Fused AdamW (torch.optim.AdamW semantics; reference clip_scorefusion/train.py:52-61 param
 * groups, engine.py:41-45 step + zero_grad): p,g,m,v fp32 flat segments; g is multiplied by
 * grad_scale, then zeroed; p_bf16 (optional) receives the bf16 shadow the GEMMs read. */
int uniir_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int32_t step,
                     float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Attention core softmax(QK^T/sqrt(64) [+causal]) V, head_dim 64, bf16, flash-style.
 * qkv bf16 [B*S, 3*H*64] (q | k | v, head h at columns h*64 of each third -- the layout
 * nn.MultiheadAttention.in_proj produces); out bf16 [B*S, H*64]; lse fp32 [B,H,S].
 * Replaces the core of nn.MultiheadAttention in upstream CLIP ResidualAttentionBlock.
 * ---------------------------------------------------------------------------------- */
int uniir_attn_fwd(const void* qkv, int32_t B, int32_t S, int32_t H, int32_t causal, void* out, float* lse,
                   void* stream);
/* General form: separate query and key/value tensors (cross-attention: BLIP med.py:150-232 with
 * encoder_hidden_states) and an optional per-batch valid-key count kv_len[B] (BERT padding mask,
 * med.py:687-688: keys >= kv_len[b] get the reference's -10000 additive mask, i.e. weight 0).
 * q rows b*Sq+s at stride ldq, k/v rows b*Skv+t at stride ldkv (head h at columns h*64), out stride ldo. */
int uniir_attn_fwd_ex(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int32_t B,
                      int32_t Sq, int32_t Skv, int32_t H, int32_t causal, const int32_t* kv_len, void* out,
                      int64_t ldo, float* lse, void* stream);
/* Implementation switch (bit mask): bit 0 = tcgen05 forward (scores of a 128-query tile in TMEM, P staged in
 * shared memory, K/V of the next head prefetched by TMA) wherever Skv <= 272 and not causal; bit 1 = tcgen05
 * backward (attention_bwd_tc.cu: S^T / dP^T of a 128-key tile in TMEM, P^T / dS^T staged in shared memory,
 * dK / dV / dQ accumulated in TMEM) wherever Sq <= 260; bit 2 = take the tcgen05 kernels also for the shapes
 * where they measured slower (causal forward); 0 = the mma.sync head-resident kernels everywhere, which are
 * also the fallback for shapes the tcgen05 kernels do not take.  Default 3; environment UNIIR_ATTN_TC
 * overrides.  Returns the previous mode.  Both families compute the same function; tests compare them. */
int uniir_attn_set_mode(int32_t mode);
/* Diagnostic: device buffer of 3 x 256 int64 that receives SM-clock timestamps of CTA 0's phases in the
 * tcgen05 forward (softmax warp / MMA thread / tail warp; scripts/trace_attn.py prints them); null = off. */
int uniir_attn_set_trace(void* device_buf_3x256_int64);
/* Attention forward with an additive score bias and a caller-chosen scale: softmax(scale * QK^T + bias) V, bias fp32
 * [H, Sq, Skv] shared by all batch rows (T5 relative-position bias of the CLIP_FF fusion encoder, clip_ff.py:82-98,
 * 178-183: transformers' T5Attention adds `position_bias` to unscaled scores, i.e. scale = 1).  Addressing as
 * uniir_attn_fwd_ex; mma.sync kernel (Skv <= 768). */
int uniir_attn_fwd_bias(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int32_t B, int32_t Sq,
                        int32_t Skv, int32_t H, float scale, const float* bias, void* out, int64_t ldo, float* lse,
                        void* stream);
/* Backward of uniir_attn_fwd_bias; dbias fp32 [H, Sq, Skv] += the gradient of the bias (sum over the batch of
 * P o (dP - delta); zero it first).  mma.sync kernels.  delta_ws: B * H * Sq floats. */
int uniir_attn_bwd_bias(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* out,
                        int64_t ldo, const void* dout, int64_t lddo, const float* lse, int32_t B, int32_t Sq, int32_t Skv,
                        int32_t H, float scale, const float* bias, float* dbias, float* delta_ws, void* dq, int64_t lddq,
                        void* dk, void* dv, int64_t lddkv, void* stream);
/* Backward workspace (fp32 elements, 16-byte aligned): delta = rowsum(dO o O), the base-2 LSE and the dS of
 * the key columns beyond the last full 128-key tile, each padded to 16 rows per head. */
int64_t uniir_attn_bwd_workspace_floats(int32_t B, int32_t Sq, int32_t Skv, int32_t H);
int uniir_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, int32_t B,
                   int32_t S, int32_t H, int32_t causal,
                   float* delta_ws /* uniir_attn_bwd_workspace_floats(B, S, S, H) floats */, void* dqkv,
                   void* stream);

/* ------------------------------------------------------------------------------------
 * In-batch contrastive loss of CLIPScoreFusion (clip_sf.py:88-103,134-147), fp32.
 *   normalize -> [host: all-gather p over ranks] -> fwd (one fused kernel: logits tiles,
 *   online softmax, CE mean, argmax accuracy) -> bwd (dQ, d all_p, d logit_scale)
 *   -> [host: reduce-scatter d all_p] -> normalize_bwd.
 * ---------------------------------------------------------------------------------- */
size_t uniir_infonce_workspace_bytes(int32_t bs, int32_t n_cols); /* zero-fill once */
int uniir_infonce_normalize(const float* emb, const int32_t* q_idx, const int32_t* p_idx, int32_t bs,
                            int32_t E, float* qn, float* pn, float* q_norm, float* p_norm, void* stream);
int uniir_infonce_fwd(const float* qn, const float* all_p, int32_t bs, int32_t n_cols, int32_t E,
                      const float* logit_scale_log, int32_t target_offset, float* loss, float* accuracy,
                      float* lse, void* workspace, size_t workspace_bytes, void* stream);
/* G: scratch fp32 [bs, n_cols].  grad_out: device scalar dL (NULL = 1).  d_logit_scale += . */
int uniir_infonce_bwd(const float* qn, const float* all_p, int32_t bs, int32_t n_cols, int32_t E,
                      const float* logit_scale_log, int32_t target_offset, const float* lse,
                      const float* grad_out, float* G, float* d_qn, float* d_all_p,
                      float* d_logit_scale, void* stream);
int uniir_infonce_normalize_bwd(const int32_t* q_idx, const int32_t* p_idx, const float* qn,
                                const float* pn, const float* q_norm, const float* p_norm,
                                const float* d_qn, const float* d_pn, int32_t bs, int32_t E, float* d_emb,
                                void* stream);

/* Hard-negative branch of the loss (clip_sf.py:105-131), forward + backward fused:
 * logits_i = [q.p, q.n_1..q.n_m, q.p x k] * exp(logit_scale), k = min(bs-1, in_batch_neg_num).
 * loss_sum / correct_sum (device scalars, zeroed by the caller) accumulate sum_i(-log_softmax_i[0])
 * and sum_i[argmax_i == 0]; d_emb (zeroed by the caller, [items,E]) receives d(mean loss)/d emb;
 * d_logit_scale += .  n_idx int32 [bs, neg_num]. */
int uniir_hardneg_loss(const float* emb, const int32_t* q_idx, const int32_t* p_idx, const int32_t* n_idx,
                       int32_t bs, int32_t neg_num, int32_t in_batch_neg_num, int32_t E,
                       const float* logit_scale_log, float* loss_sum, float* correct_sum, float* d_emb,
                       float* d_logit_scale, void* stream);

/* Backward of uniir_attn_fwd_ex (same addressing); dq rows at stride lddq, dk/dv rows at stride lddkv;
 * delta_ws: uniir_attn_bwd_workspace_floats(B, Sq, Skv, H) floats. */
int uniir_attn_bwd_ex(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* out,
                      int64_t ldo, const void* dout, int64_t lddo, const float* lse, int32_t B, int32_t Sq,
                      int32_t Skv, int32_t H, int32_t causal, const int32_t* kv_len, float* delta_ws, void* dq,
                      int64_t lddq, void* dk, void* dv, int64_t lddkv, void* stream);
/* Momentum-encoder EMA over a flat segment (blip_ff.py:289-292) + its bf16 shadow. */
int uniir_ema_update(float* p_m, const float* p, int64_t n, float momentum, void* p_m_bf16, void* stream);
/* dz = dy * (1 - y^2): backward of the BERT pooler tanh (med.py:505-511); dz bf16. */
int uniir_tanh_bwd(const float* dy, const float* y, int64_t n, void* dz_bf16, void* stream);
/* fp32 out[M,N] = A[M,K] . B / (*temp); B stored [K,N] (b_kn=1) or [N,K] (b_kn=0).  BLIP logits
 * q @ [pc_m^T | cand_queue] / temp (blip_ff.py:215-216,225-226) and the matching input gradients. */
int uniir_sgemm_temp(const float* A, const float* B, int32_t b_kn, int32_t M, int32_t N, int32_t K,
                     const float* temp, float* out, void* stream);
/* Momentum-distillation soft-label cross entropy (blip_ff.py:164-178, 219-231, 250-252), one direction:
 * S (online logits [bs,n], overwritten with dLoss/dS * gscale), Sm (momentum logits), pc_id [bs], id_all [n]
 * (int64), loss_sum += sum_rows loss, acc_sum (optional) += sum_rows pos[argmax S], d_temp += dLoss/dtemp. */
int uniir_blip_softce(float* S, const float* Sm, const int64_t* pc_id, const int64_t* id_all, int32_t bs, int32_t n,
                      float alpha, float gscale, const float* temp, float* loss_sum, float* acc_sum, float* d_temp,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * Dropout / stochastic depth of the reference's training arms: BERT "MED" hidden dropout and attention-probability
 * dropout (uniir_blip/backbone/med.py:63,97 embeddings; :123,217 attention probabilities; :240,244 and :321,325 the
 * dense outputs; config hidden_dropout_prob = attention_probs_dropout_prob = 0.1), timm DropPath of the ViT-L blocks
 * (backbone/vit.py:136,151-152,217; blip.py:253 drop_path_rate 0.1) and the T5 fusion stack of CLIP_FF
 * (clip_ff.py:82-98, T5Config dropout_rate 0.1).  torch draws these masks from its stateful Philox stream; here a
 * mask is a pure function of (ctx[0] = seed of the step, site, element index within the WHOLE batch):
 *     keep = mix32(idx ^ mix32(seed + site * 0x9E3779B1)) >= floor(p * 2^32)
 * with mix32 the 32-bit finaliser x^=x>>17; x*=0xed5ad4bb; x^=x>>11; x*=0xac4c1b51; x^=x>>15; x*=0x31848bab; x^=x>>14 and
 * idx = ctx[1] * per_item + i (mod 2^32), ctx[1] = index of this micro-batch's first item in the batch.  Nothing is
 * saved: the gradient-cache re-forward and the backward regenerate the same bits under any micro-batch tiling, and
 * a captured CUDA graph replays with a new seed because ctx (uint32[2]) is device memory.
 * ---------------------------------------------------------------------------------- */
/* out[i] = (keep(i) ? x[i] / (1 - p) : 0) + (residual ? residual[i] : 0), i < n; x / out fp32 or bf16 (UNIIR_F32 /
 * UNIIR_BF16), residual fp32; out may alias x or residual.  Forward: the dense output before the residual + LayerNorm;
 * backward: the same call on the gradient (no residual). */
int uniir_dropout(const void* x, int32_t x_dtype, const float* residual, void* out, int32_t out_dtype, int64_t n,
                  const uint32_t* ctx, uint32_t site, int64_t per_item, float p, void* stream);
/* keep[i] (uint8, 1 = kept) of the same mask: what the tests and an external checker read. */
int uniir_dropout_mask(const uint32_t* ctx, uint32_t site, int64_t per_item, int64_t n, float p, uint8_t* keep,
                       void* stream);
/* DropPath: scales[s * mb + b] = keep ? 1 / (1 - p_sites[s]) : 0 for site site_base + s and batch item ctx[1] + b
 * (one draw per sample and residual branch, timm scale_by_keep).  p_sites: device fp32 [n_sites]. */
int uniir_droppath_scales(const uint32_t* ctx, uint32_t site_base, int32_t n_sites, int32_t mb, const float* p_sites,
                          float* scales, void* stream);
/* x[r, :] *= scales[r / rows_per_scale], bf16 in place: the branch gradient under DropPath. */
int uniir_row_scale(void* x_bf16, int64_t rows, int32_t cols, int64_t ld, const float* scales, int32_t rows_per_scale,
                    void* stream);
/* softmax(scale * QK^T [+ bias] [key-padding / causal mask]) with dropout on the probabilities, then V: element
 * (b, h, q, key) of the batch-wide [items, H, Sq, Skv] tensor is element ((ctx[1] + b) * H + h) * Sq * Skv + q * Skv + key
 * of site `site`.  Addressing as uniir_attn_fwd_ex; bias (fp32 [H, Sq, Skv]) may be NULL; lse is that of the undropped
 * softmax.  mma.sync head-resident kernels (Skv <= 768). */
int uniir_attn_fwd_drop(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, int32_t B, int32_t Sq,
                        int32_t Skv, int32_t H, int32_t causal, const int32_t* kv_len, float scale, const float* bias,
                        void* out, int64_t ldo, float* lse, const uint32_t* drop_ctx, uint32_t site, float p_drop,
                        void* stream);
/* Backward of uniir_attn_fwd_drop with the same (ctx, site, p); dbias as uniir_attn_bwd_bias (NULL iff bias is NULL);
 * delta_ws: B * H * Sq floats. */
int uniir_attn_bwd_drop(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* out,
                        int64_t ldo, const void* dout, int64_t lddo, const float* lse, int32_t B, int32_t Sq, int32_t Skv,
                        int32_t H, int32_t causal, const int32_t* kv_len, float scale, const float* bias, float* dbias,
                        float* delta_ws, void* dq, int64_t lddq, void* dk, void* dv, int64_t lddkv,
                        const uint32_t* drop_ctx, uint32_t site, float p_drop, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIIR_B200_H_ */
