// In-batch contrastive (InfoNCE) loss of CLIPScoreFusion, reference
// clip_scorefusion/clip_sf.py:88-103,134-147:
//   q, p = normalize(emb[q_idx]), normalize(emb[p_idx]);  all_p = all_gather(p)
//   logits = q @ all_p.T * exp(logit_scale);  loss = CE(logits, rank*bs + arange(bs)).mean()
//   accuracy = mean(argmax(logits) == target)
// Forward is ONE kernel: 64x64 logit tiles (fp32 FFMA, operands L2-resident) -> per-tile
// online-softmax partials -> the last CTA of each row band folds them into LSE / loss /
// accuracy, the last CTA overall publishes the means.  Logits are never written.
// Backward recomputes the tiles into G = (softmax - onehot) * gout / bs (8 MB, L2-resident)
// and runs the two small products dQ = s*G@P, dP = s*G^T@Q.
// fp32 on purpose: the problem is 25.8 GFLOP per step against ~7e15 for the towers, and
// fp32 keeps the loss within 1e-5 of the reference arithmetic.
#include <float.h>

#include "host_util.cuh"

namespace uniir {

constexpr int kT = 64;   // tile edge
constexpr int kKC = 32;  // k chunk

// acc[i][j] = sum_k A(m0 + ty*4+i, k) * B(n0 + tx*4+j, k);  A(m,k) at A[m*a_rs + k*a_cs]
__device__ __forceinline__ void sgemm_tile(const float* __restrict__ A, long long a_rs, long long a_cs,
                                           const float* __restrict__ B, long long b_rs, long long b_cs,
                                           int M, int N, int K, int m0, int n0, float (&acc)[4][4],
                                           float (*As)[kT + 1], float (*Bs)[kT + 1]) {
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kKC) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m, k;
      if (a_cs == 1) { k = t & 31; m = (t >> 5) + 8 * i; } else { m = t & 63; k = (t >> 6) + 4 * i; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
      if (b_cs == 1) { k = t & 31; m = (t >> 5) + 8 * i; } else { m = t & 63; k = (t >> 6) + 4 * i; }
      const int gn = n0 + m, gk2 = k0 + k;
      Bs[k][m] = (gn < N && gk2 < K) ? B[gn * b_rs + gk2 * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kKC; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
}

// qn/pn = normalize(emb[idx]) (eps 1e-12 like F.normalize), norms saved for backward
__global__ void nce_normalize_kernel(const float* __restrict__ emb, const int* __restrict__ q_idx,
                                     const int* __restrict__ p_idx, int bs, int E,
                                     float* __restrict__ qn, float* __restrict__ pn,
                                     float* __restrict__ q_norm, float* __restrict__ p_norm) {
  const int r = blockIdx.x >> 1, which = blockIdx.x & 1;
  const float* src = emb + static_cast<long long>(which ? p_idx[r] : q_idx[r]) * E;
  float* dst = (which ? pn : qn) + static_cast<long long>(r) * E;
  __shared__ float sh[33];
  float s = 0.f;
  for (int c = threadIdx.x; c < E; c += blockDim.x) s += src[c] * src[c];
  // block reduce
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sh[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    sh[32] = fmaxf(sqrtf(t), 1e-12f);
  }
  __syncthreads();
  const float nrm = sh[32];
  for (int c = threadIdx.x; c < E; c += blockDim.x) dst[c] = src[c] / nrm;
  if (threadIdx.x == 0) (which ? p_norm : q_norm)[r] = nrm;
}

// d_emb[idx[r]] = (d_n - n * <n, d_n>) / ||x||  for the q rows and the p rows
__global__ void nce_normalize_bwd_kernel(const int* __restrict__ q_idx, const int* __restrict__ p_idx,
                                         const float* __restrict__ qn, const float* __restrict__ pn,
                                         const float* __restrict__ q_norm, const float* __restrict__ p_norm,
                                         const float* __restrict__ d_qn, const float* __restrict__ d_pn,
                                         int bs, int E, float* __restrict__ d_emb) {
  const int r = blockIdx.x >> 1, which = blockIdx.x & 1;
  const float* n = (which ? pn : qn) + static_cast<long long>(r) * E;
  const float* d = (which ? d_pn : d_qn) + static_cast<long long>(r) * E;
  float* dst = d_emb + static_cast<long long>(which ? p_idx[r] : q_idx[r]) * E;
  const float nrm = (which ? p_norm : q_norm)[r];
  __shared__ float sh[33];
  float s = 0.f;
  for (int c = threadIdx.x; c < E; c += blockDim.x) s += n[c] * d[c];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sh[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
    sh[32] = t;
  }
  __syncthreads();
  const float dot = sh[32];
  for (int c = threadIdx.x; c < E; c += blockDim.x) dst[c] = (d[c] - n[c] * dot) / nrm;
}

struct NceWs {          // device workspace layout (floats/ints), zero-initialised once
  float* part_m;        // [bs][nct]
  float* part_l;        // [bs][nct]
  float* part_av;       // [bs][nct] arg-max value
  int* part_ai;         // [bs][nct] arg-max column
  float* tgt;           // [bs] target logit
  float* sums;          // [2] loss sum, correct sum
  int* band_count;      // [nrt]
  int* done_count;      // [1]
};

__global__ void __launch_bounds__(256)
nce_fwd_kernel(const float* __restrict__ qn, const float* __restrict__ allp, int bs, int n, int E,
               const float* __restrict__ logit_scale_log, int target_offset, NceWs ws, int nct,
               float* __restrict__ lse_out, float* __restrict__ loss_out, float* __restrict__ acc_out) {
  __shared__ float As[kKC][kT + 1];
  __shared__ float Bs[kKC][kT + 1];
  __shared__ float red_m[kT][17];
  __shared__ float red_v[kT][17];
  __shared__ int red_i[kT][17];
  __shared__ int is_last;
  const int ct = blockIdx.x, rt = blockIdx.y;
  const int m0 = rt * kT, n0 = ct * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float scale = __expf(*logit_scale_log);
  float acc[4][4];
  sgemm_tile(qn, E, 1, allp, E, 1, bs, n, E, m0, n0, acc, As, Bs);
  // per-thread partials over its 4 columns, for each of its 4 rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    float mx = -INFINITY;
    int mi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      const float s = (col < n) ? acc[i][j] * scale : -INFINITY;
      acc[i][j] = s;
      if (s > mx) { mx = s; mi = col; }
      if (row < bs && col == target_offset + row) ws.tgt[row] = s;
    }
    red_m[ty * 4 + i][tx] = mx;
    red_i[ty * 4 + i][tx] = mi;
  }
  __syncthreads();
  // row max / argmax over the 16 column groups (first index wins ties, like torch.max)
  if (t < kT) {
    float mx = -INFINITY;
    int mi = 0x7fffffff;
    for (int g = 0; g < 16; ++g) {
      const float v = red_m[t][g];
      const int ix = red_i[t][g];
      if (v > mx || (v == mx && ix < mi)) { mx = v; mi = ix; }
    }
    red_m[t][16] = mx;
    red_i[t][16] = mi;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float mx = red_m[ty * 4 + i][16];
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) l += (acc[i][j] == -INFINITY) ? 0.f : __expf(acc[i][j] - mx);
    red_v[ty * 4 + i][tx] = l;
  }
  __syncthreads();
  if (t < kT && m0 + t < bs) {
    float l = 0.f;
    for (int g = 0; g < 16; ++g) l += red_v[t][g];
    const long long o = static_cast<long long>(m0 + t) * nct + ct;
    ws.part_m[o] = red_m[t][16];
    ws.part_l[o] = l;
    ws.part_av[o] = red_m[t][16];
    ws.part_ai[o] = red_i[t][16];
  }
  // ---- last CTA of this row band folds the partials ----
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = (atomicAdd(&ws.band_count[rt], 1) == nct - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float loss_r = 0.f, corr_r = 0.f;
  if (t < kT && m0 + t < bs) {
    const int row = m0 + t;
    float mx = -INFINITY, av = -INFINITY;
    int ai = 0x7fffffff;
    for (int c = 0; c < nct; ++c) {
      const long long o = static_cast<long long>(row) * nct + c;
      const float pm = __ldcg(&ws.part_m[o]);
      mx = fmaxf(mx, pm);
      const float v = __ldcg(&ws.part_av[o]);
      const int ix = __ldcg(&ws.part_ai[o]);
      if (v > av || (v == av && ix < ai)) { av = v; ai = ix; }
    }
    float l = 0.f;
    for (int c = 0; c < nct; ++c) {
      const long long o = static_cast<long long>(row) * nct + c;
      l += __ldcg(&ws.part_l[o]) * __expf(__ldcg(&ws.part_m[o]) - mx);
    }
    const float lse = mx + logf(l);
    lse_out[row] = lse;
    loss_r = lse - __ldcg(&ws.tgt[row]);
    corr_r = (ai == target_offset + row) ? 1.f : 0.f;
  }
  __syncthreads();
  if (t < kT) { red_v[t][0] = loss_r; red_v[t][1] = corr_r; }
  __syncthreads();
  if (t == 0) {
    float ls = 0.f, cs = 0.f;
    for (int i = 0; i < kT; ++i) { ls += red_v[i][0]; cs += red_v[i][1]; }
    atomicAdd(&ws.sums[0], ls);
    atomicAdd(&ws.sums[1], cs);
    ws.band_count[rt] = 0;  // leave the workspace clean for the next launch
    __threadfence();
    if (atomicAdd(ws.done_count, 1) == static_cast<int>(gridDim.y) - 1) {
      __threadfence();
      *loss_out = __ldcg(&ws.sums[0]) / bs;
      *acc_out = __ldcg(&ws.sums[1]) / bs;
      ws.sums[0] = 0.f;
      ws.sums[1] = 0.f;
      *ws.done_count = 0;
    }
  }
}

// G[row, col] = (exp(s - lse[row]) - [col == target]) * gscale ; d_logit_scale += sum G*s
__global__ void __launch_bounds__(256)
nce_bwd_g_kernel(const float* __restrict__ qn, const float* __restrict__ allp, int bs, int n, int E,
                 const float* __restrict__ logit_scale_log, int target_offset,
                 const float* __restrict__ lse, const float* __restrict__ gout, float inv_bs,
                 float* __restrict__ G, float* __restrict__ d_logit_scale) {
  __shared__ float As[kKC][kT + 1];
  __shared__ float Bs[kKC][kT + 1];
  __shared__ float red[8];
  const int m0 = blockIdx.y * kT, n0 = blockIdx.x * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float scale = __expf(*logit_scale_log);
  const float gs = (gout ? *gout : 1.f) * inv_bs;
  float acc[4][4];
  sgemm_tile(qn, E, 1, allp, E, 1, bs, n, E, m0, n0, acc, As, Bs);
  float dsc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= bs) continue;
    const float l = lse[row];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= n) continue;
      const float s = acc[i][j] * scale;
      const float g = (__expf(s - l) - (col == target_offset + row ? 1.f : 0.f)) * gs;
      G[static_cast<long long>(row) * n + col] = g;
      dsc += g * s;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dsc += __shfl_xor_sync(0xffffffffu, dsc, o);
  if ((t & 31) == 0) red[t >> 5] = dsc;
  __syncthreads();
  if (t == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(d_logit_scale, s);
  }
}

// C[M,N] = scale * A . B^T with general strides (fp32), used for dQ and dP
__global__ void __launch_bounds__(256)
nce_sgemm_kernel(const float* __restrict__ A, long long a_rs, long long a_cs, const float* __restrict__ B,
                 long long b_rs, long long b_cs, int M, int N, int K,
                 const float* __restrict__ logit_scale_log, float* __restrict__ C, int recip = 0) {
  __shared__ float As[kKC][kT + 1];
  __shared__ float Bs[kKC][kT + 1];
  const int m0 = blockIdx.y * kT, n0 = blockIdx.x * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float scale = recip ? 1.f / *logit_scale_log : __expf(*logit_scale_log);
  float acc[4][4];
  sgemm_tile(A, a_rs, a_cs, B, b_rs, b_cs, M, N, K, m0, n0, acc, As, Bs);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col < N) C[static_cast<long long>(row) * N + col] = acc[i][j] * scale;
    }
  }
}


// ---------------------------------------------------------------------------------
// Hard-negative branch, reference clip_sf.py:105-131.  Per query row i the logits are
//   [ q.p , q.n_1 .. q.n_m , (q.p repeated k times) ] * exp(logit_scale)
// (the reference's "in-batch negatives" index `p.unsqueeze(1).expand(-1,bs,-1)[mask]`,
// whose [i,j] entry is p_i, so the k extra columns repeat the row's own positive -- restated
// faithfully), loss = mean(-log_softmax[:,0]), accuracy = mean(argmax == 0).
// One CTA per row: normalise q, p, n_j, form the logits, softmax, and write the gradient with
// respect to the raw embeddings (normalisation backward included) and to logit_scale.
// ---------------------------------------------------------------------------------
constexpr int kMaxNeg = 32;

__device__ __forceinline__ float cta_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(128)
nce_hardneg_kernel(const float* __restrict__ emb, const int* __restrict__ q_idx, const int* __restrict__ p_idx,
                   const int* __restrict__ n_idx, int bs, int neg, int k, int E,
                   const float* __restrict__ logit_scale_log, float* __restrict__ loss_sum,
                   float* __restrict__ correct_sum, float* __restrict__ d_emb,
                   float* __restrict__ d_logit_scale) {
  __shared__ float sh[8];
  __shared__ float s_dot[kMaxNeg + 1], s_nrm[kMaxNeg + 2], s_g[kMaxNeg + 1];
  const int i = blockIdx.x;
  const float scale = __expf(*logit_scale_log);
  const float* q = emb + static_cast<long long>(q_idx[i]) * E;
  // norms of q, p, n_j and raw dots q.p, q.n_j
  for (int j = -1; j <= neg; ++j) {  // j=-1: q ; j=0: p ; j>=1: negatives
    const float* x = (j < 0) ? q : emb + static_cast<long long>(j == 0 ? p_idx[i] : n_idx[i * neg + j - 1]) * E;
    float nn = 0.f, dd = 0.f;
    for (int c = threadIdx.x; c < E; c += blockDim.x) { nn += x[c] * x[c]; dd += x[c] * q[c]; }
    nn = cta_sum(nn, sh);
    dd = cta_sum(dd, sh);
    if (threadIdx.x == 0) {
      s_nrm[j + 1] = fmaxf(sqrtf(nn), 1e-12f);
      if (j >= 0) s_dot[j] = dd;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float qn = s_nrm[0];
    float z[kMaxNeg + 1], mx = -INFINITY;
    for (int j = 0; j <= neg; ++j) { z[j] = s_dot[j] / (qn * s_nrm[j + 1]) * scale; mx = fmaxf(mx, z[j]); }
    float Z = (1.f + k) * __expf(z[0] - mx);
    for (int j = 1; j <= neg; ++j) Z += __expf(z[j] - mx);
    const float lse = mx + logf(Z);
    atomicAdd(loss_sum, lse - z[0]);
    bool best = true;
    for (int j = 1; j <= neg; ++j) best = best && (z[0] >= z[j]);   // torch.max: first index wins ties
    atomicAdd(correct_sum, best ? 1.f : 0.f);
    float dls = 0.f;
    for (int j = 0; j <= neg; ++j) {
      const float P = __expf(z[j] - lse);
      const float g = (j == 0) ? ((1.f + k) * P - 1.f) : P;       // dL_i/dz_j (copies folded into j=0)
      s_g[j] = g / bs;                                             // mean over rows
      dls += s_g[j] * z[j];
    }
    atomicAdd(d_logit_scale, dls);
  }
  __syncthreads();
  // cos_j = <qh, xh_j>;  d qh = sum_j g_j s xh_j ; d xh_j = g_j s qh ; then normalise-backward
  const float qn = s_nrm[0];
  float dq_dot = 0.f;  // <qh, d qh>
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j <= neg; ++j) {
      const float* x = emb + static_cast<long long>(j == 0 ? p_idx[i] : n_idx[i * neg + j - 1]) * E;
      acc += s_g[j] * scale * x[c] / s_nrm[j + 1];
    }
    dq_dot += acc * q[c] / qn;
  }
  dq_dot = cta_sum(dq_dot, sh);
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j <= neg; ++j) {
      const float* x = emb + static_cast<long long>(j == 0 ? p_idx[i] : n_idx[i * neg + j - 1]) * E;
      acc += s_g[j] * scale * x[c] / s_nrm[j + 1];
    }
    atomicAdd(d_emb + static_cast<long long>(q_idx[i]) * E + c, (acc - (q[c] / qn) * dq_dot) / qn);
  }
  for (int j = 0; j <= neg; ++j) {
    const long long row = (j == 0) ? p_idx[i] : n_idx[i * neg + j - 1];
    const float* x = emb + row * E;
    const float xn = s_nrm[j + 1], gs = s_g[j] * scale;
    const float cosv = s_dot[j] / (qn * xn);   // <xh, d xh> = gs * cos
    for (int c = threadIdx.x; c < E; c += blockDim.x)
      atomicAdd(d_emb + row * E + c, gs * (q[c] / qn - (x[c] / xn) * cosv) / xn);
  }
}


// ---------------------------------------------------------------------------------
// BLIP momentum-distillation loss (reference blip_ff.py:204-231, 250-252), one CTA per row:
//   S  = x . M / temp   (online logits, given)      Sm = x_m . M / temp  (momentum logits, given)
//   t  = alpha * softmax(Sm) + (1-alpha) * pos / sum(pos),   pos_j = [pc_id[row] == id_all[j]]
//   loss_row = -sum_j t_j log_softmax(S)_j = lse(S) - sum_j t_j S_j
//   G  (in place of S) = (softmax(S) - t) * gscale        d temp += sum_j G_j * (-S_j / temp)
//   accuracy term (q2pc direction only): pos[argmax_j S_j]
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float cta_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t = fmaxf(t, sh[i]);
  return t;
}

__global__ void __launch_bounds__(256)
blip_softce_kernel(float* __restrict__ S, const float* __restrict__ Sm, const long long* __restrict__ pc_id,
                   const long long* __restrict__ id_all, int n, float alpha, float gscale,
                   const float* __restrict__ temp, float* __restrict__ loss_sum, float* __restrict__ acc_sum,
                   float* __restrict__ d_temp) {
  __shared__ float sh[8];
  __shared__ int sh_i[8];
  const int row = blockIdx.x;
  float* s = S + static_cast<long long>(row) * n;
  const float* sm = Sm + static_cast<long long>(row) * n;
  const long long my = pc_id[row];
  float mx = -INFINITY, mxm = -INFINITY, npos = 0.f;
  int amax = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = s[j];
    if (v > mx) { mx = v; amax = j; }
    mxm = fmaxf(mxm, sm[j]);
    npos += (id_all[j] == my) ? 1.f : 0.f;
  }
  // arg-max with first-index tie break (torch.max)
  const float gmx = cta_max(mx, sh);
  int cand = (mx == gmx) ? amax : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh_i[threadIdx.x >> 5] = cand;
  __syncthreads();
  int gam = 0x7fffffff;
  for (int i = 0; i < (blockDim.x >> 5); ++i) gam = min(gam, sh_i[i]);
  const float gmxm = cta_max(mxm, sh);
  npos = cta_sum(npos, sh);
  float z = 0.f, zm = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) { z += __expf(s[j] - gmx); zm += __expf(sm[j] - gmxm); }
  z = cta_sum(z, sh);
  zm = cta_sum(zm, sh);
  const float lse = gmx + logf(z), inv_zm = 1.f / zm, inv_np = 1.f / npos, tmp = *temp;
  float dot = 0.f, dt = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = s[j];
    const float t = alpha * __expf(sm[j] - gmxm) * inv_zm + (1.f - alpha) * ((id_all[j] == my) ? inv_np : 0.f);
    dot += t * v;
    const float g = (__expf(v - lse) - t) * gscale;
    s[j] = g;
    dt -= g * v / tmp;
  }
  dot = cta_sum(dot, sh);
  dt = cta_sum(dt, sh);
  if (threadIdx.x == 0) {
    atomicAdd(loss_sum, lse - dot);
    atomicAdd(d_temp, dt);
    if (acc_sum) atomicAdd(acc_sum, (id_all[gam] == my) ? 1.f : 0.f);
  }
}

static int nce_carve(void* workspace, size_t bytes, int bs, int nct, int nrt, NceWs* ws) {
  char* p = static_cast<char*>(workspace);
  const size_t per = static_cast<size_t>(bs) * nct;
  size_t off = 0;
  ws->sums = reinterpret_cast<float*>(p + off); off += 16;
  ws->done_count = reinterpret_cast<int*>(p + off); off += 16;
  ws->band_count = reinterpret_cast<int*>(p + off); off += ((static_cast<size_t>(nrt) * 4 + 15) / 16) * 16;
  ws->tgt = reinterpret_cast<float*>(p + off); off += ((static_cast<size_t>(bs) * 4 + 15) / 16) * 16;
  ws->part_m = reinterpret_cast<float*>(p + off); off += per * 4;
  ws->part_l = reinterpret_cast<float*>(p + off); off += per * 4;
  ws->part_av = reinterpret_cast<float*>(p + off); off += per * 4;
  ws->part_ai = reinterpret_cast<int*>(p + off); off += per * 4;
  return off <= bytes ? 0 : -1;
}

}  // namespace uniir

using namespace uniir;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" size_t uniir_infonce_workspace_bytes(int32_t bs, int32_t n_cols) {
  const size_t nct = (n_cols + kT - 1) / kT, nrt = (bs + kT - 1) / kT;
  return 32 + ((nrt * 4 + 15) / 16) * 16 + ((static_cast<size_t>(bs) * 4 + 15) / 16) * 16 +
         static_cast<size_t>(bs) * nct * 16 + 64;
}

extern "C" int uniir_infonce_normalize(const float* emb, const int32_t* q_idx, const int32_t* p_idx,
                                       int32_t bs, int32_t E, float* qn, float* pn, float* q_norm,
                                       float* p_norm, void* stream) {
  UNIIR_CHECK_ARG(emb && q_idx && p_idx && qn && pn && q_norm && p_norm && bs > 0 && E > 0,
                  "uniir_infonce_normalize: bad args");
  nce_normalize_kernel<<<2 * bs, 256, 0, ST(stream)>>>(emb, q_idx, p_idx, bs, E, qn, pn, q_norm, p_norm);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_infonce_fwd(const float* qn, const float* all_p, int32_t bs, int32_t n_cols, int32_t E,
                                 const float* logit_scale_log, int32_t target_offset, float* loss,
                                 float* accuracy, float* lse, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  UNIIR_CHECK_ARG(qn && all_p && logit_scale_log && loss && accuracy && lse && workspace && bs > 0 &&
                      n_cols > 0 && E > 0,
                  "uniir_infonce_fwd: bad args");
  UNIIR_CHECK_ARG(target_offset >= 0 && target_offset + bs <= n_cols,
                  "uniir_infonce_fwd: targets [%d,%d) outside the %d gathered columns", target_offset,
                  target_offset + bs, n_cols);
  const int nct = (n_cols + kT - 1) / kT, nrt = (bs + kT - 1) / kT;
  NceWs ws;
  UNIIR_CHECK_ARG(nce_carve(workspace, workspace_bytes, bs, nct, nrt, &ws) == 0,
                  "uniir_infonce_fwd: workspace too small");
  nce_fwd_kernel<<<dim3(nct, nrt), 256, 0, ST(stream)>>>(qn, all_p, bs, n_cols, E, logit_scale_log,
                                                          target_offset, ws, nct, lse, loss, accuracy);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_infonce_bwd(const float* qn, const float* all_p, int32_t bs, int32_t n_cols, int32_t E,
                                 const float* logit_scale_log, int32_t target_offset, const float* lse,
                                 const float* grad_out, float* G, float* d_qn, float* d_all_p,
                                 float* d_logit_scale, void* stream) {
  UNIIR_CHECK_ARG(qn && all_p && logit_scale_log && lse && G && d_qn && d_all_p && d_logit_scale,
                  "uniir_infonce_bwd: null pointer");
  const int nct = (n_cols + kT - 1) / kT, nrt = (bs + kT - 1) / kT;
  nce_bwd_g_kernel<<<dim3(nct, nrt), 256, 0, ST(stream)>>>(qn, all_p, bs, n_cols, E, logit_scale_log,
                                                            target_offset, lse, grad_out, 1.f / bs, G,
                                                            d_logit_scale);
  // dQ[bs,E] = s * G[bs,n] . P[n,E]      : A = G (K contiguous), B(e,k) = P[k*E + e]
  nce_sgemm_kernel<<<dim3((E + kT - 1) / kT, nrt), 256, 0, ST(stream)>>>(
      G, n_cols, 1, all_p, 1, E, bs, E, n_cols, logit_scale_log, d_qn);
  // dP[n,E]  = s * G^T[n,bs] . Q[bs,E]   : A(c,k) = G[k*n + c], B(e,k) = Q[k*E + e]
  nce_sgemm_kernel<<<dim3((E + kT - 1) / kT, nct), 256, 0, ST(stream)>>>(
      G, 1, n_cols, qn, 1, E, n_cols, E, bs, logit_scale_log, d_all_p);
  count_launch(3);
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_infonce_normalize_bwd(const int32_t* q_idx, const int32_t* p_idx, const float* qn,
                                           const float* pn, const float* q_norm, const float* p_norm,
                                           const float* d_qn, const float* d_pn, int32_t bs, int32_t E,
                                           float* d_emb, void* stream) {
  UNIIR_CHECK_ARG(q_idx && p_idx && qn && pn && q_norm && p_norm && d_qn && d_pn && d_emb,
                  "uniir_infonce_normalize_bwd: null pointer");
  nce_normalize_bwd_kernel<<<2 * bs, 256, 0, ST(stream)>>>(q_idx, p_idx, qn, pn, q_norm, p_norm, d_qn,
                                                            d_pn, bs, E, d_emb);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_hardneg_loss(const float* emb, const int32_t* q_idx, const int32_t* p_idx,
                                  const int32_t* n_idx, int32_t bs, int32_t neg_num,
                                  int32_t in_batch_neg_num, int32_t E, const float* logit_scale_log,
                                  float* loss_sum, float* correct_sum, float* d_emb,
                                  float* d_logit_scale, void* stream) {
  UNIIR_CHECK_ARG(emb && q_idx && p_idx && n_idx && logit_scale_log && loss_sum && correct_sum && d_emb &&
                      d_logit_scale && bs > 0 && E > 0,
                  "uniir_hardneg_loss: bad args");
  UNIIR_CHECK_ARG(neg_num >= 1 && neg_num <= kMaxNeg, "uniir_hardneg_loss: neg_num must be in [1,%d]", kMaxNeg);
  const int k = in_batch_neg_num < bs - 1 ? in_batch_neg_num : bs - 1;  // clip_sf.py:110
  nce_hardneg_kernel<<<bs, 128, 0, ST(stream)>>>(emb, q_idx, p_idx, n_idx, bs, neg_num, k < 0 ? 0 : k, E,
                                                  logit_scale_log, loss_sum, correct_sum, d_emb, d_logit_scale);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

/* out[M,N] = (A[M,K] . B)/(*temp) with B given as [K,N] (b_kn=1) or [N,K] (b_kn=0), fp32 */
extern "C" int uniir_sgemm_temp(const float* A, const float* B, int32_t b_kn, int32_t M, int32_t N, int32_t K,
                                const float* temp, float* out, void* stream) {
  UNIIR_CHECK_ARG(A && B && temp && out && M > 0 && N > 0 && K > 0, "uniir_sgemm_temp: bad args");
  nce_sgemm_kernel<<<dim3((N + kT - 1) / kT, (M + kT - 1) / kT), 256, 0, ST(stream)>>>(
      A, K, 1, B, b_kn ? 1 : K, b_kn ? N : 1, M, N, K, temp, out, 1);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}

extern "C" int uniir_blip_softce(float* S, const float* Sm, const int64_t* pc_id, const int64_t* id_all, int32_t bs,
                                 int32_t n, float alpha, float gscale, const float* temp, float* loss_sum,
                                 float* acc_sum, float* d_temp, void* stream) {
  UNIIR_CHECK_ARG(S && Sm && pc_id && id_all && temp && loss_sum && d_temp && bs > 0 && n > 0, "uniir_blip_softce: bad args");
  blip_softce_kernel<<<bs, 256, 0, ST(stream)>>>(S, Sm, reinterpret_cast<const long long*>(pc_id),
                                                 reinterpret_cast<const long long*>(id_all), n, alpha, gscale, temp,
                                                 loss_sum, acc_sum, d_temp);
  count_launch();
  UNIIR_CUDA_OK(cudaGetLastError());
  return UNIIR_OK;
}
