"""Torch-tensor front-ends for the C-ABI kernels.  PyTorch is plumbing here: it owns
device memory and the stream; all arithmetic happens in libuniir_b200.so."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L

_DT = {torch.bfloat16: L.BF16, torch.float16: L.F16, torch.float32: L.F32}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.UniirError("uniir_b200 ops need CUDA tensors (no CPU fallback)")


def _ptr(t):
    return None if t is None else t.data_ptr()


_GEMM_PROF = None
_GEMM_EVENTS = False
_GRAPH_LAUNCHES = 0


def graph_launches():
    """uniir kernels executed through CUDA-graph replays (each replay re-runs the captured launches)."""
    return _GRAPH_LAUNCHES


def add_graph_launches(n):
    global _GRAPH_LAUNCHES
    _GRAPH_LAUNCHES += n


def set_gemm_profiler(sink, events=False):
    """sink: list that receives one descriptor dict per GEMM launch (shape + epilogue flags), or None.
    events=True: every launch is also bracketed by two CUDA events on the launching stream (entry["events"]); after
    a synchronize, `e0.elapsed_time(e1)` is that launch's duration IN ITS PLACE in the step -- same predecessor
    kernels, same L2 contents -- which is what a roofline against the sustained peak has to be measured on."""
    global _GEMM_PROF, _GEMM_EVENTS
    _GEMM_PROF = sink
    _GEMM_EVENTS = bool(events) and sink is not None


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, act=L.ACT_NONE,
         residual=None, aux_out=None, act_grad_src=None, alpha=1.0, accumulate=False, split_k=1, row_scale=None,
         row_scale_rows=1):
    """out[M,N] (+)= epilogue(alpha * A . B^T); see include/uniir_b200.h.
    row_scale (fp32 [ceil(M / row_scale_rows)]): the value is multiplied by row_scale[row // row_scale_rows] before the
    residual is added (DropPath of a residual branch).

    a: [M,K] (a_mn False) or [K,M] (a_mn True);  b: [N,K] or [K,N] likewise."""
    _require_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float16)
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype or a.dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    d.B, d.ldb, d.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    d.in_dtype = _DT[a.dtype]
    d.out, d.ldo, d.out_dtype = out.data_ptr(), out.stride(0), _DT[out.dtype]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    d.bias = _ptr(bias)
    d.act = act
    if residual is not None:
        assert residual.shape == (M, N) and residual.stride(1) == 1
        d.residual, d.ldr, d.res_dtype = residual.data_ptr(), residual.stride(0), _DT[residual.dtype]
    if aux_out is not None:
        assert aux_out.shape == (M, N) and aux_out.dtype == a.dtype and aux_out.stride(1) == 1
        d.aux_out, d.ld_aux = aux_out.data_ptr(), aux_out.stride(0)
    if act_grad_src is not None:
        assert act_grad_src.shape == (M, N) and act_grad_src.dtype == a.dtype
        d.act_grad_src, d.ld_ags = act_grad_src.data_ptr(), act_grad_src.stride(0)
    d.alpha = alpha
    d.accumulate = int(accumulate)
    d.split_k = split_k
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous() and row_scale.numel() * row_scale_rows >= M
        d.row_scale, d.row_scale_rows = row_scale.data_ptr(), int(row_scale_rows)
    if _GEMM_PROF is not None:
        _GEMM_PROF.append(dict(M=M, N=N, K=K, a_mn=bool(a_mn), b_mn=bool(b_mn), dtype=a.dtype,
                               out_dtype=out.dtype, bias=bias is not None, act=act,
                               residual=None if residual is None else residual.dtype,
                               aux=aux_out is not None, ags=act_grad_src is not None,
                               accumulate=bool(accumulate), split_k=split_k))
        if _GEMM_EVENTS:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.load().uniir_gemm_epi(C.byref(d), _stream()), "uniir_gemm_epi")
            e1.record()
            _GEMM_PROF[-1]["events"] = (e0, e1)
            return out
    L.check(L.load().uniir_gemm_epi(C.byref(d), _stream()), "uniir_gemm_epi")
    return out


_SMS = None


def wgrad(dy, x, out):
    """out[N_w, K_w] += dy^T . x   (dy [T, N_w], x [T, K_w] bf16; out fp32).  Contraction over
    tokens with both operands MN-major; splits the contraction (atomic fp32 accumulate) when
    the output has too few tiles to fill the SMs."""
    global _SMS
    if _SMS is None:
        _SMS = torch.cuda.get_device_properties(dy.device).multi_processor_count
    Nw, Kw, T = dy.shape[1], x.shape[1], dy.shape[0]
    tiles = ((Nw + 127) // 128) * ((Kw + 255) // 256)
    split = 1
    if tiles * 2 <= _SMS:
        split = max(1, min(_SMS // tiles, (T + 63) // 64))
    return gemm(dy, x, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=split)


def row_inv_norm(x):
    _require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    inv = torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
    L.check(L.load().uniir_row_inv_norm(x.data_ptr(), _DT[x.dtype], x.shape[0], x.shape[1],
                                        x.stride(0), inv.data_ptr(), _stream()), "uniir_row_inv_norm")
    return inv


def ip_topk_plan(Q, N, k):
    s = C.c_int32()
    ws = C.c_size_t()
    L.check(L.load().uniir_ip_topk_plan(Q, N, k, C.byref(s), C.byref(ws)), "uniir_ip_topk_plan")
    return s.value, ws.value


def ip_topk_partial(q, q_inv, c, c_inv, k, workspace=None, out=None):
    """Fused cosine GEMM + top-k over one candidate shard.  Returns (scores [Q,S,k] fp32,
    idx [Q,S,k] int32) partial lists, one per internal candidate split S."""
    _require_cuda(q, c)
    assert q.dtype == torch.float16 and c.dtype == torch.float16 and q.is_contiguous() and c.is_contiguous()
    Q, d = q.shape
    N = c.shape[0]
    splits, ws_bytes = ip_topk_plan(Q, N, k)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    if out is None:
        ps = torch.empty(Q, splits, k, device=q.device, dtype=torch.float32)
        pi = torch.empty(Q, splits, k, device=q.device, dtype=torch.int32)
    else:
        ps, pi = out
    L.check(L.load().uniir_ip_topk(q.data_ptr(), q_inv.data_ptr(), c.data_ptr(), c_inv.data_ptr(),
                                   Q, N, d, k, splits, ps.data_ptr(), pi.data_ptr(),
                                   workspace.data_ptr(), workspace.numel(), _stream()), "uniir_ip_topk")
    return ps, pi


def topk_merge(scores, idx, k, list_offset=None, id_table=None):
    """scores/idx: [Q, L, k] partial lists -> ([Q,k] fp32 desc, [Q,k] int64)."""
    _require_cuda(scores, idx)
    assert scores.dim() == 3 and scores.shape == idx.shape and scores.is_contiguous() and idx.is_contiguous()
    assert idx.dtype in (torch.int32, torch.int64)
    Q, Ls, kk = scores.shape
    assert kk == k
    os_ = torch.empty(Q, k, device=scores.device, dtype=torch.float32)
    oi = torch.empty(Q, k, device=scores.device, dtype=torch.int64)
    L.check(L.load().uniir_topk_merge(scores.data_ptr(), idx.data_ptr(), int(idx.dtype == torch.int64),
                                      Q, Ls, k, Ls * k, k, _ptr(list_offset), _ptr(id_table),
                                      os_.data_ptr(), oi.data_ptr(), _stream()), "uniir_topk_merge")
    return os_, oi


# ----------------------------------------------------------------------------------------
# row-wise kernels
# ----------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, row_index=None, out=None,
                  out_dtype=torch.bfloat16, stats=None, save_stats=False):
    """x fp32 viewed as [rows, D] with row stride ldx (or gathered through row_index).
    stats=(mean, rstd) preallocated fp32 [rows] to save the statistics for backward."""
    _require_cuda(x)
    D = gamma.numel()
    if rows is None:
        rows = x.numel() // D if row_index is None else row_index.numel()
    if ldx is None:
        ldx = D
    if out is None:
        out = torch.empty(rows, D, device=x.device, dtype=out_dtype)
    mean = rstd = None
    if stats is not None:
        mean, rstd = stats
    elif save_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    L.check(L.load().uniir_layernorm_fwd(x.data_ptr(), ldx, _ptr(row_index), gamma.data_ptr(),
                                         beta.data_ptr(), eps, rows, D, out.data_ptr(), D,
                                         _DT[out.dtype], _ptr(mean), _ptr(rstd), _stream()),
            "uniir_layernorm_fwd")
    return (out, mean, rstd) if save_stats else out


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, *, rows=None, ldx=None, lddx=None,
                  row_index=None, accumulate=True, dx_bf16=None, dcolsum=None):
    """dcolsum fp32 [D] (optional) += column sums of the updated dx: the bias gradient of the linear layer
    whose output gradient this dx is (saves the separate colsum_accum pass over dx_bf16)."""
    D = gamma.numel()
    if rows is None:
        rows = mean.numel()
    L.check(L.load().uniir_layernorm_bwd(dy.data_ptr(), D, _DT[dy.dtype], x.data_ptr(), ldx or D,
                                         _ptr(row_index), mean.data_ptr(), rstd.data_ptr(),
                                         gamma.data_ptr(), rows, D, dx.data_ptr(), lddx or (ldx or D),
                                         int(accumulate), dgamma.data_ptr(), dbeta.data_ptr(), _ptr(dx_bf16),
                                         _ptr(dcolsum), _stream()),
            "uniir_layernorm_bwd")


def rmsnorm_fwd(x, gamma, eps, out):
    """T5LayerNorm: out = x * rsqrt(mean(x^2) + eps) * gamma; x fp32 [rows, D] contiguous, out bf16 or fp32."""
    _require_cuda(x)
    D = gamma.numel()
    rows = x.numel() // D
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and out.numel() == x.numel()
    L.check(L.load().uniir_rmsnorm_fwd(x.data_ptr(), D, gamma.data_ptr(), eps, rows, D, out.data_ptr(), D, _DT[out.dtype],
                                       _stream()), "uniir_rmsnorm_fwd")
    return out


def token_mean(x, out=None):
    """x fp32 [B, S, D] contiguous -> mean over the S tokens, fp32 [B, D]."""
    B, S, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty(B, D, device=x.device, dtype=torch.float32)
    L.check(L.load().uniir_token_mean(x.data_ptr(), B, S, D, out.data_ptr(), _stream()), "uniir_token_mean")
    return out


def rmsnorm_bwd(dy, x, gamma, eps, dx, dgamma, accumulate=True):
    """Backward of rmsnorm_fwd: dx (+)= d x, dgamma += d gamma; dy / x / dx fp32 [rows, D] contiguous."""
    D = gamma.numel()
    rows = x.numel() // D
    assert dy.dtype == x.dtype == dx.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous()
    L.check(L.load().uniir_rmsnorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr(), eps, rows, D, dx.data_ptr(),
                                       int(accumulate), dgamma.data_ptr(), _stream()), "uniir_rmsnorm_bwd")


def token_mean_bwd(d_out, S, out):
    """out fp32 [B, S, D] <- d_out[B, D] / S broadcast over the tokens."""
    B, D = d_out.shape
    assert d_out.dtype == torch.float32 and d_out.is_contiguous() and out.is_contiguous() and out.numel() == B * S * D
    L.check(L.load().uniir_token_mean_bwd(d_out.data_ptr(), B, S, D, out.data_ptr(), _stream()), "uniir_token_mean_bwd")
    return out


def patch_im2col(img, P, Kpad, out=None):
    B, C3, H, W = img.shape
    assert C3 == 3 and img.dtype == torch.float32 and img.is_contiguous()
    g = H // P
    if out is None:
        out = torch.empty(B * g * g, Kpad, device=img.device, dtype=torch.bfloat16)
    L.check(L.load().uniir_patch_im2col(img.data_ptr(), B, H, W, P, Kpad, out.data_ptr(), _stream()),
            "uniir_patch_im2col")
    return out


def vision_embed_fwd(patch, cls, pos, B, S, D, out=None):
    if out is None:
        out = torch.empty(B, S, D, device=patch.device, dtype=torch.float32)
    L.check(L.load().uniir_vision_embed_fwd(patch.data_ptr(), cls.data_ptr(), pos.data_ptr(), B, S, D,
                                            out.data_ptr(), _stream()), "uniir_vision_embed_fwd")
    return out


def vision_embed_bwd(dx0, B, S, D, d_patch, d_cls, d_pos):
    L.check(L.load().uniir_vision_embed_bwd(dx0.data_ptr(), B, S, D, d_patch.data_ptr(), d_cls.data_ptr(),
                                            d_pos.data_ptr(), _stream()), "uniir_vision_embed_bwd")


def text_embed_fwd(ids, tok, pos, out=None, eot=None):
    B, S = ids.shape
    D = tok.shape[1]
    assert ids.dtype == torch.int32 and ids.is_contiguous()
    if out is None:
        out = torch.empty(B, S, D, device=ids.device, dtype=torch.float32)
    if eot is None:
        eot = torch.empty(B, device=ids.device, dtype=torch.int32)
    L.check(L.load().uniir_text_embed_fwd(ids.data_ptr(), tok.data_ptr(), pos.data_ptr(), B, S, D,
                                          out.data_ptr(), eot.data_ptr(), _stream()), "uniir_text_embed_fwd")
    return out, eot


def text_embed_bwd(dx, ids, d_tok, d_pos):
    B, S = ids.shape
    D = d_tok.shape[1]
    L.check(L.load().uniir_text_embed_bwd(dx.data_ptr(), ids.data_ptr(), B, S, D, d_tok.data_ptr(),
                                          d_pos.data_ptr(), _stream()), "uniir_text_embed_bwd")


def score_fuse_fwd(t, v, mt, mv, out=None):
    B, E = t.shape
    if out is None:
        out = torch.empty(B, E, device=t.device, dtype=torch.float32)
    L.check(L.load().uniir_score_fuse_fwd(t.data_ptr(), v.data_ptr(), mt.data_ptr(), mv.data_ptr(), B, E,
                                          out.data_ptr(), _stream()), "uniir_score_fuse_fwd")
    return out


def score_fuse_bwd(d, mt, mv, out=None):
    B, E = d.shape
    if out is None:
        dt = torch.empty(B, E, device=d.device, dtype=torch.bfloat16)
        dv = torch.empty(B, E, device=d.device, dtype=torch.bfloat16)
    else:
        dt, dv = out
    L.check(L.load().uniir_score_fuse_bwd(d.data_ptr(), mt.data_ptr(), mv.data_ptr(), B, E, dt.data_ptr(),
                                          dv.data_ptr(), _stream()), "uniir_score_fuse_bwd")
    return dt, dv


def score_fuse_scatter(emb, rows, mask, out):
    """out[rows[r]] += mask[rows[r]] * emb[r] for the first rows.numel() rows of emb (compacted tower output)."""
    n, E = rows.numel(), out.shape[1]
    assert emb.dtype == torch.float32 and out.dtype == torch.float32 and rows.dtype == torch.int32 and mask.dtype == torch.int64
    L.check(L.load().uniir_score_fuse_scatter(emb.data_ptr(), rows.data_ptr(), mask.data_ptr(), n, E, out.data_ptr(),
                                              _stream()), "uniir_score_fuse_scatter")


def score_fuse_gather_bwd(d_out, rows, mask, out):
    """out[r] = bf16(mask[rows[r]] * d_out[rows[r]]) for r < rows.numel(), zero for the padding rows of out."""
    n, (n_pad, E) = rows.numel(), out.shape
    assert d_out.dtype == torch.float32 and out.dtype == torch.bfloat16 and rows.dtype == torch.int32
    L.check(L.load().uniir_score_fuse_gather_bwd(d_out.data_ptr(), rows.data_ptr(), mask.data_ptr(), n, n_pad, E,
                                                 out.data_ptr(), _stream()), "uniir_score_fuse_gather_bwd")


def colsum_accum(dy, out):
    M, N = dy.shape
    assert dy.dtype == torch.bfloat16 and dy.stride(1) == 1 and out.dtype == torch.float32
    L.check(L.load().uniir_colsum_accum(dy.data_ptr(), M, N, dy.stride(0), out.data_ptr(), _stream()),
            "uniir_colsum_accum")


def pad_cast_bf16(src, Kpad, out=None):
    rows, K = src.shape
    if out is None:
        out = torch.empty(rows, Kpad, device=src.device, dtype=torch.bfloat16)
    L.check(L.load().uniir_pad_cast_bf16(src.data_ptr(), rows, K, Kpad, out.data_ptr(), _stream()),
            "uniir_pad_cast_bf16")
    return out


def unpad_accum(src, dst):
    rows, K = dst.shape
    L.check(L.load().uniir_unpad_accum(src.data_ptr(), rows, K, src.shape[1], dst.data_ptr(), _stream()),
            "uniir_unpad_accum")


def cast_bf16(src, out=None):
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    L.check(L.load().uniir_cast_bf16(src.data_ptr(), src.numel(), out.data_ptr(), _stream()), "uniir_cast_bf16")
    return out


def adamw_step(p, g, m, v, p16, lr, betas, eps, wd, step, grad_scale=1.0):
    L.check(L.load().uniir_adamw_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), _ptr(p16),
                                      p.numel(), lr, betas[0], betas[1], eps, wd, step, grad_scale,
                                      _stream()), "uniir_adamw_step")


# ----------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------
def attn_set_mode(mode):
    """Bit mask: 1 = tcgen05 forward where Skv <= 272 (non-causal), 2 = tcgen05 backward where Sq <= 260, 4 = also
    the shapes where tcgen05 measured slower (causal forward); 0 = mma.sync kernels.  Default 3.  Returns the
    previous mode."""
    return L.load().uniir_attn_set_mode(int(mode))


def attn_bwd_ws_floats(B, Sq, Skv, H):
    """fp32 elements of the backward workspace (delta, base-2 LSE, tail-column dS)."""
    return int(L.load().uniir_attn_bwd_workspace_floats(B, Sq, Skv, H))


def attn_fwd(qkv, B, S, H, causal, out=None, lse=None, need_lse=True):
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * S, 3 * H * 64)
    if out is None:
        out = torch.empty(B * S, H * 64, device=qkv.device, dtype=torch.bfloat16)
    if lse is None and need_lse:
        lse = torch.empty(B, H, S, device=qkv.device, dtype=torch.float32)
    L.check(L.load().uniir_attn_fwd(qkv.data_ptr(), B, S, H, int(causal), out.data_ptr(), _ptr(lse),
                                    _stream()), "uniir_attn_fwd")
    return out, lse


def attn_fwd_ex(q, k, v, B, Sq, Skv, H, *, causal=False, kv_len=None, out=None, lse=None, drop=None):
    """q [B*Sq, >=H*64] / k, v [B*Skv, >=H*64] bf16 row-strided views (column slices of packed projections are
    fine), optional int32 kv_len[B] key-padding mask.  drop = (ctx, site, p): dropout on the probabilities."""
    assert q.dtype == torch.bfloat16 and q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1
    assert k.stride(0) == v.stride(0)
    if out is None:
        out = torch.empty(B * Sq, H * 64, device=q.device, dtype=torch.bfloat16)
    if drop is not None:
        ctx, site, p = drop
        L.check(L.load().uniir_attn_fwd_drop(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), B, Sq, Skv,
                                             H, int(causal), _ptr(kv_len), 0.125, None, out.data_ptr(), out.stride(0),
                                             _ptr(lse), ctx.data_ptr(), int(site), float(p), _stream()),
                "uniir_attn_fwd_drop")
        return out
    L.check(L.load().uniir_attn_fwd_ex(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), B, Sq, Skv, H,
                                       int(causal), _ptr(kv_len), out.data_ptr(), out.stride(0), _ptr(lse), _stream()),
            "uniir_attn_fwd_ex")
    return out


def attn_fwd_bias(q, k, v, B, Sq, Skv, H, bias, scale=1.0, out=None, lse=None, drop=None):
    """softmax(scale * q k^T + bias) v with bias fp32 [H, Sq, Skv] shared by all batch rows (T5 relative-position
    bias; T5 does not scale the scores).  q / k / v: bf16 row-strided views as in attn_fwd_ex."""
    assert q.dtype == torch.bfloat16 and q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1
    assert k.stride(0) == v.stride(0)
    assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.shape == (H, Sq, Skv)
    if out is None:
        out = torch.empty(B * Sq, H * 64, device=q.device, dtype=torch.bfloat16)
    if drop is not None:
        ctx, site, p = drop
        L.check(L.load().uniir_attn_fwd_drop(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), B, Sq, Skv,
                                             H, 0, None, float(scale), bias.data_ptr(), out.data_ptr(), out.stride(0),
                                             _ptr(lse), ctx.data_ptr(), int(site), float(p), _stream()),
                "uniir_attn_fwd_drop")
        return out
    L.check(L.load().uniir_attn_fwd_bias(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), B, Sq, Skv, H,
                                         float(scale), bias.data_ptr(), out.data_ptr(), out.stride(0), _ptr(lse), _stream()),
            "uniir_attn_fwd_bias")
    return out


def attn_bwd_bias(q, k, v, out, dout, lse, B, Sq, Skv, H, bias, dbias, dq, dk, dv, scale=1.0, delta=None, drop=None):
    """Backward of attn_fwd_bias; dbias fp32 [H, Sq, Skv] += gradient of the bias (zero it first)."""
    if delta is None:
        delta = torch.empty(B * H * Sq, device=q.device, dtype=torch.float32)
    assert dk.stride(0) == dv.stride(0) and k.stride(0) == v.stride(0)
    assert bias.shape == dbias.shape == (H, Sq, Skv) and dbias.dtype == torch.float32 and dbias.is_contiguous()
    if drop is not None:
        ctx, site, p = drop
        L.check(L.load().uniir_attn_bwd_drop(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                             out.data_ptr(), out.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(),
                                             B, Sq, Skv, H, 0, None, float(scale), bias.data_ptr(), dbias.data_ptr(),
                                             delta.data_ptr(), dq.data_ptr(), dq.stride(0), dk.data_ptr(), dv.data_ptr(),
                                             dk.stride(0), ctx.data_ptr(), int(site), float(p), _stream()),
                "uniir_attn_bwd_drop")
        return
    L.check(L.load().uniir_attn_bwd_bias(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
                                         out.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(), B, Sq, Skv, H,
                                         float(scale), bias.data_ptr(), dbias.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                         dq.stride(0), dk.data_ptr(), dv.data_ptr(), dk.stride(0), _stream()),
            "uniir_attn_bwd_bias")


def attn_bwd(qkv, out, dout, lse, B, S, H, causal, dqkv=None, delta=None):
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    if delta is None:
        delta = torch.empty(attn_bwd_ws_floats(B, S, S, H), device=qkv.device, dtype=torch.float32)
    L.check(L.load().uniir_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, S, H,
                                    int(causal), delta.data_ptr(), dqkv.data_ptr(), _stream()),
            "uniir_attn_bwd")
    return dqkv


# ----------------------------------------------------------------------------------------
# InfoNCE
# ----------------------------------------------------------------------------------------
class InfoNCE:
    """Stateful helper holding the scratch of the fused loss for fixed (bs, n_cols, E)."""

    def __init__(self, bs, n_cols, E, device):
        self.bs, self.n, self.E = bs, n_cols, E
        f = dict(device=device, dtype=torch.float32)
        self.ws = torch.zeros(L.load().uniir_infonce_workspace_bytes(bs, n_cols), device=device, dtype=torch.uint8)
        self.qn = torch.empty(bs, E, **f)
        self.all_p = torch.empty(n_cols, E, **f)
        self.q_norm = torch.empty(bs, **f)
        self.p_norm = torch.empty(bs, **f)
        self.lse = torch.empty(bs, **f)
        self.loss = torch.zeros((), **f)
        self.acc = torch.zeros((), **f)
        self.G = torch.empty(bs, n_cols, **f)
        self.d_qn = torch.empty(bs, E, **f)
        self.d_all_p = torch.empty(n_cols, E, **f)

    def normalize(self, emb, q_idx, p_idx, target_offset):
        """qn <- normalize(emb[q_idx]); all_p[target_offset : +bs] <- normalize(emb[p_idx])."""
        lib = L.load()
        pn = self.all_p[target_offset:target_offset + self.bs]
        L.check(lib.uniir_infonce_normalize(emb.data_ptr(), q_idx.data_ptr(), p_idx.data_ptr(), self.bs,
                                            self.E, self.qn.data_ptr(), pn.data_ptr(),
                                            self.q_norm.data_ptr(), self.p_norm.data_ptr(), _stream()),
                "uniir_infonce_normalize")
        return pn

    def forward(self, logit_scale, target_offset):
        L.check(L.load().uniir_infonce_fwd(self.qn.data_ptr(), self.all_p.data_ptr(), self.bs, self.n, self.E,
                                           logit_scale.data_ptr(), target_offset, self.loss.data_ptr(),
                                           self.acc.data_ptr(), self.lse.data_ptr(), self.ws.data_ptr(),
                                           self.ws.numel(), _stream()), "uniir_infonce_fwd")
        return self.loss, self.acc

    def backward(self, logit_scale, target_offset, d_logit_scale, grad_out=None):
        L.check(L.load().uniir_infonce_bwd(self.qn.data_ptr(), self.all_p.data_ptr(), self.bs, self.n, self.E,
                                           logit_scale.data_ptr(), target_offset, self.lse.data_ptr(),
                                           _ptr(grad_out), self.G.data_ptr(), self.d_qn.data_ptr(),
                                           self.d_all_p.data_ptr(), d_logit_scale.data_ptr(), _stream()),
                "uniir_infonce_bwd")
        return self.d_qn, self.d_all_p

    def normalize_bwd(self, q_idx, p_idx, d_pn, d_emb, target_offset):
        pn = self.all_p[target_offset:target_offset + self.bs]
        L.check(L.load().uniir_infonce_normalize_bwd(q_idx.data_ptr(), p_idx.data_ptr(), self.qn.data_ptr(),
                                                     pn.data_ptr(), self.q_norm.data_ptr(),
                                                     self.p_norm.data_ptr(), self.d_qn.data_ptr(),
                                                     d_pn.data_ptr(), self.bs, self.E, d_emb.data_ptr(),
                                                     _stream()), "uniir_infonce_normalize_bwd")


def time_gemm_config(cfg, iters=10, device="cuda"):
    """Average duration (ms) of one launch of the GEMM described by a trace entry, timed with CUDA
    events over `iters` back-to-back launches on fresh (L2-cold on first touch) operands."""
    M, N, K = cfg["M"], cfg["N"], cfg["K"]
    dt = cfg["dtype"]
    a = torch.randn((K, M) if cfg["a_mn"] else (M, K), device=device).to(dt)
    b = torch.randn((K, N) if cfg["b_mn"] else (N, K), device=device).to(dt)
    out = torch.zeros(M, N, device=device, dtype=cfg["out_dtype"])
    kw = dict(a_mn=cfg["a_mn"], b_mn=cfg["b_mn"], out=out, act=cfg["act"], accumulate=cfg["accumulate"],
              split_k=cfg["split_k"])
    if cfg["bias"]:
        kw["bias"] = torch.randn(N, device=device)
    if cfg["residual"] is not None:
        kw["residual"] = out if cfg["residual"] == out.dtype else torch.randn(M, N, device=device).to(cfg["residual"])
    if cfg["aux"]:
        kw["aux_out"] = torch.empty(M, N, device=device, dtype=dt)
    if cfg["ags"]:
        kw["act_grad_src"] = torch.randn(M, N, device=device).to(dt)
    for _ in range(2):
        gemm(a, b, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        gemm(a, b, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def hardneg_loss(emb, q_idx, p_idx, n_idx, in_batch_neg_num, logit_scale, d_logit_scale):
    """Fused forward+backward of the hard-negative loss branch.  Returns (loss, acc, d_emb)."""
    bs, neg = n_idx.shape
    E = emb.shape[1]
    sums = torch.zeros(2, device=emb.device, dtype=torch.float32)
    d_emb = torch.zeros_like(emb)
    L.check(L.load().uniir_hardneg_loss(emb.data_ptr(), q_idx.data_ptr(), p_idx.data_ptr(), n_idx.data_ptr(), bs, neg,
                                        in_batch_neg_num, E, logit_scale.data_ptr(), sums[0:1].data_ptr(),
                                        sums[1:2].data_ptr(), d_emb.data_ptr(), d_logit_scale.data_ptr(), _stream()),
            "uniir_hardneg_loss")
    return sums[0] / bs, sums[1] / bs, d_emb


def attn_bwd_ex(q, k, v, out, dout, lse, B, Sq, Skv, H, dq, dk, dv, *, causal=False, kv_len=None, delta=None, drop=None):
    """Backward of attn_fwd_ex; dq/dk/dv are preallocated bf16 row-strided views (dk, dv share a stride)."""
    if delta is None:
        delta = torch.empty(attn_bwd_ws_floats(B, Sq, Skv, H), device=q.device, dtype=torch.float32)
    assert dk.stride(0) == dv.stride(0) and k.stride(0) == v.stride(0)
    if drop is not None:
        ctx, site, p = drop
        L.check(L.load().uniir_attn_bwd_drop(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                             out.data_ptr(), out.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(),
                                             B, Sq, Skv, H, int(causal), _ptr(kv_len), 0.125, None, None, delta.data_ptr(),
                                             dq.data_ptr(), dq.stride(0), dk.data_ptr(), dv.data_ptr(), dk.stride(0),
                                             ctx.data_ptr(), int(site), float(p), _stream()), "uniir_attn_bwd_drop")
        return
    L.check(L.load().uniir_attn_bwd_ex(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
                                       out.stride(0), dout.data_ptr(), dout.stride(0), lse.data_ptr(), B, Sq, Skv, H,
                                       int(causal), _ptr(kv_len), delta.data_ptr(), dq.data_ptr(), dq.stride(0),
                                       dk.data_ptr(), dv.data_ptr(), dk.stride(0), _stream()), "uniir_attn_bwd_ex")


def ema_update(p_m, p, momentum, p_m16=None):
    L.check(L.load().uniir_ema_update(p_m.data_ptr(), p.data_ptr(), p.numel(), momentum, _ptr(p_m16), _stream()),
            "uniir_ema_update")


def tanh_bwd(dy, y, out):
    L.check(L.load().uniir_tanh_bwd(dy.data_ptr(), y.data_ptr(), y.numel(), out.data_ptr(), _stream()), "uniir_tanh_bwd")


def sgemm_temp(a, b, b_kn, temp, out=None):
    """fp32 (a @ b or a @ b.T) / temp with b stored [K,N] (b_kn) or [N,K]."""
    M, K = a.shape
    N = b.shape[1] if b_kn else b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    L.check(L.load().uniir_sgemm_temp(a.data_ptr(), b.data_ptr(), int(b_kn), M, N, K, temp.data_ptr(), out.data_ptr(),
                                      _stream()), "uniir_sgemm_temp")
    return out


def blip_softce(S, Sm, pc_id, id_all, alpha, gscale, temp, loss_sum, acc_sum, d_temp):
    bs, n = S.shape
    L.check(L.load().uniir_blip_softce(S.data_ptr(), Sm.data_ptr(), pc_id.data_ptr(), id_all.data_ptr(), bs, n, alpha,
                                       gscale, temp.data_ptr(), loss_sum.data_ptr(), _ptr(acc_sum), d_temp.data_ptr(),
                                       _stream()), "uniir_blip_softce")


# ---- dropout / stochastic depth (include/uniir_b200.h; counter-based masks, nothing saved) -------------------------
def drop_ctx(device):
    """uint32[2] device buffer {seed of the step, first item of the micro-batch}; int32 storage (torch has no uint32
    arithmetic), read as uint32 by the kernels."""
    return torch.zeros(2, device=device, dtype=torch.int32)


def dropout(x, ctx, site, per_item, p, *, residual=None, out=None):
    """out = dropout(x) (+ residual): keep with probability 1 - p, scale 1 / (1 - p); x / out fp32 or bf16 contiguous,
    residual fp32; out defaults to x (in place)."""
    out = x if out is None else out
    assert x.is_contiguous() and out.is_contiguous() and x.numel() == out.numel()
    assert residual is None or (residual.dtype == torch.float32 and residual.is_contiguous() and residual.numel() == x.numel())
    L.check(L.load().uniir_dropout(x.data_ptr(), _DT[x.dtype], _ptr(residual), out.data_ptr(), _DT[out.dtype], x.numel(),
                                   ctx.data_ptr(), int(site), int(per_item), float(p), _stream()), "uniir_dropout")
    return out


def dropout_mask(ctx, site, per_item, n, p):
    keep = torch.empty(n, device=ctx.device, dtype=torch.uint8)
    L.check(L.load().uniir_dropout_mask(ctx.data_ptr(), int(site), int(per_item), n, float(p), keep.data_ptr(), _stream()),
            "uniir_dropout_mask")
    return keep


def droppath_scales(ctx, site_base, p_sites, mb, out=None):
    """fp32 [n_sites, mb]: 1 / (1 - p_sites[s]) where item b keeps residual branch s, else 0."""
    n = p_sites.numel()
    if out is None:
        out = torch.empty(n, mb, device=ctx.device, dtype=torch.float32)
    assert p_sites.dtype == torch.float32 and p_sites.is_contiguous() and out.is_contiguous()
    L.check(L.load().uniir_droppath_scales(ctx.data_ptr(), int(site_base), n, mb, p_sites.data_ptr(), out.data_ptr(),
                                           _stream()), "uniir_droppath_scales")
    return out


def row_scale(x, scales, rows_per_scale):
    """x[r, :] *= scales[r // rows_per_scale]; x bf16 [rows, cols] row-strided, in place."""
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and scales.dtype == torch.float32 and scales.is_contiguous()
    L.check(L.load().uniir_row_scale(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), scales.data_ptr(),
                                     int(rows_per_scale), _stream()), "uniir_row_scale")
    return x
