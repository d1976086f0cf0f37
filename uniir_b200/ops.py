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


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, act=L.ACT_NONE,
         residual=None, aux_out=None, act_grad_src=None, alpha=1.0, accumulate=False, split_k=1):
    """out[M,N] (+)= epilogue(alpha * A . B^T); see include/uniir_b200.h.

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
    L.check(L.load().uniir_gemm_epi(C.byref(d), _stream()), "uniir_gemm_epi")
    return out


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
