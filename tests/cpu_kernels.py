"""Plain-torch statements of the C-ABI kernel contracts (include/uniir_b200.h), one per wrapper in uniir_b200.ops.

Test infrastructure: `install(monkeypatch)` swaps them in for the ctypes wrappers so that the product's HOST logic
(kernel sequencing, buffer wiring, weight views, gradient routing) runs on a machine without a GPU and can be
compared against the oracle's autograd.  The kernels themselves are covered by the `-m gpu` tests; nothing here is
reachable from the product package."""
import torch
import torch.nn.functional as F


def _act(y, act, L):
    if act == L.ACT_QUICKGELU:
        return y * torch.sigmoid(1.702 * y)
    if act == L.ACT_GELU:
        return F.gelu(y)
    if act == L.ACT_TANH:
        return torch.tanh(y)
    if act == L.ACT_RELU:
        return torch.relu(y)
    return y


def _act_grad(x, act, L):
    if act == L.ACT_QUICKGELU:
        s = torch.sigmoid(1.702 * x)
        return s * (1 + 1.702 * x * (1 - s))
    if act == L.ACT_GELU:
        return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * 0.3989422804014327 * torch.exp(-0.5 * x * x)
    if act == L.ACT_RELU:
        return (x > 0).float()
    return torch.ones_like(x)


# ---- counter-based dropout masks (include/uniir_b200.h, csrc/dropout.cuh), restated with int64 arithmetic ------------
_M32 = 0xFFFFFFFF


def mix32(x):
    x = x & _M32
    x = x ^ (x >> 17); x = (x * 0xed5ad4bb) & _M32
    x = x ^ (x >> 11); x = (x * 0xac4c1b51) & _M32
    x = x ^ (x >> 15); x = (x * 0x31848bab) & _M32
    return x ^ (x >> 14)


def drop_thresh(p):
    import numpy as np
    t = float(np.float32(p)) * 4294967296.0
    return 0xFFFFFFFF if t >= 4294967295.0 else max(0, int(t))


def drop_scale(p):
    import numpy as np
    return float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))


def keep_mask(seed, site, idx, p):
    """idx: int64 tensor of element indices within the whole batch (taken mod 2^32) -> bool keep."""
    key = mix32(torch.tensor((int(seed) + int(site) * 0x9E3779B1) & _M32, dtype=torch.int64))
    return mix32((idx & _M32) ^ key) >= drop_thresh(p)


def _ctx(ctx):
    return int(ctx[0]) & _M32, int(ctx[1]) & _M32


def install(monkeypatch):
    from uniir_b200 import lib as L
    from uniir_b200 import ops

    def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, act=L.ACT_NONE, residual=None,
             aux_out=None, act_grad_src=None, alpha=1.0, accumulate=False, split_k=1, row_scale=None, row_scale_rows=1):
        # epilogue order of gemm_kernel.cuh: scale / bias / (aux copy) / act | act-grad / row scale / residual / accumulate / store
        A = a.float().t() if a_mn else a.float()
        Bm = b.float() if b_mn else b.float().t()
        y = alpha * (A @ Bm)
        if bias is not None:
            y = y + bias
        if aux_out is not None:
            aux_out.copy_(y)
        if act_grad_src is not None:
            y = y * _act_grad(act_grad_src.float(), act, L)
        else:
            y = _act(y, act, L)
        if row_scale is not None:
            y = y * row_scale[torch.arange(y.shape[0]) // row_scale_rows][:, None]
        if residual is not None:
            y = y + residual.float()
        if out is None:
            out = torch.zeros(y.shape, dtype=out_dtype or a.dtype)
        out.copy_(out.float() + y if accumulate else y)
        return out

    def wgrad(dy, x, out):
        out.add_(dy.float().t() @ x.float())
        return out

    def colsum_accum(dy, out):
        out.add_(dy.float().sum(0))

    def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, row_index=None, out=None, out_dtype=torch.bfloat16,
                      stats=None, save_stats=False):
        D = gamma.numel()
        xr = x.reshape(-1, D).float()
        if row_index is not None:
            xr = xr[row_index.long()]
        mean = xr.mean(-1)
        rstd = torch.rsqrt(xr.var(-1, unbiased=False) + eps)
        y = (xr - mean[:, None]) * rstd[:, None] * gamma + beta
        if out is None:
            out = torch.empty(y.shape, dtype=out_dtype)
        out.view(-1, D).copy_(y)
        if stats is not None:
            stats[0].copy_(mean); stats[1].copy_(rstd)
        return out

    def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, *, rows=None, ldx=None, lddx=None, row_index=None,
                      accumulate=True, dx_bf16=None, dcolsum=None):
        D = gamma.numel()
        xr = x.reshape(-1, D).float()
        idx = row_index.long() if row_index is not None else torch.arange(mean.numel())
        xr = xr[idx]
        dyf = dy.reshape(-1, D).float()
        xhat = (xr - mean[:, None]) * rstd[:, None]
        g = dyf * gamma
        d = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
        dxv = dx.view(-1, D)
        new = dxv[idx] + d if accumulate else d
        dxv[idx] = new
        dgamma.add_((dyf * xhat).sum(0)); dbeta.add_(dyf.sum(0))
        if dx_bf16 is not None:
            dx_bf16.view(-1, D)[idx] = new.to(dx_bf16.dtype)
        if dcolsum is not None:
            dcolsum.add_(new.sum(0))

    def _split(t, B, S, H):
        return t.float().reshape(B, S, H, 64).transpose(1, 2)

    def _scores(q, k, B, Sq, Skv, H, scale, causal, kv_len, bias=None):
        s = _split(q, B, Sq, H) @ _split(k, B, Skv, H).transpose(-1, -2) * scale
        if bias is not None:
            s = s + bias
        if causal:
            s = s + torch.full((Sq, Skv), float("-inf")).triu_(1)
        if kv_len is not None:
            s = s.masked_fill(torch.arange(Skv)[None, None, None, :] >= kv_len.long()[:, None, None, None], float("-inf"))
        return s

    def _prob_mask(drop, B, Sq, Skv, H):
        """mask / (1 - p) over [B, H, Sq, Skv]: element ((item0 + b) * H + h) * Sq * Skv + q * Skv + key of `site`."""
        if drop is None:
            return None
        ctx, site, p = drop
        seed, item0 = _ctx(ctx)
        idx = torch.arange(B * H * Sq * Skv, dtype=torch.int64) + item0 * H * Sq * Skv
        return keep_mask(seed, site, idx, p).view(B, H, Sq, Skv).float() * drop_scale(p)

    def _attn_fwd(q, k, v, B, Sq, Skv, H, scale, causal, kv_len, out, lse, bias=None, drop=None):
        s = _scores(q, k, B, Sq, Skv, H, scale, causal, kv_len, bias)
        pr = s.softmax(-1)
        if drop is not None:
            pr = pr * _prob_mask(drop, B, Sq, Skv, H)
        y = (pr @ _split(v, B, Skv, H)).transpose(1, 2).reshape(B * Sq, H * 64)
        if out is None:
            out = torch.empty(B * Sq, H * 64, dtype=torch.bfloat16)
        out.copy_(y)
        if lse is not None:
            lse.copy_(torch.logsumexp(s, -1))
        return out

    def _attn_bwd(q, k, v, dout, B, Sq, Skv, H, scale, causal, kv_len, drop=None):
        qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
        s = _scores(qf, kf, B, Sq, Skv, H, scale, causal, kv_len)
        pr = s.softmax(-1)
        if drop is not None:
            pr = pr * _prob_mask(drop, B, Sq, Skv, H)
        y = (pr @ _split(vf, B, Skv, H)).transpose(1, 2).reshape(B * Sq, H * 64)
        y.backward(dout.float())
        return qf.grad, kf.grad, vf.grad

    def attn_fwd(qkv, B, S, H, causal, out=None, lse=None, need_lse=True):
        D = H * 64
        if lse is None and need_lse:
            lse = torch.empty(B, H, S)
        return _attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, S, S, H, 0.125, causal, None, out, lse), lse

    def attn_fwd_ex(q, k, v, B, Sq, Skv, H, *, causal=False, kv_len=None, out=None, lse=None, drop=None):
        return _attn_fwd(q, k, v, B, Sq, Skv, H, 0.125, causal, kv_len, out, lse, drop=drop)

    def attn_fwd_bias(q, k, v, B, Sq, Skv, H, bias, scale=1.0, out=None, lse=None, drop=None):
        return _attn_fwd(q, k, v, B, Sq, Skv, H, scale, False, None, out, lse, bias, drop=drop)

    def attn_bwd(qkv, out, dout, lse, B, S, H, causal, dqkv=None, delta=None):
        D = H * 64
        dq, dk, dv = _attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], dout, B, S, S, H, 0.125, causal, None)
        if dqkv is None:
            dqkv = torch.empty_like(qkv)
        dqkv[:, :D] = dq.to(dqkv.dtype); dqkv[:, D:2 * D] = dk.to(dqkv.dtype); dqkv[:, 2 * D:] = dv.to(dqkv.dtype)
        return dqkv

    def attn_bwd_ex(q, k, v, out, dout, lse, B, Sq, Skv, H, dq, dk, dv, *, causal=False, kv_len=None, delta=None, drop=None):
        a, b_, c = _attn_bwd(q, k, v, dout, B, Sq, Skv, H, 0.125, causal, kv_len, drop=drop)
        dq.copy_(a); dk.copy_(b_); dv.copy_(c)

    def attn_bwd_ws_floats(B, Sq, Skv, H):
        return B * H * ((Sq + 15) // 16 * 16) * 3

    def rmsnorm_fwd(x, gamma, eps, out):
        D = gamma.numel()
        xr = x.reshape(-1, D).float()
        out.view(-1, D).copy_(xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + eps) * gamma)
        return out

    def rmsnorm_bwd(dy, x, gamma, eps, dx, dgamma, accumulate=True):
        D = gamma.numel()
        xr, dyr = x.reshape(-1, D).float(), dy.reshape(-1, D).float()
        rstd = torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + eps)
        xh = xr * rstd
        g = dyr * gamma
        d = rstd * (g - xh * (g * xh).mean(-1, keepdim=True))
        dxv = dx.view(-1, D)
        dxv.copy_(dxv + d if accumulate else d)
        dgamma.add_((dyr * xh).sum(0))

    def token_mean_bwd(d_out, S, out):
        B, D = d_out.shape
        out.view(B, S, D).copy_((d_out / S)[:, None, :].expand(B, S, D))
        return out

    def attn_bwd_bias(q, k, v, out, dout, lse, B, Sq, Skv, H, bias, dbias, dq, dk, dv, scale=1.0, delta=None, drop=None):
        qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
        bf = bias.detach().clone().requires_grad_(True)
        s = _scores(qf, kf, B, Sq, Skv, H, scale, False, None, bf)
        pr = s.softmax(-1)
        if drop is not None:
            pr = pr * _prob_mask(drop, B, Sq, Skv, H)
        y = (pr @ _split(vf, B, Skv, H)).transpose(1, 2).reshape(B * Sq, H * 64)
        y.backward(dout.float())
        dq.copy_(qf.grad); dk.copy_(kf.grad); dv.copy_(vf.grad)
        dbias.add_(bf.grad)

    def token_mean(x, out=None):
        y = x.mean(1)
        return y if out is None else out.copy_(y)

    def patch_im2col(img, P, Kpad, out=None):
        col = F.unfold(img, P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
        if out is None:
            out = torch.zeros(col.shape[0], Kpad, dtype=torch.bfloat16)
        out.zero_()
        out[:, :col.shape[1]].copy_(col)
        return out

    def vision_embed_fwd(patch, cls, pos, B, S, D, out=None):
        y = torch.cat([cls.expand(B, 1, D), patch.float().view(B, S - 1, D)], 1) + pos
        return y if out is None else out.copy_(y.view(out.shape))

    def vision_embed_bwd(dx0, B, S, D, d_patch, d_cls, d_pos):
        d = dx0.view(B, S, D)
        d_patch.view(B, S - 1, D).copy_(d[:, 1:])
        d_cls.add_(d[:, 0].sum(0)); d_pos.add_(d.sum(0))

    def text_embed_fwd(ids, tok, pos, out=None, eot=None):
        y = tok[ids.long()] + pos
        if out is not None:
            out.copy_(y.view(out.shape))
        if eot is not None:
            eot.copy_(torch.arange(ids.shape[0]) * ids.shape[1] + ids.argmax(-1))
        return out, eot

    def text_embed_bwd(dx, ids, d_tok, d_pos):
        D = d_tok.shape[1]
        d_tok.index_add_(0, ids.long().flatten(), dx.reshape(-1, D))
        d_pos.add_(dx.reshape(ids.shape[0], ids.shape[1], D).sum(0))

    def score_fuse_fwd(t, v, mt, mv, out=None):
        y = t * mt[:, None].float() + v * mv[:, None].float()
        return y if out is None else out.copy_(y)

    def score_fuse_bwd(d, mt, mv, out=None):
        dt, dv = out if out is not None else (torch.empty(d.shape, dtype=torch.bfloat16), torch.empty(d.shape, dtype=torch.bfloat16))
        dt.copy_(d * mt[:, None].float()); dv.copy_(d * mv[:, None].float())
        return dt, dv

    def pad_cast_bf16(src, Kpad, out=None):
        if out is None:
            out = torch.zeros(src.shape[0], Kpad, dtype=torch.bfloat16)
        out.zero_()
        out[:, :src.shape[1]].copy_(src)
        return out

    def unpad_accum(src, dst):
        dst.add_(src[:, :dst.shape[1]])

    def cast_bf16(src, out=None):
        return src.bfloat16() if out is None else out.copy_(src)

    def adamw_step(p, g, m, v, p16, lr, betas, eps, wd, step, grad_scale=1.0):
        # rowops.cu::adamw_kernel: decoupled decay, bias correction, bf16 shadow refresh, gradient zeroed
        b1, b2 = betas
        gr = g * grad_scale
        p.mul_(1.0 - lr * wd)
        m.mul_(b1).add_(gr, alpha=1 - b1)
        v.mul_(b2).addcmul_(gr, gr, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        p.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
        g.zero_()
        if p16 is not None:
            p16.copy_(p)

    def tanh_bwd(dy, y, out):
        out.copy_(dy * (1 - y * y))

    def drop_ctx(device):
        return torch.zeros(2, dtype=torch.int32)

    def dropout(x, ctx, site, per_item, p, *, residual=None, out=None):
        out = x if out is None else out
        seed, item0 = _ctx(ctx)
        idx = torch.arange(x.numel(), dtype=torch.int64) + item0 * per_item
        y = torch.where(keep_mask(seed, site, idx, p).view(x.shape), x.float() * drop_scale(p), torch.zeros(()))
        if residual is not None:
            y = y + residual.view(x.shape)
        out.copy_(y.view(out.shape))
        return out

    def dropout_mask(ctx, site, per_item, n, p):
        seed, item0 = _ctx(ctx)
        return keep_mask(seed, site, torch.arange(n, dtype=torch.int64) + item0 * per_item, p).to(torch.uint8)

    def droppath_scales(ctx, site_base, p_sites, mb, out=None):
        seed, item0 = _ctx(ctx)
        rows = []
        for s_, p in enumerate(p_sites.tolist()):
            keep = keep_mask(seed, site_base + s_, torch.arange(mb, dtype=torch.int64) + item0, p)
            rows.append(keep.float() * drop_scale(p))
        y = torch.stack(rows)
        return y if out is None else out.copy_(y)

    def row_scale(x, scales, rows_per_scale):
        x.copy_(x.float() * scales[torch.arange(x.shape[0]) // rows_per_scale][:, None])
        return x

    fns = dict(gemm=gemm, wgrad=wgrad, colsum_accum=colsum_accum, layernorm_fwd=layernorm_fwd, layernorm_bwd=layernorm_bwd,
               attn_fwd=attn_fwd, attn_fwd_ex=attn_fwd_ex, attn_fwd_bias=attn_fwd_bias, attn_bwd=attn_bwd,
               attn_bwd_ex=attn_bwd_ex, attn_bwd_ws_floats=attn_bwd_ws_floats, rmsnorm_fwd=rmsnorm_fwd,
               rmsnorm_bwd=rmsnorm_bwd, token_mean_bwd=token_mean_bwd, attn_bwd_bias=attn_bwd_bias,
               token_mean=token_mean, patch_im2col=patch_im2col, vision_embed_fwd=vision_embed_fwd,
               vision_embed_bwd=vision_embed_bwd, text_embed_fwd=text_embed_fwd, text_embed_bwd=text_embed_bwd,
               score_fuse_fwd=score_fuse_fwd, score_fuse_bwd=score_fuse_bwd, pad_cast_bf16=pad_cast_bf16,
               unpad_accum=unpad_accum, cast_bf16=cast_bf16, tanh_bwd=tanh_bwd, adamw_step=adamw_step,
               drop_ctx=drop_ctx, dropout=dropout, dropout_mask=dropout_mask, droppath_scales=droppath_scales,
               row_scale=row_scale)
    for name, fn in fns.items():
        monkeypatch.setattr(ops, name, fn)
