"""Host logic of the CLIP_FF encode path (uniir_b200/clip_ff.py::CLIPFFEncoder) on CPU: every kernel wrapper in
uniir_b200.ops is replaced by a plain-torch statement of its documented contract (include/uniir_b200.h), so what runs
is exactly the product's sequencing -- buffer wiring, weight views (the fused q|k|v projection, x @ W vs Linear
layouts), the relative-position bias table, the text-then-image token order -- and it must reproduce the oracle.
The kernels themselves are covered by the GPU tests; this isolates the orchestration, which needs no GPU."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle.clip_ff_oracle import CLIPFeatureFusionOracle


def _install_cpu_ops(monkeypatch):
    from uniir_b200 import lib as L
    from uniir_b200 import ops

    def act_fn(y, act):
        if act == L.ACT_QUICKGELU:
            return y * torch.sigmoid(1.702 * y)
        if act == L.ACT_GELU:
            return F.gelu(y)
        if act == L.ACT_TANH:
            return torch.tanh(y)
        if act == L.ACT_RELU:
            return torch.relu(y)
        return y

    def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=None, bias=None, act=L.ACT_NONE, residual=None,
             aux_out=None, act_grad_src=None, alpha=1.0, accumulate=False, split_k=1):
        A = a.float().t() if a_mn else a.float()
        Bm = b.float() if b_mn else b.float().t()
        y = alpha * (A @ Bm)
        if bias is not None:
            y = y + bias
        if aux_out is not None:
            aux_out.copy_(y)
        y = act_fn(y, act)
        if residual is not None:
            y = y + residual.float()
        if out is None:
            out = torch.empty(y.shape, dtype=out_dtype or a.dtype)
        out.copy_(out.float() + y if accumulate else y)
        return out

    def layernorm_fwd(x, gamma, beta, eps, *, rows=None, ldx=None, row_index=None, out=None, out_dtype=torch.bfloat16,
                      stats=None, save_stats=False):
        D = gamma.numel()
        xr = x.reshape(-1, D)
        if row_index is not None:
            xr = xr[row_index.long()]
        y = F.layer_norm(xr.float(), (D,), gamma, beta, eps)
        if out is None:
            out = torch.empty(y.shape, dtype=out_dtype)
        out.view(-1, D).copy_(y)
        return out

    def attn_core(q, k, v, B, Sq, Skv, H, scale, add):
        sp = lambda t, S: t.float().reshape(B, S, H, 64).transpose(1, 2)
        s = sp(q, Sq) @ sp(k, Skv).transpose(-1, -2) * scale
        if add is not None:
            s = s + add
        return (s.softmax(-1) @ sp(v, Skv)).transpose(1, 2).reshape(B * Sq, H * 64)

    def attn_fwd(qkv, B, S, H, causal, out=None, lse=None, need_lse=True):
        D = H * 64
        add = torch.full((S, S), float("-inf")).triu_(1) if causal else None
        y = attn_core(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, S, S, H, 0.125, add)
        if out is None:
            out = torch.empty(B * S, D, dtype=torch.bfloat16)
        out.copy_(y)
        return out, lse

    def attn_fwd_bias(q, k, v, B, Sq, Skv, H, bias, scale=1.0, out=None, lse=None):
        y = attn_core(q, k, v, B, Sq, Skv, H, scale, bias)
        if out is None:
            out = torch.empty(B * Sq, H * 64, dtype=torch.bfloat16)
        out.copy_(y)
        return out

    def rmsnorm_fwd(x, gamma, eps, out):
        D = gamma.numel()
        xr = x.reshape(-1, D).float()
        out.view(-1, D).copy_(xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + eps) * gamma)
        return out

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

    def text_embed_fwd(ids, tok, pos, out=None, eot=None):
        y = tok[ids.long()] + pos
        if out is not None:
            out.copy_(y.view(out.shape))
        if eot is not None:
            eot.copy_(torch.arange(ids.shape[0]) * ids.shape[1] + ids.argmax(-1))
        return out, eot

    def pad_cast_bf16(src, Kpad, out=None):
        if out is None:
            out = torch.zeros(src.shape[0], Kpad, dtype=torch.bfloat16)
        out.zero_()
        out[:, :src.shape[1]].copy_(src)
        return out

    def cast_bf16(src, out=None):
        return src.bfloat16() if out is None else out.copy_(src)

    for name, fn in dict(gemm=gemm, layernorm_fwd=layernorm_fwd, attn_fwd=attn_fwd, attn_fwd_bias=attn_fwd_bias,
                         rmsnorm_fwd=rmsnorm_fwd, token_mean=token_mean, patch_im2col=patch_im2col,
                         vision_embed_fwd=vision_embed_fwd, text_embed_fwd=text_embed_fwd, pad_cast_bf16=pad_cast_bf16,
                         cast_bf16=cast_bf16).items():
        monkeypatch.setattr(ops, name, fn)


def test_clip_ff_encoder_sequencing_reproduces_the_oracle(monkeypatch):
    _install_cpu_ops(monkeypatch)
    from uniir_b200.clip_ff import CLIPFFEncoder, clip_ff_param_specs, t5_heads
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule
    name = "tiny-test"
    torch.manual_seed(3)
    orc = CLIPFeatureFusionOracle(name).eval()
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    d = CLIP_CONFIGS[name]
    root = torch.nn.Module()
    P = FlatParamModule(clip_ff_param_specs(d, t5_heads(name)), "cpu", root=root)
    missing, unexpected = root.load_state_dict(orc.state_dict(), strict=True)
    P.refresh_shadow(force=True)
    mb = 3
    enc = CLIPFFEncoder(P, d, t5_heads(name), mb, "cpu")
    enc.refresh_weights()
    g = torch.Generator().manual_seed(4)
    ids = torch.zeros(mb, d.context, dtype=torch.int32)
    for i in range(mb):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g).int(); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(mb, 3, d.image_size, d.image_size, generator=g)
    with torch.no_grad():
        ref = orc.encode_multimodal_input(ids.long(), img)
        out = enc.forward(ids, img).clone()
    cos = F.cosine_similarity(out, ref, dim=-1)
    assert cos.min().item() > 1 - 1e-3, cos            # bf16 operand storage is the only difference
    assert (out - ref).abs().max().item() < 3e-2 * ref.abs().max().item()
    # the intermediate features too: text tokens first, then image tokens (clip_ff.py:177)
    with torch.no_grad():
        t_ref, v_ref = orc.encode_text(ids.long()), orc.encode_image(img)
    assert F.cosine_similarity(enc.t_feat.flatten(1), t_ref.flatten(1), dim=-1).min() > 1 - 1e-3
    assert F.cosine_similarity(enc.v_feat.flatten(1), v_ref.flatten(1), dim=-1).min() > 1 - 1e-3
