"""Drop-in for the reference ``CLIPFeatureFusion`` (src/models/uniir_clip/clip_featurefusion/clip_ff.py:62-300;
SURVEY.md 8f-3) on the B200-native kernels: the encode path (``encode_multimodal_input``, ``forward(batch,
encode_mbeir_batch=True)`` / ``encode_mbeir_batch``) and the in-batch contrastive training step.

STATUS: the host logic (kernel sequencing of the forward and of the hand-written backward) is verified on CPU against
the oracle's autograd with every kernel wrapper replaced by its torch contract (tests/test_clip_ff_host_logic.py); the
kernels (RMS norm fwd/bwd, attention with score bias fwd/bwd + bias gradient, ReLU epilogue, token mean fwd/bwd),
the encode path and the training step are asserted against torch / the oracle on the GPU by tests/test_clip_ff_gpu.py
(green on a B200 since the end of round 1); the oracle is pinned to the reference's own class
(tests/test_oracle_clip_ref.py).
Training: ``engine.train_step(model, batch, FlatAdamW(model._flat))`` (gradient-cached, like CLIP_SF).

Arithmetic (restated in oracle/clip_ff_oracle.py, whose T5 stack is pinned against transformers' own ``T5Stack``):
  vision  upstream CLIP ViT *without pooling* (clip_ff.py:23-59): ln_post and ``@ proj`` on all 257 tokens;
  text    CLIP text trunk to ln_final on all 77 tokens, no text_projection (:110, :148-156);
  fusion  cat([text, image]) -> 2-layer T5 encoder (RMS norm, q/k/v/o without bias, unscaled scores + relative-
          position bias, ReLU FFN 768 -> 2048 -> 768) -> final RMS norm -> mean over the 334 tokens (:161-192);
          ``txt_mask`` / ``img_mask`` are accepted and ignored, exactly like the reference (:173-176).
Train mode: the T5 stack's dropout (T5Config default ``dropout_rate`` 0.1, which clip_ff.py:82-98 keeps) at
transformers' sites -- stack input, attention probabilities, both sub-layer outputs, relu(wi x), final norm -- with
counter-based masks (include/uniir_b200.h) that the gradient-cache re-forward and the backward regenerate; the oracle's
sites are pinned against transformers' T5Stack in train mode (tests/test_oracle_clip_ff.py).  ``model.eval()`` or
``set_dropout(0)`` give the deterministic encoder the parity tests use.
Kernels: the CLIP towers of clip_towers.py (tcgen05 GEMMs, tcgen05 / mma.sync attention, LayerNorm), plus
``uniir_rmsnorm_fwd``, ``uniir_attn_fwd_bias`` (S = 334 > the tcgen05 forward's window: mma.sync kernel with the bias
added in score units), the ReLU GEMM epilogue and ``uniir_token_mean``.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import lib as L
from . import ops
from .clip_towers import LN_EPS, Tower
from .params import CLIP_CONFIGS, CLIPDims, FlatParamModule, clip_param_specs

T5_FF, T5_BUCKETS, T5_MAX_DIST, T5_EPS, T5_DKV = 2048, 32, 128, 1e-6, 64


def t5_heads(model_name):
    return 2 if model_name == "tiny-test" else 12       # clip_ff.py:84-97: num_heads = 12 for both supported models


def clip_ff_param_specs(d: CLIPDims, heads: int, layers: int = 2):
    """Reference state-dict layout: ``clip_model.*`` without ``text_projection`` (deleted at clip_ff.py:110) and
    ``t5_layers.*`` with transformers' T5Stack names (no ``embed_tokens``: the stack runs on ``inputs_embeds``).
    q / k / v weights of a block are adjacent, so the fused [3 * inner, d_model] projection is a plain view."""
    specs = [("clip_model." + n, s, dc, std) for n, s, dc, std in clip_param_specs(d) if n != "text_projection"]
    D, inner = d.embed_dim, heads * T5_DKV
    # every t5 parameter is in the reference's weight-decay-0.2 group, norms and bias table included
    # (clip_featurefusion/train.py:200-206: the no-decay rule applies only to names without "t5")
    add = lambda n, shape, std: specs.append((n, tuple(shape), True, std))
    for i in range(layers):
        a, f = f"t5_layers.block.{i}.layer.0", f"t5_layers.block.{i}.layer.1"
        for nm in ("q", "k", "v"):
            add(f"{a}.SelfAttention.{nm}.weight", (inner, D), 0.02)
        add(f"{a}.SelfAttention.o.weight", (D, inner), 0.02)
        if i == 0:
            add(f"{a}.SelfAttention.relative_attention_bias.weight", (T5_BUCKETS, heads), 0.02)
        add(f"{a}.layer_norm.weight", (D,), -1.0)
        add(f"{f}.DenseReluDense.wi.weight", (T5_FF, D), 0.02)
        add(f"{f}.DenseReluDense.wo.weight", (D, T5_FF), 0.02)
        add(f"{f}.layer_norm.weight", (D,), -1.0)
    add("t5_layers.final_layer_norm.weight", (D,), -1.0)
    return specs


def t5_relative_position_bucket(S, num_buckets=T5_BUCKETS, max_distance=T5_MAX_DIST):
    """int64 [S, S] bucket of (key position - query position), bidirectional (transformers T5Attention.
    _relative_position_bucket): index arithmetic on the host, done once per sequence length."""
    pos = torch.arange(S)
    rel = pos[None, :] - pos[:, None]
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return out + torch.where(rel < max_exact, rel, large)


class _T5Acts:
    """Saved activations of one T5 block for one micro-batch."""

    def __init__(self, T, D, inner, B, H, S, dev):
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.h_a = torch.empty(T, D, **bf)                 # rmsnorm(x_a): operand of the q|k|v weight gradient
        self.qkv = torch.empty(T, 3 * inner, **bf)
        self.att = torch.empty(T, inner, **bf)
        self.lse = torch.empty(B, H, S, device=dev, dtype=torch.float32)
        self.h_f = torch.empty(T, D, **bf)                 # rmsnorm(x_f)
        self.ff = torch.empty(T, T5_FF, **bf)              # relu(wi h_f): also the ReLU gradient mask (ff > 0)


class CLIPFFEncoder:
    """Un-pooled CLIP towers + T5 fusion stack for a fixed micro-batch: forward, and (train=True) forward with saved
    activations + the hand-written backward.  The residual stream of the T5 stack flows through 2 * layers + 1
    buffers when activations are kept (the inputs of the RMS norms are needed again), in place otherwise."""

    def __init__(self, P: FlatParamModule, d: CLIPDims, heads: int, mb: int, dev, layers: int = 2, train: bool = False):
        self.P, self.d, self.heads, self.mb, self.dev, self.layers, self.train = P, d, heads, mb, dev, layers, train
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        cm = "clip_model."
        self.vis = Tower(P, cm + "visual.transformer", d.v_width, d.v_layers, d.v_heads, d.v_tokens, False, mb, dev, train)
        self.txt = Tower(P, cm + "transformer", d.t_width, d.t_layers, d.t_heads, d.context, True, mb, dev, train)
        self.conv_w16 = torch.zeros(d.v_width, d.patch_k_pad, **bf)
        self.col = torch.empty(mb * d.grid * d.grid, d.patch_k_pad, **bf)
        self.patch = torch.empty(mb * d.grid * d.grid, d.v_width, **bf)
        self.x0 = torch.empty(mb, d.v_tokens, d.v_width, **f32)
        self.v_tok = torch.empty(mb * d.v_tokens, d.v_width, **bf)
        self.v_feat = torch.empty(mb, d.v_tokens, d.embed_dim, **f32)
        self.t_feat = torch.empty(mb, d.context, d.t_width, **f32)
        self.eot = torch.zeros(mb, device=dev, dtype=torch.int32)
        self.S = d.context + d.v_tokens                      # 77 + 257 = 334 fused tokens
        T, D, inner = mb * self.S, d.embed_dim, heads * T5_DKV
        self.inner = inner
        n_x = 2 * layers + 1 if train else 1
        self.xs = [torch.empty(mb, self.S, D, **f32) for _ in range(n_x)]      # fp32 residual stream of the T5 stack
        self.x = self.xs[0]
        self.h = torch.empty(T, D, **bf)
        self.qkv = torch.empty(T, 3 * inner, **bf)
        self.att = torch.empty(T, inner, **bf)
        self.ff = torch.empty(T, T5_FF, **bf)
        self.y = torch.empty(mb, self.S, D, **f32)
        self.emb = torch.empty(mb, D, **f32)
        self.bucket = t5_relative_position_bucket(self.S).to(dev)
        self.bias = torch.empty(heads, self.S, self.S, **f32)
        self.dctx, self.p_drop, self._drop = None, 0.0, False      # T5 dropout (set_dropout / forward(drop_seed=...))
        if train:
            Tv, Tt = mb * d.v_tokens, mb * d.context
            self.acts = [_T5Acts(T, D, inner, mb, heads, self.S, dev) for _ in range(layers)]
            self.pre_m = torch.empty(Tv, **f32); self.pre_r = torch.empty(Tv, **f32)      # ln_pre
            self.vp_m = torch.empty(Tv, **f32); self.vp_r = torch.empty(Tv, **f32)        # ln_post, all tokens
            self.tp_m = torch.empty(Tt, **f32); self.tp_r = torch.empty(Tt, **f32)        # ln_final, all tokens
            self.dX = torch.empty(mb, self.S, D, **f32)          # gradient of the residual stream
            self.dXb = torch.empty(T, D, **bf)
            self.d_h = torch.empty(T, D, **f32)                  # gradient of an RMS-norm output
            self.d_ff = torch.empty(T, T5_FF, **bf)
            self.d_att = torch.empty(T, inner, **bf)
            self.dqkv = torch.empty(T, 3 * inner, **bf)
            self.dbias = torch.zeros(heads, self.S, self.S, **f32)
            self.delta = torch.empty(mb * heads * self.S, **f32)
            self.d_t = torch.empty(Tt, d.t_width, **f32)
            self.d_vb = torch.empty(Tv, d.embed_dim, **bf)
            self.d_vtok = torch.empty(Tv, d.v_width, **bf)
            self.d_patch = torch.empty(mb * d.grid * d.grid, d.v_width, **bf)
            self.d_conv_pad = torch.zeros(d.v_width, d.patch_k_pad, **f32)
            self.dx0 = torch.empty(mb, d.v_tokens, d.v_width, **f32)
            self.ids = None

    # ---- dropout of the T5 stack in train mode (transformers T5Config.dropout_rate = 0.1, clip_ff.py:82-98 leaves the
    # default): site 0 = the stack input (T5Stack.forward: dropout(inputs_embeds)); 1 + 4 i + {0: attention probabilities,
    # 1: attention output, 2: relu(wi x), 3: wo output} of block i (T5Attention / T5LayerSelfAttention / T5DenseActDense /
    # T5LayerFF); 1 + 4 layers = the output of the final layer norm.  Counter-based masks: include/uniir_b200.h. ----
    def set_dropout(self, p):
        self.p_drop = float(p)
        self.dctx = ops.drop_ctx(self.dev) if p > 0 else None

    def refresh_weights(self):
        """bf16 copy of the (K-padded) patch-embedding weight and the [H, S, S] relative-position bias table
        (a gather of the [32, H] embedding: once per weight update, not per batch)."""
        d = self.d
        ops.pad_cast_bf16(self.P.w("clip_model.visual.conv1.weight").view(d.v_width, d.patch_k), d.patch_k_pad, out=self.conv_w16)
        w = self.P.w("t5_layers.block.0.layer.0.SelfAttention.relative_attention_bias.weight")      # [32, H]
        self.bias.copy_(w[self.bucket].permute(2, 0, 1))

    def _fused_qkv(self, i, bufs=None):
        decay, off, n, shape = self.P._index[f"t5_layers.block.{i}.layer.0.SelfAttention.q.weight"]
        inner, D = shape
        return (bufs or self.P.flat16)[decay][off:off + 3 * inner * D].view(3 * inner, D)

    def forward(self, ids, img, save=False, drop_seed=None, item0=0):
        """ids int32 [mb, 77], img fp32 [mb, 3, H, W] -> fused embedding fp32 [mb, embed_dim] (buffer reused).
        drop_seed (int, needs set_dropout): train-mode dropout masks of that seed; item0 = index of the micro-batch's
        first item in the batch."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        dr = self._drop = drop_seed is not None and self.dctx is not None
        if dr:
            self.dctx.copy_(torch.tensor([int(drop_seed) & 0x7FFFFFFF, int(item0)], dtype=torch.int32), non_blocking=True)
        ctx, pd = self.dctx, self.p_drop
        w, w16 = P.w, P.w16
        cm = "clip_model."
        assert not save or self.train
        # ---- vision tower, all tokens (clip_ff.py:35-59)
        ops.patch_im2col(img, d.patch, d.patch_k_pad, out=self.col)
        ops.gemm(self.col, self.conv_w16, out=self.patch)
        ops.vision_embed_fwd(self.patch, w(cm + "visual.class_embedding"), w(cm + "visual.positional_embedding"),
                             mb, d.v_tokens, d.v_width, out=self.x0)
        ops.layernorm_fwd(self.x0, w(cm + "visual.ln_pre.weight"), w(cm + "visual.ln_pre.bias"), LN_EPS, out=self.vis.x,
                          stats=((self.pre_m, self.pre_r) if save else None))
        xo = self.vis.forward(save)
        ops.layernorm_fwd(xo, w(cm + "visual.ln_post.weight"), w(cm + "visual.ln_post.bias"), LN_EPS, out=self.v_tok,
                          stats=((self.vp_m, self.vp_r) if save else None))
        ops.gemm(self.v_tok, w16(cm + "visual.proj"), b_mn=True, out=self.v_feat.view(mb * d.v_tokens, d.embed_dim))
        # ---- text trunk, all tokens, no projection (clip_ff.py:148-156)
        ops.text_embed_fwd(ids, w(cm + "token_embedding.weight"), w(cm + "positional_embedding"),
                           out=self.txt.x.view(mb, d.context, d.t_width), eot=self.eot)
        if save:
            self.ids = ids
        xo = self.txt.forward(save)
        ops.layernorm_fwd(xo, w(cm + "ln_final.weight"), w(cm + "ln_final.bias"), LN_EPS,
                          out=self.t_feat.view(mb * d.context, d.t_width), stats=((self.tp_m, self.tp_r) if save else None))
        # ---- cat([text, image], dim=1) (clip_ff.py:177)
        self.x[:, :d.context].copy_(self.t_feat)
        self.x[:, d.context:].copy_(self.v_feat)
        # ---- T5 encoder stack (clip_ff.py:178-183)
        inner, H, D = self.inner, self.heads, d.embed_dim
        x_in = self.x.view(mb * S, D)
        if dr:
            ops.dropout(x_in, ctx, 0, S * D, pd)
        for i in range(self.layers):
            a, f = f"t5_layers.block.{i}.layer.0", f"t5_layers.block.{i}.layer.1"
            act = self.acts[i] if save else None
            h_a, qkv, att, h_f, ff = (act.h_a, act.qkv, act.att, act.h_f, act.ff) if save else (self.h, self.qkv, self.att, self.h, self.ff)
            x_mid = self.xs[2 * i + 1].view(mb * S, D) if save else x_in
            x_out = self.xs[2 * i + 2].view(mb * S, D) if save else x_in
            ops.rmsnorm_fwd(x_in, w(f"{a}.layer_norm.weight"), T5_EPS, out=h_a)
            ops.gemm(h_a, self._fused_qkv(i), out=qkv)
            ops.attn_fwd_bias(qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:], mb, S, S, H, self.bias,
                              scale=1.0, out=att, lse=(act.lse if save else None), drop=((ctx, 1 + 4 * i, pd) if dr else None))
            if dr:      # hidden = hidden + dropout(o(attn)); when x_mid aliases x_in the projection goes through self.y
                tmp = self.y.view(mb * S, D) if x_mid is x_in else x_mid
                ops.gemm(att, w16(f"{a}.SelfAttention.o.weight"), out=tmp)
                ops.dropout(tmp, ctx, 2 + 4 * i, S * D, pd, residual=x_in, out=x_mid)
            else:
                ops.gemm(att, w16(f"{a}.SelfAttention.o.weight"), residual=x_in, out=x_mid)
            ops.rmsnorm_fwd(x_mid, w(f"{f}.layer_norm.weight"), T5_EPS, out=h_f)
            ops.gemm(h_f, w16(f"{f}.DenseReluDense.wi.weight"), act=L.ACT_RELU, out=ff)
            if dr:
                ops.dropout(ff, ctx, 3 + 4 * i, S * T5_FF, pd)          # the saved ff is the dropped one (operand of wo)
                tmp = self.y.view(mb * S, D) if x_out is x_mid else x_out
                ops.gemm(ff, w16(f"{f}.DenseReluDense.wo.weight"), out=tmp)
                ops.dropout(tmp, ctx, 4 + 4 * i, S * D, pd, residual=x_mid, out=x_out)
            else:
                ops.gemm(ff, w16(f"{f}.DenseReluDense.wo.weight"), residual=x_mid, out=x_out)
            x_in = x_out
        ops.rmsnorm_fwd(x_in, w("t5_layers.final_layer_norm.weight"), T5_EPS, out=self.y)
        if dr:
            ops.dropout(self.y.view(mb * S, D), ctx, 1 + 4 * self.layers, S * D, pd)
        return ops.token_mean(self.y, out=self.emb)           # mean pooling (clip_ff.py:185-192)

    def backward(self, d_emb):
        """d_emb fp32 [mb, embed_dim] = dLoss/d(embedding) of the last forward(save=True).  Accumulates parameter
        gradients into P.grad."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        w, w16, g = P.w, P.w16, P.g
        cm = "clip_model."
        inner, H, D = self.inner, self.heads, d.embed_dim
        T = mb * S
        dX = self.dX.view(T, D)
        dr, ctx, pd = self._drop, self.dctx, self.p_drop
        # ---- mean pooling and the final RMS norm
        ops.token_mean_bwd(d_emb, S, self.d_h)
        if dr:
            ops.dropout(self.d_h, ctx, 1 + 4 * self.layers, S * D, pd)
        ops.rmsnorm_bwd(self.d_h, self.xs[2 * self.layers].view(T, D), w("t5_layers.final_layer_norm.weight"), T5_EPS, dX,
                        g("t5_layers.final_layer_norm.weight"), accumulate=False)
        self.dbias.zero_()
        for i in reversed(range(self.layers)):
            a, f = f"t5_layers.block.{i}.layer.0", f"t5_layers.block.{i}.layer.1"
            act = self.acts[i]
            x_a, x_f = self.xs[2 * i].view(T, D), self.xs[2 * i + 1].view(T, D)
            # ---- feed-forward: x_out = x_f + wo(relu(wi(rmsnorm(x_f))))   [train mode: + dropout(wo(dropout(relu(.))))]
            if dr:
                ops.dropout(dX, ctx, 4 + 4 * i, S * D, pd, out=self.dXb)      # branch gradient; dX itself is the residual's
            else:
                ops.cast_bf16(dX, out=self.dXb)
            ops.wgrad(self.dXb, act.ff, g(f"{f}.DenseReluDense.wo.weight"))
            ops.gemm(self.dXb, w16(f"{f}.DenseReluDense.wo.weight"), b_mn=True, act=L.ACT_RELU, act_grad_src=act.ff, out=self.d_ff)
            if dr:      # act.ff > 0 already zeroes the dropped units; this applies the 1 / (1 - p) of the kept ones
                ops.dropout(self.d_ff, ctx, 3 + 4 * i, S * T5_FF, pd)
            ops.wgrad(self.d_ff, act.h_f, g(f"{f}.DenseReluDense.wi.weight"))
            ops.gemm(self.d_ff, w16(f"{f}.DenseReluDense.wi.weight"), b_mn=True, out=self.d_h)
            ops.rmsnorm_bwd(self.d_h, x_f, w(f"{f}.layer_norm.weight"), T5_EPS, dX, g(f"{f}.layer_norm.weight"))
            # ---- attention: x_f = x_a + o(attn(q|k|v(rmsnorm(x_a))))
            if dr:
                ops.dropout(dX, ctx, 2 + 4 * i, S * D, pd, out=self.dXb)
            else:
                ops.cast_bf16(dX, out=self.dXb)
            ops.wgrad(self.dXb, act.att, g(f"{a}.SelfAttention.o.weight"))
            ops.gemm(self.dXb, w16(f"{a}.SelfAttention.o.weight"), b_mn=True, out=self.d_att)
            q, k, v = act.qkv[:, :inner], act.qkv[:, inner:2 * inner], act.qkv[:, 2 * inner:]
            ops.attn_bwd_bias(q, k, v, act.att, self.d_att, act.lse, mb, S, S, H, self.bias, self.dbias,
                              self.dqkv[:, :inner], self.dqkv[:, inner:2 * inner], self.dqkv[:, 2 * inner:], scale=1.0,
                              delta=self.delta, drop=((ctx, 1 + 4 * i, pd) if dr else None))
            ops.wgrad(self.dqkv, act.h_a, self._fused_qkv(i, P.grad))
            ops.gemm(self.dqkv, self._fused_qkv(i), b_mn=True, out=self.d_h)
            ops.rmsnorm_bwd(self.d_h, x_a, w(f"{a}.layer_norm.weight"), T5_EPS, dX, g(f"{a}.layer_norm.weight"))
        # the position bias is computed once (block 0) and shared by both blocks: one gradient table, scattered back
        # into the [32, H] bucket embedding (index arithmetic; 334 x 334 x 12 values)
        g("t5_layers.block.0.layer.0.SelfAttention.relative_attention_bias.weight").index_add_(
            0, self.bucket.flatten(), self.dbias.permute(1, 2, 0).reshape(-1, H))
        # ---- split the gradient of cat([text, image])
        if dr:
            ops.dropout(dX, ctx, 0, S * D, pd)
        self.d_t.view(mb, d.context, D).copy_(self.dX[:, :d.context])
        self.d_vb.view(mb, d.v_tokens, D).copy_(self.dX[:, d.context:])
        # ---- text trunk
        ops.layernorm_bwd(self.d_t, self.txt.x_out, self.tp_m, self.tp_r, w(cm + "ln_final.weight"), self.txt.dx,
                          g(cm + "ln_final.weight"), g(cm + "ln_final.bias"), accumulate=False, dx_bf16=self.txt.dxb,
                          dcolsum=self.txt.top_bias_grad())
        self.txt.backward(top_bias_done=True)
        ops.text_embed_bwd(self.txt.dx, self.ids, g(cm + "token_embedding.weight"), g(cm + "positional_embedding"))
        # ---- vision tower: v_feat = ln_post(x_out) @ proj on every token
        ops.wgrad(self.v_tok, self.d_vb, g(cm + "visual.proj"))                    # [v_width, E] += v_tok^T d_v
        ops.gemm(self.d_vb, w16(cm + "visual.proj"), out=self.d_vtok)              # d_vtok = d_v @ proj^T
        ops.layernorm_bwd(self.d_vtok, self.vis.x_out, self.vp_m, self.vp_r, w(cm + "visual.ln_post.weight"), self.vis.dx,
                          g(cm + "visual.ln_post.weight"), g(cm + "visual.ln_post.bias"), accumulate=False,
                          dx_bf16=self.vis.dxb, dcolsum=self.vis.top_bias_grad())
        self.vis.backward(top_bias_done=True)
        ops.layernorm_bwd(self.vis.dx, self.x0, self.pre_m, self.pre_r, w(cm + "visual.ln_pre.weight"), self.dx0,
                          g(cm + "visual.ln_pre.weight"), g(cm + "visual.ln_pre.bias"), accumulate=False)
        ops.vision_embed_bwd(self.dx0, mb, d.v_tokens, d.v_width, self.d_patch, g(cm + "visual.class_embedding"),
                             g(cm + "visual.positional_embedding"))
        self.d_conv_pad.zero_()
        ops.wgrad(self.d_patch, self.col, self.d_conv_pad)
        ops.unpad_accum(self.d_conv_pad, g(cm + "visual.conv1.weight").view(d.v_width, d.patch_k))


class CLIPFeatureFusion(nn.Module):
    def __init__(self, model_name="ViT-B/32", device="cuda", jit=False, download_root=None, config=None,
                 micro_batch=64, seed=2023):
        super().__init__()
        if model_name not in ("ViT-B/32", "ViT-L/14", "tiny-test"):
            raise NotImplementedError("Only ViT-B/32 and ViT-L/14 are supported.")     # clip_ff.py:100
        L.load()  # fail loudly if the CUDA extension is missing: there is no fallback path
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.UniirError("uniir_b200.CLIPFeatureFusion runs on CUDA only (no CPU fallback); "
                               "use oracle/ for CPU reference numbers")
        self.model_name, self.dims, self.device_ = model_name, CLIP_CONFIGS[model_name], dev
        if self.dims.t_width != self.dims.embed_dim:
            raise ValueError("CLIP_FF concatenates text tokens (text width) and projected image tokens (embed_dim)")
        self.heads = t5_heads(model_name)
        self._flat = FlatParamModule(clip_ff_param_specs(self.dims, self.heads), dev, seed=seed, root=self)
        self.loss_function = nn.CrossEntropyLoss()        # attribute parity (clip_ff.py:79)
        if config is not None:                            # clip_ff.py:102-107
            self.gather_embeddings = config.model.gather_embeddings
            self.in_batch_neg_num = config.data_config.in_batch_neg_num
        else:
            self.gather_embeddings = None
            self.in_batch_neg_num = None
        self.micro_batch = micro_batch
        self.train_micro_batch = 32
        self._enc = {}
        self._nce = None
        self._train_state = None
        # transformers' T5Config() default, which clip_ff.py:82-98 keeps: dropout_rate 0.1 on the stack input, the
        # attention probabilities, both sub-layer outputs, relu(wi x) and the final norm -- active in train mode
        self.t5_dropout_rate = 0.1
        self._drop_gen, self._drop_seed0 = None, seed

    def set_dropout(self, rate):
        """Override the T5 stack's train-mode dropout rate (0 switches it off; `model.eval()` does too)."""
        self.t5_dropout_rate = float(rate)

    def _dropout_active(self):
        return self.training and self.t5_dropout_rate > 0

    def _next_drop_seed(self):
        import torch.distributed as dist
        if self._drop_gen is None:
            rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
            self._drop_gen = torch.Generator(device="cpu").manual_seed(self._drop_seed0 * 1000003 + rank)
        return int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self._drop_gen))

    def _with_dropout(self, enc):
        if getattr(enc, "_drop_rate", None) != self.t5_dropout_rate:
            enc.set_dropout(self.t5_dropout_rate)
            enc._drop_rate = self.t5_dropout_rate
        return enc

    def build_optimizer(self, learning_rate=1e-5, t5_learning_rate=5e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2):
        """The reference's three AdamW groups (clip_featurefusion/train.py:49-63, 200-206) on the fused optimizer: CLIP
        gains / biases without decay, the other CLIP parameters and every T5 parameter with 0.2, T5 at its own rate."""
        from .engine import FlatAdamW
        scale = t5_learning_rate / learning_rate
        return FlatAdamW(self._flat, lr=learning_rate, betas=betas, eps=eps, weight_decay=weight_decay,
                         lr_scale=lambda name: scale if name.startswith("t5_layers.") else 1.0)

    # ---- reference surface ---------------------------------------------------------
    def get_img_preprocess_fn(self):
        from .preprocess import clip_image_preprocess
        return clip_image_preprocess(self.dims.image_size)

    def get_tokenizer(self):
        from .preprocess import clip_tokenizer
        return clip_tokenizer(context_length=77, vocab=self.dims.vocab)

    def get_logit_scale(self):
        return self.get_parameter("clip_model.logit_scale").exp()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat.rebind_after_apply()
        return out

    def _encoder(self, mb):
        key = ("embed", mb)
        if key not in self._enc:
            self._enc = {key: CLIPFFEncoder(self._flat, self.dims, self.heads, mb, self.device_)}
        return self._with_dropout(self._enc[key])

    @torch.no_grad()
    def encode_multimodal_input(self, txt_tensor, img_tensor, txt_mask=None, img_mask=None):
        """[B,77] int, [B,3,H,W] fp32 -> fp32 [B, embed_dim]; the masks are ignored like in the reference.  In train mode
        the T5 stack runs with dropout (fresh masks per call), like the reference's module."""
        dev = self.device_
        seed = self._next_drop_seed() if self._dropout_active() else None
        ids = txt_tensor.to(dev, torch.int32).contiguous()
        img = img_tensor.to(dev, torch.float32).contiguous()
        n = ids.shape[0]
        mb = min(self.micro_batch, n)
        enc = self._encoder(mb)
        self._flat.refresh_shadow()
        enc.refresh_weights()
        out = torch.empty(n, self.dims.embed_dim, device=dev, dtype=torch.float32)
        for s in range(0, n, mb):
            e = min(n, s + mb)
            i_c, im_c = ids[s:e], img[s:e]
            if e - s < mb:      # ragged tail: zero-padded items, dropped below
                i_c = torch.cat([i_c, i_c.new_zeros(mb - (e - s), ids.shape[1])])
                im_c = torch.cat([im_c, im_c.new_zeros(mb - (e - s), *img.shape[1:])])
            out[s:e] = enc.forward(i_c.contiguous(), im_c.contiguous(), drop_seed=seed, item0=s)[:e - s]
        return out

    # ---- training (clip_ff.py:197-270): the in-batch / hard-negative contrastive loss of CLIP_SF on the fused
    # embeddings, with embedding-level gradient caching like clip_sf.py: phase A embeds every micro-batch without
    # keeping activations, the loss kernels give dLoss/d(embedding), phase C re-runs each micro-batch with saved
    # activations and backpropagates.  Driven by engine.train_step(model, batch, FlatAdamW(model._flat)). ----
    def compute_inbatch_contrastive_loss(self, batch):
        """Reference-compatible entry: {"loss", "accuracy"} (0-d tensors).  With autograd enabled the loss is a graph
        node whose ``.backward()`` runs ``_loss_backward`` and fills ``p.grad`` (the reference loop,
        uniir_clip/engine.py:26-45); ``engine.train_step`` calls the two halves directly."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd_bridge import bridged_loss
            loss, acc = bridged_loss(self._flat, lambda: self._loss_forward(batch), lambda g: self._loss_backward(batch, g))
            return {"loss": loss, "accuracy": acc}
        loss, acc = self._loss_forward(batch)
        return {"loss": loss.clone(), "accuracy": acc.clone()}

    def _world(self):
        from . import dist_utils as du   # noqa: F401
        import torch.distributed as dist
        if getattr(self, "gather_embeddings", False):
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("gather_embeddings=True needs an initialised process group "
                                   "(reference clip_ff.py:231-232 calls torch.distributed.nn.all_gather)")
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _train_encoder(self, n_items):
        mb = min(self.train_micro_batch, n_items)
        if n_items % mb != 0 or mb % 8 != 0:
            mb = next((m for m in range(mb - mb % 8, 7, -8) if n_items % m == 0), None)
            if mb is None:
                raise ValueError(f"encoder batch of {n_items} items cannot be tiled by a micro-batch that is a multiple of 8")
        key = ("train", mb)
        if key not in self._enc:
            self._enc = {key: CLIPFFEncoder(self._flat, self.dims, self.heads, mb, self.device_, train=True)}
        return self._with_dropout(self._enc[key])

    @torch.no_grad()
    def _loss_forward(self, batch):
        from . import dist_utils as du
        im = batch["index_mapping"]
        dev, E, P = self.device_, self.dims.embed_dim, self._flat
        q_idx = torch.tensor(im["query"], dtype=torch.int32).flatten().to(dev)
        p_idx = torch.tensor(im["pos_cand"], dtype=torch.int32).flatten().to(dev)
        ids = batch["txt_batched"].to(dev, torch.int32).contiguous()
        img = batch["image_batched"].to(dev, torch.float32).contiguous()
        bs, n_items = q_idx.numel(), ids.shape[0]
        rank, world = self._world()
        enc = self._train_encoder(n_items)
        P.refresh_shadow()
        enc.refresh_weights()
        emb = torch.empty(n_items, E, device=dev, dtype=torch.float32)
        seed = self._next_drop_seed() if self._dropout_active() else None      # phase A and the phase-C re-forward share it
        for s in range(0, n_items, enc.mb):                     # phase A
            emb[s:s + enc.mb] = enc.forward(ids[s:s + enc.mb], img[s:s + enc.mb], save=False, drop_seed=seed, item0=s)
        st = dict(q_idx=q_idx, p_idx=p_idx, bs=bs, n_items=n_items, rank=rank, world=world, enc=enc, ids=ids, img=img, seed=seed)
        P.ensure_grad()
        ls = P.w("clip_model.logit_scale")
        if "neg_cand_list" in im:       # hard-negative branch (clip_ff.py:234-258): no gather, fused forward + backward
            n_idx = torch.tensor(im["neg_cand_list"], dtype=torch.int32).to(dev).contiguous()
            st["hard_dls"] = torch.zeros((), device=dev, dtype=torch.float32)
            loss, acc, st["hard_d_emb"] = ops.hardneg_loss(emb, q_idx, p_idx, n_idx, int(self.in_batch_neg_num or 0), ls,
                                                           st["hard_dls"])
        else:
            nce = self._nce
            if nce is None or (nce.bs, nce.n, nce.E) != (bs, bs * world, E):
                nce = self._nce = ops.InfoNCE(bs, bs * world, E, dev)
            pn = nce.normalize(emb, q_idx, p_idx, rank * bs)
            if world > 1:
                du.all_gather_rows(pn.clone(), out=nce.all_p)
            loss, acc = nce.forward(ls, rank * bs)
        self._train_state = st
        return loss, acc

    @torch.no_grad()
    def _loss_backward(self, batch=None, g_loss=None):
        from . import dist_utils as du
        st = self._train_state
        assert st is not None, "backward without forward"
        P, enc, E = self._flat, st["enc"], self.dims.embed_dim
        bs, off = st["bs"], st["rank"] * st["bs"]
        go = None if g_loss is None else g_loss.to(torch.float32).reshape(()).contiguous()
        if "hard_d_emb" in st:
            d_emb = st["hard_d_emb"] if go is None else st["hard_d_emb"] * go
            P.g("clip_model.logit_scale").add_(st["hard_dls"] if go is None else st["hard_dls"] * go)
        else:
            nce = self._nce
            d_qn, d_all_p = nce.backward(P.w("clip_model.logit_scale"), off, P.g("clip_model.logit_scale"), grad_out=go)
            d_pn = du.reduce_scatter_rows(d_all_p) if st["world"] > 1 else d_all_p
            d_emb = torch.zeros(st["n_items"], E, device=self.device_, dtype=torch.float32)
            nce.normalize_bwd(st["q_idx"], st["p_idx"], d_pn, d_emb, off)
        for s in range(0, st["n_items"], enc.mb):               # phase C
            enc.forward(st["ids"][s:s + enc.mb], st["img"][s:s + enc.mb], save=True, drop_seed=st["seed"], item0=s)
            enc.backward(d_emb[s:s + enc.mb].contiguous())
        self._train_state = None

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        assert id_list is not None, "id_list must be provided."
        assert isinstance(id_list[0], int), "id_list must be hashed to int."
        embeddings = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"],
                                                  batch["image_mask_batched"])
        assert embeddings.size(0) == len(id_list), "embeddings and id_batched must have the same batch size."
        return embeddings, id_list

    def forward(self, batch, encode_mbeir_batch=False):
        if encode_mbeir_batch:
            return self.encode_mbeir_batch(batch)
        return self.compute_inbatch_contrastive_loss(batch)
