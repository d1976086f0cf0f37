"""Forward / backward orchestration of the CLIP towers on the C-ABI kernels.

Host code only sequences kernels (PyTorch = device memory + stream); every FLOP happens in
libuniir_b200.so.  Semantics follow upstream openai/CLIP as restated in the reference
(vision: clip_featurefusion/clip_ff.py:35-59 with CLS pooling before ln_post; text:
clip_ff.py:148-156 + EOT pooling + text_projection), called from clip_sf.py:43-47.

Numerics: bf16 GEMM operands, fp32 accumulate, fp32 residual stream, fp32 LayerNorm
statistics, fp32 softmax -- the BASELINE "bf16" configuration; parity vs the fp32 oracle is
checked in tests/ with an explicit tolerance.
"""
from __future__ import annotations

import torch

from . import lib as L
from . import ops
from .params import CLIPDims, FlatParamModule

LN_EPS = 1e-5


class _LayerActs:
    """Saved activations of one residual block for one micro-batch (allocated once, reused)."""

    def __init__(self, T, D, B, H, S, dev):
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.x_in = torch.empty(T, D, **f32)
        self.h1 = torch.empty(T, D, **bf)
        self.m1 = torch.empty(T, **f32); self.r1 = torch.empty(T, **f32)
        self.qkv = torch.empty(T, 3 * D, **bf)
        self.att = torch.empty(T, D, **bf)
        self.lse = torch.empty(B, H, S, **f32)
        self.x_mid = torch.empty(T, D, **f32)
        self.h2 = torch.empty(T, D, **bf)
        self.m2 = torch.empty(T, **f32); self.r2 = torch.empty(T, **f32)
        self.pre = torch.empty(T, 4 * D, **bf)
        self.act = torch.empty(T, 4 * D, **bf)


class Tower:
    """A stack of upstream-CLIP ResidualAttentionBlocks: x += attn(ln_1(x)); x += mlp(ln_2(x))."""

    def __init__(self, params: FlatParamModule, prefix, width, layers, heads, S, causal, mb, dev,
                 train):
        self.P, self.prefix = params, prefix
        self.D, self.nl, self.H, self.S, self.causal, self.mb = width, layers, heads, S, causal, mb
        self.T = mb * S
        self.dev = dev
        T, D = self.T, width
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.x = torch.empty(T, D, **f32)          # residual stream
        if train:
            self.acts = [_LayerActs(T, D, mb, heads, S, dev) for _ in range(layers)]
            self.dx = torch.empty(T, D, **f32)
            self.dxb = torch.empty(T, D, **bf)
            self.d_big = torch.empty(T, 4 * D, **bf)   # d_pre
            self.d_h = torch.empty(T, D, **bf)         # d_h1 / d_h2 / d_att
            self.dqkv = torch.empty(T, 3 * D, **bf)
            self.delta = torch.empty(mb * heads * S, **f32)
            # the no-save forward (gradient-cache phase A) borrows block 0's buffers as scratch
            a0 = self.acts[0]
            self.h, self.qkv, self.att, self.big = a0.h1, a0.qkv, a0.att, a0.act
        else:
            self.acts = None
            self.h = torch.empty(T, D, **bf)
            self.qkv = torch.empty(T, 3 * D, **bf)
            self.att = torch.empty(T, D, **bf)
            self.big = torch.empty(T, 4 * D, **bf)

    def _n(self, i, leaf):
        return f"{self.prefix}.resblocks.{i}.{leaf}"

    # ---- forward over all blocks, in place on self.x --------------------------------
    def forward(self, save):
        P, x, B, S, H = self.P, self.x, self.mb, self.S, self.H
        for i in range(self.nl):
            if save:
                a = self.acts[i]
                a.x_in.copy_(x)
                h1, qkv, att, h2, pre, act = a.h1, a.qkv, a.att, a.h2, a.pre, a.act
                ops.layernorm_fwd(x, P.w(self._n(i, "ln_1.weight")), P.w(self._n(i, "ln_1.bias")), LN_EPS,
                                  out=h1, stats=(a.m1, a.r1))
            else:
                h1 = h2 = self.h
                qkv, att, pre, act = self.qkv, self.att, None, self.big
                ops.layernorm_fwd(x, P.w(self._n(i, "ln_1.weight")), P.w(self._n(i, "ln_1.bias")), LN_EPS, out=h1)
            ops.gemm(h1, P.w16(self._n(i, "attn.in_proj_weight")), bias=P.w(self._n(i, "attn.in_proj_bias")), out=qkv)
            ops.attn_fwd(qkv, B, S, H, self.causal, out=att, lse=(self.acts[i].lse if save else None),
                         need_lse=save)
            ops.gemm(att, P.w16(self._n(i, "attn.out_proj.weight")), bias=P.w(self._n(i, "attn.out_proj.bias")),
                     residual=x, out=x)
            if save:
                a.x_mid.copy_(x)
                ops.layernorm_fwd(x, P.w(self._n(i, "ln_2.weight")), P.w(self._n(i, "ln_2.bias")), LN_EPS,
                                  out=h2, stats=(a.m2, a.r2))
            else:
                ops.layernorm_fwd(x, P.w(self._n(i, "ln_2.weight")), P.w(self._n(i, "ln_2.bias")), LN_EPS, out=h2)
            ops.gemm(h2, P.w16(self._n(i, "mlp.c_fc.weight")), bias=P.w(self._n(i, "mlp.c_fc.bias")),
                     act=L.ACT_QUICKGELU, aux_out=pre, out=act)
            ops.gemm(act, P.w16(self._n(i, "mlp.c_proj.weight")), bias=P.w(self._n(i, "mlp.c_proj.bias")),
                     residual=x, out=x)

    # ---- backward: self.dx holds dL/dx_out on entry, dL/dx_in on exit ----------------
    def backward(self):
        P, dx, dxb, B, S, H = self.P, self.dx, self.dxb, self.mb, self.S, self.H
        for i in reversed(range(self.nl)):
            a = self.acts[i]
            g = lambda leaf: P.g(self._n(i, leaf))
            w16 = lambda leaf: P.w16(self._n(i, leaf))
            # ---- MLP branch: x_out = x_mid + c_proj(quickgelu(c_fc(ln_2(x_mid))))
            ops.cast_bf16(dx, out=dxb)
            ops.wgrad(dxb, a.act, g("mlp.c_proj.weight"))
            ops.colsum_accum(dxb, g("mlp.c_proj.bias"))
            ops.gemm(dxb, w16("mlp.c_proj.weight"), b_mn=True, act=L.ACT_QUICKGELU, act_grad_src=a.pre,
                     out=self.d_big)
            ops.wgrad(self.d_big, a.h2, g("mlp.c_fc.weight"))
            ops.colsum_accum(self.d_big, g("mlp.c_fc.bias"))
            ops.gemm(self.d_big, w16("mlp.c_fc.weight"), b_mn=True, out=self.d_h)
            ops.layernorm_bwd(self.d_h, a.x_mid, a.m2, a.r2, P.w(self._n(i, "ln_2.weight")), dx,
                              g("ln_2.weight"), g("ln_2.bias"))
            # ---- attention branch: x_mid = x_in + out_proj(attn(in_proj(ln_1(x_in))))
            ops.cast_bf16(dx, out=dxb)
            ops.wgrad(dxb, a.att, g("attn.out_proj.weight"))
            ops.colsum_accum(dxb, g("attn.out_proj.bias"))
            ops.gemm(dxb, w16("attn.out_proj.weight"), b_mn=True, out=self.d_h)
            ops.attn_bwd(a.qkv, a.att, self.d_h, a.lse, B, S, H, self.causal, dqkv=self.dqkv, delta=self.delta)
            ops.wgrad(self.dqkv, a.h1, g("attn.in_proj_weight"))
            ops.colsum_accum(self.dqkv, g("attn.in_proj_bias"))
            ops.gemm(self.dqkv, w16("attn.in_proj_weight"), b_mn=True, out=self.d_h)
            ops.layernorm_bwd(self.d_h, a.x_in, a.m1, a.r1, P.w(self._n(i, "ln_1.weight")), dx,
                              g("ln_1.weight"), g("ln_1.bias"))


class CLIPEncoder:
    """Both CLIP towers + pooling/projection for a fixed micro-batch size."""

    def __init__(self, params: FlatParamModule, dims: CLIPDims, mb: int, dev, train: bool):
        self.P, self.d, self.mb, self.dev, self.train = params, dims, mb, dev, train
        d = dims
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.vis = Tower(params, "visual.transformer", d.v_width, d.v_layers, d.v_heads, d.v_tokens, False,
                         mb, dev, train)
        self.txt = Tower(params, "transformer", d.t_width, d.t_layers, d.t_heads, d.context, True, mb, dev,
                         train)
        self.conv_w16 = torch.zeros(d.v_width, d.patch_k_pad, **bf)
        self.col = torch.empty(mb * d.grid * d.grid, d.patch_k_pad, **bf)
        self.patch = torch.empty(mb * d.grid * d.grid, d.v_width, **bf)
        self.x0 = torch.empty(mb, d.v_tokens, d.v_width, **f32)
        self.cls_rows = (torch.arange(mb, device=dev, dtype=torch.int32) * d.v_tokens).contiguous()
        self.v_pool = torch.empty(mb, d.v_width, **bf)
        self.t_pool = torch.empty(mb, d.t_width, **bf)
        self.v_emb = torch.empty(mb, d.embed_dim, **f32)
        self.t_emb = torch.empty(mb, d.embed_dim, **f32)
        if train:
            self.vp_m = torch.empty(mb, **f32); self.vp_r = torch.empty(mb, **f32)
            self.tp_m = torch.empty(mb, **f32); self.tp_r = torch.empty(mb, **f32)
            self.pre_m = torch.empty(mb * d.v_tokens, **f32); self.pre_r = torch.empty(mb * d.v_tokens, **f32)
            self.d_pool_v = torch.empty(mb, d.v_width, **bf)
            self.d_pool_t = torch.empty(mb, d.t_width, **bf)
            self.d_patch = torch.empty(mb * d.grid * d.grid, d.v_width, **bf)
            self.d_conv_pad = torch.zeros(d.v_width, d.patch_k_pad, **f32)
            self.dx0 = torch.empty(mb, d.v_tokens, d.v_width, **f32)
        self.eot_rows = torch.zeros(mb, device=dev, dtype=torch.int32)
        self.ids = None
        # static I/O so that forward / forward+backward can be captured into CUDA graphs
        self.in_ids = torch.zeros(mb, d.context, device=dev, dtype=torch.int32)
        self.in_img = torch.zeros(mb, 3, d.image_size, d.image_size, **f32)
        self.in_mt = torch.ones(mb, device=dev, dtype=torch.int64)
        self.in_mv = torch.ones(mb, device=dev, dtype=torch.int64)
        self.out_emb = torch.zeros(mb, d.embed_dim, **f32)
        if train:
            self.in_demb = torch.zeros(mb, d.embed_dim, **f32)
            self.d_t = torch.zeros(mb, d.embed_dim, **bf)
            self.d_v = torch.zeros(mb, d.embed_dim, **bf)
        self._graphs = {}

    def refresh_weights(self):
        d = self.d
        ops.pad_cast_bf16(self.P.w("visual.conv1.weight").view(d.v_width, d.patch_k), d.patch_k_pad,
                          out=self.conv_w16)

    # ---- forward: returns (t_emb, v_emb) fp32 [mb, E] (buffers reused per call) -----
    def forward(self, ids, img, save):
        P, d, mb = self.P, self.d, self.mb
        assert ids.shape == (mb, d.context) and img.shape == (mb, 3, d.image_size, d.image_size)
        # vision
        ops.patch_im2col(img, d.patch, d.patch_k_pad, out=self.col)
        ops.gemm(self.col, self.conv_w16, out=self.patch)
        ops.vision_embed_fwd(self.patch, P.w("visual.class_embedding"), P.w("visual.positional_embedding"),
                             mb, d.v_tokens, d.v_width, out=self.x0)
        ops.layernorm_fwd(self.x0, P.w("visual.ln_pre.weight"), P.w("visual.ln_pre.bias"), LN_EPS,
                          out=self.vis.x, stats=((self.pre_m, self.pre_r) if save else None))
        self.vis.forward(save)
        ops.layernorm_fwd(self.vis.x, P.w("visual.ln_post.weight"), P.w("visual.ln_post.bias"), LN_EPS,
                          row_index=self.cls_rows, out=self.v_pool,
                          stats=((self.vp_m, self.vp_r) if save else None))
        ops.gemm(self.v_pool, P.w16("visual.proj"), b_mn=True, out=self.v_emb)
        # text
        ops.text_embed_fwd(ids, P.w("token_embedding.weight"), P.w("positional_embedding"),
                           out=self.txt.x.view(mb, d.context, d.t_width), eot=self.eot_rows)
        self.ids = ids
        self.txt.forward(save)
        ops.layernorm_fwd(self.txt.x, P.w("ln_final.weight"), P.w("ln_final.bias"), LN_EPS,
                          row_index=self.eot_rows, out=self.t_pool,
                          stats=((self.tp_m, self.tp_r) if save else None))
        ops.gemm(self.t_pool, P.w16("text_projection"), b_mn=True, out=self.t_emb)
        return self.t_emb, self.v_emb

    # ---- backward from d t_emb / d v_emb (bf16 [mb, E]); accumulates into P.grad -----
    def backward(self, d_t, d_v):
        P, d, mb = self.P, self.d, self.mb
        # ---------------- vision ----------------
        ops.wgrad(self.v_pool, d_v, P.g("visual.proj"))                 # [D,E] += pool^T d_v
        ops.gemm(d_v, P.w16("visual.proj"), out=self.d_pool_v)          # d_pool = d_v @ proj^T
        self.vis.dx.zero_()
        ops.layernorm_bwd(self.d_pool_v, self.vis.x, self.vp_m, self.vp_r, P.w("visual.ln_post.weight"),
                          self.vis.dx, P.g("visual.ln_post.weight"), P.g("visual.ln_post.bias"),
                          row_index=self.cls_rows, accumulate=False)
        self.vis.backward()
        ops.layernorm_bwd(self.vis.dx, self.x0, self.pre_m, self.pre_r, P.w("visual.ln_pre.weight"),
                          self.dx0, P.g("visual.ln_pre.weight"), P.g("visual.ln_pre.bias"), accumulate=False)
        ops.vision_embed_bwd(self.dx0, mb, d.v_tokens, d.v_width, self.d_patch, P.g("visual.class_embedding"),
                             P.g("visual.positional_embedding"))
        self.d_conv_pad.zero_()
        ops.wgrad(self.d_patch, self.col, self.d_conv_pad)
        ops.unpad_accum(self.d_conv_pad, P.g("visual.conv1.weight").view(d.v_width, d.patch_k))
        # ---------------- text ----------------
        ops.wgrad(self.t_pool, d_t, P.g("text_projection"))
        ops.gemm(d_t, P.w16("text_projection"), out=self.d_pool_t)
        self.txt.dx.zero_()
        ops.layernorm_bwd(self.d_pool_t, self.txt.x, self.tp_m, self.tp_r, P.w("ln_final.weight"),
                          self.txt.dx, P.g("ln_final.weight"), P.g("ln_final.bias"),
                          row_index=self.eot_rows, accumulate=False)
        self.txt.backward()
        ops.text_embed_bwd(self.txt.dx, self.ids, P.g("token_embedding.weight"), P.g("positional_embedding"))

    # ---- CUDA-graph front-ends over the static I/O buffers ----------------------------------
    def _body_embed(self):
        t, v = self.forward(self.in_ids, self.in_img, save=False)
        ops.score_fuse_fwd(t, v, self.in_mt, self.in_mv, out=self.out_emb)

    def _body_fwdbwd(self):
        self.forward(self.in_ids, self.in_img, save=True)
        ops.score_fuse_bwd(self.in_demb, self.in_mt, self.in_mv, out=(self.d_t, self.d_v))
        self.backward(self.d_t, self.d_v)

    def _run(self, key, body, use_graph):
        """First call runs eagerly (sets kernel attributes, warms allocators), second call
        captures, later calls replay: ~2000 launches per micro-batch become one graph launch."""
        if not use_graph:
            body()
            return
        st = self._graphs.get(key, 0)
        if st == 0:
            body()
            self._graphs[key] = 1
        elif st == 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = (g, L.launch_count() - n0)   # kernels captured in this graph
            # NB: capture does not execute -- replay once so this call has its effect
            g.replay()
            ops.add_graph_launches(self._graphs[key][1])
        else:
            st[0].replay()
            ops.add_graph_launches(st[1])

    def set_inputs(self, ids, img, mt, mv, d_emb=None):
        n = ids.shape[0]
        self.in_ids[:n].copy_(ids, non_blocking=True)
        self.in_img[:n].copy_(img, non_blocking=True)
        self.in_mt[:n].copy_(mt, non_blocking=True)
        self.in_mv[:n].copy_(mv, non_blocking=True)
        if n < self.mb:
            self.in_ids[n:].zero_(); self.in_img[n:].zero_()
        if d_emb is not None:
            self.in_demb.copy_(d_emb, non_blocking=True)

    def embed(self, use_graph=True):
        self._run("embed", self._body_embed, use_graph)
        return self.out_emb

    def fwdbwd(self, use_graph=True):
        self._run("fwdbwd", self._body_fwdbwd, use_graph)
