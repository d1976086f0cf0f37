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

    CLIP_NAMES = dict(ln1="ln_1", qkv_w="attn.in_proj_weight", qkv_b="attn.in_proj_bias", out="attn.out_proj",
                      ln2="ln_2", fc1="mlp.c_fc", fc2="mlp.c_proj", blocks="resblocks")
    TIMM_NAMES = dict(ln1="norm1", qkv_w="attn.qkv.weight", qkv_b="attn.qkv.bias", out="attn.proj",
                      ln2="norm2", fc1="mlp.fc1", fc2="mlp.fc2", blocks="blocks")

    def __init__(self, params: FlatParamModule, prefix, width, layers, heads, S, causal, mb, dev,
                 train, names=None, act=L.ACT_QUICKGELU, eps=LN_EPS):
        self.P, self.prefix = params, prefix
        self.nm = names or self.CLIP_NAMES
        self.act, self.eps = act, eps
        self.D, self.nl, self.H, self.S, self.causal, self.mb = width, layers, heads, S, causal, mb
        self.T = mb * S
        self.dev = dev
        T, D = self.T, width
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.x = torch.empty(T, D, **f32)          # residual stream
        self.dp = None                             # DropPath (set_drop_path); dp_on: active for the current run
        self.dp_on = False
        if train:
            self.acts = [_LayerActs(T, D, mb, heads, S, dev) for _ in range(layers)]
            self.x_out = torch.empty(T, D, **f32)      # tower output when the stream runs through acts
            self.dx = torch.empty(T, D, **f32)
            self.dxb = torch.empty(T, D, **bf)
            self.d_big = torch.empty(T, 4 * D, **bf)   # d_pre
            self.d_h = torch.empty(T, D, **bf)         # d_h1 / d_h2 / d_att
            self.dqkv = torch.empty(T, 3 * D, **bf)
            self.delta = torch.empty(ops.attn_bwd_ws_floats(mb, S, S, heads), **f32)
            # the no-save forward (gradient-cache phase A) borrows block 0's buffers as scratch
            a0 = self.acts[0]
            self.h, self.qkv, self.att, self.big = a0.h1, a0.qkv, a0.att, a0.act
        else:
            self.acts = None
            self.h = torch.empty(T, D, **bf)
            self.qkv = torch.empty(T, 3 * D, **bf)
            self.att = torch.empty(T, D, **bf)
            self.big = torch.empty(T, 4 * D, **bf)

    def set_drop_path(self, rate, ctx, site_base=4096):
        """timm DropPath on both residual branches (vit.py:136,151-152): drop probability linspace(0, rate, depth)[layer]
        (vit.py:217), one draw per sample and branch, kept branches scaled by 1 / (1 - p).  ctx: the encoder's
        ops.drop_ctx buffer.  Active only while `dp_on` is set (training forward / re-forward / backward)."""
        if rate <= 0:
            self.dp = None
            return
        p_layer = torch.linspace(0, rate, self.nl)
        p_sites = p_layer.repeat_interleave(2).contiguous()              # site 2l = attention branch, 2l + 1 = MLP branch
        self.dp = dict(ctx=ctx, site_base=site_base, p=p_sites.tolist(), p_dev=p_sites.to(self.dev),
                       scales=torch.ones(2 * self.nl, self.mb, device=self.dev, dtype=torch.float32))

    def _branch_scale(self, site):
        """Per-item scale vector of residual branch `site`, or None where the branch is never dropped."""
        if not self.dp_on or self.dp["p"][site] == 0.0:
            return None
        return self.dp["scales"][site]

    def _n(self, i, leaf):
        """leaf is written in upstream-CLIP terms; translated for timm-style (BLIP) towers."""
        nm = self.nm
        for a, b in (("ln_1", nm["ln1"]), ("ln_2", nm["ln2"]), ("attn.in_proj_weight", nm["qkv_w"]),
                     ("attn.in_proj_bias", nm["qkv_b"]), ("attn.out_proj", nm["out"]), ("mlp.c_fc", nm["fc1"]),
                     ("mlp.c_proj", nm["fc2"])):
            if leaf.startswith(a):
                leaf = b + leaf[len(a):]
                break
        return f"{self.prefix}.{nm['blocks']}.{i}.{leaf}"

    # ---- forward over all blocks.  Input: self.x.  Output: self.x ------------------------------
    # Inference / gradient-cache phase A: the residual stream is updated in place.  Training
    # (save=True): it flows through the saved per-block buffers (x_in -> x_mid -> next x_in), so
    # saving the block inputs costs no extra copies.
    def forward(self, save):
        P, B, S, H = self.P, self.mb, self.S, self.H
        w16, w = self.P.w16, self.P.w
        if self.dp_on:
            ops.droppath_scales(self.dp["ctx"], self.dp["site_base"], self.dp["p_dev"], B, out=self.dp["scales"])
        for i in range(self.nl):
            n = lambda leaf: self._n(i, leaf)
            sa, sm = (self._branch_scale(2 * i), self._branch_scale(2 * i + 1)) if self.dp_on else (None, None)
            if save:
                a = self.acts[i]
                x_in = self.x if i == 0 else a.x_in
                if i == 0:
                    a.x_in = self.x          # block 0 reads the tower input in place
                x_mid = a.x_mid
                x_out = self.acts[i + 1].x_in if i + 1 < self.nl else self.x_out
                h1, qkv, att, h2, pre, act, lse = a.h1, a.qkv, a.att, a.h2, a.pre, a.act, a.lse
                st1, st2 = (a.m1, a.r1), (a.m2, a.r2)
            else:
                x_in = x_mid = x_out = self.x
                h1 = h2 = self.h
                qkv, att, pre, act, lse = self.qkv, self.att, None, self.big, None
                st1 = st2 = None
            ops.layernorm_fwd(x_in, w(n("ln_1.weight")), w(n("ln_1.bias")), self.eps, out=h1, stats=st1)
            ops.gemm(h1, w16(n("attn.in_proj_weight")), bias=w(n("attn.in_proj_bias")), out=qkv)
            ops.attn_fwd(qkv, B, S, H, self.causal, out=att, lse=lse, need_lse=save)
            ops.gemm(att, w16(n("attn.out_proj.weight")), bias=w(n("attn.out_proj.bias")), residual=x_in, out=x_mid,
                     row_scale=sa, row_scale_rows=S)
            ops.layernorm_fwd(x_mid, w(n("ln_2.weight")), w(n("ln_2.bias")), self.eps, out=h2, stats=st2)
            ops.gemm(h2, w16(n("mlp.c_fc.weight")), bias=w(n("mlp.c_fc.bias")), act=self.act, aux_out=pre, out=act)
            ops.gemm(act, w16(n("mlp.c_proj.weight")), bias=w(n("mlp.c_proj.bias")), residual=x_mid, out=x_out,
                     row_scale=sm, row_scale_rows=S)
        return self.x_out if save else self.x

    def top_bias_grad(self):
        """Gradient buffer of the last block's c_proj bias: pass it as `dcolsum` to the LayerNorm backward
        that produces this tower's incoming dx / dxb and call backward(top_bias_done=True)."""
        return self.P.g(self._n(self.nl - 1, "mlp.c_proj.bias"))

    # ---- backward: self.dx holds dL/dx_out on entry, dL/dx_in on exit; self.dxb = bf16(self.dx).
    # The bias gradients of out_proj / c_proj are the column sums of the residual-stream gradient;
    # the LayerNorm backward that produces that gradient emits them (dcolsum), so only the wide
    # d_pre / d_qkv matrices need a separate column-sum pass.
    def backward(self, top_bias_done=False):
        P, dx, dxb, B, S, H = self.P, self.dx, self.dxb, self.mb, self.S, self.H
        dp = self.dp_on
        assert not (dp and top_bias_done), "under DropPath the bias gradients are summed after the per-item scaling"
        for i in reversed(range(self.nl)):
            a = self.acts[i]
            g = lambda leaf: P.g(self._n(i, leaf))
            w16 = lambda leaf: P.w16(self._n(i, leaf))
            # ---- MLP branch: x_out = x_mid + c_proj(quickgelu(c_fc(ln_2(x_mid))))
            if dp:      # branch gradient = per-item scale o residual-stream gradient; dx itself passes unscaled
                sm = self._branch_scale(2 * i + 1)
                if sm is not None:
                    ops.row_scale(dxb, sm, S)
                ops.colsum_accum(dxb, g("mlp.c_proj.bias"))
            ops.wgrad(dxb, a.act, g("mlp.c_proj.weight"))
            if i == self.nl - 1 and not top_bias_done and not dp:
                ops.colsum_accum(dxb, g("mlp.c_proj.bias"))
            ops.gemm(dxb, w16("mlp.c_proj.weight"), b_mn=True, act=self.act, act_grad_src=a.pre,
                     out=self.d_big)
            ops.wgrad(self.d_big, a.h2, g("mlp.c_fc.weight"))
            ops.colsum_accum(self.d_big, g("mlp.c_fc.bias"))
            ops.gemm(self.d_big, w16("mlp.c_fc.weight"), b_mn=True, out=self.d_h)
            ops.layernorm_bwd(self.d_h, a.x_mid, a.m2, a.r2, P.w(self._n(i, "ln_2.weight")), dx,
                              g("ln_2.weight"), g("ln_2.bias"), dx_bf16=dxb, dcolsum=(None if dp else g("attn.out_proj.bias")))
            # ---- attention branch: x_mid = x_in + out_proj(attn(in_proj(ln_1(x_in))))
            if dp:
                sa = self._branch_scale(2 * i)
                if sa is not None:
                    ops.row_scale(dxb, sa, S)
                ops.colsum_accum(dxb, g("attn.out_proj.bias"))
            ops.wgrad(dxb, a.att, g("attn.out_proj.weight"))
            ops.gemm(dxb, w16("attn.out_proj.weight"), b_mn=True, out=self.d_h)
            ops.attn_bwd(a.qkv, a.att, self.d_h, a.lse, B, S, H, self.causal, dqkv=self.dqkv, delta=self.delta)
            ops.wgrad(self.dqkv, a.h1, g("attn.in_proj_weight"))
            ops.colsum_accum(self.dqkv, g("attn.in_proj_bias"))
            ops.gemm(self.dqkv, w16("attn.in_proj_weight"), b_mn=True, out=self.d_h)
            ops.layernorm_bwd(self.d_h, a.x_in, a.m1, a.r1, P.w(self._n(i, "ln_1.weight")), dx,
                              g("ln_1.weight"), g("ln_1.bias"), dx_bf16=dxb,
                              dcolsum=(P.g(self._n(i - 1, "mlp.c_proj.bias")) if i > 0 and not dp else None))


class CLIPEncoder:
    """Both CLIP towers + pooling/projection for a fixed micro-batch size."""

    def __init__(self, params: FlatParamModule, dims: CLIPDims, mb: int, dev, train: bool):
        self.P, self.d, self.mb, self.dev, self.train = params, dims, mb, dev, train
        d = dims
        on_gpu = torch.device(dev).type == "cuda"       # (CPU only in the host-logic tests, which swap the kernels out)
        mem0 = torch.cuda.memory_allocated(dev) if on_gpu else 0
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
        self._side = None
        self.bytes_allocated = (torch.cuda.memory_allocated(dev) - mem0) if on_gpu else 0   # what one more encoder like this costs

    def refresh_weights(self):
        d = self.d
        ops.pad_cast_bf16(self.P.w("visual.conv1.weight").view(d.v_width, d.patch_k), d.patch_k_pad,
                          out=self.conv_w16)

    # ---- the two towers are independent until score fusion: the text tower runs on a side
    # stream (fork/join with events; captured as parallel branches of the CUDA graph) so its
    # small GEMMs fill the SMs the vision kernels leave idle ------------------------------------
    serial_towers = False   # diagnostics (bench roofline pass): run the text tower on the main stream too, so that a
    #                         kernel timed with events in its place in the step is not sharing the SMs with the other tower

    def _fork(self):
        main = torch.cuda.current_stream()
        if self.serial_towers:
            self._side = main
            return main
        if self._side is None or self._side is main:
            self._side = torch.cuda.Stream(device=self.dev)
        self._side.wait_stream(main)
        return main

    def _join(self, main):
        if self._side is not main:
            main.wait_stream(self._side)

    def _vision_fwd(self, img, save):
        P, d, mb = self.P, self.d, self.mb
        ops.patch_im2col(img, d.patch, d.patch_k_pad, out=self.col)
        ops.gemm(self.col, self.conv_w16, out=self.patch)
        ops.vision_embed_fwd(self.patch, P.w("visual.class_embedding"), P.w("visual.positional_embedding"),
                             mb, d.v_tokens, d.v_width, out=self.x0)
        ops.layernorm_fwd(self.x0, P.w("visual.ln_pre.weight"), P.w("visual.ln_pre.bias"), LN_EPS,
                          out=self.vis.x, stats=((self.pre_m, self.pre_r) if save else None))
        xo = self.vis.forward(save)
        ops.layernorm_fwd(xo, P.w("visual.ln_post.weight"), P.w("visual.ln_post.bias"), LN_EPS,
                          row_index=self.cls_rows, out=self.v_pool,
                          stats=((self.vp_m, self.vp_r) if save else None))
        ops.gemm(self.v_pool, P.w16("visual.proj"), b_mn=True, out=self.v_emb)

    def _text_fwd(self, ids, save):
        P, d, mb = self.P, self.d, self.mb
        ops.text_embed_fwd(ids, P.w("token_embedding.weight"), P.w("positional_embedding"),
                           out=self.txt.x.view(mb, d.context, d.t_width), eot=self.eot_rows)
        self.ids = ids
        xo = self.txt.forward(save)
        ops.layernorm_fwd(xo, P.w("ln_final.weight"), P.w("ln_final.bias"), LN_EPS,
                          row_index=self.eot_rows, out=self.t_pool,
                          stats=((self.tp_m, self.tp_r) if save else None))
        ops.gemm(self.t_pool, P.w16("text_projection"), b_mn=True, out=self.t_emb)

    # ---- forward: returns (t_emb, v_emb) fp32 [mb, E] (buffers reused per call) -----
    def forward(self, ids, img, save):
        d, mb = self.d, self.mb
        assert ids.shape == (mb, d.context) and img.shape == (mb, 3, d.image_size, d.image_size)
        main = self._fork()
        with torch.cuda.stream(self._side):
            self._text_fwd(ids, save)
        self._vision_fwd(img, save)
        self._join(main)
        return self.t_emb, self.v_emb

    def _vision_bwd(self, d_v):
        P, d, mb = self.P, self.d, self.mb
        ops.wgrad(self.v_pool, d_v, P.g("visual.proj"))                 # [D,E] += pool^T d_v
        ops.gemm(d_v, P.w16("visual.proj"), out=self.d_pool_v)          # d_pool = d_v @ proj^T
        self.vis.dx.zero_()
        self.vis.dxb.zero_()
        ops.layernorm_bwd(self.d_pool_v, self.vis.x_out, self.vp_m, self.vp_r, P.w("visual.ln_post.weight"),
                          self.vis.dx, P.g("visual.ln_post.weight"), P.g("visual.ln_post.bias"),
                          row_index=self.cls_rows, accumulate=False, dx_bf16=self.vis.dxb,
                          dcolsum=self.vis.top_bias_grad())
        self.vis.backward(top_bias_done=True)
        ops.layernorm_bwd(self.vis.dx, self.x0, self.pre_m, self.pre_r, P.w("visual.ln_pre.weight"),
                          self.dx0, P.g("visual.ln_pre.weight"), P.g("visual.ln_pre.bias"), accumulate=False)
        ops.vision_embed_bwd(self.dx0, mb, d.v_tokens, d.v_width, self.d_patch, P.g("visual.class_embedding"),
                             P.g("visual.positional_embedding"))
        self.d_conv_pad.zero_()
        ops.wgrad(self.d_patch, self.col, self.d_conv_pad)
        ops.unpad_accum(self.d_conv_pad, P.g("visual.conv1.weight").view(d.v_width, d.patch_k))

    def _text_bwd(self, d_t):
        P = self.P
        ops.wgrad(self.t_pool, d_t, P.g("text_projection"))
        ops.gemm(d_t, P.w16("text_projection"), out=self.d_pool_t)
        self.txt.dx.zero_()
        self.txt.dxb.zero_()
        ops.layernorm_bwd(self.d_pool_t, self.txt.x_out, self.tp_m, self.tp_r, P.w("ln_final.weight"),
                          self.txt.dx, P.g("ln_final.weight"), P.g("ln_final.bias"),
                          row_index=self.eot_rows, accumulate=False, dx_bf16=self.txt.dxb,
                          dcolsum=self.txt.top_bias_grad())
        self.txt.backward(top_bias_done=True)
        ops.text_embed_bwd(self.txt.dx, self.ids, P.g("token_embedding.weight"), P.g("positional_embedding"))

    # ---- backward from d t_emb / d v_emb (bf16 [mb, E]); accumulates into P.grad -----
    def backward(self, d_t, d_v):
        main = self._fork()
        with torch.cuda.stream(self._side):
            self._text_bwd(d_t)
        self._vision_bwd(d_v)
        self._join(main)

    # ---- CUDA-graph front-ends over the static I/O buffers ----------------------------------
    def _body_embed(self):
        t, v = self.forward(self.in_ids, self.in_img, save=False)
        ops.score_fuse_fwd(t, v, self.in_mt, self.in_mv, out=self.out_emb)

    def _body_fwdbwd(self):
        self.forward(self.in_ids, self.in_img, save=True)
        ops.score_fuse_bwd(self.in_demb, self.in_mt, self.in_mv, out=(self.d_t, self.d_v))
        self.backward(self.d_t, self.d_v)

    # ---- retained micro-batches: an encoder that is not reused between the embedding pass and the
    # backward pass keeps its activations, so the gradient-cache re-forward is skipped for it ----
    def _body_fwd_save(self):
        t, v = self.forward(self.in_ids, self.in_img, save=True)
        ops.score_fuse_fwd(t, v, self.in_mt, self.in_mv, out=self.out_emb)

    def _body_bwd_only(self):
        ops.score_fuse_bwd(self.in_demb, self.in_mt, self.in_mv, out=(self.d_t, self.d_v))
        self.backward(self.d_t, self.d_v)

    # ---- single-tower passes over COMPACTED rows.  The reference encodes "" / a black
    # image for a missing modality and multiplies the result by its 0 mask (clip_sf.py:53-63); here
    # the caller hands each tower only the rows whose mask is non-zero. ---------------------------
    def _body_embed_text(self):
        self._text_fwd(self.in_ids, save=False)

    def _body_embed_vision(self):
        self._vision_fwd(self.in_img, save=False)

    def _body_fwdbwd_text(self):
        self._text_fwd(self.in_ids, save=True)
        self._text_bwd(self.d_t)

    def _body_fwdbwd_vision(self):
        self._vision_fwd(self.in_img, save=True)
        self._vision_bwd(self.d_v)

    def set_tower_input(self, tower, src, rows):
        """Gather batch rows `rows` (int32, <= mb of them) of the token ids / images into the static input
        buffer of one tower; the unused tail is zeroed (finite activations, zero gradient)."""
        buf = self.in_ids if tower == "text" else self.in_img
        n = rows.numel()
        torch.index_select(src, 0, rows, out=buf[:n])
        if n < self.mb:
            buf[n:].zero_()

    def run_tower(self, tower, train, use_graph=True):
        """train=False: forward -> fp32 [mb, E] embedding of that tower.  train=True: forward with saved
        activations + backward from self.d_t / self.d_v (bf16 [mb, E], set by the caller)."""
        body = {("text", False): self._body_embed_text, ("vision", False): self._body_embed_vision,
                ("text", True): self._body_fwdbwd_text, ("vision", True): self._body_fwdbwd_vision}[(tower, train)]
        self._run(f"{'fwdbwd' if train else 'embed'}_{tower}", body, use_graph)
        return self.t_emb if tower == "text" else self.v_emb

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

    def fwd_save(self, use_graph=True):
        """Forward that keeps the activations (this encoder must not run anything else until bwd_only)."""
        self._run("fwd_save", self._body_fwd_save, use_graph)
        return self.out_emb

    def bwd_only(self, d_emb, use_graph=True):
        """Backward from the activations kept by the last fwd_save()."""
        self.in_demb.copy_(d_emb, non_blocking=True)
        self._run("bwd_only", self._body_bwd_only, use_graph)
