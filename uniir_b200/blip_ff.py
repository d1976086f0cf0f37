"""Drop-in for the reference ``BLIPFeatureFusion`` (src/models/uniir_blip/blip_featurefusion/
blip_ff.py:11-332) on the B200-native kernels, and the shared machinery of
``BLIPScoreFusion`` (blip_sf.py; ``BLIPEncoder(mode="sf")``).

Built: constructor / state-dict layout (online + momentum encoders, queues, temp: reference checkpoints load
with a plain ``load_state_dict``), ``encode_multimodal_input``, ``forward(batch, encode_mbeir_batch=True)`` /
``encode_mbeir_batch``, and the momentum-distillation training step (``compute_contrastive_loss`` +
``backward_contrastive``, driven by ``engine.train_step_blip`` / ``blip_engine.train_one_epoch``, or through
autograd via ``autograd_bridge``): EMA, queue, soft-label cross-entropy, gradient caching, and the hard-negative variant
of the loss (blip_ff.py:126-131,160-171,208-218,233-246).

Train mode (``model.train()``, the reference's engine.py:16): BERT ``hidden_dropout_prob`` / ``attention_probs_dropout_prob``
(0.1 in med_config.json; med.py:97,217,244,325) and timm DropPath 0.1 in the large ViT (vit.py:151-152, blip.py:253), in the
online AND the momentum encoders, with counter-based masks (include/uniir_b200.h): a pure function of (step seed, site,
element index within the batch), so nothing is stored, the gradient-cache re-forward and the backward regenerate them
under any micro-batch tiling, and CUDA graphs replay with a new seed.  The oracle's dropout sites are pinned against the
reference's modules in train mode; product == oracle given the same masks (tests/test_dropout_gpu.py).  ``model.eval()``
(or ``set_dropout(0, 0, 0)``) gives the deterministic encoder.  The mask BITS differ from torch's Philox stream, as they
would between two torch seeds.

Arithmetic (vendored reference code, restated in oracle/blip_oracle.py and pinned against the shimmed
reference modules): ViT-L/16 with biased patch-embed, pre-LN blocks (eps 1e-6), exact GELU, final norm, all
197 tokens kept (backbone/vit.py:254-268); BERT "MED" with word+position embeddings -> LN(1e-12), 12 post-LN
layers of self-attention (key-padding mask), cross-attention into the image tokens (K/V = Linear(1024->768)),
GELU FFN, then pooler tanh(W h[:,0] + b) (backbone/med.py:52-98,101-232,330-396,498-511).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist
from torch import nn

from . import lib as L
from . import ops
from .clip_towers import Tower
from .params import FlatParamModule


@dataclass(frozen=True)
class BLIPDims:
    image_size: int = 224
    patch: int = 16
    v_width: int = 1024
    v_layers: int = 24
    v_heads: int = 16
    t_width: int = 768
    t_layers: int = 12
    t_heads: int = 12
    t_inter: int = 3072
    vocab: int = 30524
    max_pos: int = 512

    @property
    def grid(self):
        return self.image_size // self.patch

    @property
    def v_tokens(self):
        return self.grid ** 2 + 1


BLIP_CONFIGS = {
    "large": BLIPDims(),
    "base": BLIPDims(v_width=768, v_layers=12, v_heads=12),
    "tiny-test": BLIPDims(image_size=64, patch=16, v_width=128, v_layers=2, v_heads=2, t_width=64, t_layers=2,
                          t_heads=1, t_inter=256, vocab=600, max_pos=128),
}


def blip_param_specs(d: BLIPDims, suffix="", mode="ff", embed_dim=None):
    """[(name, shape, decay, std)].  q/k/v (and cross k/v) weights, then their biases, are adjacent so the
    fused [3D, D] / [2D, Dv] projection matrices are plain views of the flat buffer.  `decay` is True for
    every trained parameter: the reference uses one AdamW group with wd 0.05 (uniir_blip/train.py:193-197).
    mode "sf" (BLIPScoreFusion, blip_sf.py:36-50): no pooler, vision_proj / text_proj heads, and the
    cross-attention parameters -- present in the checkpoint, frozen (:75-78) and never used in mode="text" --
    go to the second flat segment, which gets no weight decay and never sees a gradient, so the fused AdamW
    leaves them exactly as loaded (torch's AdamW skips parameters without a gradient)."""
    sp = []
    add = lambda n, shape, std: sp.append((n, tuple(shape), not (mode == "sf" and ".crossattention." in n), std))
    v, t = f"visual_encoder{suffix}", f"text_encoder{suffix}"
    D = d.v_width
    add(f"{v}.cls_token", (1, 1, D), 0.02)
    add(f"{v}.pos_embed", (1, d.v_tokens, D), 0.02)
    add(f"{v}.patch_embed.proj.weight", (D, 3, d.patch, d.patch), 0.02)
    add(f"{v}.patch_embed.proj.bias", (D,), 0.02)
    for i in range(d.v_layers):
        b = f"{v}.blocks.{i}"
        add(f"{b}.norm1.weight", (D,), -1.0); add(f"{b}.norm1.bias", (D,), 0.0)
        add(f"{b}.attn.qkv.weight", (3 * D, D), 0.02); add(f"{b}.attn.qkv.bias", (3 * D,), 0.02)
        add(f"{b}.attn.proj.weight", (D, D), 0.02); add(f"{b}.attn.proj.bias", (D,), 0.02)
        add(f"{b}.norm2.weight", (D,), -1.0); add(f"{b}.norm2.bias", (D,), 0.0)
        add(f"{b}.mlp.fc1.weight", (4 * D, D), 0.02); add(f"{b}.mlp.fc1.bias", (4 * D,), 0.02)
        add(f"{b}.mlp.fc2.weight", (D, 4 * D), 0.02); add(f"{b}.mlp.fc2.bias", (D,), 0.02)
    add(f"{v}.norm.weight", (D,), -1.0); add(f"{v}.norm.bias", (D,), 0.0)
    H = d.t_width
    add(f"{t}.embeddings.word_embeddings.weight", (d.vocab, H), 0.02)
    add(f"{t}.embeddings.position_embeddings.weight", (d.max_pos, H), 0.02)
    add(f"{t}.embeddings.LayerNorm.weight", (H,), -1.0); add(f"{t}.embeddings.LayerNorm.bias", (H,), 0.0)
    for i in range(d.t_layers):
        l = f"{t}.encoder.layer.{i}"
        for nm in ("query", "key", "value"):
            add(f"{l}.attention.self.{nm}.weight", (H, H), 0.02)
        for nm in ("query", "key", "value"):
            add(f"{l}.attention.self.{nm}.bias", (H,), 0.02)
        add(f"{l}.attention.output.dense.weight", (H, H), 0.02); add(f"{l}.attention.output.dense.bias", (H,), 0.02)
        add(f"{l}.attention.output.LayerNorm.weight", (H,), -1.0); add(f"{l}.attention.output.LayerNorm.bias", (H,), 0.0)
        add(f"{l}.crossattention.self.query.weight", (H, H), 0.02); add(f"{l}.crossattention.self.query.bias", (H,), 0.02)
        add(f"{l}.crossattention.self.key.weight", (H, D), 0.02); add(f"{l}.crossattention.self.value.weight", (H, D), 0.02)
        add(f"{l}.crossattention.self.key.bias", (H,), 0.02); add(f"{l}.crossattention.self.value.bias", (H,), 0.02)
        add(f"{l}.crossattention.output.dense.weight", (H, H), 0.02); add(f"{l}.crossattention.output.dense.bias", (H,), 0.02)
        add(f"{l}.crossattention.output.LayerNorm.weight", (H,), -1.0); add(f"{l}.crossattention.output.LayerNorm.bias", (H,), 0.0)
        add(f"{l}.intermediate.dense.weight", (d.t_inter, H), 0.02); add(f"{l}.intermediate.dense.bias", (d.t_inter,), 0.02)
        add(f"{l}.output.dense.weight", (H, d.t_inter), 0.02); add(f"{l}.output.dense.bias", (H,), 0.02)
        add(f"{l}.output.LayerNorm.weight", (H,), -1.0); add(f"{l}.output.LayerNorm.bias", (H,), 0.0)
    if mode == "ff":
        add(f"{t}.pooler.dense.weight", (H, H), 0.02); add(f"{t}.pooler.dense.bias", (H,), 0.02)
    else:
        add(f"vision_proj{suffix}.weight", (embed_dim, D), 0.02); add(f"vision_proj{suffix}.bias", (embed_dim,), 0.02)
        add(f"text_proj{suffix}.weight", (embed_dim, H), 0.02); add(f"text_proj{suffix}.bias", (embed_dim,), 0.02)
    if not suffix:
        add("temp", (), -3.0)   # LAST, so the momentum set's flat buffer is an exact prefix (EMA = one kernel)
    return sp


class _BertActs:
    """Saved activations of one BERT layer for one micro-batch."""

    def __init__(self, T, H, I, Tv, B, nh, S, Sv, dev, cross=True):
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.hb_in = torch.empty(T, H, **bf)
        self.qkv = torch.empty(T, 3 * H, **bf); self.ctx1 = torch.empty(T, H, **bf); self.lse1 = torch.empty(B, nh, S, **f32)
        self.t1 = torch.empty(T, H, **f32); self.m1 = torch.empty(T, **f32); self.r1 = torch.empty(T, **f32)
        self.h1b = torch.empty(T, H, **bf)
        if cross:
            self.q2 = torch.empty(T, H, **bf); self.kv2 = torch.empty(Tv, 2 * H, **bf); self.ctx2 = torch.empty(T, H, **bf)
            self.lse2 = torch.empty(B, nh, S, **f32)
            self.t2 = torch.empty(T, H, **f32); self.m2 = torch.empty(T, **f32); self.r2 = torch.empty(T, **f32)
            self.h2b = torch.empty(T, H, **bf)
        else:
            self.h2b = self.h1b          # mode="text": the FFN reads the self-attention block's output
        self.pre = torch.empty(T, I, **bf); self.act = torch.empty(T, I, **bf)
        self.t3 = torch.empty(T, H, **f32); self.m3 = torch.empty(T, **f32); self.r3 = torch.empty(T, **f32)


class BLIPEncoder:
    """ViT + cross-attention BERT for a fixed (micro-batch, text length): forward, and (train=True) forward
    with saved activations + backward."""

    def __init__(self, P: FlatParamModule, d: BLIPDims, mb, S, dev, suffix="", train=False, mode="ff", embed_dim=None):
        """mode "ff": feature fusion (cross-attention BERT over all image tokens, pooler).  mode "sf": score fusion
        (blip_sf.py:97-172): BERT in mode="text", [CLS] -> text_proj; ViT class token -> vision_proj; masked sum."""
        self.P, self.d, self.mb, self.S, self.dev, self.train = P, d, mb, S, dev, train
        self.mode, self.sfx = mode, suffix
        self.v, self.t = f"visual_encoder{suffix}", f"text_encoder{suffix}"
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.vis = Tower(P, self.v, d.v_width, d.v_layers, d.v_heads, d.v_tokens, False, mb, dev, train,
                         names=Tower.TIMM_NAMES, act=L.ACT_GELU, eps=1e-6)
        Kp = 3 * d.patch * d.patch
        assert Kp % 8 == 0
        Tv = mb * d.v_tokens
        self.col = torch.empty(mb * d.grid ** 2, Kp, **bf)
        self.patch = torch.empty(mb * d.grid ** 2, d.v_width, **bf)
        self.img_tok = torch.empty(Tv, d.v_width, **bf)
        T, H = mb * S, d.t_width
        self.dctx = None         # dropout (set_dropout): ops.drop_ctx buffer {seed, first item}; _drop: on for this run
        self.p_h = self.p_a = 0.0
        self._drop = False
        self.h = torch.empty(T, H, **f32)
        self.tmp = torch.empty(T, H, **f32)
        self.hb = torch.empty(T, H, **bf)
        self.qkv = torch.empty(T, 3 * H, **bf)
        self.ctx = torch.empty(T, H, **bf)
        self.q = torch.empty(T, H, **bf)
        self.kv = torch.empty(Tv, 2 * H, **bf)
        self.inter = torch.empty(T, d.t_inter, **bf)
        self.pool = torch.empty(mb, H, **bf)
        self.emb = torch.empty(mb, H if mode == "ff" else embed_dim, **f32)
        if mode == "sf":
            self.cls_rows = (torch.arange(mb, device=dev, dtype=torch.int32) * d.v_tokens).contiguous()
            self.v_pool = torch.empty(mb, d.v_width, **bf)
            self.t_emb = torch.empty(mb, embed_dim, **f32)
            self.v_emb = torch.empty(mb, embed_dim, **f32)
            self.ones = torch.ones(mb, device=dev, dtype=torch.int64)
            self.mt = self.mv = self.ones
            if train:
                self.d_t = torch.empty(mb, embed_dim, **bf); self.d_v = torch.empty(mb, embed_dim, **bf)
                self.d_vpool = torch.empty(mb, d.v_width, **bf)
        if train:
            self.acts = [_BertActs(T, H, d.t_inter, Tv, mb, d.t_heads, S, d.v_tokens, dev, cross=(mode == "ff"))
                         for _ in range(d.t_layers)]
            self.emb_pre = torch.empty(T, H, **f32); self.me = torch.empty(T, **f32); self.re = torch.empty(T, **f32)
            self.vn_m = torch.empty(Tv, **f32); self.vn_r = torch.empty(Tv, **f32)
            self.dh = torch.empty(T, H, **f32)       # gradient of the fp32 hidden stream
            self.dt = torch.empty(T, H, **f32)       # LN-backward output (grad of the pre-LN sum)
            self.dtb = torch.empty(T, H, **bf)
            self.d_ctx = torch.empty(T, H, **bf)
            self.dqkv = torch.empty(T, 3 * H, **bf)
            self.dq2 = torch.empty(T, H, **bf)
            self.dkv2 = torch.empty(Tv, 2 * H, **bf)
            self.d_pre = torch.empty(T, d.t_inter, **bf)
            self.d_img = torch.empty(Tv, d.v_width, **f32)
            self.dz = torch.empty(mb, H, **bf)
            self.d_pool = torch.empty(mb, H, **bf)
            self.d_patch = torch.empty(mb * d.grid ** 2, d.v_width, **bf)
            self.delta = torch.empty(max(ops.attn_bwd_ws_floats(mb, S, S, d.t_heads),
                                         ops.attn_bwd_ws_floats(mb, S, Tv // mb, d.t_heads)), **f32)
            self.att_len = None
            self.ids = None

    # ---- dropout of the reference's train mode.  Sites: 0 = embeddings (med.py:97); 1 + 5 i + {0: self-attention
    # probabilities, 1: self-attention output, 2: cross-attention probabilities, 3: cross-attention output, 4: FFN output}
    # of BERT layer i (med.py:217,244,325); 4096 + 2 l + {0, 1}: DropPath of ViT block l (vit.py:151-152).  The masks are
    # functions of (seed, site, element index in the whole batch): see include/uniir_b200.h. ----
    def set_dropout(self, p_hidden, p_attn, drop_path_rate):
        self.p_h, self.p_a = float(p_hidden), float(p_attn)
        self.dctx = ops.drop_ctx(self.dev)
        self.vis.set_drop_path(float(drop_path_rate), self.dctx)
        if getattr(self, "_graphs", None):      # captured graphs carry the old probabilities as kernel arguments
            self._graphs = {k: v for k, v in self._graphs.items() if not k.endswith("+drop")}

    def _site(self, layer, k):
        return 1 + 5 * layer + k

    # ---- CUDA-graph front-end (BLIPFeatureFusion.use_cuda_graphs, default on).  The eager step issues 120-150 k kernel
    # launches per 2048-pair batch; captured, a micro-batch is one graph launch.  Static I/O buffers, first call eager
    # (kernel attributes, allocator warm-up), second call captures, later calls replay -- as CLIPEncoder._run. ----
    def _static_io(self):
        if getattr(self, "in_ids", None) is None:
            dev, d, mb, S = self.dev, self.d, self.mb, self.S
            self.in_ids = torch.zeros(mb, S, device=dev, dtype=torch.int32)
            self.in_len = torch.ones(mb, device=dev, dtype=torch.int32)
            self.in_img = torch.zeros(mb, 3, d.image_size, d.image_size, device=dev, dtype=torch.float32)
            self.in_mt = torch.ones(mb, device=dev, dtype=torch.int64)
            self.in_mv = torch.ones(mb, device=dev, dtype=torch.int64)
            self.in_demb = torch.zeros_like(self.emb)
            self._graphs = {}

    def run(self, ids, att_len, img, mt=None, mv=None, d_emb=None, use_graph=False, drop_seed=None, item0=0):
        """forward (d_emb None: no activations kept) or forward-with-activations + backward (d_emb given), eagerly or
        through a captured CUDA graph over the static I/O buffers.  Returns the embedding buffer.
        drop_seed (int, needs set_dropout): run with the train-mode dropout masks of that seed; item0 = index of this
        micro-batch's first item in the batch (the masks are indexed by batch position, not micro-batch position)."""
        self._drop = drop_seed is not None and self.dctx is not None
        self.vis.dp_on = self._drop and self.vis.dp is not None
        if self._drop:
            self.dctx.copy_(torch.tensor([int(drop_seed) & 0x7FFFFFFF, int(item0)], dtype=torch.int32), non_blocking=True)
        if not use_graph:
            out = self.forward(ids, att_len, img, save=d_emb is not None, mt=mt, mv=mv)
            if d_emb is not None:
                self.backward(d_emb)
            return out
        self._static_io()
        self.in_ids.copy_(ids, non_blocking=True); self.in_len.copy_(att_len, non_blocking=True)
        self.in_img.copy_(img, non_blocking=True)
        if mt is not None:
            self.in_mt.copy_(mt, non_blocking=True); self.in_mv.copy_(mv, non_blocking=True)
        sf = self.mode == "sf"
        if d_emb is None:
            key = "fwd+drop" if self._drop else "fwd"
            body = lambda: self.forward(self.in_ids, self.in_len, self.in_img, save=False,
                                        mt=self.in_mt if sf else None, mv=self.in_mv if sf else None)
        else:
            key = "fwdbwd+drop" if self._drop else "fwdbwd"
            self.in_demb.copy_(d_emb, non_blocking=True)

            def body():
                self.forward(self.in_ids, self.in_len, self.in_img, save=True, mt=self.in_mt if sf else None,
                             mv=self.in_mv if sf else None)
                self.backward(self.in_demb)
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
            self._graphs[key] = (g, L.launch_count() - n0)
            g.replay()                         # capture does not execute
            ops.add_graph_launches(self._graphs[key][1])
        else:
            st[0].replay()
            ops.add_graph_launches(st[1])
        return self.emb

    def _fused(self, first, rows, cols, bufs=None):
        """[rows, cols] view starting at parameter `first` (adjacent parameters, see blip_param_specs)."""
        decay, off, n, shape = self.P._index[first]
        return (bufs or self.P.flat16)[decay][off:off + rows * cols].view(rows, cols)

    def _fused1(self, first, n, bufs=None):
        decay, off, _, _ = self.P._index[first]
        return (bufs or self.P.flat)[decay][off:off + n]

    def forward(self, ids, att_len, img, save=False, mt=None, mv=None):
        """ids int32 [mb,S]; att_len int32 [mb] (valid tokens: attention_mask is a prefix mask); img fp32
        [mb,3,H,W]; mt / mv int64 [mb] modality masks (mode "sf" only; default all ones).  Returns the embedding
        fp32 [mb, H] (ff: pooled fused feature) or [mb, embed_dim] (sf: masked score fusion); buffer reused."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        w, w16 = P.w, P.w16
        v, t = self.v, self.t
        ff = self.mode == "ff"
        assert not save or self.train
        # ---------------- ViT ----------------
        ops.patch_im2col(img, d.patch, self.col.shape[1], out=self.col)
        ops.gemm(self.col, w16(f"{v}.patch_embed.proj.weight").view(d.v_width, -1), bias=w(f"{v}.patch_embed.proj.bias"),
                 out=self.patch)
        ops.vision_embed_fwd(self.patch, w(f"{v}.cls_token").view(-1), w(f"{v}.pos_embed").view(d.v_tokens, d.v_width),
                             mb, d.v_tokens, d.v_width, out=self.vis.x.view(mb, d.v_tokens, d.v_width))
        xo = self.vis.forward(save)
        if ff:
            ops.layernorm_fwd(xo, w(f"{v}.norm.weight"), w(f"{v}.norm.bias"), 1e-6, out=self.img_tok,
                              stats=((self.vn_m, self.vn_r) if save else None))
        else:   # only the class token is used (blip_sf.py:136-143): final norm on those rows, then vision_proj
            ops.layernorm_fwd(xo, w(f"{v}.norm.weight"), w(f"{v}.norm.bias"), 1e-6, row_index=self.cls_rows,
                              out=self.v_pool, stats=((self.vn_m[:mb], self.vn_r[:mb]) if save else None))
            ops.gemm(self.v_pool, w16(f"vision_proj{self.sfx}.weight"), bias=w(f"vision_proj{self.sfx}.bias"), out=self.v_emb)
        # ---------------- BERT ----------------
        H, nh = d.t_width, d.t_heads
        ctx, per = self.dctx, S * H
        dh = self._drop and self.p_h > 0           # hidden dropout / attention-probability dropout on for this run
        da = self._drop and self.p_a > 0
        e_pre = self.emb_pre if save else self.tmp
        ops.text_embed_fwd(ids, w(f"{t}.embeddings.word_embeddings.weight"),
                           w(f"{t}.embeddings.position_embeddings.weight")[:S], out=e_pre.view(mb, S, H))
        ops.layernorm_fwd(e_pre, w(f"{t}.embeddings.LayerNorm.weight"), w(f"{t}.embeddings.LayerNorm.bias"), 1e-12,
                          out=self.h, stats=((self.me, self.re) if save else None))
        if dh:
            ops.dropout(self.h, ctx, 0, per, self.p_h)                                        # med.py:97
        if save:
            self.att_len, self.ids = att_len, ids
        for i in range(d.t_layers):
            l = f"{t}.encoder.layer.{i}"
            a = self.acts[i] if save else None
            hb_in = a.hb_in if save else self.hb
            ops.cast_bf16(self.h, out=hb_in)
            qkv, ctx1 = (a.qkv, a.ctx1) if save else (self.qkv, self.ctx)
            # self-attention (q|k|v fused), key-padding mask
            ops.gemm(hb_in, self._fused(f"{l}.attention.self.query.weight", 3 * H, H),
                     bias=self._fused1(f"{l}.attention.self.query.bias", 3 * H), out=qkv)
            ops.attn_fwd_ex(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], mb, S, S, nh, kv_len=att_len, out=ctx1,
                            lse=(a.lse1 if save else None), drop=((ctx, self._site(i, 0), self.p_a) if da else None))
            t1 = a.t1 if save else self.tmp
            ops.gemm(ctx1, w16(f"{l}.attention.output.dense.weight"), bias=w(f"{l}.attention.output.dense.bias"),
                     residual=(None if dh else self.h), out=t1)
            if dh:     # med.py:244: dropout on the dense output, then the residual
                ops.dropout(t1, ctx, self._site(i, 1), per, self.p_h, residual=self.h)
            ops.layernorm_fwd(t1, w(f"{l}.attention.output.LayerNorm.weight"), w(f"{l}.attention.output.LayerNorm.bias"),
                              1e-12, out=self.h, stats=((a.m1, a.r1) if save else None))
            h1b = a.h1b if save else self.hb
            ops.cast_bf16(self.h, out=h1b)
            if ff:
                # cross-attention into the image tokens (encoder mask is all ones, blip_ff.py:107)
                q2, kv2, ctx2 = (a.q2, a.kv2, a.ctx2) if save else (self.q, self.kv, self.ctx)
                ops.gemm(h1b, w16(f"{l}.crossattention.self.query.weight"), bias=w(f"{l}.crossattention.self.query.bias"), out=q2)
                ops.gemm(self.img_tok, self._fused(f"{l}.crossattention.self.key.weight", 2 * H, d.v_width),
                         bias=self._fused1(f"{l}.crossattention.self.key.bias", 2 * H), out=kv2)
                ops.attn_fwd_ex(q2, kv2[:, :H], kv2[:, H:], mb, S, d.v_tokens, nh, out=ctx2, lse=(a.lse2 if save else None),
                                drop=((ctx, self._site(i, 2), self.p_a) if da else None))
                t2 = a.t2 if save else self.tmp
                ops.gemm(ctx2, w16(f"{l}.crossattention.output.dense.weight"), bias=w(f"{l}.crossattention.output.dense.bias"),
                         residual=(None if dh else self.h), out=t2)
                if dh:
                    ops.dropout(t2, ctx, self._site(i, 3), per, self.p_h, residual=self.h)
                ops.layernorm_fwd(t2, w(f"{l}.crossattention.output.LayerNorm.weight"),
                                  w(f"{l}.crossattention.output.LayerNorm.bias"), 1e-12, out=self.h,
                                  stats=((a.m2, a.r2) if save else None))
                h2b = a.h2b if save else self.hb
                ops.cast_bf16(self.h, out=h2b)
            else:
                h2b = h1b        # mode="text" (med.py:368-383): no cross-attention, the FFN follows directly
            # FFN
            pre, act = (a.pre, a.act) if save else (None, self.inter)
            ops.gemm(h2b, w16(f"{l}.intermediate.dense.weight"), bias=w(f"{l}.intermediate.dense.bias"), act=L.ACT_GELU,
                     aux_out=pre, out=act)
            t3 = a.t3 if save else self.tmp
            ops.gemm(act, w16(f"{l}.output.dense.weight"), bias=w(f"{l}.output.dense.bias"), residual=(None if dh else self.h),
                     out=t3)
            if dh:     # med.py:325
                ops.dropout(t3, ctx, self._site(i, 4), per, self.p_h, residual=self.h)
            ops.layernorm_fwd(t3, w(f"{l}.output.LayerNorm.weight"), w(f"{l}.output.LayerNorm.bias"), 1e-12, out=self.h,
                              stats=((a.m3, a.r3) if save else None))
        ops.cast_bf16(self.h, out=self.hb)
        self.pool.copy_(self.hb.view(mb, S, H)[:, 0])
        if ff:
            ops.gemm(self.pool, w16(f"{t}.pooler.dense.weight"), bias=w(f"{t}.pooler.dense.bias"), act=L.ACT_TANH, out=self.emb)
            return self.emb
        ops.gemm(self.pool, w16(f"text_proj{self.sfx}.weight"), bias=w(f"text_proj{self.sfx}.bias"), out=self.t_emb)
        self.mt = self.ones if mt is None else mt
        self.mv = self.ones if mv is None else mv
        ops.score_fuse_fwd(self.t_emb, self.v_emb, self.mt, self.mv, out=self.emb)   # txt * txt_mask + img * img_mask
        return self.emb

    def backward(self, d_emb):
        """d_emb fp32 [mb, H] = dLoss/d(pooled embedding) of the last forward(save=True).  Accumulates
        parameter gradients into P.grad."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        w, w16, g = P.w, P.w16, P.g
        v, t = self.v, self.t
        H, nh, Sv = d.t_width, d.t_heads, d.v_tokens
        G = P.grad
        ff = self.mode == "ff"
        ctx, per = self.dctx, S * H
        dh = self._drop and self.p_h > 0
        da = self._drop and self.p_a > 0

        def dense_out_bwd(t_saved, m, r, ln, dense_bias, site):
            """Backward of LN(dropout(dense(.)) + residual): self.dt = gradient of the pre-LN sum (it is also the residual
            gradient), self.dtb = bf16 gradient of the dense output, bias gradient accumulated."""
            if not dh:
                ops.layernorm_bwd(self.dh, t_saved, m, r, w(f"{ln}.weight"), self.dt, g(f"{ln}.weight"), g(f"{ln}.bias"),
                                  accumulate=False, dx_bf16=self.dtb, dcolsum=g(dense_bias))
                return
            ops.layernorm_bwd(self.dh, t_saved, m, r, w(f"{ln}.weight"), self.dt, g(f"{ln}.weight"), g(f"{ln}.bias"),
                              accumulate=False)
            ops.dropout(self.dt, ctx, site, per, self.p_h, out=self.dtb)
            ops.colsum_accum(self.dtb, g(dense_bias))
        if ff:
            # ---- pooler: emb = tanh(W h[:,0] + b)
            ops.tanh_bwd(d_emb, self.emb, self.dz)
            ops.wgrad(self.dz, self.pool, g(f"{t}.pooler.dense.weight"))
            ops.colsum_accum(self.dz, g(f"{t}.pooler.dense.bias"))
            ops.gemm(self.dz, w16(f"{t}.pooler.dense.weight"), b_mn=True, out=self.d_pool)
            self.d_img.zero_()
        else:
            # ---- emb = mt * text_proj(h[:,0]) + mv * vision_proj(norm(x)[:,0])
            ops.score_fuse_bwd(d_emb, self.mt, self.mv, out=(self.d_t, self.d_v))
            ops.wgrad(self.d_t, self.pool, g(f"text_proj{self.sfx}.weight"))
            ops.colsum_accum(self.d_t, g(f"text_proj{self.sfx}.bias"))
            ops.gemm(self.d_t, w16(f"text_proj{self.sfx}.weight"), b_mn=True, out=self.d_pool)
        self.dh.zero_()
        self.dh.view(mb, S, H)[:, 0] = self.d_pool.float()
        for i in reversed(range(d.t_layers)):
            l = f"{t}.encoder.layer.{i}"
            a = self.acts[i]
            # ---- FFN: h3 = LN(dense2(gelu(dense1(h2))) + h2)
            dense_out_bwd(a.t3, a.m3, a.r3, f"{l}.output.LayerNorm", f"{l}.output.dense.bias", self._site(i, 4))
            ops.wgrad(self.dtb, a.act, g(f"{l}.output.dense.weight"))
            ops.gemm(self.dtb, w16(f"{l}.output.dense.weight"), b_mn=True, act=L.ACT_GELU, act_grad_src=a.pre, out=self.d_pre)
            ops.wgrad(self.d_pre, a.h2b, g(f"{l}.intermediate.dense.weight"))
            ops.colsum_accum(self.d_pre, g(f"{l}.intermediate.dense.bias"))
            ops.gemm(self.d_pre, w16(f"{l}.intermediate.dense.weight"), b_mn=True, residual=self.dt, out=self.dh)
            # ---- cross-attention: h2 = LN(dense(attn(q(h1), kv(img))) + h1)   (mode "sf": absent, dh passes through)
            if ff:
                dense_out_bwd(a.t2, a.m2, a.r2, f"{l}.crossattention.output.LayerNorm", f"{l}.crossattention.output.dense.bias",
                              self._site(i, 3))
                ops.wgrad(self.dtb, a.ctx2, g(f"{l}.crossattention.output.dense.weight"))
                ops.gemm(self.dtb, w16(f"{l}.crossattention.output.dense.weight"), b_mn=True, out=self.d_ctx)
                ops.attn_bwd_ex(a.q2, a.kv2[:, :H], a.kv2[:, H:], a.ctx2, self.d_ctx, a.lse2, mb, S, Sv, nh,
                                self.dq2, self.dkv2[:, :H], self.dkv2[:, H:], delta=self.delta,
                                drop=((ctx, self._site(i, 2), self.p_a) if da else None))
                ops.wgrad(self.dq2, a.h1b, g(f"{l}.crossattention.self.query.weight"))
                ops.colsum_accum(self.dq2, g(f"{l}.crossattention.self.query.bias"))
                ops.gemm(self.dq2, w16(f"{l}.crossattention.self.query.weight"), b_mn=True, residual=self.dt, out=self.dh)
                ops.wgrad(self.dkv2, self.img_tok, self._fused(f"{l}.crossattention.self.key.weight", 2 * H, d.v_width, G))
                ops.colsum_accum(self.dkv2, self._fused1(f"{l}.crossattention.self.key.bias", 2 * H, G))
                ops.gemm(self.dkv2, self._fused(f"{l}.crossattention.self.key.weight", 2 * H, d.v_width), b_mn=True,
                         out=self.d_img, accumulate=True)
            # ---- self-attention: h1 = LN(dense(attn(qkv(h_in))) + h_in)
            dense_out_bwd(a.t1, a.m1, a.r1, f"{l}.attention.output.LayerNorm", f"{l}.attention.output.dense.bias", self._site(i, 1))
            ops.wgrad(self.dtb, a.ctx1, g(f"{l}.attention.output.dense.weight"))
            ops.gemm(self.dtb, w16(f"{l}.attention.output.dense.weight"), b_mn=True, out=self.d_ctx)
            ops.attn_bwd_ex(a.qkv[:, :H], a.qkv[:, H:2 * H], a.qkv[:, 2 * H:], a.ctx1, self.d_ctx, a.lse1, mb, S, S, nh,
                            self.dqkv[:, :H], self.dqkv[:, H:2 * H], self.dqkv[:, 2 * H:], kv_len=self.att_len, delta=self.delta,
                            drop=((ctx, self._site(i, 0), self.p_a) if da else None))
            ops.wgrad(self.dqkv, a.hb_in, self._fused(f"{l}.attention.self.query.weight", 3 * H, H, G))
            ops.colsum_accum(self.dqkv, self._fused1(f"{l}.attention.self.query.bias", 3 * H, G))
            ops.gemm(self.dqkv, self._fused(f"{l}.attention.self.query.weight", 3 * H, H), b_mn=True, residual=self.dt, out=self.dh)
        # ---- embeddings: h0 = dropout(LN(word[ids] + pos))
        if dh:
            ops.dropout(self.dh, ctx, 0, per, self.p_h)
        ops.layernorm_bwd(self.dh, self.emb_pre, self.me, self.re, w(f"{t}.embeddings.LayerNorm.weight"), self.dt,
                          g(f"{t}.embeddings.LayerNorm.weight"), g(f"{t}.embeddings.LayerNorm.bias"), accumulate=False)
        ops.text_embed_bwd(self.dt.view(mb, S, H), self.ids, g(f"{t}.embeddings.word_embeddings.weight"),
                           g(f"{t}.embeddings.position_embeddings.weight")[:S])
        if ff:
            # ---- ViT: img_tok = LN_final(blocks(patch-embed(img)))
            ops.layernorm_bwd(self.d_img, self.vis.x_out, self.vn_m, self.vn_r, w(f"{v}.norm.weight"), self.vis.dx,
                              g(f"{v}.norm.weight"), g(f"{v}.norm.bias"), accumulate=False, dx_bf16=self.vis.dxb,
                              dcolsum=(None if self.vis.dp_on else self.vis.top_bias_grad()))
        else:
            # ---- ViT: v_emb = vision_proj(LN_final(x_out)[class-token rows])
            ops.wgrad(self.d_v, self.v_pool, g(f"vision_proj{self.sfx}.weight"))
            ops.colsum_accum(self.d_v, g(f"vision_proj{self.sfx}.bias"))
            ops.gemm(self.d_v, w16(f"vision_proj{self.sfx}.weight"), b_mn=True, out=self.d_vpool)
            self.vis.dx.zero_()
            self.vis.dxb.zero_()
            ops.layernorm_bwd(self.d_vpool, self.vis.x_out, self.vn_m[:mb], self.vn_r[:mb], w(f"{v}.norm.weight"), self.vis.dx,
                              g(f"{v}.norm.weight"), g(f"{v}.norm.bias"), row_index=self.cls_rows, accumulate=False,
                              dx_bf16=self.vis.dxb, dcolsum=(None if self.vis.dp_on else self.vis.top_bias_grad()))
        self.vis.backward(top_bias_done=not self.vis.dp_on)
        ops.vision_embed_bwd(self.vis.dx.view(mb, Sv, d.v_width), mb, Sv, d.v_width, self.d_patch,
                             g(f"{v}.cls_token").view(-1), g(f"{v}.pos_embed").view(Sv, d.v_width))
        ops.wgrad(self.d_patch, self.col, g(f"{v}.patch_embed.proj.weight").view(d.v_width, -1))
        ops.colsum_accum(self.d_patch, g(f"{v}.patch_embed.proj.bias"))


class BLIPFeatureFusion(nn.Module):
    MODE = "ff"     # BLIPScoreFusion (blip_sf.py) overrides: same loss / queues / EMA, different encoder head
    # Train-mode regularisation of the reference: BertConfig hidden_dropout_prob / attention_probs_dropout_prob (0.1 in
    # backbone/configs/med_config.json) and timm DropPath in the ViT (blip.py:241,253: rate 0.1 for "large", 0 for "base").
    # Active whenever the module is in train mode, like nn.Dropout; `model.eval()` or `set_dropout(0, 0, 0)` turn it off.
    DROP_PATH_RATE = {"large": 0.1}

    def __init__(self, med_config="backbone/configs/med_config.json", image_size=224, vit="base", vit_grad_ckpt=False,
                 vit_ckpt_layer=0, embed_dim=768, queue_size=57600, momentum=0.995, config=None, device="cuda",
                 micro_batch=256, seed=2023):
        super().__init__()
        L.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise L.UniirError("uniir_b200.BLIPFeatureFusion runs on CUDA only (no CPU fallback)")
        if vit not in BLIP_CONFIGS:
            raise ValueError(f"vit must be one of {list(BLIP_CONFIGS)}")
        d = BLIP_CONFIGS[vit]
        if image_size != d.image_size:
            d = BLIPDims(**{**d.__dict__, "image_size": image_size})
        p_hidden = p_attn = 0.1            # backbone/configs/med_config.json:5,7
        if isinstance(med_config, dict):   # the reference reads a BertConfig json; a dict is accepted for tests
            med_config = dict(med_config)
            p_hidden = med_config.pop("hidden_dropout_prob", p_hidden)
            p_attn = med_config.pop("attention_probs_dropout_prob", p_attn)
            d = BLIPDims(**{**d.__dict__, **med_config})
        elif isinstance(med_config, str) and os.path.isfile(med_config):
            # The reference builds its BERT from this BertConfig json (blip_ff.py:34-36).  The kernels here are laid out
            # for the dimensions of BLIPDims: read the file and refuse a configuration they would silently ignore.
            with open(med_config) as f:
                mc = json.load(f)
            want = {"hidden_size": d.t_width, "num_hidden_layers": d.t_layers, "num_attention_heads": d.t_heads,
                    "intermediate_size": d.t_inter, "vocab_size": d.vocab, "max_position_embeddings": d.max_pos}
            bad = {k: (mc[k], v) for k, v in want.items() if k in mc and mc[k] != v}
            if bad:
                raise ValueError(f"med_config {med_config} does not match the built-in BERT dimensions "
                                 f"(file value, built-in value): {bad}; pass the overrides as a dict")
            if mc.get("hidden_act", "gelu") != "gelu" or abs(mc.get("layer_norm_eps", 1e-12) - 1e-12) > 1e-15:
                raise ValueError("med_config: only hidden_act='gelu' and layer_norm_eps=1e-12 are implemented")
            p_hidden = mc.get("hidden_dropout_prob", p_hidden)
            p_attn = mc.get("attention_probs_dropout_prob", p_attn)
        self.dims, self.device_, self.micro_batch = d, dev, micro_batch
        self.image_size, self.queue_size, self.momentum, self.embed_dim, self.config = image_size, queue_size, momentum, embed_dim, config
        # online + momentum encoders as two flat parameter sets with identical layouts (EMA = one kernel)
        self._out_dim = d.t_width if self.MODE == "ff" else embed_dim     # width of the fused embedding
        self._online = FlatParamModule(blip_param_specs(d, "", self.MODE, embed_dim), dev, seed=seed, root=self)
        self._mom = FlatParamModule(blip_param_specs(d, "_m", self.MODE, embed_dim), dev, seed=seed, root=self)
        for name, *_ in self._mom._specs:
            self.get_parameter(name).requires_grad_(False)          # blip_ff.py:283-287 copy_params
        if self.MODE == "sf":                                       # blip_sf.py:75-78
            for name, *_ in self._online._specs:
                if ".crossattention." in name:
                    self.get_parameter(name).requires_grad_(False)
        for k in self._mom.flat:
            self._mom.flat[k].copy_(self._online.flat[k][:self._mom.flat[k].numel()])
        for enc in ("text_encoder", "text_encoder_m"):
            self.get_submodule(f"{enc}.embeddings").register_buffer(
                "position_ids", torch.arange(d.max_pos, device=dev).expand((1, -1)))
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.register_buffer("query_queue", nn.functional.normalize(torch.randn(embed_dim, queue_size, generator=g), dim=0).to(dev))
        self.register_buffer("cand_queue", nn.functional.normalize(torch.randn(embed_dim, queue_size, generator=g), dim=0).to(dev))
        self.register_buffer("idx_queue", torch.full((1, queue_size), -100, device=dev))
        self.register_buffer("new_ptr_queue", torch.zeros(1, dtype=torch.long, device=dev))
        self._enc = {}
        self._train_state = None
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = float(p_hidden), float(p_attn)
        self.drop_path_rate = float(self.DROP_PATH_RATE.get(vit, 0.0))
        self._drop_gen = None            # host generator of the per-step mask seeds (seeded with seed + rank, lazily)
        self._drop_seed0 = seed
        self.train_micro_batch = 64
        self.use_cuda_graphs = True      # BLIPEncoder.run: one graph launch per micro-batch instead of ~2400 kernel launches

    # ---- reference surface -------------------------------------------------------------
    def get_img_preprocess_fn(self):
        """blip_ff.py:65-68: the train augmentation while the module is in train mode, the eval transform otherwise."""
        from .preprocess import blip_image_preprocess
        return blip_image_preprocess(self.image_size, is_train=self.training, min_scale=0.5)

    def get_tokenizer(self):
        """blip_ff.py:70-80: BertTokenizer("bert-base-uncased") + the "[DEC]" / "[ENC]" special tokens (blip.py:221-226),
        wrapped with padding="max_length", truncation, max_length=config.tokenizer_max_length, return_tensors="pt".
        The vocabulary is not in the reference tree and there is no network: UNIIR_BERT_VOCAB must point at
        bert-base-uncased's vocab.txt (or a directory holding it, e.g. a HuggingFace snapshot)."""
        from .preprocess import blip_tokenizer
        wrapper, self.tokenizer = blip_tokenizer(self.config.tokenizer_max_length if self.config is not None else 100,
                                                 model_vocab=self.dims.vocab)
        return wrapper

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._online.rebind_after_apply()
        self._mom.rebind_after_apply()
        return out

    def _encoder(self, mb, S, momentum, train=False):
        key = (mb, S, momentum, train)
        if key not in self._enc:
            self._enc = {k: v for k, v in self._enc.items() if (k[2], k[3]) != (momentum, train)}
            self._enc[key] = BLIPEncoder(self._mom if momentum else self._online, self.dims, mb, S, self.device_,
                                         "_m" if momentum else "", train=train, mode=self.MODE, embed_dim=self.embed_dim)
        enc = self._enc[key]
        probs = (self.hidden_dropout_prob, self.attention_probs_dropout_prob, self.drop_path_rate)
        if getattr(enc, "_drop_probs", None) != probs:
            enc.set_dropout(*probs)
            enc._drop_probs = probs
        return enc

    def set_dropout(self, hidden=0.0, attention=0.0, drop_path=0.0):
        """Override the train-mode dropout probabilities (the reference's come from med_config / create_vit)."""
        self.hidden_dropout_prob, self.attention_probs_dropout_prob, self.drop_path_rate = float(hidden), float(attention), float(drop_path)

    def _dropout_active(self):
        return self.training and (self.hidden_dropout_prob > 0 or self.attention_probs_dropout_prob > 0 or self.drop_path_rate > 0)

    def _next_drop_seed(self):
        if self._drop_gen is None:
            rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
            self._drop_gen = torch.Generator(device="cpu").manual_seed(self._drop_seed0 * 1000003 + rank)   # seed + rank, as train.py
        return int(torch.randint(0, 2 ** 31 - 1, (1,), generator=self._drop_gen))

    def _masks(self, txt_mask, img_mask, n):
        """Modality masks as int64 [n] device tensors (score fusion); None for feature fusion, which ignores them."""
        if self.MODE == "ff":
            return None, None
        one = torch.ones(n, dtype=torch.int64, device=self.device_)
        mt = one if txt_mask is None else txt_mask.to(self.device_, torch.int64).reshape(n).contiguous()
        mv = one if img_mask is None else img_mask.to(self.device_, torch.int64).reshape(n).contiguous()
        return mt, mv



    def _prep_txt(self, txt_dict_batched):
        ids = txt_dict_batched["input_ids"] if isinstance(txt_dict_batched, dict) else txt_dict_batched.input_ids
        att = txt_dict_batched["attention_mask"] if isinstance(txt_dict_batched, dict) else txt_dict_batched.attention_mask
        dev = self.device_
        ids = ids.to(dev, torch.int32).contiguous()
        att = att.to(dev)
        if not bool((att.cumsum(1) == torch.arange(1, att.shape[1] + 1, device=dev)).logical_or(att == 0).all()):
            raise ValueError("attention_mask must be a prefix mask (padding='max_length' tokenisation)")
        return ids, att.sum(1).to(torch.int32).contiguous()

    @torch.no_grad()
    def encode_multimodal_input(self, txt_dict_batched, image_batched, txt_mask=None, img_mask=None, use_momentum=False,
                                _drop_seed=None):
        """blip_ff.py:82-116: txt_mask / img_mask are accepted and ignored, exactly like the reference.
        blip_sf.py:158-172 (MODE "sf"): text_proj(text [CLS]) * txt_mask + vision_proj(image class token) * img_mask.
        In train mode the encoders run with dropout / DropPath like the reference's modules (fresh masks per call;
        the training step passes its own seed so that the gradient-cache re-forward sees the same masks)."""
        if _drop_seed is None and self._dropout_active():
            _drop_seed = self._next_drop_seed()
        ids = txt_dict_batched["input_ids"] if isinstance(txt_dict_batched, dict) else txt_dict_batched.input_ids
        att = txt_dict_batched["attention_mask"] if isinstance(txt_dict_batched, dict) else txt_dict_batched.attention_mask
        dev = self.device_
        ids = ids.to(dev, torch.int32).contiguous()
        att = att.to(dev)
        lens = att.sum(1).to(torch.int32)
        if not bool((att.cumsum(1) == torch.arange(1, att.shape[1] + 1, device=dev)).logical_or(att == 0).all()):
            raise ValueError("attention_mask must be a prefix mask (padding='max_length' tokenisation)")
        img = image_batched.to(dev, torch.float32).contiguous()
        n, S = ids.shape
        mb = min(self.micro_batch, n)
        enc = self._encoder(mb, S, use_momentum)
        P = self._mom if use_momentum else self._online
        P.refresh_shadow()
        out = torch.empty(n, self._out_dim, device=dev, dtype=torch.float32)
        mt, mv = self._masks(txt_mask, img_mask, n)
        for s in range(0, n, mb):
            e = min(n, s + mb)
            i_c, l_c, im_c = ids[s:e], lens[s:e], img[s:e]
            mt_c, mv_c = (None, None) if mt is None else (mt[s:e], mv[s:e])
            if e - s < mb:
                pad = mb - (e - s)
                i_c = torch.cat([i_c, i_c.new_zeros(pad, S)]); l_c = torch.cat([l_c, l_c.new_ones(pad)])
                im_c = torch.cat([im_c, im_c.new_zeros(pad, *im_c.shape[1:])])
                if mt is not None:
                    mt_c = torch.cat([mt_c, mt_c.new_zeros(pad)]); mv_c = torch.cat([mv_c, mv_c.new_zeros(pad)])
            out[s:e] = enc.run(i_c.contiguous(), l_c.contiguous(), im_c.contiguous(),
                               mt=None if mt_c is None else mt_c.contiguous(),
                               mv=None if mv_c is None else mv_c.contiguous(), use_graph=self.use_cuda_graphs,
                               drop_seed=_drop_seed, item0=s)[:e - s]
        return out

    def forward(self, batch, alpha=None, encode_mbeir_batch=False):
        if encode_mbeir_batch:
            return self.encode_mbeir_batch(batch)
        return self.compute_contrastive_loss(batch, alpha)

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        if id_list is None:
            raise ValueError("id_list not found in batch.")
        embeddings = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch.get("txt_mask_batched"),
                                                  batch.get("image_mask_batched"), use_momentum=False)
        assert embeddings.size(0) == len(id_list), "embeddings and id_batched must have the same batch size."
        return embeddings, id_list

    # ---- training: momentum distillation + queue (blip_ff.py:118-257) ---------------------------
    def compute_contrastive_loss(self, batch, alpha):
        """Reference-compatible entry: returns {"loss", "accuracy"}.  With autograd enabled the loss is a graph node
        whose ``.backward()`` (the reference loop's ``scaler.scale(loss).backward()``, uniir_blip/engine.py:45) runs
        ``backward_contrastive`` and fills ``p.grad`` of the online parameters; the fused engine path
        (``engine.train_step_blip`` / ``blip_engine.train_one_epoch``) calls the two halves directly."""
        if torch.is_grad_enabled() and self.training:
            from .autograd_bridge import bridged_loss

            def bwd(g_loss):
                self.backward_contrastive()
                g = float(g_loss)
                if g != 1.0:
                    for buf in self._online.grad.values():
                        buf.mul_(g)
            loss, acc = bridged_loss(self._online, lambda: self._loss_forward(batch, alpha), bwd)
        else:
            loss, acc = self._loss_forward(batch, alpha)
        return {"loss": loss, "accuracy": acc}

    @torch.no_grad()
    def _momentum_update(self):
        for k in self._mom.flat:
            ops.ema_update(self._mom.flat[k], self._online.flat[k][:self._mom.flat[k].numel()], self.momentum,
                           self._mom.flat16[k])
        self._mom.mark_shadow_fresh()

    @torch.no_grad()
    def _dequeue_and_enqueue(self, q_m, pc_m, pc_ids):
        from . import dist_utils as du
        idxs = du.all_gather_rows(pc_ids.view(-1, 1))
        q_m, pc_m = du.all_gather_rows(q_m), du.all_gather_rows(pc_m)
        n = q_m.shape[0]
        ptr = int(self.new_ptr_queue)
        assert self.queue_size % n == 0  # This is important (blip_ff.py:303)
        self.query_queue[:, ptr:ptr + n] = q_m.T
        self.cand_queue[:, ptr:ptr + n] = pc_m.T
        self.idx_queue[:, ptr:ptr + n] = idxs.T
        self.new_ptr_queue[0] = (ptr + n) % self.queue_size

    @torch.no_grad()
    def _loss_forward(self, batch, alpha):
        im = batch["index_mapping"]
        hard = "neg_cand_list" in im                                      # blip_ff.py:126-131
        dev, E = self.device_, self._out_dim
        P = self._online
        P.w("temp").clamp_(0.001, 0.5)                                   # blip_ff.py:132-133
        ids, lens = self._prep_txt(batch["txt_batched"])
        img = batch["image_batched"].to(dev, torch.float32).contiguous()
        q_idx = torch.tensor(im["query"], dtype=torch.int32).flatten().to(dev)
        p_idx = torch.tensor(im["pos_cand"], dtype=torch.int32).flatten().to(dev)
        pc_ids = torch.as_tensor(batch["p_did_list"]).to(dev, torch.int64).contiguous()
        bs, n_items, S = q_idx.numel(), ids.shape[0], ids.shape[1]
        txt = {"input_ids": ids, "attention_mask": (torch.arange(S, device=dev)[None, :] < lens[:, None]).long()}
        mt, mv = self._masks(batch.get("txt_mask_batched"), batch.get("image_mask_batched"), n_items)
        # train mode: one mask seed for the online encoder (phase A and the phase-C re-forward share it), one for the
        # momentum encoder (its modules are in train mode too in the reference)
        seed, seed_m = (self._next_drop_seed(), self._next_drop_seed()) if self._dropout_active() else (None, None)
        emb = self.encode_multimodal_input(txt, img, mt, mv, _drop_seed=seed)   # phase A, online (no activations kept)
        self._momentum_update()                                           # blip_ff.py:182
        emb_m = self.encode_multimodal_input(txt, img, mt, mv, use_momentum=True, _drop_seed=seed_m)
        f32 = dict(device=dev, dtype=torch.float32)
        qn, pn, qm, pm = (torch.empty(bs, E, **f32) for _ in range(4))
        qnorm, pnorm, tmpn = torch.empty(bs, **f32), torch.empty(bs, **f32), torch.empty(bs, **f32)
        lib = L.load()
        L.check(lib.uniir_infonce_normalize(emb.data_ptr(), q_idx.data_ptr(), p_idx.data_ptr(), bs, E, qn.data_ptr(),
                                            pn.data_ptr(), qnorm.data_ptr(), pnorm.data_ptr(), ops._stream()), "normalize")
        L.check(lib.uniir_infonce_normalize(emb_m.data_ptr(), q_idx.data_ptr(), p_idx.data_ptr(), bs, E, qm.data_ptr(),
                                            pm.data_ptr(), tmpn.data_ptr(), tmpn.data_ptr(), ops._stream()), "normalize")
        q_all = torch.cat([qm.t(), self.query_queue], dim=1).contiguous()     # [E, bs+Qz]  (data movement only)
        if hard:
            # blip_ff.py:160-171, 208-218: the negatives' ids and (un-normalised, :196) momentum features take the place
            # of the first bs*neg_num queue columns
            n_rows = torch.tensor(im["neg_cand_list"], dtype=torch.int64).to(dev)         # [bs, neg_num]
            nc_ids = torch.as_tensor(batch["nc_dids_list"]).to(dev, torch.int64).reshape(bs, -1)
            hard_n = nc_ids.numel()
            nc_m = emb_m[n_rows.flatten()]                                                  # [bs*neg_num, E]
            pc_all = torch.cat([pm.t(), nc_m.t(), self.cand_queue[:, hard_n:]], dim=1).contiguous()
            id_all = torch.cat([pc_ids, nc_ids.flatten(), self.idx_queue[0, hard_n:]]).contiguous()
        else:
            pc_all = torch.cat([pm.t(), self.cand_queue], dim=1).contiguous()
            id_all = torch.cat([pc_ids, self.idx_queue[0]]).contiguous()
        temp = P.w("temp")
        S1, S1m = ops.sgemm_temp(qn, pc_all, True, temp), ops.sgemm_temp(qm, pc_all, True, temp)
        S2, S2m = ops.sgemm_temp(pn, q_all, True, temp), ops.sgemm_temp(pm, q_all, True, temp)
        acc_l = torch.zeros(4, **f32)     # loss q2pc, loss pc2q, acc, d_temp
        gs = 0.5 / bs
        ops.blip_softce(S1, S1m, pc_ids, id_all, float(alpha), gs, temp, acc_l[0:1], acc_l[2:3], acc_l[3:4])
        ops.blip_softce(S2, S2m, pc_ids, id_all, float(alpha), gs, temp, acc_l[1:2], None, acc_l[3:4])
        loss = (acc_l[0] + acc_l[1]) * gs
        acc = acc_l[2] / bs
        d_qn = ops.sgemm_temp(S1, pc_all, False, temp)                    # G @ pc_all^T / temp
        d_pn = ops.sgemm_temp(S2, q_all, False, temp)
        d_emb = torch.zeros(n_items, E, **f32)
        L.check(lib.uniir_infonce_normalize_bwd(q_idx.data_ptr(), p_idx.data_ptr(), qn.data_ptr(), pn.data_ptr(),
                                                qnorm.data_ptr(), pnorm.data_ptr(), d_qn.data_ptr(), d_pn.data_ptr(), bs, E,
                                                d_emb.data_ptr(), ops._stream()), "normalize_bwd")
        if hard and not bool(torch.rand(1) < 0.5):                        # blip_ff.py:233-244: coin flip (host RNG, like the reference)
            self._dequeue_and_enqueue(qm, nc_m.view(bs, -1, E)[:, 0, :].contiguous(), nc_ids[:, 0].contiguous())
        else:
            self._dequeue_and_enqueue(qm, pm, pc_ids)                     # blip_ff.py:247
        self._train_state = dict(ids=ids, lens=lens, img=img, d_emb=d_emb, d_temp=acc_l[3].clone(), mt=mt, mv=mv, seed=seed)
        return loss.clone(), acc.clone()

    @torch.no_grad()
    def backward_contrastive(self):
        """Phase C of the gradient cache: re-forward every micro-batch with saved activations and
        backpropagate the cached embedding gradients into the flat gradient buffers."""
        st = self._train_state
        assert st is not None, "backward without forward"
        P = self._online
        P.ensure_grad()
        P.g("temp").add_(st["d_temp"])
        n, S = st["ids"].shape
        mb = min(self.train_micro_batch, n)
        if n % mb or mb % 8:
            mb = next((m for m in range(mb - mb % 8, 7, -8) if n % m == 0), None)
            if mb is None:
                raise ValueError(f"batch of {n} items cannot be tiled by a micro-batch that is a multiple of 8")
        enc = self._encoder(mb, S, False, train=True)
        P.refresh_shadow()
        for s in range(0, n, mb):
            e = s + mb
            enc.run(st["ids"][s:e].contiguous(), st["lens"][s:e].contiguous(), st["img"][s:e].contiguous(),
                    mt=None if st["mt"] is None else st["mt"][s:e].contiguous(),
                    mv=None if st["mv"] is None else st["mv"][s:e].contiguous(), d_emb=st["d_emb"][s:e].contiguous(),
                    use_graph=self.use_cuda_graphs, drop_seed=st["seed"], item0=s)
        self._train_state = None
