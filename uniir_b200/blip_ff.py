"""Drop-in for the reference ``BLIPFeatureFusion`` (src/models/uniir_blip/blip_featurefusion/
blip_ff.py:11-332) on the B200-native kernels -- ENCODE PATH.

Built this round: constructor / state-dict layout (online + momentum encoders, queues, temp: reference
checkpoints load with a plain ``load_state_dict``), ``encode_multimodal_input``, ``forward(batch,
encode_mbeir_batch=True)`` / ``encode_mbeir_batch`` (embedding extraction and retrieval with BLIP_FF).
Not built yet: ``compute_contrastive_loss`` (momentum distillation + queue, row a9) raises
NotImplementedError.

Arithmetic (vendored reference code, restated in oracle/blip_oracle.py and pinned against the shimmed
reference modules): ViT-L/16 with biased patch-embed, pre-LN blocks (eps 1e-6), exact GELU, final norm, all
197 tokens kept (backbone/vit.py:254-268); BERT "MED" with word+position embeddings -> LN(1e-12), 12 post-LN
layers of self-attention (key-padding mask), cross-attention into the image tokens (K/V = Linear(1024->768)),
GELU FFN, then pooler tanh(W h[:,0] + b) (backbone/med.py:52-98,101-232,330-396,498-511).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
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


def blip_param_specs(d: BLIPDims, suffix=""):
    """[(name, shape, decay, std)].  q/k/v (and cross k/v) weights, then their biases, are adjacent so the
    fused [3D, D] / [2D, Dv] projection matrices are plain views of the flat buffer.  `decay` is True for
    everything: the reference uses one AdamW group with wd 0.05 (uniir_blip/train.py:193-197)."""
    sp = []
    add = lambda n, shape, std: sp.append((n, tuple(shape), True, std))
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
    add(f"{t}.pooler.dense.weight", (H, H), 0.02); add(f"{t}.pooler.dense.bias", (H,), 0.02)
    return sp


class BLIPEncoder:
    """ViT + cross-attention BERT forward for a fixed (micro-batch, text length)."""

    def __init__(self, P: FlatParamModule, d: BLIPDims, mb, S, dev, suffix=""):
        self.P, self.d, self.mb, self.S, self.dev = P, d, mb, S, dev
        self.v, self.t = f"visual_encoder{suffix}", f"text_encoder{suffix}"
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        self.vis = Tower(P, self.v, d.v_width, d.v_layers, d.v_heads, d.v_tokens, False, mb, dev, False,
                         names=Tower.TIMM_NAMES, act=L.ACT_GELU, eps=1e-6)
        Kp = 3 * d.patch * d.patch
        assert Kp % 8 == 0
        self.col = torch.empty(mb * d.grid ** 2, Kp, **bf)
        self.patch = torch.empty(mb * d.grid ** 2, d.v_width, **bf)
        self.img_tok = torch.empty(mb * d.v_tokens, d.v_width, **bf)
        T, H = mb * S, d.t_width
        self.h = torch.empty(T, H, **f32)
        self.tmp = torch.empty(T, H, **f32)
        self.hb = torch.empty(T, H, **bf)
        self.qkv = torch.empty(T, 3 * H, **bf)
        self.ctx = torch.empty(T, H, **bf)
        self.q = torch.empty(T, H, **bf)
        self.kv = torch.empty(mb * d.v_tokens, 2 * H, **bf)
        self.inter = torch.empty(T, d.t_inter, **bf)
        self.pool = torch.empty(mb, H, **bf)
        self.emb = torch.empty(mb, H, **f32)

    def _fused(self, first, rows, cols):
        """[rows, cols] bf16 view starting at parameter `first` (adjacent parameters, see blip_param_specs)."""
        decay, off, n, shape = self.P._index[first]
        return self.P.flat16[decay][off:off + rows * cols].view(rows, cols)

    def _fused32(self, first, n):
        decay, off, _, _ = self.P._index[first]
        return self.P.flat[decay][off:off + n]

    def forward(self, ids, att_len, img):
        """ids int32 [mb,S]; att_len int32 [mb] (number of valid tokens: attention_mask is a prefix mask);
        img fp32 [mb,3,H,W].  Returns pooled fused embedding fp32 [mb, 768] (buffer reused)."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        w, w16 = P.w, P.w16
        v, t = self.v, self.t
        # ---------------- ViT ----------------
        ops.patch_im2col(img, d.patch, self.col.shape[1], out=self.col)
        ops.gemm(self.col, w16(f"{v}.patch_embed.proj.weight").view(d.v_width, -1), bias=w(f"{v}.patch_embed.proj.bias"),
                 out=self.patch)
        ops.vision_embed_fwd(self.patch, w(f"{v}.cls_token").view(-1), w(f"{v}.pos_embed").view(d.v_tokens, d.v_width),
                             mb, d.v_tokens, d.v_width, out=self.vis.x.view(mb, d.v_tokens, d.v_width))
        xo = self.vis.forward(False)
        ops.layernorm_fwd(xo, w(f"{v}.norm.weight"), w(f"{v}.norm.bias"), 1e-6, out=self.img_tok)
        # ---------------- BERT ----------------
        H, nh = d.t_width, d.t_heads
        ops.text_embed_fwd(ids, w(f"{t}.embeddings.word_embeddings.weight"),
                           w(f"{t}.embeddings.position_embeddings.weight")[:S], out=self.tmp.view(mb, S, H))
        ops.layernorm_fwd(self.tmp, w(f"{t}.embeddings.LayerNorm.weight"), w(f"{t}.embeddings.LayerNorm.bias"), 1e-12, out=self.h)
        ops.cast_bf16(self.h, out=self.hb)
        for i in range(d.t_layers):
            l = f"{t}.encoder.layer.{i}"
            # self-attention (q|k|v fused), key-padding mask
            ops.gemm(self.hb, self._fused(f"{l}.attention.self.query.weight", 3 * H, H),
                     bias=self._fused32(f"{l}.attention.self.query.bias", 3 * H), out=self.qkv)
            ops.attn_fwd_ex(self.qkv[:, :H], self.qkv[:, H:2 * H], self.qkv[:, 2 * H:], mb, S, S, nh, kv_len=att_len, out=self.ctx)
            ops.gemm(self.ctx, w16(f"{l}.attention.output.dense.weight"), bias=w(f"{l}.attention.output.dense.bias"),
                     residual=self.h, out=self.tmp)
            ops.layernorm_fwd(self.tmp, w(f"{l}.attention.output.LayerNorm.weight"), w(f"{l}.attention.output.LayerNorm.bias"),
                              1e-12, out=self.h)
            ops.cast_bf16(self.h, out=self.hb)
            # cross-attention into the image tokens (encoder mask is all ones, blip_ff.py:107)
            ops.gemm(self.hb, w16(f"{l}.crossattention.self.query.weight"), bias=w(f"{l}.crossattention.self.query.bias"), out=self.q)
            ops.gemm(self.img_tok, self._fused(f"{l}.crossattention.self.key.weight", 2 * H, d.v_width),
                     bias=self._fused32(f"{l}.crossattention.self.key.bias", 2 * H), out=self.kv)
            ops.attn_fwd_ex(self.q, self.kv[:, :H], self.kv[:, H:], mb, S, d.v_tokens, nh, out=self.ctx)
            ops.gemm(self.ctx, w16(f"{l}.crossattention.output.dense.weight"), bias=w(f"{l}.crossattention.output.dense.bias"),
                     residual=self.h, out=self.tmp)
            ops.layernorm_fwd(self.tmp, w(f"{l}.crossattention.output.LayerNorm.weight"),
                              w(f"{l}.crossattention.output.LayerNorm.bias"), 1e-12, out=self.h)
            ops.cast_bf16(self.h, out=self.hb)
            # FFN
            ops.gemm(self.hb, w16(f"{l}.intermediate.dense.weight"), bias=w(f"{l}.intermediate.dense.bias"), act=L.ACT_GELU,
                     out=self.inter)
            ops.gemm(self.inter, w16(f"{l}.output.dense.weight"), bias=w(f"{l}.output.dense.bias"), residual=self.h, out=self.tmp)
            ops.layernorm_fwd(self.tmp, w(f"{l}.output.LayerNorm.weight"), w(f"{l}.output.LayerNorm.bias"), 1e-12, out=self.h)
            ops.cast_bf16(self.h, out=self.hb)
        self.pool.copy_(self.hb.view(mb, S, H)[:, 0])
        ops.gemm(self.pool, w16(f"{t}.pooler.dense.weight"), bias=w(f"{t}.pooler.dense.bias"), act=L.ACT_TANH, out=self.emb)
        return self.emb


class BLIPFeatureFusion(nn.Module):
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
        if isinstance(med_config, dict):   # the reference reads a BertConfig json; a dict is accepted for tests
            d = BLIPDims(**{**d.__dict__, **med_config})
        self.dims, self.device_, self.micro_batch = d, dev, micro_batch
        self.image_size, self.queue_size, self.momentum, self.embed_dim, self.config = image_size, queue_size, momentum, embed_dim, config
        # online + momentum encoders as two flat parameter sets with identical layouts (EMA = one kernel)
        self._online = FlatParamModule(blip_param_specs(d), dev, seed=seed, root=self)
        self._mom = FlatParamModule(blip_param_specs(d, "_m"), dev, seed=seed, root=self)
        for name, *_ in self._mom._specs:
            self.get_parameter(name).requires_grad_(False)          # blip_ff.py:283-287 copy_params
        for k in self._mom.flat:
            self._mom.flat[k].copy_(self._online.flat[k])
        for enc in ("text_encoder", "text_encoder_m"):
            self.get_submodule(f"{enc}.embeddings").register_buffer(
                "position_ids", torch.arange(d.max_pos, device=dev).expand((1, -1)))
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.register_buffer("query_queue", nn.functional.normalize(torch.randn(embed_dim, queue_size, generator=g), dim=0).to(dev))
        self.register_buffer("cand_queue", nn.functional.normalize(torch.randn(embed_dim, queue_size, generator=g), dim=0).to(dev))
        self.register_buffer("idx_queue", torch.full((1, queue_size), -100, device=dev))
        self.register_buffer("new_ptr_queue", torch.zeros(1, dtype=torch.long, device=dev))
        self.temp = nn.Parameter(0.07 * torch.ones([], device=dev))
        self._enc = {}

    # ---- reference surface -------------------------------------------------------------
    def get_img_preprocess_fn(self):
        from .preprocess import blip_image_preprocess
        return blip_image_preprocess(self.image_size)

    def get_tokenizer(self):
        raise FileNotFoundError("the bert-base-uncased vocabulary (BertTokenizer, blip.py:221-226) is not available offline; "
                                "pass pre-tokenised {'input_ids','attention_mask'} batches")

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._online.rebind_after_apply()
        self._mom.rebind_after_apply()
        return out

    def _encoder(self, mb, S, momentum):
        key = (mb, S, momentum)
        if key not in self._enc:
            self._enc = {k: v for k, v in self._enc.items() if k[2] != momentum}
            self._enc[key] = BLIPEncoder(self._mom if momentum else self._online, self.dims, mb, S, self.device_,
                                         "_m" if momentum else "")
        return self._enc[key]

    @torch.no_grad()
    def encode_multimodal_input(self, txt_dict_batched, image_batched, txt_mask=None, img_mask=None, use_momentum=False):
        """blip_ff.py:82-116.  txt_mask / img_mask are accepted and ignored, exactly like the reference."""
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
        out = torch.empty(n, self.dims.t_width, device=dev, dtype=torch.float32)
        for s in range(0, n, mb):
            e = min(n, s + mb)
            i_c, l_c, im_c = ids[s:e], lens[s:e], img[s:e]
            if e - s < mb:
                pad = mb - (e - s)
                i_c = torch.cat([i_c, i_c.new_zeros(pad, S)]); l_c = torch.cat([l_c, l_c.new_ones(pad)])
                im_c = torch.cat([im_c, im_c.new_zeros(pad, *im_c.shape[1:])])
            out[s:e] = enc.forward(i_c.contiguous(), l_c.contiguous(), im_c.contiguous())[:e - s]
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

    def compute_contrastive_loss(self, batch, alpha):
        raise NotImplementedError("BLIP_FF momentum-distillation / queue loss (blip_ff.py:118-257) is not built yet "
                                  "(SURVEY.md 8 row a9); the encode path is available")
