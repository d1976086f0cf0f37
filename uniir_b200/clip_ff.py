"""Drop-in for the reference ``CLIPFeatureFusion`` (src/models/uniir_clip/clip_featurefusion/clip_ff.py:62-300;
SURVEY.md 8f-3) on the B200-native kernels -- ENCODE PATH (``encode_multimodal_input``, ``forward(batch,
encode_mbeir_batch=True)`` / ``encode_mbeir_batch``: embedding extraction and retrieval with CLIP_FF checkpoints).

STATUS: written and compiled in a session that had no GPU time left; parity against the oracle is asserted by
tests/test_zz_unconfirmed_gpu.py, which is marked xfail(strict=False) until it has been seen green on a B200.
Training (``compute_inbatch_contrastive_loss``) raises NotImplementedError: the backward of the T5 stack (RMS-norm
backward, the bias gradient of the attention) is not built.

Arithmetic (restated in oracle/clip_ff_oracle.py, whose T5 stack is pinned against transformers' own ``T5Stack``):
  vision  upstream CLIP ViT *without pooling* (clip_ff.py:23-59): ln_post and ``@ proj`` on all 257 tokens;
  text    CLIP text trunk to ln_final on all 77 tokens, no text_projection (:110, :148-156);
  fusion  cat([text, image]) -> 2-layer T5 encoder (RMS norm, q/k/v/o without bias, unscaled scores + relative-
          position bias, ReLU FFN 768 -> 2048 -> 768) -> final RMS norm -> mean over the 334 tokens (:161-192);
          ``txt_mask`` / ``img_mask`` are accepted and ignored, exactly like the reference (:173-176).
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
    add = lambda n, shape, std: specs.append((n, tuple(shape), len(shape) >= 2 and "layer_norm" not in n, std))
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


class CLIPFFEncoder:
    """Un-pooled CLIP towers + T5 fusion stack for a fixed micro-batch (inference)."""

    def __init__(self, P: FlatParamModule, d: CLIPDims, heads: int, mb: int, dev, layers: int = 2):
        self.P, self.d, self.heads, self.mb, self.dev, self.layers = P, d, heads, mb, dev, layers
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        cm = "clip_model."
        self.vis = Tower(P, cm + "visual.transformer", d.v_width, d.v_layers, d.v_heads, d.v_tokens, False, mb, dev, False)
        self.txt = Tower(P, cm + "transformer", d.t_width, d.t_layers, d.t_heads, d.context, True, mb, dev, False)
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
        self.x = torch.empty(mb, self.S, D, **f32)           # fp32 residual stream of the T5 stack
        self.h = torch.empty(T, D, **bf)
        self.qkv = torch.empty(T, 3 * inner, **bf)
        self.att = torch.empty(T, inner, **bf)
        self.ff = torch.empty(T, T5_FF, **bf)
        self.y = torch.empty(mb, self.S, D, **f32)
        self.emb = torch.empty(mb, D, **f32)
        self.bucket = t5_relative_position_bucket(self.S).to(dev)
        self.bias = torch.empty(heads, self.S, self.S, **f32)

    def refresh_weights(self):
        """bf16 copy of the (K-padded) patch-embedding weight and the [H, S, S] relative-position bias table
        (a gather of the [32, H] embedding: once per weight update, not per batch)."""
        d = self.d
        ops.pad_cast_bf16(self.P.w("clip_model.visual.conv1.weight").view(d.v_width, d.patch_k), d.patch_k_pad, out=self.conv_w16)
        w = self.P.w("t5_layers.block.0.layer.0.SelfAttention.relative_attention_bias.weight")      # [32, H]
        self.bias.copy_(w[self.bucket].permute(2, 0, 1))

    def _fused_qkv(self, i):
        decay, off, n, shape = self.P._index[f"t5_layers.block.{i}.layer.0.SelfAttention.q.weight"]
        inner, D = shape
        return self.P.flat16[decay][off:off + 3 * inner * D].view(3 * inner, D)

    def forward(self, ids, img):
        """ids int32 [mb, 77], img fp32 [mb, 3, H, W] -> fused embedding fp32 [mb, embed_dim] (buffer reused)."""
        P, d, mb, S = self.P, self.d, self.mb, self.S
        w, w16 = P.w, P.w16
        cm = "clip_model."
        # ---- vision tower, all tokens (clip_ff.py:35-59)
        ops.patch_im2col(img, d.patch, d.patch_k_pad, out=self.col)
        ops.gemm(self.col, self.conv_w16, out=self.patch)
        ops.vision_embed_fwd(self.patch, w(cm + "visual.class_embedding"), w(cm + "visual.positional_embedding"),
                             mb, d.v_tokens, d.v_width, out=self.x0)
        ops.layernorm_fwd(self.x0, w(cm + "visual.ln_pre.weight"), w(cm + "visual.ln_pre.bias"), LN_EPS, out=self.vis.x)
        xo = self.vis.forward(False)
        ops.layernorm_fwd(xo, w(cm + "visual.ln_post.weight"), w(cm + "visual.ln_post.bias"), LN_EPS, out=self.v_tok)
        ops.gemm(self.v_tok, w16(cm + "visual.proj"), b_mn=True, out=self.v_feat.view(mb * d.v_tokens, d.embed_dim))
        # ---- text trunk, all tokens, no projection (clip_ff.py:148-156)
        ops.text_embed_fwd(ids, w(cm + "token_embedding.weight"), w(cm + "positional_embedding"),
                           out=self.txt.x.view(mb, d.context, d.t_width), eot=self.eot)
        xo = self.txt.forward(False)
        ops.layernorm_fwd(xo, w(cm + "ln_final.weight"), w(cm + "ln_final.bias"), LN_EPS,
                          out=self.t_feat.view(mb * d.context, d.t_width))
        # ---- cat([text, image], dim=1) (clip_ff.py:177)
        self.x[:, :d.context].copy_(self.t_feat)
        self.x[:, d.context:].copy_(self.v_feat)
        # ---- T5 encoder stack (clip_ff.py:178-183)
        x2 = self.x.view(mb * S, d.embed_dim)
        inner, H = self.heads * T5_DKV, self.heads
        for i in range(self.layers):
            a, f = f"t5_layers.block.{i}.layer.0", f"t5_layers.block.{i}.layer.1"
            ops.rmsnorm_fwd(x2, w(f"{a}.layer_norm.weight"), T5_EPS, out=self.h)
            ops.gemm(self.h, self._fused_qkv(i), out=self.qkv)
            ops.attn_fwd_bias(self.qkv[:, :inner], self.qkv[:, inner:2 * inner], self.qkv[:, 2 * inner:], mb, S, S, H,
                              self.bias, scale=1.0, out=self.att)
            ops.gemm(self.att, w16(f"{a}.SelfAttention.o.weight"), residual=x2, out=x2)
            ops.rmsnorm_fwd(x2, w(f"{f}.layer_norm.weight"), T5_EPS, out=self.h)
            ops.gemm(self.h, w16(f"{f}.DenseReluDense.wi.weight"), act=L.ACT_RELU, out=self.ff)
            ops.gemm(self.ff, w16(f"{f}.DenseReluDense.wo.weight"), residual=x2, out=x2)
        ops.rmsnorm_fwd(x2, w("t5_layers.final_layer_norm.weight"), T5_EPS, out=self.y)
        return ops.token_mean(self.y, out=self.emb)           # mean pooling (clip_ff.py:185-192)


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
        self._enc = {}

    # ---- reference surface ---------------------------------------------------------
    def get_img_preprocess_fn(self):
        from .preprocess import clip_image_preprocess
        return clip_image_preprocess(self.dims.image_size)

    def get_tokenizer(self):
        from .preprocess import clip_tokenizer
        return clip_tokenizer(context_length=77)

    def get_logit_scale(self):
        return self.get_parameter("clip_model.logit_scale").exp()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flat.rebind_after_apply()
        return out

    def _encoder(self, mb):
        if mb not in self._enc:
            self._enc = {mb: CLIPFFEncoder(self._flat, self.dims, self.heads, mb, self.device_)}
        return self._enc[mb]

    @torch.no_grad()
    def encode_multimodal_input(self, txt_tensor, img_tensor, txt_mask=None, img_mask=None):
        """[B,77] int, [B,3,H,W] fp32 -> fp32 [B, embed_dim]; the masks are ignored like in the reference."""
        dev = self.device_
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
            out[s:e] = enc.forward(i_c.contiguous(), im_c.contiguous())[:e - s]
        return out

    def compute_inbatch_contrastive_loss(self, batch):
        raise NotImplementedError("CLIP_FF training (clip_ff.py:197-270) is not built: the T5 stack has no backward yet; "
                                  "the encode path (embedding extraction / retrieval) is")

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
