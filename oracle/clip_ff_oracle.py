"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU fp32 restatement of the reference's CLIPFeatureFusion
(src/models/uniir_clip/clip_featurefusion/clip_ff.py:62-270), SURVEY.md 8f-3 -- the ORACLE for the next row to be
built; there is no product path for CLIP_FF yet.

What the reference does (clip_ff.py):
  * vision: upstream CLIP VisionTransformer WITHOUT pooling (:23-59): ln_post and `@ proj` on all 257 tokens;
  * text: the CLIP text trunk up to ln_final on all 77 tokens, NO text_projection (deleted at :110, :148-156);
  * fusion (:161-192): cat([text tokens, image tokens], dim=1) -> a 2-layer HuggingFace T5 *encoder* stack
    (`T5Stack(T5Config(num_layers=2, num_heads=12, d_model=768 | 512, d_kv=64))`, d_ff 2048, ReLU, RMS "T5LayerNorm"
    eps 1e-6, relative-position bias with 32 buckets / max distance 128 shared from block 0, no 1/sqrt(d) scaling,
    no attention mask -- `txt_mask` / `img_mask` are ignored, :173-176) -> mean over the 334 tokens;
  * loss (:197-270): the same in-batch / hard-negative contrastive loss as CLIP_SF (clip_sf.py:68-147), reused
    here from oracle/clip_oracle.py.

The T5 arithmetic lives in the third-party dependency `transformers` (unpinned in src/models/uniir_env.yml; 5.5.0 is
installed here): tests/test_oracle_clip_ff.py pins T5StackOracle against `transformers.models.t5.modeling_t5.T5Stack`
with copied weights, the rest against the CLIP oracle that tests/test_oracle_clip.py pins.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .clip_oracle import CLIP_CONFIGS, CLIPOracle, contrastive_loss_from_embeddings


class _T5Norm(nn.Module):
    """T5LayerNorm: x * rsqrt(mean(x^2) + eps) * weight -- no mean subtraction, no bias."""

    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps

    def forward(self, x):
        var = x.float().pow(2).mean(-1, keepdim=True)
        return self.weight * (x * torch.rsqrt(var + self.eps))


# Dropout sites of transformers' encoder T5Stack in train mode (T5Config.dropout_rate = 0.1; the reference keeps the
# default, clip_ff.py:82-98): T5Stack.forward `dropout(inputs_embeds)` and `dropout(final_layer_norm(h))`; T5Attention
# `dropout(softmax(scores))`; T5LayerSelfAttention `h + dropout(attn)`; T5DenseActDense `wo(dropout(relu(wi(x))))`;
# T5LayerFF `h + dropout(ff)`.  The restatement is the eval-mode arithmetic (identity at every site) unless a test
# installs DROPOUT, a callable (kind, layer, x) -> x with kind in "input", "probs", "attn_out", "ff_inner", "ff_out",
# "final": the masks then come from the caller (tests/test_oracle_clip_ff.py pins the SITES against transformers'
# T5Stack in train mode by handing both the same masks).
DROPOUT = None


def _drop(kind, layer, x):
    return x if DROPOUT is None else DROPOUT(kind, layer, x)


def t5_relative_position_bucket(rel, num_buckets=32, max_distance=128):
    """Bidirectional bucketing of `memory_position - query_position` (modeling_t5 `_relative_position_bucket`)."""
    num_buckets //= 2
    out = (rel > 0).long() * num_buckets
    rel = rel.abs()
    max_exact = num_buckets // 2
    small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact)
                         * (num_buckets - max_exact)).long()
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return out + torch.where(small, rel, large)


class _T5SelfAttention(nn.Module):
    def __init__(self, d_model, heads, d_kv, has_bias, num_buckets):
        super().__init__()
        inner = heads * d_kv
        self.heads, self.d_kv = heads, d_kv
        self.q = nn.Linear(d_model, inner, bias=False)
        self.k = nn.Linear(d_model, inner, bias=False)
        self.v = nn.Linear(d_model, inner, bias=False)
        self.o = nn.Linear(inner, d_model, bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    def forward(self, x, position_bias, layer=0):
        B, S, _ = x.shape
        sp = lambda t: t.view(B, S, self.heads, self.d_kv).transpose(1, 2)
        q, k, v = sp(self.q(x)), sp(self.k(x)), sp(self.v(x))
        scores = q @ k.transpose(-1, -2) + position_bias          # NB: no 1/sqrt(d_kv) (folded into the T5 init)
        w = _drop("probs", layer, F.softmax(scores.float(), dim=-1).type_as(scores))
        return self.o((w @ v).transpose(1, 2).reshape(B, S, self.heads * self.d_kv))


class _T5LayerSelfAttention(nn.Module):
    def __init__(self, d_model, heads, d_kv, has_bias, num_buckets):
        super().__init__()
        self.SelfAttention = _T5SelfAttention(d_model, heads, d_kv, has_bias, num_buckets)
        self.layer_norm = _T5Norm(d_model)


class _T5DenseReluDense(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.wi = nn.Linear(d_model, d_ff, bias=False)
        self.wo = nn.Linear(d_ff, d_model, bias=False)


class _T5LayerFF(nn.Module):
    def __init__(self, d_model, d_ff):
        super().__init__()
        self.DenseReluDense = _T5DenseReluDense(d_model, d_ff)
        self.layer_norm = _T5Norm(d_model)


class _T5Block(nn.Module):
    def __init__(self, d_model, heads, d_kv, d_ff, has_bias, num_buckets):
        super().__init__()
        self.layer = nn.ModuleList([_T5LayerSelfAttention(d_model, heads, d_kv, has_bias, num_buckets),
                                    _T5LayerFF(d_model, d_ff)])


class T5StackOracle(nn.Module):
    """Encoder `T5Stack` on `inputs_embeds` without mask -- HF parameter names; dropout: see DROPOUT above."""

    def __init__(self, d_model=768, num_layers=2, heads=12, d_kv=64, d_ff=2048, num_buckets=32, max_distance=128):
        super().__init__()
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.block = nn.ModuleList([_T5Block(d_model, heads, d_kv, d_ff, i == 0, num_buckets) for i in range(num_layers)])
        self.final_layer_norm = _T5Norm(d_model)

    def position_bias(self, S):
        pos = torch.arange(S)
        bucket = t5_relative_position_bucket(pos[None, :] - pos[:, None], self.num_buckets, self.max_distance)
        emb = self.block[0].layer[0].SelfAttention.relative_attention_bias(bucket)       # [S, S, heads]
        return emb.permute(2, 0, 1).unsqueeze(0)                                            # [1, heads, S, S]

    def forward(self, x):
        bias = self.position_bias(x.shape[1])
        x = _drop("input", 0, x)
        for i, blk in enumerate(self.block):
            att, ff = blk.layer
            x = x + _drop("attn_out", i, att.SelfAttention(att.layer_norm(x), bias, i))
            x = x + _drop("ff_out", i, ff.DenseReluDense.wo(_drop("ff_inner", i, F.relu(ff.DenseReluDense.wi(ff.layer_norm(x))))))
        return _drop("final", 0, self.final_layer_norm(x))


class CLIPFeatureFusionOracle(nn.Module):
    """Reference CLIPFeatureFusion restated; state-dict keys `clip_model.*` (no `text_projection`) and `t5_layers.*`."""

    def __init__(self, model_name="ViT-B/32", config=None, rank=0, world=1, gather_fn=None):
        super().__init__()
        self.clip_model = CLIPOracle(model_name)
        d = self.clip_model.dims
        del self.clip_model.text_projection                                   # clip_ff.py:110
        if model_name not in ("ViT-B/32", "ViT-L/14", "tiny-test"):
            raise NotImplementedError("Only ViT-B/32 and ViT-L/14 are supported.")  # clip_ff.py:100
        if d.t_width != d.embed_dim:
            raise ValueError("text width must equal the vision projection width (tokens are concatenated)")
        self.t5_layers = T5StackOracle(d_model=d.embed_dim, heads=12 if model_name != "tiny-test" else 2)
        self.gather_embeddings = config.model.gather_embeddings if config is not None else None
        self.in_batch_neg_num = config.data_config.in_batch_neg_num if config is not None else None
        self.rank, self.world, self._gather = rank, world, gather_fn     # gather_fn(p) -> all_p (clip_ff.py:231-232)

    def encode_text(self, ids):                                                # clip_ff.py:148-156
        m = self.clip_model
        x = m.token_embedding(ids.long()) + m.positional_embedding
        return m.ln_final(m.transformer(x, m.attn_mask))                      # [B, 77, t_width]

    def encode_image(self, img):                                               # clip_ff.py:35-59
        v = self.clip_model.visual
        x = v.conv1(img.float()).flatten(2).transpose(1, 2)
        x = torch.cat([v.class_embedding.expand(x.shape[0], 1, -1), x], dim=1) + v.positional_embedding
        x = v.ln_post(v.transformer(v.ln_pre(x)))
        return x @ v.proj                                                      # [B, tokens, embed_dim]

    def encode_multimodal_input(self, txt, img, txt_mask=None, img_mask=None):   # masks ignored (:173-176)
        feats = torch.cat([self.encode_text(txt), self.encode_image(img)], dim=1)
        return self.t5_layers(feats).mean(dim=1)

    def get_logit_scale(self):
        return self.clip_model.logit_scale.exp()

    def compute_inbatch_contrastive_loss(self, batch):
        emb = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"],
                                           batch["image_mask_batched"])
        return contrastive_loss_from_embeddings(emb, batch["index_mapping"], self.get_logit_scale(),
                                                in_batch_neg_num=self.in_batch_neg_num or 0,
                                                all_p_fn=self._gather if self.gather_embeddings else None, rank=self.rank)

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        assert id_list is not None, "id_list must be provided."
        assert isinstance(id_list[0], int), "id_list must be hashed to int."
        emb = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"],
                                           batch["image_mask_batched"])
        assert emb.size(0) == len(id_list), "embeddings and id_batched must have the same batch size."
        return emb, id_list

    def forward(self, batch, encode_mbeir_batch=False):
        return self.encode_mbeir_batch(batch) if encode_mbeir_batch else self.compute_inbatch_contrastive_loss(batch)
