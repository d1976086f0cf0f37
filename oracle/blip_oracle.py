"""fp32 CPU restatement of the BLIP feature-fusion encoder (reference BLIPFeatureFusion.
encode_multimodal_input, src/models/uniir_blip/blip_featurefusion/blip_ff.py:82-116).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the vendored reference code, whose arithmetic IS in the tree but is written against
transformers 4.15 / timm / fairscale (not importable on the GPU box):
  * ViT (timm style)  : uniir_blip/backbone/vit.py:52-106 (Attention), :109-153 (Block, pre-LN,
    eps 1e-6 via blip.py:243-253 `partial(nn.LayerNorm, eps=1e-6)`, exact GELU), :254-268 (forward:
    conv patch-embed WITH bias, cls token, +pos_embed, blocks, final norm, all tokens returned)
  * BERT "MED"        : uniir_blip/backbone/med.py:52-98 (embeddings: word + position -> LN 1e-12),
    :101-232 (self / cross attention; additive mask (1-m)*-10000 from :687-688), :235-246 and
    :313-327 (dense + residual + post-LN), :330-396 (layer: self-attn -> cross-attn -> FFN, mode
    "multimodal"), :498-511 (pooler tanh(W h[:,0] + b))
State-dict keys are the reference's.  tests/test_oracle_blip.py pins this
restatement against the shimmed import of the reference modules (only possible where
/root/reference exists) and commits golden vectors for the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn


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
    def v_tokens(self):
        return (self.image_size // self.patch) ** 2 + 1


BLIP_CONFIGS = {
    "large": BLIPDims(),                                                       # blip.py:247-253
    "base": BLIPDims(v_width=768, v_layers=12, v_heads=12),                    # blip.py:235-245
    "tiny-test": BLIPDims(image_size=64, patch=16, v_width=128, v_layers=2, v_heads=2, t_width=64, t_layers=2,
                          t_heads=1, t_inter=256, vocab=600, max_pos=128),
}


class _VitAttn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.heads, C // self.heads).permute(2, 0, 3, 1, 4)
        a = ((q @ k.transpose(-2, -1)) * (C // self.heads) ** -0.5).softmax(-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, N, C))


class _VitMlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _VitBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _VitAttn(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _VitMlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.proj = nn.Conv2d(3, d.v_width, d.patch, d.patch)


class BLIPViTOracle(nn.Module):
    def __init__(self, d: BLIPDims):
        super().__init__()
        self.patch_embed = _PatchEmbed(d)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, d.v_width))
        self.pos_embed = nn.Parameter(torch.zeros(1, d.v_tokens, d.v_width))
        self.blocks = nn.ModuleList([_VitBlock(d.v_width, d.v_heads) for _ in range(d.v_layers)])
        self.norm = nn.LayerNorm(d.v_width, eps=1e-6)

    def forward(self, img):
        x = self.patch_embed.proj(img).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], 1) + self.pos_embed[:, :x.shape[1] + 1]
        for b in self.blocks:
            x = b(x)
        return self.norm(x)


class _SelfAttn(nn.Module):
    def __init__(self, hidden, heads, kv_width):
        super().__init__()
        self.heads = heads
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_width, hidden)
        self.value = nn.Linear(kv_width, hidden)

    def forward(self, h, kv, add_mask):
        B, S, D = h.shape
        H = self.heads
        sp = lambda t: t.view(B, -1, H, D // H).permute(0, 2, 1, 3)
        q, k, v = sp(self.query(h)), sp(self.key(kv)), sp(self.value(kv))
        s = q @ k.transpose(-1, -2) / math.sqrt(D // H)
        if add_mask is not None:
            s = s + add_mask
        return (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, S, D)


class _AttnOut(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=1e-12)

    def forward(self, ctx, inp):
        return self.LayerNorm(self.dense(ctx) + inp)


class _BertAttention(nn.Module):
    def __init__(self, hidden, heads, kv_width):
        super().__init__()
        self.self = _SelfAttn(hidden, heads, kv_width)
        self.output = _AttnOut(hidden)

    def forward(self, h, kv, mask):
        return self.output(self.self(h, kv, mask), h)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _FfnOut(nn.Module):
    def __init__(self, inter, hidden):
        super().__init__()
        self.dense = nn.Linear(inter, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=1e-12)


class _BertLayer(nn.Module):
    def __init__(self, d: BLIPDims):
        super().__init__()
        self.attention = _BertAttention(d.t_width, d.t_heads, d.t_width)
        self.crossattention = _BertAttention(d.t_width, d.t_heads, d.v_width)
        self.intermediate = _Dense(d.t_width, d.t_inter)
        self.output = _FfnOut(d.t_inter, d.t_width)

    def forward(self, h, mask, enc, enc_mask):
        a = self.attention(h, h, mask)
        a = self.crossattention(a, enc, enc_mask)
        return self.output.LayerNorm(self.output.dense(F.gelu(self.intermediate.dense(a))) + a)


class _Embeddings(nn.Module):
    def __init__(self, d: BLIPDims):
        super().__init__()
        self.word_embeddings = nn.Embedding(d.vocab, d.t_width, padding_idx=0)
        self.position_embeddings = nn.Embedding(d.max_pos, d.t_width)
        self.LayerNorm = nn.LayerNorm(d.t_width, eps=1e-12)
        self.register_buffer("position_ids", torch.arange(d.max_pos).expand((1, -1)))

    def forward(self, ids):
        S = ids.shape[1]
        return self.LayerNorm(self.word_embeddings(ids) + self.position_embeddings(self.position_ids[:, :S]))


class _Encoder(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(d) for _ in range(d.t_layers)])


class BLIPBertOracle(nn.Module):
    def __init__(self, d: BLIPDims):
        super().__init__()
        self.embeddings = _Embeddings(d)
        self.encoder = _Encoder(d)
        self.pooler = _Dense(d.t_width, d.t_width)

    def forward(self, input_ids, attention_mask, encoder_hidden_states):
        mask = (1.0 - attention_mask[:, None, None, :].float()) * -10000.0        # med.py:687-688
        h = self.embeddings(input_ids.long())
        for lyr in self.encoder.layer:
            h = lyr(h, mask, encoder_hidden_states, None)   # image mask is all ones (blip_ff.py:107)
        return torch.tanh(self.pooler.dense(h[:, 0]))


class BLIPFeatureFusionOracle(nn.Module):
    """encode path of reference BLIPFeatureFusion (blip_ff.py:82-116, 259-287)."""

    def __init__(self, size="large"):
        super().__init__()
        self.dims = BLIP_CONFIGS[size]
        self.visual_encoder = BLIPViTOracle(self.dims)
        self.text_encoder = BLIPBertOracle(self.dims)
        for p in self.parameters():
            if p.ndim >= 2:
                nn.init.normal_(p, 0.0, 0.02)
            else:
                nn.init.normal_(p, 0.0, 0.02)
        for m in self.modules():
            if isinstance(m, nn.LayerNorm):
                nn.init.normal_(m.weight, 1.0, 0.05)

    def encode_multimodal_input(self, input_ids, attention_mask, image, txt_mask=None, img_mask=None):
        img = self.visual_encoder(image.float())
        return self.text_encoder(input_ids, attention_mask, img)

    def encode_mbeir_batch(self, batch):   # blip_ff.py:264-287 (ValueError, not assert)
        id_list = batch.get("did_list") or batch.get("qid_list")
        if id_list is None:
            raise ValueError("id_list not found in batch.")
        t = batch["txt_batched"]
        emb = self.encode_multimodal_input(t["input_ids"], t["attention_mask"], batch["image_batched"])
        assert emb.size(0) == len(id_list)
        return emb, id_list
