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


# Dropout sites of the reference in train mode: med.py:97 (embeddings, after the LayerNorm), :217 (attention
# probabilities, self and cross), :244 and :325 (dense outputs, before the residual + LayerNorm), vit.py:151-152 (timm
# DropPath on both residual branches; rate linspace(0, 0.1, depth)[layer] for ViT-L, 0 for ViT-B: blip.py:241,253).
# The restatement is the eval-mode arithmetic (identity at every site) unless a test installs DROPOUT, a callable
# (tag, kind, layer, x) -> x with tag "" / "_m" (online / momentum encoders) and kind one of "emb", "self_probs",
# "self_out", "cross_probs", "cross_out", "ffn_out", "path_attn", "path_mlp": the masks then come from the caller
# (the reference's own come from torch's RNG stream; tests/test_oracle_blip.py pins the SITES against the reference's
# modules by handing both the same masks).
DROPOUT = None
_CTX = {"tag": "", "layer": 0}


def _drop(kind, x):
    return x if DROPOUT is None else DROPOUT(_CTX["tag"], kind, _CTX["layer"], x)


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
        x = x + _drop("path_attn", self.attn(self.norm1(x)))
        return x + _drop("path_mlp", self.mlp(self.norm2(x)))


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
        for i, b in enumerate(self.blocks):
            _CTX["layer"] = i
            x = b(x)
        return self.norm(x)


class _SelfAttn(nn.Module):
    def __init__(self, hidden, heads, kv_width):
        super().__init__()
        self.heads = heads
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_width, hidden)
        self.value = nn.Linear(kv_width, hidden)

    def forward(self, h, kv, add_mask, kind="self"):
        B, S, D = h.shape
        H = self.heads
        sp = lambda t: t.view(B, -1, H, D // H).permute(0, 2, 1, 3)
        q, k, v = sp(self.query(h)), sp(self.key(kv)), sp(self.value(kv))
        s = q @ k.transpose(-1, -2) / math.sqrt(D // H)
        if add_mask is not None:
            s = s + add_mask
        return (_drop(kind + "_probs", s.softmax(-1)) @ v).permute(0, 2, 1, 3).reshape(B, S, D)


class _AttnOut(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=1e-12)

    def forward(self, ctx, inp, kind="self"):
        return self.LayerNorm(_drop(kind + "_out", self.dense(ctx)) + inp)


class _BertAttention(nn.Module):
    def __init__(self, hidden, heads, kv_width):
        super().__init__()
        self.self = _SelfAttn(hidden, heads, kv_width)
        self.output = _AttnOut(hidden)

    def forward(self, h, kv, mask, kind="self"):
        return self.output(self.self(h, kv, mask, kind), h, kind)


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
        if enc is not None:          # mode="multimodal"; mode="text" skips the cross-attention (med.py:368-383)
            a = self.crossattention(a, enc, enc_mask, "cross")
        return self.output.LayerNorm(_drop("ffn_out", self.output.dense(F.gelu(self.intermediate.dense(a)))) + a)


class _Embeddings(nn.Module):
    def __init__(self, d: BLIPDims):
        super().__init__()
        self.word_embeddings = nn.Embedding(d.vocab, d.t_width, padding_idx=0)
        self.position_embeddings = nn.Embedding(d.max_pos, d.t_width)
        self.LayerNorm = nn.LayerNorm(d.t_width, eps=1e-12)
        self.register_buffer("position_ids", torch.arange(d.max_pos).expand((1, -1)))

    def forward(self, ids):
        S = ids.shape[1]
        return _drop("emb", self.LayerNorm(self.word_embeddings(ids) + self.position_embeddings(self.position_ids[:, :S])))


class _Encoder(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(d) for _ in range(d.t_layers)])


class BLIPBertOracle(nn.Module):
    def __init__(self, d: BLIPDims, add_pooling_layer=True):
        super().__init__()
        self.embeddings = _Embeddings(d)
        self.encoder = _Encoder(d)
        if add_pooling_layer:
            self.pooler = _Dense(d.t_width, d.t_width)

    def hidden(self, input_ids, attention_mask, encoder_hidden_states=None):
        """last_hidden_state; encoder_hidden_states=None is the reference's mode="text" (no cross-attention)."""
        mask = (1.0 - attention_mask[:, None, None, :].float()) * -10000.0        # med.py:687-688
        _CTX["layer"] = 0
        h = self.embeddings(input_ids.long())
        for i, lyr in enumerate(self.encoder.layer):
            _CTX["layer"] = i
            h = lyr(h, mask, encoder_hidden_states, None)   # image mask is all ones (blip_ff.py:107)
        return h

    def forward(self, input_ids, attention_mask, encoder_hidden_states):
        return torch.tanh(self.pooler.dense(self.hidden(input_ids, attention_mask, encoder_hidden_states)[:, 0]))


class BLIPFeatureFusionOracle(nn.Module):
    """Reference BLIPFeatureFusion restated: encode path (blip_ff.py:82-116, 259-287), momentum encoders
    and queues (:41-63), contrastive loss with momentum distillation (:118-257), EMA (:289-292) and queue
    update (:295-310), including the hard-negative variant (:126-131, 160-171, 208-218, 233-246: the negatives'
    momentum features -- NOT normalised in the reference -- and ids replace the first bs*neg_num queue columns; a
    coin flip decides whether the positives or each query's first negative are enqueued)."""

    def __init__(self, size="large", embed_dim=None, queue_size=57600, momentum=0.995, dims=None, gather_fn=None):
        super().__init__()
        self.dims = dims if dims is not None else BLIP_CONFIGS[size]
        embed_dim = embed_dim or self.dims.t_width
        self.visual_encoder = BLIPViTOracle(self.dims)
        self.text_encoder = BLIPBertOracle(self.dims)
        for p in self.parameters():
            nn.init.normal_(p, 0.0, 0.02)
        for m in self.modules():
            if isinstance(m, nn.LayerNorm):
                nn.init.normal_(m.weight, 1.0, 0.05)
        self.visual_encoder_m = BLIPViTOracle(self.dims)
        self.text_encoder_m = BLIPBertOracle(self.dims)
        self.model_pairs = [[self.visual_encoder, self.visual_encoder_m], [self.text_encoder, self.text_encoder_m]]
        self.copy_params()
        self.register_buffer("query_queue", F.normalize(torch.randn(embed_dim, queue_size), dim=0))
        self.register_buffer("cand_queue", F.normalize(torch.randn(embed_dim, queue_size), dim=0))
        self.register_buffer("idx_queue", torch.full((1, queue_size), -100))
        self.register_buffer("new_ptr_queue", torch.zeros(1, dtype=torch.long))
        self.queue_size, self.momentum = queue_size, momentum
        self.temp = nn.Parameter(0.07 * torch.ones([]))
        self._gather = gather_fn or (lambda t: t)   # concat_all_gather (:323-332); identity for world size 1

    @torch.no_grad()
    def copy_params(self):
        for a, b in self.model_pairs:
            for p, pm in zip(a.parameters(), b.parameters()):
                pm.data.copy_(p.data)
                pm.requires_grad = False

    @torch.no_grad()
    def _momentum_update(self):
        for a, b in self.model_pairs:
            for p, pm in zip(a.parameters(), b.parameters()):
                pm.data = pm.data * self.momentum + p.data * (1.0 - self.momentum)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, q, c, idxs):
        idxs, q, c = self._gather(idxs), self._gather(q), self._gather(c)
        n = q.shape[0]
        ptr = int(self.new_ptr_queue)
        assert self.queue_size % n == 0
        self.query_queue[:, ptr:ptr + n] = q.T
        self.cand_queue[:, ptr:ptr + n] = c.T
        self.idx_queue[:, ptr:ptr + n] = idxs.T
        self.new_ptr_queue[0] = (ptr + n) % self.queue_size

    def encode_multimodal_input(self, input_ids, attention_mask, image, txt_mask=None, img_mask=None, use_momentum=False):
        v, t = (self.visual_encoder_m, self.text_encoder_m) if use_momentum else (self.visual_encoder, self.text_encoder)
        _CTX["tag"] = "_m" if use_momentum else ""
        return t(input_ids, attention_mask, v(image.float()))

    def compute_contrastive_loss(self, batch, alpha):
        t = batch["txt_batched"]
        ids, att, img = t["input_ids"], t["attention_mask"], batch["image_batched"]
        pc_idx = torch.tensor(batch["p_did_list"]).view(-1, 1)
        im = batch["index_mapping"]
        hard = "neg_cand_list" in im
        if hard:
            nc_idx = torch.tensor(batch["nc_dids_list"]).view(-1, 1)          # [bs * neg_num, 1]
            hard_n = nc_idx.size(0)
        with torch.no_grad():
            self.temp.clamp_(0.001, 0.5)
        mt, mv = batch.get("txt_mask_batched"), batch.get("image_mask_batched")   # ignored by FF, applied by SF
        emb = self.encode_multimodal_input(ids, att, img, mt, mv)
        qi = torch.tensor(im["query"]).flatten()
        pi = torch.tensor(im["pos_cand"]).flatten()
        q, pc = F.normalize(emb[qi], dim=-1), F.normalize(emb[pi], dim=-1)
        if hard:
            idx_all = torch.cat([pc_idx.t(), nc_idx.t(), self.idx_queue.clone()[:, hard_n:]], dim=1)
        else:
            idx_all = torch.cat([pc_idx.t(), self.idx_queue.clone()], dim=1)
        pos = torch.eq(pc_idx, idx_all).float()
        sim_targets = pos / pos.sum(1, keepdim=True)
        with torch.no_grad():
            self._momentum_update()
            emb_m = self.encode_multimodal_input(ids, att, img, mt, mv, use_momentum=True)
            q_m, pc_m = F.normalize(emb_m[qi], dim=-1), F.normalize(emb_m[pi], dim=-1)
            q_all = torch.cat([q_m.t(), self.query_queue.clone()], dim=1)
            if hard:
                nc_m = emb_m[torch.tensor(im["neg_cand_list"])]                # [bs, neg_num, E], un-normalised (:196)
                pc_all = torch.cat([pc_m.t(), nc_m.view(hard_n, -1).t(), self.cand_queue.clone()[:, hard_n:]], dim=1)
            else:
                pc_all = torch.cat([pc_m.t(), self.cand_queue.clone()], dim=1)
            t_q2pc = alpha * F.softmax(q_m @ pc_all / self.temp, dim=1) + (1 - alpha) * sim_targets
            t_pc2q = alpha * F.softmax(pc_m @ q_all / self.temp, dim=1) + (1 - alpha) * sim_targets
        s_q2pc = q @ pc_all / self.temp
        s_pc2q = pc @ q_all / self.temp
        loss = (-(F.log_softmax(s_q2pc, dim=1) * t_q2pc).sum(1).mean() - (F.log_softmax(s_pc2q, dim=1) * t_pc2q).sum(1).mean()) / 2
        if hard and not bool(torch.rand(1) < 0.5):                             # :233-244
            self._dequeue_and_enqueue(q_m, nc_m[:, 0, :].contiguous(), nc_idx.view(q.size(0), -1)[:, 0].contiguous())
        else:
            self._dequeue_and_enqueue(q_m, pc_m, pc_idx)
        acc = pos.gather(1, s_q2pc.max(1)[1].unsqueeze(1)).squeeze().mean()
        return {"loss": loss, "accuracy": acc}

    def encode_mbeir_batch(self, batch):   # blip_ff.py:264-287 (ValueError, not assert)
        id_list = batch.get("did_list") or batch.get("qid_list")
        if id_list is None:
            raise ValueError("id_list not found in batch.")
        t = batch["txt_batched"]
        emb = self.encode_multimodal_input(t["input_ids"], t["attention_mask"], batch["image_batched"])
        assert emb.size(0) == len(id_list)
        return emb, id_list

    def forward(self, batch, alpha=None, encode_mbeir_batch=False):
        return self.encode_mbeir_batch(batch) if encode_mbeir_batch else self.compute_contrastive_loss(batch, alpha)


class BLIPScoreFusionOracle(BLIPFeatureFusionOracle):
    """Reference BLIPScoreFusion restated (blip_scorefusion/blip_sf.py:11-332): ViT class token -> vision_proj,
    BERT in mode="text" (cross-attention layers present but skipped and frozen, no pooler) [CLS] -> text_proj,
    each multiplied by its modality mask and summed (:97-172).  Momentum encoders, queues, loss, EMA and queue
    update are the BLIP_FF ones (:174-330 is the same code), inherited."""

    def __init__(self, size="large", embed_dim=768, queue_size=57600, momentum=0.995, dims=None, gather_fn=None):
        nn.Module.__init__(self)
        self.dims = dims if dims is not None else BLIP_CONFIGS[size]
        d = self.dims
        self.visual_encoder = BLIPViTOracle(d)
        self.text_encoder = BLIPBertOracle(d, add_pooling_layer=False)
        self.vision_proj = nn.Linear(d.v_width, embed_dim)
        self.text_proj = nn.Linear(d.t_width, embed_dim)
        for p in self.parameters():
            nn.init.normal_(p, 0.0, 0.02)
        for m in self.modules():
            if isinstance(m, nn.LayerNorm):
                nn.init.normal_(m.weight, 1.0, 0.05)
        for n, p in self.text_encoder.named_parameters():      # blip_sf.py:75-78
            if "crossattention" in n:
                p.requires_grad = False
        self.visual_encoder_m = BLIPViTOracle(d)
        self.vision_proj_m = nn.Linear(d.v_width, embed_dim)
        self.text_encoder_m = BLIPBertOracle(d, add_pooling_layer=False)
        self.text_proj_m = nn.Linear(d.t_width, embed_dim)
        self.model_pairs = [[self.visual_encoder, self.visual_encoder_m], [self.vision_proj, self.vision_proj_m],
                            [self.text_encoder, self.text_encoder_m], [self.text_proj, self.text_proj_m]]
        self.copy_params()
        self.register_buffer("query_queue", F.normalize(torch.randn(embed_dim, queue_size), dim=0))
        self.register_buffer("cand_queue", F.normalize(torch.randn(embed_dim, queue_size), dim=0))
        self.register_buffer("idx_queue", torch.full((1, queue_size), -100))
        self.register_buffer("new_ptr_queue", torch.zeros(1, dtype=torch.long))
        self.queue_size, self.momentum = queue_size, momentum
        self.temp = nn.Parameter(0.07 * torch.ones([]))
        self._gather = gather_fn or (lambda t: t)

    def encode_multimodal_input(self, input_ids, attention_mask, image, txt_mask=None, img_mask=None, use_momentum=False):
        v, vp, t, tp = ((self.visual_encoder_m, self.vision_proj_m, self.text_encoder_m, self.text_proj_m) if use_momentum
                        else (self.visual_encoder, self.vision_proj, self.text_encoder, self.text_proj))
        _CTX["tag"] = "_m" if use_momentum else ""
        txt = tp(t.hidden(input_ids, attention_mask)[:, 0, :]) * txt_mask.unsqueeze(-1)        # blip_sf.py:97-121,168
        img = vp(v(image.float())[:, 0, :]) * img_mask.unsqueeze(-1)                           # :123-143,169
        return img + txt                                                                       # :145-156

    def encode_mbeir_batch(self, batch):
        id_list = batch.get("did_list") or batch.get("qid_list")
        if id_list is None:
            raise ValueError("id_list not found in batch.")
        t = batch["txt_batched"]
        emb = self.encode_multimodal_input(t["input_ids"], t["attention_mask"], batch["image_batched"],
                                           batch["txt_mask_batched"], batch["image_mask_batched"])
        assert emb.size(0) == len(id_list)
        return emb, id_list
