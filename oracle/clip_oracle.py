"""fp32 CPU restatement of the OpenAI-CLIP towers and of UniIR's CLIPScoreFusion.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  * vision tower : upstream ``clip.model.VisionTransformer.forward`` as restated in the
    reference at src/models/uniir_clip/clip_featurefusion/clip_ff.py:35-59 (we pool
    ``x[:, 0, :]`` before ``ln_post`` -- the commented-out line clip_ff.py:54 -- which is
    what ``clip_model.encode_image`` does, called at clip_sf.py:46-47).
  * text tower   : upstream ``CLIP.encode_text``; trunk restated at clip_ff.py:148-156,
    pooling at ``argmax(token_ids)`` + ``@ text_projection`` (called at clip_sf.py:43-44).
  * score fusion : clip_sf.py:49-63.
  * in-batch loss: clip_sf.py:68-147 (incl. hard-negative branch :105-131).
  * embed API    : clip_sf.py:149-168.

State-dict key names are the upstream ones so that reference
checkpoints (prefix ``clip_model.``) load into both the oracle and the CUDA model.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import nn


@dataclass(frozen=True)
class CLIPDims:
    image_size: int
    patch: int
    v_width: int
    v_layers: int
    v_heads: int
    t_width: int
    t_layers: int
    t_heads: int
    embed_dim: int
    context: int = 77
    vocab: int = 49408

    @property
    def v_tokens(self) -> int:
        return (self.image_size // self.patch) ** 2 + 1


CLIP_CONFIGS = {
    # dims per clip_ff.py:114-131 + upstream
    "ViT-B/32": CLIPDims(224, 32, 768, 12, 12, 512, 12, 8, 512),
    "ViT-L/14": CLIPDims(224, 14, 1024, 24, 16, 768, 12, 12, 768),
    # tiny config used only by unit tests (same code path, seconds on CPU)
    "tiny-test": CLIPDims(56, 14, 128, 2, 2, 64, 2, 1, 64, context=77, vocab=512),
}


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


class _LN(nn.LayerNorm):
    """upstream LayerNorm subclass: computes in fp32, eps 1e-5."""

    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class _Attn(nn.Module):
    """nn.MultiheadAttention parameters (in_proj_weight/in_proj_bias/out_proj) with the
    math written out; batch-first [B, S, D]."""

    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)

    def forward(self, x, mask=None):
        B, S, D = x.shape
        H = self.heads
        qkv = F.linear(x, self.in_proj_weight, self.in_proj_bias)
        q, k, v = qkv.view(B, S, 3, H, D // H).permute(2, 0, 3, 1, 4)
        att = (q @ k.transpose(-1, -2)) / math.sqrt(D // H)
        if mask is not None:
            att = att + mask
        att = att.softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, D)
        return self.out_proj(o)


class _MLP(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.c_fc = nn.Linear(width, 4 * width)
        self.c_proj = nn.Linear(4 * width, width)

    def forward(self, x):
        return self.c_proj(quick_gelu(self.c_fc(x)))


class _Block(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.attn = _Attn(width, heads)
        self.ln_1 = _LN(width)
        self.mlp = _MLP(width)
        self.ln_2 = _LN(width)

    def forward(self, x, mask=None):
        x = x + self.attn(self.ln_1(x), mask)
        x = x + self.mlp(self.ln_2(x))
        return x


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.ModuleList([_Block(width, heads) for _ in range(layers)])

    def forward(self, x, mask=None):
        for blk in self.resblocks:
            x = blk(x, mask)
        return x


class _Visual(nn.Module):
    def __init__(self, d: CLIPDims):
        super().__init__()
        self.conv1 = nn.Conv2d(3, d.v_width, d.patch, d.patch, bias=False)
        self.class_embedding = nn.Parameter(torch.empty(d.v_width))
        self.positional_embedding = nn.Parameter(torch.empty(d.v_tokens, d.v_width))
        self.ln_pre = _LN(d.v_width)
        self.transformer = _Transformer(d.v_width, d.v_layers, d.v_heads)
        self.ln_post = _LN(d.v_width)
        self.proj = nn.Parameter(torch.empty(d.v_width, d.embed_dim))

    def forward(self, img):
        x = self.conv1(img)  # [B, width, g, g]
        x = x.flatten(2).transpose(1, 2)  # [B, g*g, width]
        cls = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding
        x = self.ln_pre(x)
        x = self.transformer(x)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class CLIPOracle(nn.Module):
    """The object ``clip.load`` returns (upstream ``clip.model.CLIP``), restated."""

    def __init__(self, name="ViT-B/32"):
        super().__init__()
        d = CLIP_CONFIGS[name]
        self.dims = d
        self.visual = _Visual(d)
        self.transformer = _Transformer(d.t_width, d.t_layers, d.t_heads)
        self.token_embedding = nn.Embedding(d.vocab, d.t_width)
        self.positional_embedding = nn.Parameter(torch.empty(d.context, d.t_width))
        self.ln_final = _LN(d.t_width)
        self.text_projection = nn.Parameter(torch.empty(d.t_width, d.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        mask = torch.full((d.context, d.context), float("-inf")).triu_(1)
        self.register_buffer("attn_mask", mask, persistent=False)
        self.init_parameters()

    def init_parameters(self, generator=None):
        """upstream ``initialize_parameters`` scales."""
        d = self.dims
        g = generator

        def n_(t, std):
            with torch.no_grad():
                t.normal_(0.0, std, generator=g)

        n_(self.token_embedding.weight, 0.02)
        n_(self.positional_embedding, 0.01)
        for trunk, width, layers in ((self.transformer, d.t_width, d.t_layers),
                                     (self.visual.transformer, d.v_width, d.v_layers)):
            proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
            for blk in trunk.resblocks:
                n_(blk.attn.in_proj_weight, width ** -0.5)
                n_(blk.attn.out_proj.weight, proj_std)
                n_(blk.mlp.c_fc.weight, (2 * width) ** -0.5)
                n_(blk.mlp.c_proj.weight, proj_std)
                for b in (blk.attn.in_proj_bias, blk.attn.out_proj.bias, blk.mlp.c_fc.bias,
                          blk.mlp.c_proj.bias):
                    n_(b, 0.01)  # upstream leaves zeros; non-zero exercises the bias paths
        n_(self.text_projection, d.t_width ** -0.5)
        sc = d.v_width ** -0.5
        n_(self.visual.class_embedding, sc)
        n_(self.visual.positional_embedding, sc)
        n_(self.visual.proj, sc)
        n_(self.visual.conv1.weight, (3 * d.patch * d.patch) ** -0.5)
        with torch.no_grad():
            self.logit_scale.fill_(math.log(100.0))  # pretrained value

    def encode_image(self, img):
        return self.visual(img.float())

    def encode_text(self, ids):
        x = self.token_embedding(ids.long()) + self.positional_embedding
        x = self.transformer(x, self.attn_mask)
        x = self.ln_final(x)
        eot = ids.argmax(dim=-1)
        x = x[torch.arange(x.shape[0]), eot]
        return x @ self.text_projection


class CLIPScoreFusionOracle(nn.Module):
    """Restatement of reference ``CLIPScoreFusion`` (clip_sf.py:13-168) minus ``clip.load``."""

    def __init__(self, model_name="ViT-B/32", config=None, rank=0, world=1, gather_fn=None):
        super().__init__()
        self.clip_model = CLIPOracle(model_name)
        self.gather_embeddings = False
        self.in_batch_neg_num = 0
        if config is not None:  # clip_sf.py:28-30
            self.gather_embeddings = config.model.gather_embeddings
            self.in_batch_neg_num = config.data_config.in_batch_neg_num
        self._rank, self._world, self._gather_fn = rank, world, gather_fn

    # clip_sf.py:53-63
    def encode_multimodal_input(self, txt, img, txt_mask, img_mask):
        t = self.clip_model.encode_text(txt) * txt_mask.unsqueeze(-1)
        v = self.clip_model.encode_image(img) * img_mask.unsqueeze(-1)
        return v + t

    # clip_sf.py:68-147
    def compute_inbatch_contrastive_loss(self, batch):
        im = batch["index_mapping"]
        emb = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"],
                                           batch["txt_mask_batched"], batch["image_mask_batched"])
        return contrastive_loss_from_embeddings(
            emb, im, self.clip_model.logit_scale.exp(), self.in_batch_neg_num,
            all_p_fn=(self._gather_fn if self.gather_embeddings else None), rank=self._rank)

    def encode_mbeir_batch(self, batch):  # clip_sf.py:154-168
        id_list = batch.get("did_list") or batch.get("qid_list")
        assert id_list is not None, "id_list must be provided."
        assert isinstance(id_list[0], int), "id_list must be hashed to int."
        emb = self.encode_multimodal_input(batch["txt_batched"], batch["image_batched"],
                                           batch["txt_mask_batched"], batch["image_mask_batched"])
        assert emb.size(0) == len(id_list)
        return emb, id_list

    def forward(self, batch, encode_mbeir_batch=False):
        if encode_mbeir_batch:
            return self.encode_mbeir_batch(batch)
        return self.compute_inbatch_contrastive_loss(batch)


def contrastive_loss_from_embeddings(emb, index_mapping, logit_scale, in_batch_neg_num=0,
                                     all_p_fn=None, rank=0):
    """clip_sf.py:88-147 on already-computed fused embeddings.

    ``all_p_fn(p) -> all_p`` plays the role of ``torch.cat(torch.distributed.nn.all_gather(p))``
    (clip_sf.py:103); ``None`` means gather_embeddings=False.
    """
    q = emb[torch.tensor(index_mapping["query"]).flatten()]
    p = emb[torch.tensor(index_mapping["pos_cand"]).flatten()]
    hard = "neg_cand_list" in index_mapping
    bs = q.size(0)
    q = F.normalize(q, dim=-1)
    p = F.normalize(p, dim=-1)
    if hard:
        n = F.normalize(emb[torch.tensor(index_mapping["neg_cand_list"])], dim=-1)
        k = min(bs - 1, in_batch_neg_num)
        off = torch.eye(bs) == 0
        inb = p.unsqueeze(1).expand(-1, bs, -1)
        # row i of `inb[off]` = all p_j with j != i ... NB reference indexes expand(-1,bs,-1)
        # whose [i, j] entry is p_i; masked-select keeps (i, j!=i) -> p_i repeated bs-1
        # times (clip_sf.py:113-115).  We restate that arithmetic faithfully.
        inb = inb[off].reshape(bs, bs - 1, -1)[:, :k, :]
        aug = torch.cat([n, inb], dim=1)
        pos = (q * p).sum(-1) * logit_scale
        neg = (q.unsqueeze(1) * aug).sum(-1) * logit_scale
        logits = torch.cat([pos.unsqueeze(-1), neg], 1)
        loss = torch.mean(-F.log_softmax(logits, dim=1)[:, 0])
        acc = (logits.max(1)[1] == 0).sum() / bs
    else:
        if all_p_fn is not None:
            all_p = all_p_fn(p)
            target = rank * bs + torch.arange(bs)
        else:
            all_p = p
            target = torch.arange(bs)
        score = q @ all_p.t() * logit_scale
        loss = F.cross_entropy(score, target)
        acc = (score.max(1)[1] == target).sum() / bs
    return {"loss": loss, "accuracy": acc}
