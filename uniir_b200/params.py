"""Flat parameter storage with the upstream OpenAI-CLIP state-dict layout.

All parameters live in two contiguous fp32 buffers (weight-decay group / no-decay group, the
split of reference clip_scorefusion/train.py:52-61) so that the optimizer step, the gradient
all-reduce and the bf16 shadow refresh are each ONE kernel / collective over a flat segment.
``nn.Parameter`` objects registered under the upstream names are views
into those buffers, so reference checkpoints (``clip_model.*`` keys, clip_sf.py:25) load with a
plain ``load_state_dict`` and ``state_dict()`` round-trips.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
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
    def grid(self):
        return self.image_size // self.patch

    @property
    def v_tokens(self):
        return self.grid ** 2 + 1

    @property
    def patch_k(self):
        return 3 * self.patch * self.patch

    @property
    def patch_k_pad(self):
        return (self.patch_k + 63) // 64 * 64


CLIP_CONFIGS = {
    "ViT-B/32": CLIPDims(224, 32, 768, 12, 12, 512, 12, 8, 512),
    "ViT-L/14": CLIPDims(224, 14, 1024, 24, 16, 768, 12, 12, 768),
    "tiny-test": CLIPDims(56, 14, 128, 2, 2, 64, 2, 1, 64, context=77, vocab=512),
}


def clip_param_specs(d: CLIPDims):
    """[(name, shape, decay, init_std)] in upstream key order; decay follows the reference
    rule: no weight decay for ndim < 2, 'bn', 'ln', 'bias', 'logit_scale' (train.py:52-61)."""
    specs = []

    def add(name, shape, std):
        nd = len(shape)
        decay = not (nd < 2 or "bn" in name or "ln" in name or "bias" in name or "logit_scale" in name)
        specs.append((name, tuple(shape), decay, std))

    def blocks(prefix, width, layers):
        attn_std = width ** -0.5
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        fc_std = (2 * width) ** -0.5
        for i in range(layers):
            p = f"{prefix}.resblocks.{i}"
            add(f"{p}.attn.in_proj_weight", (3 * width, width), attn_std)
            add(f"{p}.attn.in_proj_bias", (3 * width,), 0.0)
            add(f"{p}.attn.out_proj.weight", (width, width), proj_std)
            add(f"{p}.attn.out_proj.bias", (width,), 0.0)
            add(f"{p}.ln_1.weight", (width,), -1.0)
            add(f"{p}.ln_1.bias", (width,), 0.0)
            add(f"{p}.mlp.c_fc.weight", (4 * width, width), fc_std)
            add(f"{p}.mlp.c_fc.bias", (4 * width,), 0.0)
            add(f"{p}.mlp.c_proj.weight", (width, 4 * width), proj_std)
            add(f"{p}.mlp.c_proj.bias", (width,), 0.0)
            add(f"{p}.ln_2.weight", (width,), -1.0)
            add(f"{p}.ln_2.bias", (width,), 0.0)

    vs = d.v_width ** -0.5
    add("positional_embedding", (d.context, d.t_width), 0.01)
    add("text_projection", (d.t_width, d.embed_dim), d.t_width ** -0.5)
    add("logit_scale", (), -2.0)
    add("visual.class_embedding", (d.v_width,), vs)
    add("visual.positional_embedding", (d.v_tokens, d.v_width), vs)
    add("visual.proj", (d.v_width, d.embed_dim), vs)
    add("visual.conv1.weight", (d.v_width, 3, d.patch, d.patch), d.patch_k ** -0.5)
    add("visual.ln_pre.weight", (d.v_width,), -1.0)
    add("visual.ln_pre.bias", (d.v_width,), 0.0)
    blocks("visual.transformer", d.v_width, d.v_layers)
    add("visual.ln_post.weight", (d.v_width,), -1.0)
    add("visual.ln_post.bias", (d.v_width,), 0.0)
    blocks("transformer", d.t_width, d.t_layers)
    add("token_embedding.weight", (d.vocab, d.t_width), 0.02)
    add("ln_final.weight", (d.t_width,), -1.0)
    add("ln_final.bias", (d.t_width,), 0.0)
    return specs


class FlatParamModule(nn.Module):
    """nn.Module tree whose Parameters are views into flat fp32 buffers."""

    ALIGN = 64  # elements; keeps every parameter 256 B aligned (TMA needs 16 B)

    def __init__(self, specs, device, seed=2023, root=None):
        """root: module on which the parameter tree is built (default: this module).  With a separate
        root the Parameters appear under the root's own state-dict keys (BLIP: ``visual_encoder.*``)."""
        super().__init__()
        object.__setattr__(self, "_root", root if root is not None else self)
        self._specs = specs
        self._index = {}
        sizes = {True: 0, False: 0}
        for name, shape, decay, _ in specs:
            n = int(math.prod(shape)) if shape else 1
            self._index[name] = (decay, sizes[decay], n, shape)
            sizes[decay] += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.flat = {k: torch.zeros(max(v, self.ALIGN), device=device, dtype=torch.float32) for k, v in sizes.items()}
        self.flat16 = {k: torch.zeros_like(v, dtype=torch.bfloat16) for k, v in self.flat.items()}
        self.grad = None   # allocated lazily (training only)
        self._shadow_version = {True: -1, False: -1}
        gen = torch.Generator(device="cpu").manual_seed(seed)
        for name, shape, decay, std in specs:
            view = self._view(self.flat, name)
            if std == -1.0:
                view.fill_(1.0)
            elif std == -2.0:
                view.fill_(math.log(100.0))  # pretrained logit_scale
            elif std == -3.0:
                view.fill_(0.07)             # BLIP temperature (blip_ff.py:61)
            elif std > 0:
                view.copy_(torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=gen))
            self._register(name, nn.Parameter(view))

    # -- tree plumbing -------------------------------------------------------------
    def _register(self, dotted, param):
        mod = self._root
        parts = dotted.split(".")
        for part in parts[:-1]:
            if not hasattr(mod, part):
                mod.add_module(part, nn.Module())
            mod = getattr(mod, part)
        mod.register_parameter(parts[-1], param)

    def _view(self, bufs, name):
        decay, off, n, shape = self._index[name]
        return bufs[decay][off:off + n].view(shape)

    def w(self, name):
        """fp32 master view"""
        return self._view(self.flat, name)

    def w16(self, name):
        """bf16 shadow view (what the GEMMs read)"""
        return self._view(self.flat16, name)

    def g(self, name):
        return self._view(self.grad, name)

    def ensure_grad(self):
        if self.grad is None:
            self.grad = {k: torch.zeros_like(v) for k, v in self.flat.items()}
        return self.grad

    def rebind_after_apply(self):
        """``module.to()/float()`` may have replaced Parameter storage; re-point the flats."""
        for name, shape, decay, _ in self._specs:
            p = self._root.get_parameter(name)
            view = self._view(self.flat, name)
            if p.data_ptr() != view.data_ptr():
                view.copy_(p.data.to(view.device, torch.float32))
                p.data = view

    def refresh_shadow(self, force=False):
        """bf16 shadow <- fp32 master, if anything touched the masters since last time."""
        from . import ops
        for k, buf in self.flat.items():
            if force or buf._version != self._shadow_version[k]:
                ops.cast_bf16(buf, out=self.flat16[k])
                self._shadow_version[k] = buf._version

    def mark_shadow_fresh(self):
        for k, buf in self.flat.items():
            self._shadow_version[k] = buf._version
