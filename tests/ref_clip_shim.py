"""Import the reference's UNMODIFIED CLIP model files (uniir_clip/clip_scorefusion/clip_sf.py and
clip_featurefusion/clip_ff.py) by putting a stand-in `clip` package into sys.modules.

The real `clip` (git+https://github.com/openai/CLIP.git, unpinned, src/models/uniir_env.yml:23) is not
installable here, so the stand-in supplies what the two reference files touch -- `clip.load`,
`clip.tokenize`, `clip.model.VisionTransformer`, and on the returned model `encode_text`, `encode_image`,
`logit_scale`, `visual`, `transformer` (LND layout), `token_embedding`, `positional_embedding`, `ln_final`,
`text_projection`, `dtype`, `float()` -- built from the tower blocks of oracle/clip_oracle.py (pinned
separately against transformers.CLIPModel, tests/test_oracle_clip.py) under the UPSTREAM state-dict
names.  Everything above the towers -- masked score fusion, the T5 fusion stack (the real
transformers T5Stack), the in-batch / hard-negative / gathered contrastive loss, the embed API -- is then
the reference's own code, executed verbatim.

Only usable where /root/reference exists (the build container); the GPU box uses the golden vectors that
scripts/make_golden.py writes from these runs."""
import importlib
import os
import sys
import types

import torch
from torch import nn

REF_SRC = "/root/reference/src"


def available():
    return os.path.isfile(os.path.join(REF_SRC, "models", "uniir_clip", "clip_scorefusion", "clip_sf.py"))


def _install_fake_clip():
    if "clip" in sys.modules and getattr(sys.modules["clip"], "_uniir_shim", False):
        return sys.modules["clip"]
    from oracle.clip_oracle import CLIP_CONFIGS, CLIPDims, _LN, _Transformer
    from uniir_b200.preprocess import synthetic_clip_tokens  # noqa: F401  (tokenizer stand-in lives with the tests)

    class Transformer(_Transformer):
        """upstream signature Transformer(width, layers, heads, attn_mask=None); input / output are LND."""

        def __init__(self, width, layers, heads, attn_mask=None):
            super().__init__(width, layers, heads)
            self.attn_mask = attn_mask

        def forward(self, x):
            mask = self.attn_mask.to(x.device) if self.attn_mask is not None else None
            return super().forward(x.permute(1, 0, 2), mask).permute(1, 0, 2)

    class VisionTransformer(nn.Module):
        """upstream constructor (clip_ff.py:23-33 subclasses it and calls it with these arguments)."""

        def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
            super().__init__()
            self.input_resolution, self.output_dim = input_resolution, output_dim
            self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
            scale = width ** -0.5
            self.class_embedding = nn.Parameter(scale * torch.randn(width))
            self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
            self.ln_pre = _LN(width)
            self.transformer = Transformer(width, layers, heads)
            self.ln_post = _LN(width)
            self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

        def forward(self, x):            # pooled upstream forward: class token -> ln_post -> proj
            x = self.conv1(x)
            x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
            x = torch.cat([self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
            x = self.ln_pre(x + self.positional_embedding.to(x.dtype))
            x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
            x = self.ln_post(x[:, 0, :])
            return x @ self.proj

    class CLIP(nn.Module):
        def __init__(self, d: CLIPDims):
            super().__init__()
            self.dims = d
            self.context_length = d.context
            self.visual = VisionTransformer(d.image_size, d.patch, d.v_width, d.v_layers, d.v_heads, d.embed_dim)
            mask = torch.full((d.context, d.context), float("-inf")).triu_(1)
            self.transformer = Transformer(d.t_width, d.t_layers, d.t_heads, attn_mask=mask)
            self.token_embedding = nn.Embedding(d.vocab, d.t_width)
            self.positional_embedding = nn.Parameter(torch.empty(d.context, d.t_width).normal_(std=0.01))
            self.ln_final = _LN(d.t_width)
            self.text_projection = nn.Parameter(torch.empty(d.t_width, d.embed_dim).normal_(std=d.t_width ** -0.5))
            self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)

        @property
        def dtype(self):
            return self.visual.conv1.weight.dtype

        def encode_image(self, image):
            return self.visual(image.type(self.dtype))

        def encode_text(self, text):
            x = self.token_embedding(text).type(self.dtype)
            x = x + self.positional_embedding.type(self.dtype)
            x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)
            x = self.ln_final(x).type(self.dtype)
            return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection

    dims_by_name = dict(CLIP_CONFIGS)

    def load(name, device="cpu", jit=False, download_root=None):
        return CLIP(dims_by_name[name]).to("cpu"), (lambda pil: pil)

    def tokenize(texts, context_length=77, truncate=False):
        raise RuntimeError("the shimmed clip.tokenize is never called by the parity tests")

    clip = types.ModuleType("clip")
    clip._uniir_shim = True
    clip.load, clip.tokenize, clip.available_models = load, tokenize, (lambda: list(dims_by_name))
    model = types.ModuleType("clip.model")
    model.VisionTransformer, model.Transformer, model.CLIP = VisionTransformer, Transformer, CLIP
    clip.model = model
    sys.modules["clip"], sys.modules["clip.model"] = clip, model
    return clip


def _import(modname):
    _install_fake_clip()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    return importlib.import_module(modname)


def load_clip_sf():
    """The reference module models.uniir_clip.clip_scorefusion.clip_sf (class CLIPScoreFusion)."""
    return _import("models.uniir_clip.clip_scorefusion.clip_sf")


def load_clip_ff(tiny_heads=None):
    """The reference module models.uniir_clip.clip_featurefusion.clip_ff (class CLIPFeatureFusion).
    The class only accepts "ViT-B/32" / "ViT-L/14" (clip_ff.py:83-100, 112-133)."""
    return _import("models.uniir_clip.clip_featurefusion.clip_ff")


class Cfg:
    """config.model.gather_embeddings / config.data_config.in_batch_neg_num (clip_sf.py:28-30)."""

    def __init__(self, gather=False, in_batch_neg_num=0):
        self.model = types.SimpleNamespace(gather_embeddings=gather)
        self.data_config = types.SimpleNamespace(in_batch_neg_num=in_batch_neg_num)
