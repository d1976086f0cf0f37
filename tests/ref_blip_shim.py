"""Import the reference's vendored BLIP modules (backbone/vit.py, backbone/med.py) without editing them,
by shimming the transformers-4.15 / timm / fairscale symbols they expect (SURVEY.md 8c recipe).
Only usable where /root/reference exists (the build container); the GPU box uses the golden vectors."""
import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(os.path.join(REF_SRC, "models", "uniir_blip", "backbone"))


def load():
    import torch
    import torch.nn as nn
    import transformers  # import BEFORE stubbing
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = None

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
            super().__init__()
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class DropPath(nn.Module):
        """timm.models.layers.DropPath (timm is a third-party dependency of the reference, not in this image; the
        reference's env pins timm 0.4.12), restated from its published source: per-sample stochastic depth,
        `x.div(keep_prob) * floor(keep_prob + rand(B, 1, ..))`, identity in eval mode or at drop_prob 0.
        `mask_fn` (class attribute, None = torch's RNG) lets a test supply the per-sample keep decisions."""
        mask_fn = None

        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep_prob = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            if DropPath.mask_fn is not None:
                random_tensor = DropPath.mask_fn(shape, self.drop_prob).to(x.dtype)
            else:
                random_tensor = (keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
            return x.div(keep_prob) * random_tensor

    if "timm" not in sys.modules:
        stub("timm"); stub("timm.models")
        stub("timm.models.vision_transformer", PatchEmbed=PatchEmbed, _cfg=lambda **k: {})
        stub("timm.models.registry", register_model=lambda f: f)
        stub("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=DropPath)
        stub("timm.models.helpers", named_apply=lambda *a, **k: None, adapt_input_conv=lambda *a, **k: None)
        stub("timm.models.hub")
    if "fairscale" not in sys.modules:
        stub("fairscale"); stub("fairscale.nn"); stub("fairscale.nn.checkpoint")
        stub("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, **k: m)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    vit = importlib.import_module("models.uniir_blip.backbone.vit")
    med = importlib.import_module("models.uniir_blip.backbone.med")
    B = med.BertPreTrainedModel
    B.init_weights = lambda self: self.apply(self._init_weights)
    B.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    return vit, med


def load_blip_ff():
    """The reference BLIPFeatureFusion class itself (blip_ff.py), with the tokenizer download stubbed."""
    load()
    sys.modules["timm.models.hub"].download_cached_file = lambda *a, **k: None
    blip = importlib.import_module("models.uniir_blip.backbone.blip")
    blip.init_tokenizer = lambda: None
    bf = importlib.import_module("models.uniir_blip.blip_featurefusion.blip_ff")
    bf.init_tokenizer = lambda: None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):   # concat_all_gather at world size 1
        bf.concat_all_gather = lambda t: t
    return bf


def load_blip_sf():
    """The reference BLIPScoreFusion class itself (blip_scorefusion/blip_sf.py), tokenizer download stubbed."""
    load()
    sys.modules["timm.models.hub"].download_cached_file = lambda *a, **k: None
    blip = importlib.import_module("models.uniir_blip.backbone.blip")
    blip.init_tokenizer = lambda: None
    bs = importlib.import_module("models.uniir_blip.blip_scorefusion.blip_sf")
    bs.init_tokenizer = lambda: None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):   # concat_all_gather at world size 1
        bs.concat_all_gather = lambda t: t
    return bs
