"""Host logic of the CLIP_FF encode path (uniir_b200/clip_ff.py::CLIPFFEncoder) on CPU: every kernel wrapper in
uniir_b200.ops is replaced by a plain-torch statement of its documented contract (include/uniir_b200.h), so what runs
is exactly the product's sequencing -- buffer wiring, weight views (the fused q|k|v projection, x @ W vs Linear
layouts), the relative-position bias table, the text-then-image token order -- and it must reproduce the oracle.
The kernels themselves are covered by the GPU tests; this isolates the orchestration, which needs no GPU."""
import torch
import torch.nn.functional as F

import cpu_kernels
from oracle.clip_ff_oracle import CLIPFeatureFusionOracle


def test_clip_ff_encoder_sequencing_reproduces_the_oracle(monkeypatch):
    cpu_kernels.install(monkeypatch)
    from uniir_b200.clip_ff import CLIPFFEncoder, clip_ff_param_specs, t5_heads
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule
    name = "tiny-test"
    torch.manual_seed(3)
    orc = CLIPFeatureFusionOracle(name).eval()
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    d = CLIP_CONFIGS[name]
    root = torch.nn.Module()
    P = FlatParamModule(clip_ff_param_specs(d, t5_heads(name)), "cpu", root=root)
    missing, unexpected = root.load_state_dict(orc.state_dict(), strict=True)
    P.refresh_shadow(force=True)
    mb = 3
    enc = CLIPFFEncoder(P, d, t5_heads(name), mb, "cpu")
    enc.refresh_weights()
    g = torch.Generator().manual_seed(4)
    ids = torch.zeros(mb, d.context, dtype=torch.int32)
    for i in range(mb):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g).int(); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(mb, 3, d.image_size, d.image_size, generator=g)
    with torch.no_grad():
        ref = orc.encode_multimodal_input(ids.long(), img)
        out = enc.forward(ids, img).clone()
    cos = F.cosine_similarity(out, ref, dim=-1)
    assert cos.min().item() > 1 - 1e-3, cos            # bf16 operand storage is the only difference
    assert (out - ref).abs().max().item() < 3e-2 * ref.abs().max().item()
    # the intermediate features too: text tokens first, then image tokens (clip_ff.py:177)
    with torch.no_grad():
        t_ref, v_ref = orc.encode_text(ids.long()), orc.encode_image(img)
    assert F.cosine_similarity(enc.t_feat.flatten(1), t_ref.flatten(1), dim=-1).min() > 1 - 1e-3
    assert F.cosine_similarity(enc.v_feat.flatten(1), v_ref.flatten(1), dim=-1).min() > 1 - 1e-3


def test_clip_ff_encoder_backward_sequencing_reproduces_the_oracle_gradients(monkeypatch):
    """forward(save=True) + the hand-written backward (T5 stack: RMS-norm backward, attention backward with the bias
    gradient scattered into the bucket embedding, ReLU FFN; un-pooled towers) against the oracle's autograd."""
    cpu_kernels.install(monkeypatch)
    from uniir_b200.clip_ff import CLIPFFEncoder, clip_ff_param_specs, t5_heads
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule
    name = "tiny-test"
    torch.manual_seed(13)
    orc = CLIPFeatureFusionOracle(name).eval()
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    d = CLIP_CONFIGS[name]
    root = torch.nn.Module()
    P = FlatParamModule(clip_ff_param_specs(d, t5_heads(name)), "cpu", root=root)
    root.load_state_dict(orc.state_dict(), strict=True)
    P.refresh_shadow(force=True)
    P.ensure_grad()
    mb = 3
    enc = CLIPFFEncoder(P, d, t5_heads(name), mb, "cpu", train=True)
    enc.refresh_weights()
    g = torch.Generator().manual_seed(14)
    ids = torch.zeros(mb, d.context, dtype=torch.int32)
    for i in range(mb):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g).int(); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(mb, 3, d.image_size, d.image_size, generator=g)
    d_emb = torch.randn(mb, d.embed_dim, generator=g)
    ref = orc.encode_multimodal_input(ids.long(), img)
    ref.backward(d_emb)
    out = enc.forward(ids, img, save=True).clone()
    assert F.cosine_similarity(out, ref.detach(), dim=-1).min().item() > 1 - 1e-3
    enc.backward(d_emb)
    worst = (1.0, "")
    for n_, p_o in orc.named_parameters():
        if n_ == "clip_model.logit_scale":
            continue
        gm = P.g(n_).flatten().double()
        if p_o.grad is None:
            assert gm.abs().max().item() == 0, n_
            continue
        a = p_o.grad.flatten().double()
        cos = (torch.dot(a, gm) / (a.norm() * gm.norm() + 1e-30)).item()
        worst = min(worst, (cos, n_))
    assert worst[0] > 0.999, worst


def test_fused_adamw_parameter_groups_equal_the_reference_optimizer(monkeypatch):
    """FlatAdamW with a per-parameter learning-rate multiplier (sub-ranges of the flat segments) against torch.optim.AdamW
    with the reference's three CLIP_FF groups (clip_featurefusion/train.py:49-63, 200-206); and without the multiplier
    against the two CLIP_SF groups (clip_scorefusion/train.py:52-61)."""
    cpu_kernels.install(monkeypatch)
    from uniir_b200.clip_ff import clip_ff_param_specs, t5_heads
    from uniir_b200.engine import FlatAdamW
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule
    d = CLIP_CONFIGS["tiny-test"]
    for use_scale in (True, False):
        root = torch.nn.Module()
        P = FlatParamModule(clip_ff_param_specs(d, t5_heads("tiny-test")), "cpu", seed=3, root=root)
        ref = {n: p.detach().clone().requires_grad_(True) for n, p in root.named_parameters()}
        excl = lambda n, p: (p.ndim < 2 or any(s in n for s in ("bn", "ln", "bias", "logit_scale"))) and "t5" not in n
        groups = [{"params": [p for n, p in ref.items() if excl(n, p)], "weight_decay": 0.0},
                  {"params": [p for n, p in ref.items() if not excl(n, p) and "t5" not in n], "weight_decay": 0.2},
                  {"params": [p for n, p in ref.items() if "t5" in n], "weight_decay": 0.2, "lr": 5e-4 if use_scale else 1e-5}]
        topt = torch.optim.AdamW(groups, lr=1e-5, betas=(0.9, 0.98), eps=1e-6)
        opt = FlatAdamW(P, lr=1e-5, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2,
                        lr_scale=(lambda n: 50.0 if n.startswith("t5_layers.") else 1.0) if use_scale else None)
        if use_scale:
            assert sum(len(r) for r in opt._ranges.values()) <= 4      # a handful of launches, not one per parameter
        g = torch.Generator().manual_seed(4)
        for step in range(3):
            for n, p in ref.items():
                gr = torch.randn(p.shape, generator=g) * 0.1
                p.grad = gr.clone()
                P.g(n).copy_(gr)
            topt.step(); opt.step()
        for n, p in ref.items():
            assert torch.allclose(P.w(n), p.detach(), atol=1e-7, rtol=1e-5), n
            assert P.g(n).abs().max().item() == 0                       # the fused step zeroes the gradients


def t5_mask_hook(seed, p=0.1):
    """oracle.clip_ff_oracle.DROPOUT hook that reproduces the product's counter-based masks (CLIPFFEncoder sites)."""
    sites = {"probs": 1, "attn_out": 2, "ff_inner": 3, "ff_out": 4}

    def hook(kind, layer, x):
        site = 0 if kind == "input" else (1 + 4 * hook.layers if kind == "final" else sites[kind] + 4 * layer)
        keep = cpu_kernels.keep_mask(seed, site, torch.arange(x.numel()), p).view(x.shape)
        return x * keep.float() * cpu_kernels.drop_scale(p)
    hook.layers = 2
    return hook


def test_clip_ff_encoder_with_dropout_matches_oracle_given_the_same_masks(monkeypatch):
    """Train mode of the T5 fusion stack (dropout_rate 0.1 at transformers' sites; the oracle's hooks are pinned against
    transformers' T5Stack in tests/test_oracle_clip_ff.py): with the product's counter-based masks fed to the oracle, the
    forward and every parameter gradient agree; the no-activation forward (gradient-cache phase A, in-place residual
    stream) encodes the same thing, also when the batch is tiled differently."""
    import oracle.clip_ff_oracle as co
    cpu_kernels.install(monkeypatch)
    from uniir_b200.clip_ff import CLIPFFEncoder, clip_ff_param_specs, t5_heads
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule
    name, seed = "tiny-test", 4711
    torch.manual_seed(13)
    orc = CLIPFeatureFusionOracle(name).eval()
    with torch.no_grad():
        for n, p in orc.t5_layers.named_parameters():
            p.normal_(0.0, 0.08) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    d = CLIP_CONFIGS[name]
    root = torch.nn.Module()
    P = FlatParamModule(clip_ff_param_specs(d, t5_heads(name)), "cpu", root=root)
    root.load_state_dict(orc.state_dict(), strict=True)
    P.refresh_shadow(force=True)
    P.ensure_grad()
    mb = 4
    enc = CLIPFFEncoder(P, d, t5_heads(name), mb, "cpu", train=True)
    enc.set_dropout(0.1)
    enc.refresh_weights()
    g = torch.Generator().manual_seed(14)
    ids = torch.zeros(mb, d.context, dtype=torch.int32)
    for i in range(mb):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g).int(); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(mb, 3, d.image_size, d.image_size, generator=g)
    d_emb = torch.randn(mb, d.embed_dim, generator=g)
    plain = orc.encode_multimodal_input(ids.long(), img).detach()
    monkeypatch.setattr(co, "DROPOUT", t5_mask_hook(seed))
    ref = orc.encode_multimodal_input(ids.long(), img)
    ref.backward(d_emb)
    monkeypatch.setattr(co, "DROPOUT", None)
    assert (ref.detach() - plain).abs().max().item() > 1e-3
    out = enc.forward(ids, img, save=True, drop_seed=seed, item0=0).clone()
    assert F.cosine_similarity(out, ref.detach(), dim=-1).min().item() > 1 - 1e-3
    enc.backward(d_emb)
    worst = (1.0, "")
    for n_, p_o in orc.named_parameters():
        if n_ == "clip_model.logit_scale" or p_o.grad is None:
            continue
        a, gm = p_o.grad.flatten().double(), P.g(n_).flatten().double()
        cos = (torch.dot(a, gm) / (a.norm() * gm.norm() + 1e-30)).item()
        if n_.endswith("relative_attention_bias.weight"):
            # 32 x heads bucket sums over every (query, key) pair of the batch: the most cancellation-prone gradient
            # (0.9995 without dropout, the worst of all parameters there too); the masks add variance to each term
            assert cos > 0.997, cos
            continue
        worst = min(worst, (cos, n_))
    assert worst[0] > 0.998, worst      # bf16 storage under a random mask: 0.9990-0.9993 for the patch embedding, mask-dependent
    # phase A (no activations, in-place residual stream) with the same seed; then two micro-batches of two
    again = enc.forward(ids, img, save=False, drop_seed=seed, item0=0).clone()
    assert torch.allclose(again, out, atol=1e-5, rtol=1e-5)
    half = CLIPFFEncoder(P, d, t5_heads(name), 2, "cpu")
    half.set_dropout(0.1)
    half.refresh_weights()
    parts = [half.forward(ids[s:s + 2], img[s:s + 2], drop_seed=seed, item0=s).clone() for s in (0, 2)]
    assert torch.allclose(torch.cat(parts), out, atol=2e-2, rtol=2e-2)
    # no seed: the eval-mode encoder
    assert F.cosine_similarity(enc.forward(ids, img), plain, dim=-1).min().item() > 1 - 1e-3
