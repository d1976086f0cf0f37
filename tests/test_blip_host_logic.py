"""Host logic of the BLIP encoders (uniir_b200/blip_ff.py::BLIPEncoder, feature fusion and score fusion) on CPU: the
kernel wrappers are replaced by their torch contracts (tests/cpu_kernels.py), so the product's own forward and
hand-written backward sequencing run here and are compared with the oracle's autograd -- every parameter gradient,
including the fused q|k|v and cross k|v views and (score fusion) the untouched frozen cross-attention."""
import pytest
import torch
import torch.nn.functional as F

import cpu_kernels
from oracle.blip_oracle import BLIPFeatureFusionOracle, BLIPScoreFusionOracle
from test_oracle_blip import _inputs


@pytest.mark.parametrize("mode", ["ff", "sf"])
def test_blip_encoder_forward_and_backward_sequencing(monkeypatch, mode):
    cpu_kernels.install(monkeypatch)
    from uniir_b200.blip_ff import BLIP_CONFIGS, BLIPEncoder, blip_param_specs
    from uniir_b200.params import FlatParamModule
    d = BLIP_CONFIGS["tiny-test"]
    E = 64
    torch.manual_seed(5)
    orc = (BLIPFeatureFusionOracle if mode == "ff" else BLIPScoreFusionOracle)("tiny-test", embed_dim=E, queue_size=16).eval()
    with torch.no_grad():      # as in the GPU tests: a well-conditioned problem
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_:
                p_.mul_(4.0)
    root = torch.nn.Module()
    P = FlatParamModule(blip_param_specs(d, "", mode, E), "cpu", root=root)
    sd = {k: v for k, v in orc.state_dict().items() if k in dict(root.named_parameters())}
    missing, unexpected = root.load_state_dict(sd, strict=True)
    P.refresh_shadow(force=True)
    P.ensure_grad()
    mb, S = 4, 24
    enc = BLIPEncoder(P, d, mb, S, "cpu", train=True, mode=mode, embed_dim=E)
    ids, att, img = _inputs(d, mb, S, 7)
    lens = att.sum(1).to(torch.int32)
    mt = torch.tensor([1, 0, 1, 1]); mv = torch.tensor([1, 1, 0, 1])
    g = torch.Generator().manual_seed(8)
    d_emb = torch.randn(mb, d.t_width if mode == "ff" else E, generator=g)
    # oracle: autograd through the restated reference arithmetic
    for p_ in orc.parameters():
        p_.grad = None
    ref = orc.encode_multimodal_input(ids, att, img, mt, mv)
    ref.backward(d_emb)
    # product sequencing on the torch kernel contracts
    out = enc.forward(ids.to(torch.int32), lens, img, save=True, mt=mt if mode == "sf" else None,
                      mv=mv if mode == "sf" else None).clone()
    assert F.cosine_similarity(out, ref.detach(), dim=-1).min().item() > 1 - 1e-3
    enc.backward(d_emb)
    worst = (1.0, "")
    for n_, p_o in orc.named_parameters():
        if "_m." in n_ or n_.split(".")[0].endswith("_m") or n_ == "temp":
            continue
        gm = P.g(n_).flatten().double()
        if p_o.grad is None:
            # unused parameters: (sf) the frozen cross-attention; the positions beyond the text length never get a gradient
            assert gm.abs().max().item() == 0, n_
            continue
        a = p_o.grad.flatten().double()
        if n_.endswith("self.key.bias"):
            continue       # softmax is invariant to a per-query constant: exactly zero up to rounding
        cos = (torch.dot(a, gm) / (a.norm() * gm.norm() + 1e-30)).item()
        worst = min(worst, (cos, n_))
    assert worst[0] > 0.999, worst


def _mask_hook(seed, rate, n_layers):
    """oracle.blip_oracle.DROPOUT hook that reproduces the product's counter-based masks (BLIPEncoder site numbering)."""
    dpr = torch.linspace(0, rate, n_layers).tolist()
    kinds = {"self_probs": 0, "self_out": 1, "cross_probs": 2, "cross_out": 3, "ffn_out": 4}

    def hook(tag, kind, layer, x):
        sd = seed[tag] if isinstance(seed, dict) else seed        # {"": online seed, "_m": momentum seed} or one seed
        if kind.startswith("path"):
            p = dpr[layer]
            if p == 0.0:
                return x
            site = 4096 + 2 * layer + (kind == "path_mlp")
            keep = cpu_kernels.keep_mask(sd, site, torch.arange(x.shape[0]), p).float() * cpu_kernels.drop_scale(p)
            return x * keep.view(-1, *([1] * (x.ndim - 1)))
        site = 0 if kind == "emb" else 1 + 5 * layer + kinds[kind]
        keep = cpu_kernels.keep_mask(sd, site, torch.arange(x.numel()), 0.1).view(x.shape)
        return x * keep.float() * cpu_kernels.drop_scale(0.1)
    return hook


@pytest.mark.parametrize("mode", ["ff", "sf"])
def test_blip_encoder_with_dropout_matches_oracle_given_the_same_masks(monkeypatch, mode):
    """Train mode: hidden dropout, attention-probability dropout (self and cross) and DropPath at the reference's
    sites (oracle hooks, pinned against the reference modules in tests/test_oracle_blip.py), with the product's
    counter-based masks fed to the oracle: forward and every parameter gradient agree; and the masks follow the
    position in the BATCH, so another micro-batch tiling encodes the same thing."""
    import oracle.blip_oracle as bo
    cpu_kernels.install(monkeypatch)
    from uniir_b200.blip_ff import BLIP_CONFIGS, BLIPEncoder, blip_param_specs
    from uniir_b200.params import FlatParamModule
    d = BLIP_CONFIGS["tiny-test"]
    E, seed, rate = 64, 777, 0.3
    torch.manual_seed(5)
    orc = (BLIPFeatureFusionOracle if mode == "ff" else BLIPScoreFusionOracle)("tiny-test", embed_dim=E, queue_size=16).eval()
    with torch.no_grad():
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_:
                p_.mul_(4.0)
    root = torch.nn.Module()
    P = FlatParamModule(blip_param_specs(d, "", mode, E), "cpu", root=root)
    root.load_state_dict({k: v for k, v in orc.state_dict().items() if k in dict(root.named_parameters())}, strict=True)
    P.refresh_shadow(force=True)
    P.ensure_grad()
    mb, S = 4, 24
    enc = BLIPEncoder(P, d, mb, S, "cpu", train=True, mode=mode, embed_dim=E)
    enc.set_dropout(0.1, 0.1, rate)
    ids, att, img = _inputs(d, mb, S, 7)
    lens = att.sum(1).to(torch.int32)
    mt = torch.tensor([1, 0, 1, 1]); mv = torch.tensor([1, 1, 0, 1])
    d_emb = torch.randn(mb, d.t_width if mode == "ff" else E, generator=torch.Generator().manual_seed(8))
    plain = orc.encode_multimodal_input(ids, att, img, mt, mv).detach()
    monkeypatch.setattr(bo, "DROPOUT", _mask_hook(seed, rate, d.v_layers))
    ref = orc.encode_multimodal_input(ids, att, img, mt, mv)
    ref.backward(d_emb)
    monkeypatch.setattr(bo, "DROPOUT", None)
    assert (ref.detach() - plain).abs().max().item() > 1e-2          # the masks do something
    kw = dict(mt=mt if mode == "sf" else None, mv=mv if mode == "sf" else None)
    out = enc.run(ids.to(torch.int32), lens, img, d_emb=d_emb, drop_seed=seed, item0=0, **kw).clone()
    assert F.cosine_similarity(out, ref.detach(), dim=-1).min().item() > 1 - 1e-3
    worst = (1.0, "")
    for n_, p_o in orc.named_parameters():
        if "_m." in n_ or n_.split(".")[0].endswith("_m") or n_ == "temp" or p_o.grad is None or n_.endswith("self.key.bias"):
            continue
        a, gm = p_o.grad.flatten().double(), P.g(n_).flatten().double()
        worst = min(worst, ((torch.dot(a, gm) / (a.norm() * gm.norm() + 1e-30)).item(), n_))
    assert worst[0] > 0.999, worst
    # same seed, two micro-batches of two items: item0 carries the batch position
    half = BLIPEncoder(P, d, 2, S, "cpu", train=False, mode=mode, embed_dim=E)
    half.set_dropout(0.1, 0.1, rate)
    parts = []
    for s in (0, 2):
        kw2 = dict(mt=mt[s:s + 2] if mode == "sf" else None, mv=mv[s:s + 2] if mode == "sf" else None)
        parts.append(half.run(ids[s:s + 2].to(torch.int32), lens[s:s + 2], img[s:s + 2], drop_seed=seed, item0=s, **kw2).clone())
    assert torch.allclose(torch.cat(parts), out, atol=2e-2, rtol=2e-2)
    # and without a seed the encoder is the eval-mode one
    assert F.cosine_similarity(enc.run(ids.to(torch.int32), lens, img, **kw), plain, dim=-1).min().item() > 1 - 1e-3
