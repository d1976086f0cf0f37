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
