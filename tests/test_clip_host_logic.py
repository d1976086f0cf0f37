"""Host logic of the CLIP_SF encoder (uniir_b200/clip_towers.py::CLIPEncoder: both towers, pooling, projections, score
fusion, and the hand-written backward through all of it) on CPU with the kernel wrappers replaced by their torch
contracts (tests/cpu_kernels.py), against the oracle's autograd."""
import torch
import torch.nn.functional as F

import cpu_kernels
from oracle.clip_oracle import CLIPScoreFusionOracle


def test_clip_encoder_forward_and_backward_sequencing(monkeypatch):
    cpu_kernels.install(monkeypatch)
    from uniir_b200 import ops
    from uniir_b200.clip_towers import CLIPEncoder
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule, clip_param_specs
    monkeypatch.setattr(CLIPEncoder, "_fork", lambda self: None)      # no side stream on CPU
    monkeypatch.setattr(CLIPEncoder, "_join", lambda self, main: None)
    name = "tiny-test"
    d = CLIP_CONFIGS[name]
    torch.manual_seed(2)
    orc = CLIPScoreFusionOracle(name).eval()
    root = torch.nn.Module()
    root.clip_model = FlatParamModule(clip_param_specs(d), "cpu")
    missing, unexpected = root.load_state_dict(orc.state_dict(), strict=True)
    P = root.clip_model
    P.refresh_shadow(force=True)
    P.ensure_grad()
    mb = 4
    enc = CLIPEncoder(P, d, mb, "cpu", True)
    enc.refresh_weights()
    g = torch.Generator().manual_seed(3)
    ids = torch.zeros(mb, d.context, dtype=torch.int32)
    for i in range(mb):
        n = int(torch.randint(1, d.context - 2, (1,), generator=g))
        ids[i, 0] = d.vocab - 2; ids[i, 1:1 + n] = torch.randint(1, d.vocab - 2, (n,), generator=g).int(); ids[i, 1 + n] = d.vocab - 1
    img = torch.randn(mb, 3, d.image_size, d.image_size, generator=g)
    mt = torch.tensor([1, 1, 0, 1]); mv = torch.tensor([1, 0, 1, 1])
    d_emb = torch.randn(mb, d.embed_dim, generator=g)
    ref = orc.encode_multimodal_input(ids.long(), img, mt, mv)
    ref.backward(d_emb)
    # product: forward with saved activations, score fusion, backward (the body of the captured CUDA graph)
    enc.set_inputs(ids, img, mt, mv, d_emb)
    t, v = enc.forward(enc.in_ids, enc.in_img, save=True)
    out = ops.score_fuse_fwd(t, v, enc.in_mt, enc.in_mv)
    assert F.cosine_similarity(out, ref.detach(), dim=-1).min().item() > 1 - 1e-3
    ops.score_fuse_bwd(enc.in_demb, enc.in_mt, enc.in_mv, out=(enc.d_t, enc.d_v))
    enc.backward(enc.d_t, enc.d_v)
    worst = (1.0, "")
    for n_, p_o in orc.clip_model.named_parameters():
        if n_ == "logit_scale":
            continue
        a, c = p_o.grad.flatten().double(), P.g(n_).flatten().double()
        cos = (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item()
        worst = min(worst, (cos, n_))
    assert worst[0] > 0.999, worst
