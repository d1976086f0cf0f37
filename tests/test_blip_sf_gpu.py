"""GPU parity of BLIP_SF (reference blip_scorefusion/blip_sf.py) against the CPU oracle, which
tests/test_oracle_blip.py pins to the real reference class, and against the committed golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, NO_DROP

pytestmark = pytest.mark.gpu
COS_TOL = 1e-3


def test_blip_sf_golden_oracle_and_api():
    from oracle.blip_oracle import BLIPScoreFusionOracle
    from uniir_b200.blip_sf import BLIPScoreFusion
    from test_oracle_blip import _inputs
    g = np.load(os.path.join(GOLDEN, "blip_sf_tiny.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    m = BLIPScoreFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=64, micro_batch=4)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(("_m." in k) or k in ("query_queue", "cand_queue", "idx_queue", "new_ptr_queue", "temp") for k in missing), (missing, unexpected)
    txt = {"input_ids": torch.from_numpy(g["ids"]).cuda(), "attention_mask": torch.from_numpy(g["att"]).cuda()}
    mt, mv = torch.from_numpy(g["mt"]), torch.from_numpy(g["mv"])
    emb = m.encode_multimodal_input(txt, torch.from_numpy(g["img"]).cuda(), mt, mv).cpu()   # 6 items, mb 4: ragged tail
    ref = torch.from_numpy(g["emb"])
    assert emb.shape == ref.shape == (6, 64)
    assert torch.nn.functional.cosine_similarity(emb, ref, dim=-1).min().item() > 1 - COS_TOL
    assert (emb - ref).abs().max().item() < 3e-2
    # fresh oracle, other weights / inputs / masks, online and momentum encoders
    torch.manual_seed(9)
    orc = BLIPScoreFusionOracle("tiny-test", embed_dim=64, queue_size=64).eval()
    missing, unexpected = m.load_state_dict(orc.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)      # identical state-dict layout
    ids, att, img = _inputs(orc.dims, 5, 40, 3)
    mt = torch.tensor([1, 0, 1, 1, 0]); mv = torch.tensor([0, 1, 1, 1, 1])
    for mom in (False, True):
        with torch.no_grad():
            ref = orc.encode_multimodal_input(ids, att, img, mt, mv, use_momentum=mom)
        out = m.encode_multimodal_input({"input_ids": ids, "attention_mask": att}, img, mt, mv, use_momentum=mom).cpu()
        assert torch.nn.functional.cosine_similarity(out, ref, dim=-1).min().item() > 1 - COS_TOL
    # reference API contract (blip_sf.py:319-332): ValueError without ids; frozen cross-attention; no pooler
    b = {"txt_batched": {"input_ids": ids, "attention_mask": att}, "image_batched": img, "txt_mask_batched": mt,
         "image_mask_batched": mv, "qid_list": [3, 4, 5, 6, 7]}
    e, idl = m(b, encode_mbeir_batch=True)
    assert e.shape == (5, 64) and idl == [3, 4, 5, 6, 7]
    b.pop("qid_list")
    with pytest.raises(ValueError):
        m(b, encode_mbeir_batch=True)
    keys = set(m.state_dict())
    assert "text_encoder.pooler.dense.weight" not in keys and "vision_proj_m.bias" in keys and "text_proj.weight" in keys
    assert not m.get_parameter("text_encoder.encoder.layer.0.crossattention.self.key.weight").requires_grad
    assert m.get_parameter("text_encoder.encoder.layer.0.attention.self.key.weight").requires_grad


def test_blip_sf_training_step_matches_oracle_and_keeps_frozen_layers():
    """The momentum-distillation step of BLIP_SF (blip_sf.py:174-317) over two consecutive steps vs the oracle:
    loss, accuracy, every trained gradient, EMA, queues; then one fused AdamW step leaves the frozen cross-attention
    parameters bit-identical (torch's AdamW skips parameters without a gradient)."""
    from oracle.blip_oracle import BLIPScoreFusionOracle
    from uniir_b200.blip_sf import BLIPScoreFusion
    from uniir_b200.engine import FlatAdamW
    from test_oracle_blip import _inputs
    torch.manual_seed(3)
    orc = BLIPScoreFusionOracle("tiny-test", embed_dim=64, queue_size=16).eval()
    with torch.no_grad():          # see test_blip_ff_training_step_matches_oracle: a well-conditioned problem
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_ and "_m" != n_.split(".")[0][-2:]:
                p_.mul_(4.0)
    orc.copy_params()
    m = BLIPScoreFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=16, micro_batch=8)
    m.train_micro_batch = 8
    missing, unexpected = m.load_state_dict(orc.state_dict(), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    d = orc.dims
    mt = torch.tensor([1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1]); mv = torch.tensor([1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1])
    for step in range(2):
        ids, att, img = _inputs(d, 16, 24, 30 + step)
        b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=mt,
                 image_mask_batched=mv, p_did_list=[5, 6, 5, 7, 8, 9, 10, 11 + step],
                 index_mapping={"query": [[2 * i] for i in range(8)], "pos_cand": [[2 * i + 1] for i in range(8)]})
        orc.zero_grad()
        ref = orc.compute_contrastive_loss(b, alpha=0.3); ref["loss"].backward()
        if m._online.grad is not None:
            for gbuf in m._online.grad.values():
                gbuf.zero_()
        out = m.compute_contrastive_loss(b, alpha=0.3)
        m.backward_contrastive()
        torch.cuda.synchronize()
        assert abs(out["loss"].item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item())), (out["loss"].item(), ref["loss"].item())
        assert abs(out["accuracy"].item() - ref["accuracy"].item()) <= 1.0 / 8 + 1e-6
        worst, allcos = (1.0, ""), []
        for n_, p_o in orc.named_parameters():
            if "_m." in n_ or n_.split(".")[0].endswith("_m"):
                continue
            gm = m._online.g(n_).cpu().flatten().double()
            if p_o.grad is None:                                   # frozen cross-attention: never touched
                assert ".crossattention." in n_ and gm.abs().max().item() == 0, n_
                continue
            a = p_o.grad.flatten().double()
            if n_.endswith("self.key.bias"):
                continue   # softmax is invariant to a constant added to every key: exactly 0 + noise
            if a.numel() > 1:
                cos = (torch.dot(a, gm) / (a.norm() * gm.norm() + 1e-30)).item()
                allcos.append((round(cos, 5), n_))
                worst = min(worst, (cos, n_))
            else:
                assert abs(a.item() - gm.item()) < 0.25 * max(1e-3, abs(a.item())), (n_, a.item(), gm.item())
        assert worst[0] > 0.995, sorted(allcos)[:8]
        assert torch.equal(m.idx_queue.cpu(), orc.idx_queue) and int(m.new_ptr_queue) == int(orc.new_ptr_queue)
        assert torch.nn.functional.cosine_similarity(m.cand_queue.cpu().t(), orc.cand_queue.t(), dim=-1).min() > 1 - 1e-3
        assert torch.allclose(m.get_parameter("text_proj_m.weight").cpu(), orc.text_proj_m.weight, atol=1e-6)
        with torch.no_grad():
            for n_, p_o in orc.named_parameters():
                if p_o.grad is not None:
                    p_o -= 1e-2 * p_o.grad
                    m.get_parameter(n_).copy_(p_o)
    # fused AdamW (single group, wd 0.05): frozen parameters unchanged, trained ones move
    frozen = "text_encoder.encoder.layer.1.crossattention.output.dense.weight"
    before_f, before_t = m.get_parameter(frozen).clone(), m.get_parameter("text_proj.weight").clone()
    opt = FlatAdamW(m._online, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    opt.step()
    assert torch.equal(m.get_parameter(frozen), before_f)
    assert not torch.equal(m.get_parameter("text_proj.weight"), before_t)
