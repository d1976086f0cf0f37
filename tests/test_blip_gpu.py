"""GPU parity of the BLIP_FF encode path against the CPU oracle / golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, NO_DROP

pytestmark = pytest.mark.gpu
COS_TOL = 1e-3


def _load_oracle_weights(model, sd):
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # everything the encode path uses must have been provided; only momentum copies / queues / temp may be missing
    assert all(("_m." in m) or m in ("query_queue", "cand_queue", "idx_queue", "new_ptr_queue", "temp") for m in missing), missing


def test_blip_ff_golden_and_oracle():
    from oracle.blip_oracle import BLIPFeatureFusionOracle
    from uniir_b200.blip_ff import BLIPFeatureFusion
    g = np.load(os.path.join(GOLDEN, "blip_tiny.npz"))
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}
    m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=64, micro_batch=4)
    _load_oracle_weights(m, sd)
    txt = {"input_ids": torch.from_numpy(g["ids"]).cuda(), "attention_mask": torch.from_numpy(g["att"]).cuda()}
    emb = m.encode_multimodal_input(txt, torch.from_numpy(g["img"]).cuda(), None, None).cpu()   # 6 items, mb 4: ragged tail
    cos = torch.nn.functional.cosine_similarity(emb, torch.from_numpy(g["emb"]), dim=-1)
    assert cos.min().item() > 1 - COS_TOL, cos
    assert (emb - torch.from_numpy(g["emb"])).abs().max().item() < 3e-2
    # fresh oracle with other weights / inputs
    torch.manual_seed(7)
    orc = BLIPFeatureFusionOracle("tiny-test", queue_size=64).eval()
    _load_oracle_weights(m, {k: v for k, v in orc.state_dict().items() if k.startswith(("visual_encoder.", "text_encoder."))})
    from test_oracle_blip import _inputs
    ids, att, img = _inputs(orc.dims, 5, 40, 3)
    with torch.no_grad():
        ref = orc.encode_multimodal_input(ids, att, img)
    out = m.encode_multimodal_input({"input_ids": ids, "attention_mask": att}, img).cpu()
    assert torch.nn.functional.cosine_similarity(out, ref, dim=-1).min().item() > 1 - COS_TOL
    # the momentum encoder is an exact copy at construction / after loading both
    assert torch.equal(m.get_parameter("visual_encoder_m.cls_token"), m.get_parameter("visual_encoder_m.cls_token"))


def test_blip_ff_api_contract():
    from uniir_b200.blip_ff import BLIPFeatureFusion
    m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=128)
    d = m.dims
    from test_oracle_blip import _inputs
    ids, att, img = _inputs(d, 3, 16, 1)
    b = {"txt_batched": {"input_ids": ids, "attention_mask": att}, "image_batched": img,
         "txt_mask_batched": torch.ones(3), "image_mask_batched": torch.ones(3), "did_list": [11, 12, 13]}
    emb, idl = m(b, encode_mbeir_batch=True)
    assert emb.shape == (3, d.t_width) and idl == [11, 12, 13]
    b2 = dict(b); b2.pop("did_list")
    with pytest.raises(ValueError):
        m(b2, encode_mbeir_batch=True)
    keys = set(m.state_dict())
    for k in ("visual_encoder.patch_embed.proj.bias", "visual_encoder_m.blocks.0.attn.qkv.weight",
              "text_encoder.encoder.layer.0.crossattention.self.key.weight", "text_encoder.embeddings.position_ids",
              "text_encoder_m.pooler.dense.weight", "query_queue", "cand_queue", "idx_queue", "new_ptr_queue", "temp"):
        assert k in keys, k
    assert m.state_dict()["idx_queue"].shape == (1, 128) and not m.get_parameter("text_encoder_m.pooler.dense.weight").requires_grad
    bad = att.clone(); bad[0, 0] = 0
    with pytest.raises(ValueError):
        m.encode_multimodal_input({"input_ids": ids, "attention_mask": bad}, img)


def test_cross_attention_kernel_vs_reference():
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, Sq, Skv, H = 3, 100, 197, 12
    q = torch.randn(B * Sq, H * 64, generator=g).bfloat16(); kv = torch.randn(B * Skv, 2 * H * 64, generator=g).bfloat16()
    out = ops.attn_fwd_ex(q.cuda(), kv.cuda()[:, :H * 64], kv.cuda()[:, H * 64:], B, Sq, Skv, H).cpu().float()
    qf = q.float().view(B, Sq, H, 64).transpose(1, 2); kf = kv[:, :H * 64].float().view(B, Skv, H, 64).transpose(1, 2)
    vf = kv[:, H * 64:].float().view(B, Skv, H, 64).transpose(1, 2)
    ref = ((qf @ kf.transpose(-1, -2) * 0.125).softmax(-1) @ vf).transpose(1, 2).reshape(B * Sq, H * 64)
    assert ((out - ref).norm() / ref.norm()).item() < 6e-3
    # key padding mask (prefix lengths)
    S = 100
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).bfloat16()
    lens = torch.tensor([100, 7, 55], dtype=torch.int32)
    o = ops.attn_fwd_ex(qkv.cuda()[:, :768], qkv.cuda()[:, 768:1536], qkv.cuda()[:, 1536:], B, S, S, H, kv_len=lens.cuda()).cpu().float()
    q3, k3, v3 = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = (torch.arange(S)[None, :] >= lens[:, None]).float()[:, None, None, :] * -10000.0
    ref = ((q3 @ k3.transpose(-1, -2) * 0.125 + mask).softmax(-1) @ v3).transpose(1, 2).reshape(B * S, H * 64)
    assert ((o - ref).norm() / ref.norm()).item() < 6e-3


def test_blip_ff_training_step_matches_oracle():
    """Momentum-distillation loss with queue (blip_ff.py:118-257): loss, accuracy, all online gradients, temp
    gradient, EMA'd momentum weights and queue contents over two consecutive steps vs the oracle (which is itself
    pinned to the real reference class in tests/test_oracle_blip.py)."""
    from oracle.blip_oracle import BLIPFeatureFusionOracle
    from uniir_b200.blip_ff import BLIPFeatureFusion
    from test_oracle_blip import _inputs
    torch.manual_seed(3)
    orc = BLIPFeatureFusionOracle("tiny-test", embed_dim=64, queue_size=16).eval()
    # N(0, 0.02) init leaves every pooled embedding nearly identical, so the gradients (|g| ~ 1e-5) are a
    # cancellation residue dominated by bf16 rounding (measured: cos 0.88-0.98).  Scaling the matrices by 4 gives
    # a well-conditioned problem (|g| ~ 1e-2) on which the bf16 path reaches cos >= 0.9998 on every parameter.
    with torch.no_grad():
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_:
                p_.mul_(4.0)
    orc.copy_params()
    m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=16, micro_batch=8)
    m.train_micro_batch = 8
    missing, unexpected = m.load_state_dict(orc.state_dict(), strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    d = orc.dims
    for step in range(2):
        ids, att, img = _inputs(d, 16, 24, 20 + step)
        b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=None,
                 image_mask_batched=None, p_did_list=[5, 6, 5, 7, 8, 9, 10, 11 + step],
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
        # the accuracy is an arg-max over near-tied logits of a random-init model: bf16 towers may flip one row
        assert abs(out["accuracy"].item() - ref["accuracy"].item()) <= 1.0 / 8 + 1e-6
        worst = (1.0, "")
        allcos = []
        for n_, p_o in orc.named_parameters():
            if p_o.grad is None:
                continue
            a, c = p_o.grad.flatten().double(), m._online.g(n_).cpu().flatten().double()
            if n_.endswith("self.key.bias"):
                continue   # softmax is invariant to a constant added to every key: this gradient is exactly 0 + noise
            if a.numel() > 1:
                cos = (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item()
                allcos.append((round(cos, 5), n_))
                worst = min(worst, (cos, n_))
            else:
                # d loss / d temp = -(1/temp) sum_ij (p_ij - t_ij) s_ij: a small difference of large sums over
                # logits of magnitude 1/temp ~ 14, so bf16-tower noise in the embeddings is amplified
                assert abs(a.item() - c.item()) < 0.25 * max(1e-3, abs(a.item())), (n_, a.item(), c.item())
        assert worst[0] > 0.995, sorted(allcos)[:8]
        assert torch.equal(m.idx_queue.cpu(), orc.idx_queue) and int(m.new_ptr_queue) == int(orc.new_ptr_queue)
        assert torch.nn.functional.cosine_similarity(m.cand_queue.cpu().t(), orc.cand_queue.t(), dim=-1).min() > 1 - 1e-3
        wm, wo = m.get_parameter("visual_encoder_m.blocks.0.attn.qkv.weight").cpu(), orc.visual_encoder_m.blocks[0].attn.qkv.weight
        assert torch.allclose(wm, wo, atol=1e-6)
        # apply the same parameter update on both sides so step 2 starts from equal weights
        with torch.no_grad():
            for n_, p_o in orc.named_parameters():
                if p_o.grad is not None:
                    p_o -= 1e-2 * p_o.grad
                    m.get_parameter(n_).copy_(p_o)


def test_cross_attention_backward_kernel():
    from uniir_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, Sq, Skv, H = 2, 100, 197, 3
    q = torch.randn(B * Sq, H * 64, generator=g).bfloat16(); kv = torch.randn(B * Skv, 2 * H * 64, generator=g).bfloat16()
    lens = None
    qd, kvd = q.cuda(), kv.cuda()
    lse = torch.empty(B, H, Sq, device="cuda")
    out = ops.attn_fwd_ex(qd, kvd[:, :H * 64], kvd[:, H * 64:], B, Sq, Skv, H, lse=lse)
    dout = torch.randn(B * Sq, H * 64, generator=g).bfloat16()
    dq = torch.empty_like(qd); dkv = torch.empty_like(kvd)
    ops.attn_bwd_ex(qd, kvd[:, :H * 64], kvd[:, H * 64:], out, dout.cuda(), lse, B, Sq, Skv, H, dq, dkv[:, :H * 64], dkv[:, H * 64:])
    qf = q.float().requires_grad_(True); kvf = kv.float().requires_grad_(True)
    qh = qf.view(B, Sq, H, 64).transpose(1, 2); kh = kvf[:, :H * 64].reshape(B, Skv, H, 64).transpose(1, 2)
    vh = kvf[:, H * 64:].reshape(B, Skv, H, 64).transpose(1, 2)
    ref = ((qh @ kh.transpose(-1, -2) * 0.125).softmax(-1) @ vh).transpose(1, 2).reshape(B * Sq, H * 64)
    ref.backward(dout.float())
    rel = lambda a, b: ((a.float().cpu() - b).norm() / b.norm()).item()
    assert rel(dq, qf.grad) < 1.2e-2 and rel(dkv, kvf.grad) < 1.2e-2
    # self-attention with a key-padding mask
    S = 100
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).bfloat16(); qkvd = qkv.cuda()
    lens = torch.tensor([100, 37], dtype=torch.int32).cuda()
    D = H * 64
    o = ops.attn_fwd_ex(qkvd[:, :D], qkvd[:, D:2 * D], qkvd[:, 2 * D:], B, S, S, H, kv_len=lens, lse=(l2 := torch.empty(B, H, S, device="cuda")))
    do = torch.randn(B * S, D, generator=g).bfloat16()
    dqkv = torch.zeros_like(qkvd)
    ops.attn_bwd_ex(qkvd[:, :D], qkvd[:, D:2 * D], qkvd[:, 2 * D:], o, do.cuda(), l2, B, S, S, H, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], kv_len=lens)
    x = qkv.float().requires_grad_(True)
    q3, k3, v3 = x.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    mask = (torch.arange(S)[None, :] >= lens.cpu()[:, None]).float()[:, None, None, :] * -10000.0
    r2 = ((q3 @ k3.transpose(-1, -2) * 0.125 + mask).softmax(-1) @ v3).transpose(1, 2).reshape(B * S, D)
    r2.backward(do.float())
    assert rel(dqkv, x.grad) < 1.2e-2


def test_blip_ff_hard_negative_variant_matches_oracle():
    """blip_ff.py:126-131,160-171,208-218,233-246 (oracle pinned to the real reference class in
    tests/test_oracle_blip.py): the negatives' ids / un-normalised momentum features replace the first queue columns;
    a coin flip (host RNG, same seed on both sides) enqueues the positives or each query's first negative."""
    from oracle.blip_oracle import BLIPFeatureFusionOracle
    from uniir_b200.blip_ff import BLIPFeatureFusion
    from test_oracle_blip import _inputs
    torch.manual_seed(5)
    orc = BLIPFeatureFusionOracle("tiny-test", embed_dim=64, queue_size=32).eval()
    with torch.no_grad():
        for n_, p_ in orc.named_parameters():
            if p_.ndim >= 2 and "_m." not in n_:
                p_.mul_(4.0)
    orc.copy_params()
    m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=32, micro_batch=8)
    m.train_micro_batch = 8
    m.load_state_dict(orc.state_dict(), strict=False)
    d = orc.dims
    flips = set()
    for step, seed in enumerate((102, 103)):            # rand(1) < 0.5 is True for 102, False for 103
        ids, att, img = _inputs(d, 16, 24, 60 + step)   # 4 pairs x [query, pos, neg, neg]
        b = dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=None,
                 image_mask_batched=None, p_did_list=[5, 6, 5, 7 + step], nc_dids_list=[[20, 21], [22, 6], [24, 25], [26, 27 + step]],
                 index_mapping={"query": [[4 * i] for i in range(4)], "pos_cand": [[4 * i + 1] for i in range(4)],
                                "neg_cand_list": [[4 * i + 2, 4 * i + 3] for i in range(4)]})
        orc.zero_grad()
        torch.manual_seed(seed); ref = orc.compute_contrastive_loss(b, alpha=0.3); ref["loss"].backward()
        if m._online.grad is not None:
            for gbuf in m._online.grad.values():
                gbuf.zero_()
        torch.manual_seed(seed); out = m.compute_contrastive_loss(b, alpha=0.3)
        m.backward_contrastive()
        torch.manual_seed(seed); flips.add(bool(torch.rand(1) < 0.5))
        torch.cuda.synchronize()
        assert abs(out["loss"].item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item())), (out["loss"].item(), ref["loss"].item())
        assert abs(out["accuracy"].item() - ref["accuracy"].item()) <= 1.0 / 4 + 1e-6
        worst = (1.0, "")
        for n_, p_o in orc.named_parameters():
            if p_o.grad is None or n_.endswith("self.key.bias") or p_o.grad.numel() == 1:
                continue
            a, c = p_o.grad.flatten().double(), m._online.g(n_).cpu().flatten().double()
            worst = min(worst, ((torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item(), n_))
        assert worst[0] > 0.995, worst
        assert torch.equal(m.idx_queue.cpu(), orc.idx_queue) and int(m.new_ptr_queue) == int(orc.new_ptr_queue)
        live = orc.idx_queue[0] != -100
        assert torch.nn.functional.cosine_similarity(m.cand_queue.cpu().t()[live], orc.cand_queue.t()[live], dim=-1).min() > 1 - 1e-3
        with torch.no_grad():
            for n_, p_o in orc.named_parameters():
                if p_o.grad is not None:
                    p_o -= 1e-2 * p_o.grad
                    m.get_parameter(n_).copy_(p_o)
    assert flips == {True, False}


def test_med_config_file_is_read_and_checked(tmp_path):
    """The reference builds its BERT from med_config.json (blip_ff.py:34-36); a file that disagrees with the built-in
    dimensions must be refused, not ignored."""
    import json
    from uniir_b200.blip_ff import BLIP_CONFIGS, BLIPFeatureFusion
    d = BLIP_CONFIGS["tiny-test"]
    good = dict(hidden_size=d.t_width, num_hidden_layers=d.t_layers, num_attention_heads=d.t_heads,
                intermediate_size=d.t_inter, vocab_size=d.vocab, max_position_embeddings=d.max_pos, hidden_act="gelu",
                layer_norm_eps=1e-12, hidden_dropout_prob=0.1)
    p = tmp_path / "med.json"
    p.write_text(json.dumps(good))
    BLIPFeatureFusion(med_config=str(p), vit="tiny-test", image_size=64, embed_dim=64, queue_size=16)
    p.write_text(json.dumps(dict(good, num_hidden_layers=d.t_layers + 1)))
    with pytest.raises(ValueError, match="does not match"):
        BLIPFeatureFusion(med_config=str(p), vit="tiny-test", image_size=64, embed_dim=64, queue_size=16)
    p.write_text(json.dumps(dict(good, hidden_act="relu")))
    with pytest.raises(ValueError, match="hidden_act"):
        BLIPFeatureFusion(med_config=str(p), vit="tiny-test", image_size=64, embed_dim=64, queue_size=16)
