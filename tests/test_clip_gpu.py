"""GPU parity of the CLIPScoreFusion drop-in against the CPU oracle: embeddings within 1e-3 cosine
(north_star tolerance), loss/accuracy, gradients, golden vectors, the reference's API contract,
gradient caching, CUDA graphs, fused AdamW and the autograd bridge."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

COS_TOL = 1e-3     # north_star: embeddings within 1e-3 cosine of the reference
LOSS_RTOL = 2e-2   # bf16 towers vs fp32 oracle, logits scaled by 100


def _batch(dims, bs, seed=2023):
    from uniir_b200.preprocess import synthetic_clip_tokens
    g = torch.Generator().manual_seed(seed)
    n = 2 * bs
    mt = torch.ones(n, dtype=torch.long); mv = torch.ones(n, dtype=torch.long)
    mt[1::5] = 0; mv[2::5] = 0
    return dict(txt_batched=synthetic_clip_tokens(n, g, dims.context, dims.vocab),
                image_batched=torch.randn(n, 3, dims.image_size, dims.image_size, generator=g),
                txt_mask_batched=mt, image_mask_batched=mv,
                index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})


def _cuda(b):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}


def _pair(name, mb, seed=0):
    from oracle.clip_oracle import CLIPScoreFusionOracle
    from uniir_b200.clip_sf import CLIPScoreFusion
    torch.manual_seed(seed)
    orc = CLIPScoreFusionOracle(name)
    mine = CLIPScoreFusion(name, device="cuda", micro_batch=mb)
    mine.load_state_dict(orc.state_dict(), strict=True)
    return orc, mine


def _check_against_oracle(name, bs, mb, min_grad_cos, skip_masked=True):
    orc, mine = _pair(name, mb)
    mine.skip_masked_towers = skip_masked
    b = _batch(orc.clip_model.dims, bs)
    ref = orc(b)
    ref["loss"].backward()
    with torch.no_grad():
        emb_o = orc.encode_multimodal_input(b["txt_batched"], b["image_batched"], b["txt_mask_batched"], b["image_mask_batched"])
    gb = _cuda(b)
    emb_m = mine.encode_multimodal_input(gb["txt_batched"], gb["image_batched"], gb["txt_mask_batched"], gb["image_mask_batched"]).cpu()
    cos = torch.nn.functional.cosine_similarity(emb_m, emb_o, dim=-1)
    assert cos.min().item() > 1 - COS_TOL, cos.min().item()
    out = mine(gb)
    assert out["loss"].ndim == 0 and out["accuracy"].ndim == 0
    assert abs(out["loss"].item() - ref["loss"].item()) < LOSS_RTOL * max(1.0, abs(ref["loss"].item()))
    out["loss"].backward()
    worst = 1.0
    for n_, p_o in orc.named_parameters():
        gm = mine.get_parameter(n_).grad
        assert gm is not None, n_
        a, c = p_o.grad.flatten().double(), gm.cpu().flatten().double()
        if a.numel() > 1:
            worst = min(worst, (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item())
        else:
            assert abs(a.item() - c.item()) < 2e-2 * max(1.0, abs(a.item())), n_
    assert worst > min_grad_cos, worst


def test_tiny_model_matches_oracle_with_and_without_microbatching():
    _check_against_oracle("tiny-test", 8, 16, 0.995)    # one micro-batch
    _check_against_oracle("tiny-test", 12, 8, 0.995)    # 3 micro-batches: gradient cache is exact


def test_tiny_model_matches_oracle_without_tower_skipping():
    """_batch() has text-only and image-only items: by default they never enter the masked tower (compacted
    rows); this is the reference-shaped path that encodes every row in both towers and multiplies by the mask."""
    _check_against_oracle("tiny-test", 12, 8, 0.995, skip_masked=False)


def test_skipping_masked_towers_changes_nothing():
    """rows with mask 0 are not encoded.  Embeddings must equal the encode-everything path
    (each row's tower output does not depend on its neighbours), loss equal, gradients equal up to the order of
    the atomic accumulations.  Mix: text-only, image-only, both, neither; counts that do not fill a micro-batch."""
    from uniir_b200.clip_sf import CLIPScoreFusion
    res = {}
    for skip in (False, True):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=5)
        m.skip_masked_towers = skip
        b = _batch(m.dims, 12, seed=7)
        g = torch.Generator().manual_seed(11)
        b["txt_mask_batched"] = (torch.rand(24, generator=g) < 0.6).long()
        b["image_mask_batched"] = (torch.rand(24, generator=g) < 0.5).long()
        b["txt_mask_batched"][0] = b["image_mask_batched"][0] = 0      # an item with neither modality
        gb = _cuda(b)
        emb = m.encode_multimodal_input(gb["txt_batched"], gb["image_batched"], gb["txt_mask_batched"], gb["image_mask_batched"])
        loss, acc = m._loss_forward(gb); m._loss_backward(gb, None)
        res[skip] = (emb.clone(), loss.item(), {n: m.clip_model.g(n).clone() for n, *_ in m.clip_model._specs})
    assert res[True][0][0].abs().max().item() == 0
    assert torch.allclose(res[True][0], res[False][0], atol=1e-6, rtol=1e-6)
    assert abs(res[True][1] - res[False][1]) < 1e-5 * max(1.0, abs(res[False][1]))
    for n, ga in res[False][2].items():
        gb_ = res[True][2][n]
        assert (ga - gb_).norm() <= 2e-3 * (ga.norm() + 1e-12), n


def test_retained_micro_batches_skip_the_reforward_and_change_nothing():
    """Gradient caching re-runs each micro-batch forward before its backward; micro-batches whose activations are
    kept from the embedding pass (extra encoders, as many as HBM holds) skip that.  Same loss, same gradients (up
    to the order of the atomic accumulations), fewer kernels; eager, graph capture and graph replay."""
    from uniir_b200 import lib as L
    from uniir_b200 import ops as ops_
    from uniir_b200.clip_sf import CLIPScoreFusion
    res = {}
    for keep in (0, 2, "auto"):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=5)
        m.retain_micro_batches = keep
        b = _batch(m.dims, 16, seed=7)
        b["txt_mask_batched"] = torch.ones(32, dtype=torch.long); b["image_mask_batched"] = torch.ones(32, dtype=torch.long)
        gb = _cuda(b)
        for it in range(3):          # eager, capture, replay
            if m.clip_model.grad is not None:
                for g_ in m.clip_model.grad.values():
                    g_.zero_()
            n0 = L.launch_count() + ops_.graph_launches()
            loss, acc = m._loss_forward(gb); m._loss_backward(gb, None)
            torch.cuda.synchronize()
            launches = L.launch_count() + ops_.graph_launches() - n0   # direct + replayed from CUDA graphs
        assert m._n_retained == {0: 0, 2: 2, "auto": 4}[keep]
        res[keep] = (loss.item(), {n: m.clip_model.g(n).clone() for n, *_ in m.clip_model._specs}, launches)
    for keep in (2, "auto"):
        assert abs(res[keep][0] - res[0][0]) < 1e-6 * max(1.0, abs(res[0][0]))
        for n, ga in res[0][1].items():
            assert (ga - res[keep][1][n]).norm() <= 2e-3 * (ga.norm() + 1e-12), (keep, n)
    assert res["auto"][2] < res[2][2] < res[0][2]     # each retained micro-batch saves one forward


def test_pinned_host_batch_is_uploaded_in_chunks_and_changes_nothing():
    """End-to-end path: a pinned host batch goes straight into the public call; the images are uploaded in
    micro-batch chunks on a copy stream (overlapped with the towers).  Same loss and gradients as a device batch."""
    from uniir_b200.clip_sf import CLIPScoreFusion
    res = {}
    for where in ("device", "pinned"):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=5)
        b = _batch(m.dims, 16, seed=9)
        b["txt_mask_batched"] = torch.ones(32, dtype=torch.long); b["image_mask_batched"] = torch.ones(32, dtype=torch.long)
        gb = _cuda(b) if where == "device" else {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()}
        for it in range(3):          # eager, capture, replay
            if m.clip_model.grad is not None:
                for g_ in m.clip_model.grad.values():
                    g_.zero_()
            loss, acc = m._loss_forward(gb); m._loss_backward(gb, None)
        torch.cuda.synchronize()
        assert (m._upload is None) and (where == "device" or m._h2d_stream is not None)
        res[where] = (loss.item(), {n: m.clip_model.g(n).clone() for n, *_ in m.clip_model._specs})
    assert abs(res["pinned"][0] - res["device"][0]) < 1e-6 * max(1.0, abs(res["device"][0]))
    for n, ga in res["device"][1].items():
        assert (ga - res["pinned"][1][n]).norm() <= 2e-3 * (ga.norm() + 1e-12), n


def test_vit_b32_batch4_matches_oracle():
    """BASELINE configs[0]: CLIP_SF ViT-B/32 in-batch contrastive step, batch 4."""
    _check_against_oracle("ViT-B/32", 4, 8, 0.995)


@pytest.mark.parametrize("masks", ["mixed", "ones"])
def test_vit_l14_dims_match_oracle(masks):
    """The headline configuration's architecture (BASELINE configs[1], ViT-L/14: 24 x 1024 vision, 12 x 768 text,
    257 / 77 tokens, embed 768) at 4 pairs (the training micro-batch is a multiple of 8 items): embeddings within
    1e-3 cosine, loss, every parameter gradient."""
    orc, mine = _pair("ViT-L/14", 8)
    b = _batch(orc.clip_model.dims, 4)
    if masks == "ones":
        b["txt_mask_batched"] = torch.ones(8, dtype=torch.long); b["image_mask_batched"] = torch.ones(8, dtype=torch.long)
    ref = orc(b)
    ref["loss"].backward()
    with torch.no_grad():
        emb_o = orc.encode_multimodal_input(b["txt_batched"], b["image_batched"], b["txt_mask_batched"], b["image_mask_batched"])
    gb = _cuda(b)
    emb_m = mine.encode_multimodal_input(gb["txt_batched"], gb["image_batched"], gb["txt_mask_batched"], gb["image_mask_batched"]).cpu()
    cos = torch.nn.functional.cosine_similarity(emb_m, emb_o, dim=-1)
    assert cos.min().item() > 1 - COS_TOL, cos.min().item()
    out = mine(gb)                                      # mixed masks: compacted towers; ones: fused two-tower pass
    assert abs(out["loss"].item() - ref["loss"].item()) < LOSS_RTOL * max(1.0, abs(ref["loss"].item()))
    out["loss"].backward()
    worst, worst_name = 1.0, None
    for n_, p_o in orc.named_parameters():
        gm = mine.get_parameter(n_).grad
        assert gm is not None, n_
        a, c = p_o.grad.flatten().double(), gm.cpu().flatten().double()
        if a.numel() > 1 and a.norm() > 0:
            cs = (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item()
            if cs < worst:
                worst, worst_name = cs, n_
    assert worst > 0.99, (worst, worst_name)


def test_golden_vectors():
    from uniir_b200.clip_sf import CLIPScoreFusion
    g = np.load(os.path.join(GOLDEN, "clip_tiny.npz"))
    m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=16)
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}, strict=True)
    bs = g["ids"].shape[0] // 2
    b = dict(txt_batched=torch.from_numpy(g["ids"]).cuda(), image_batched=torch.from_numpy(g["img"]).cuda(),
             txt_mask_batched=torch.from_numpy(g["mt"]).cuda(), image_mask_batched=torch.from_numpy(g["mv"]).cuda(),
             index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})
    emb = m.encode_multimodal_input(b["txt_batched"], b["image_batched"], b["txt_mask_batched"], b["image_mask_batched"]).cpu()
    cos = torch.nn.functional.cosine_similarity(emb, torch.from_numpy(g["emb"]), dim=-1)
    assert cos.min().item() > 1 - COS_TOL
    out = m(b)
    assert abs(out["loss"].item() - float(g["loss"])) < LOSS_RTOL * max(1.0, float(g["loss"]))
    out["loss"].backward()
    for k in g.files:
        if k.startswith("g:") and g[k].size > 1:
            a, c = torch.from_numpy(g[k]).flatten().double(), m.get_parameter(k[2:]).grad.cpu().flatten().double()
            assert (torch.dot(a, c) / (a.norm() * c.norm())).item() > 0.995, k


def test_reference_api_contract():
    from uniir_b200.clip_sf import CLIPScoreFusion
    m = CLIPScoreFusion("tiny-test", device="cuda")
    d = m.dims
    b = _cuda(_batch(d, 3))
    b.pop("index_mapping")
    b["did_list"] = [10000001 + i for i in range(6)]
    emb, ids = m(b, encode_mbeir_batch=True)
    assert emb.shape == (6, d.embed_dim) and emb.dtype == torch.float32 and ids == b["did_list"]
    b2 = dict(b); b2.pop("did_list"); b2["qid_list"] = [5] * 6
    assert m(b2, encode_mbeir_batch=True)[1] == [5] * 6
    b3 = dict(b); b3.pop("did_list")
    with pytest.raises(AssertionError):
        m(b3, encode_mbeir_batch=True)
    b4 = dict(b); b4["did_list"] = ["1:1"] * 6
    with pytest.raises(AssertionError):
        m(b4, encode_mbeir_batch=True)
    assert callable(m.get_tokenizer()) and callable(m.get_img_preprocess_fn())
    assert abs(m.get_logit_scale().item() - 100.0) < 1e-3
    assert all(k.startswith("clip_model.") for k in m.state_dict())
    # masks: a text-only item equals encode_text, an image-only item equals encode_image
    bb = _cuda(_batch(d, 2))
    full = m.encode_multimodal_input(bb["txt_batched"], bb["image_batched"], 0 * bb["txt_mask_batched"] + 1, 0 * bb["image_mask_batched"])
    assert torch.allclose(full, m.encode_text(bb["txt_batched"]), atol=1e-5)


def test_train_step_graphs_equal_eager_and_loss_decreases():
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.engine import FlatAdamW, train_step
    res = {}
    for use in (False, True):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=1)
        m.use_cuda_graphs = use
        opt = FlatAdamW(m.clip_model, lr=1e-3)
        gb = _cuda(_batch(m.dims, 16))
        res[use] = [train_step(m, gb, opt)[0].item() for _ in range(5)]
    # Same kernels, same inputs: the first loss is identical.  Gradients are reproducible only to ~5e-7 relative
    # (fp32 atomics in the split-K weight gradients and column sums; scripts/check_determinism.py), and five
    # lr=1e-3 Adam steps on this tiny model amplify that: two runs of the SAME mode differ by up to ~1 % at
    # step 5 (gpurun_out/c10_modes.log), so later steps are compared loosely.
    assert res[False][0] == res[True][0]
    assert all(abs(a - b) < 1e-3 * abs(a) for a, b in zip(res[False][:2], res[True][:2]))
    assert all(abs(a - b) < 3e-2 * max(1, abs(a)) for a, b in zip(res[False], res[True]))
    assert res[True][-1] < 0.2 * res[True][0]


def test_autograd_bridge_equals_fused_path_and_trains_with_torch_adamw():
    """(a) reference-style loop: model(batch)["loss"].backward() fills p.grad for a stock torch optimizer;
    (b) fused path: flat gradient buffers.  Same model + batch -> same gradients (up to the summation order of
    the split-K / LayerNorm atomics).  Then the reference-style loop must train.  (Post-Adam parameters are
    not compared element-wise: Adam's first steps are +-lr*sign(g), so elements with |g| ~ 0 amplify
    summation-order noise; the fused AdamW kernel itself is pinned in test_kernels_gpu.py.)"""
    from uniir_b200.clip_sf import CLIPScoreFusion
    ma = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=3)
    mb_ = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=3)
    gb = _cuda(_batch(ma.dims, 8))
    la = ma(gb)["loss"]; la.backward()
    lb, _ = mb_._loss_forward(gb); mb_._loss_backward(gb, None)
    assert abs(la.item() - lb.item()) < 1e-5 * max(1, abs(la.item()))
    for n, pa in ma.clip_model.named_parameters():
        ga, gbf = pa.grad.flatten().double(), mb_.clip_model.g(n).flatten().double()
        assert (ga - gbf).norm() <= 1e-3 * (gbf.norm() + 1e-12), n
    decay = [p for n, p in ma.named_parameters() if not (p.ndim < 2 or "ln" in n or "bias" in n or "logit_scale" in n)]
    nod = [p for n, p in ma.named_parameters() if (p.ndim < 2 or "ln" in n or "bias" in n or "logit_scale" in n)]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.2}, {"params": nod, "weight_decay": 0.0}],
                            lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    losses = [la.item()]
    for _ in range(4):
        opt.step(); ma.zero_grad()
        l = ma(gb)["loss"]; l.backward(); losses.append(l.item())
    assert losses[-1] < 0.2 * losses[0], losses


def test_infonce_kernels_match_golden():
    from uniir_b200 import ops
    g = np.load(os.path.join(GOLDEN, "infonce.npz"))
    emb = torch.from_numpy(g["emb"]).cuda()
    bs, E = emb.shape[0] // 2, emb.shape[1]
    q_idx = torch.arange(0, 2 * bs, 2, device="cuda").int(); p_idx = q_idx + 1
    ls = torch.tensor(np.log(100.0), device="cuda", dtype=torch.float32)
    for world, rank, loss_k, acc_k, g_k in ((1, 0, "loss1", "acc1", "g1"), (2, 1, "loss2", "acc2", "g2")):
        nce = ops.InfoNCE(bs, bs * world, E, "cuda")
        if world == 2:
            nce.all_p[:bs] = torch.from_numpy(g["other"]).cuda()
        off = rank * bs
        nce.normalize(emb, q_idx, p_idx, off)
        for _ in range(2):   # second launch proves the workspace self-resets
            loss, acc = nce.forward(ls, off)
        assert abs(loss.item() - float(g[loss_k])) < 1e-4 and abs(acc.item() - float(g[acc_k])) < 1e-6
        dls = torch.zeros((), device="cuda")
        _, d_allp = nce.backward(ls, off, dls)
        d_emb = torch.zeros_like(emb)
        nce.normalize_bwd(q_idx, p_idx, d_allp[off:off + bs], d_emb, off)
        ref = torch.from_numpy(g[g_k])
        assert ((d_emb.cpu() - ref).norm() / ref.norm()).item() < 1e-4


def test_hard_negative_branch_matches_oracle():
    """clip_sf.py:105-131 incl. the reference's in-batch-negative indexing (own positive repeated k times)."""
    from types import SimpleNamespace as NS
    from oracle.clip_oracle import CLIPScoreFusionOracle
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.preprocess import synthetic_clip_tokens
    torch.manual_seed(0)
    cfg = NS(model=NS(gather_embeddings=False), data_config=NS(in_batch_neg_num=2))
    orc = CLIPScoreFusionOracle("tiny-test", config=cfg)
    mine = CLIPScoreFusion("tiny-test", device="cuda", config=cfg, micro_batch=16)
    mine.load_state_dict(orc.state_dict(), strict=True)
    d = orc.clip_model.dims
    bs, neg = 4, 2
    n = bs * (2 + neg)
    g = torch.Generator().manual_seed(9)
    b = dict(txt_batched=synthetic_clip_tokens(n, g, d.context, d.vocab),
             image_batched=torch.randn(n, 3, d.image_size, d.image_size, generator=g),
             txt_mask_batched=torch.ones(n, dtype=torch.long), image_mask_batched=torch.ones(n, dtype=torch.long),
             index_mapping={"query": [[4 * i] for i in range(bs)], "pos_cand": [[4 * i + 1] for i in range(bs)],
                            "neg_cand_list": [[4 * i + 2, 4 * i + 3] for i in range(bs)]})
    ref = orc(b); ref["loss"].backward()
    out = mine(_cuda(b)); out["loss"].backward()
    assert abs(out["loss"].item() - ref["loss"].item()) < LOSS_RTOL * max(1.0, abs(ref["loss"].item()))
    assert abs(out["accuracy"].item() - ref["accuracy"].item()) < 1e-6
    worst = 1.0
    for n_, p_o in orc.named_parameters():
        a, c = p_o.grad.flatten().double(), mine.get_parameter(n_).grad.cpu().flatten().double()
        if a.numel() > 1:
            worst = min(worst, (torch.dot(a, c) / (a.norm() * c.norm() + 1e-30)).item())
    assert worst > 0.99, worst


def test_hard_negative_kernel_matches_reference_golden():
    """tests/golden/infonce.npz `hn_*`: produced by the reference's own clip_sf.py:105-131 (scripts/make_golden.py
    asserts the oracle agrees) -- loss, accuracy and d_emb of the fused hard-negative kernel."""
    from uniir_b200 import ops
    g = np.load(os.path.join(GOLDEN, "infonce.npz"))
    emb = torch.from_numpy(g["hn_emb"]).cuda()
    bs = emb.shape[0] // 4
    base = torch.arange(bs, device="cuda", dtype=torch.int32) * 4
    n_idx = torch.stack([base + 2, base + 3], dim=1).contiguous()
    ls = torch.tensor(np.log(100.0), device="cuda", dtype=torch.float32)
    dls = torch.zeros((), device="cuda")
    loss, acc, d_emb = ops.hardneg_loss(emb, base, base + 1, n_idx, int(g["hn_in_batch_neg_num"]), ls, dls)
    assert abs(loss.item() - float(g["hn_loss"])) < 1e-4 and abs(acc.item() - float(g["hn_acc"])) < 1e-6
    ref = torch.from_numpy(g["hn_g"])
    assert ((d_emb.cpu() - ref).norm() / ref.norm()).item() < 1e-4
