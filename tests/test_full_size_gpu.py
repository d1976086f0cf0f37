"""BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle needs hours there):

  configs[4]  190 000 queries x 5 600 000 candidates x 768, top-50 (reference mbeir_retriever.py:188-232, FAISS Flat IP
              over L2-normalised rows): sorted / unique / valid ids over every row, planted neighbours recovered in
              order, returned scores re-computed in fp32, a sample of rows against a brute-force fp64 scan of all
              5.6 M candidates, and row independence (a row's answer does not depend on the batch around it);
  configs[1]  4096 pairs through CLIP_SF ViT-L/14 (reference clip_sf.py:68-147): loss / accuracy re-computed in fp32
              from the embeddings, the closed-form logit-scale gradient, invariance of loss and gradient under a
              permutation of the pairs and under a different micro-batch tiling of the gradient cache.
"""
import gc
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_hbm(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * (1 << 30):
        pytest.skip(f"needs {gb} GB of free HBM")


def test_retrieval_at_configs4_size():
    from uniir_b200.retrieval import FlatIPIndex
    gc.collect(); torch.cuda.empty_cache()
    _need_hbm(60)
    Q, N, d, k = 190_000, 5_600_000, 768, 50
    g = torch.Generator(device="cuda").manual_seed(4)
    c = torch.empty(N, d, device="cuda", dtype=torch.float16)
    for s in range(0, N, 700_000):
        c[s:s + 700_000] = torch.randn(min(700_000, N - s), d, generator=g, device="cuda").half()
    q = torch.randn(Q, d, generator=g, device="cuda").half()
    # plant, for the first P queries, 50 candidates at cosine 0.90, 0.89 .. 0.41 (random rows stay below ~0.25)
    P = 1024
    qn = torch.nn.functional.normalize(q[:P].float(), dim=-1)
    rows = torch.randperm(N, generator=g, device="cuda")[:P * k].view(P, k)
    cos = (0.90 - 0.01 * torch.arange(k, device="cuda")).view(1, k, 1)
    noise = torch.randn(P, k, d, generator=g, device="cuda")
    noise = torch.nn.functional.normalize(noise - (noise * qn[:, None]).sum(-1, keepdim=True) * qn[:, None], dim=-1)
    c[rows.view(-1)] = ((cos * qn[:, None] + (1 - cos ** 2).sqrt() * noise) * 3.0).view(-1, d).half()
    del noise
    ids = torch.arange(N, device="cuda", dtype=torch.int64) * 3 + 1
    index = FlatIPIndex(d, "cuda")
    index.add_with_ids(c, ids)
    D, I = index.search_device(q, k)
    assert D.shape == (Q, k) and I.shape == (Q, k) and D.dtype == torch.float32 and I.dtype == torch.int64
    # every row: sorted, finite cosines, valid ids, no id twice
    assert (D[:, :-1] >= D[:, 1:]).all() and D.abs().max().item() <= 1 + 1e-5
    assert ((I - 1) % 3 == 0).all() and (I >= 1).all() and (I < 3 * N + 1).all()
    Is = I.sort(dim=1).values
    assert (Is[:, 1:] != Is[:, :-1]).all()
    del Is
    # planted neighbours: exact ids in exact order
    assert (I[:P] == ids[rows]).all()
    assert (D[:P, 0] - 0.90).abs().max().item() < 5e-3
    # returned scores are the fp32 cosines of the rows they name (8192 sampled queries)
    sel = torch.randperm(Q, generator=g, device="cuda")[:8192]
    cn = torch.nn.functional.normalize(c[((I[sel] - 1) // 3).view(-1)].float(), dim=-1).view(-1, k, d)
    ref = (cn * torch.nn.functional.normalize(q[sel].float(), dim=-1)[:, None]).sum(-1)
    assert (ref - D[sel]).abs().max().item() < 5e-6
    del cn, ref
    # 256 unplanted queries against a brute-force fp64 scan of all 5.6 M candidates
    bf = P + torch.randperm(Q - P, generator=g, device="cuda")[:256]
    qb = torch.nn.functional.normalize(q[bf].double(), dim=-1)
    best_s = torch.zeros(256, 0, dtype=torch.float64, device="cuda"); best_i = torch.zeros(256, 0, dtype=torch.int64, device="cuda")
    for s in range(0, N, 800_000):
        sc = qb @ torch.nn.functional.normalize(c[s:s + 800_000].double(), dim=-1).T
        ts, ti = sc.topk(k, dim=1)
        best_s = torch.cat([best_s, ts], 1); best_i = torch.cat([best_i, ids[s + ti]], 1)
        ts, pos = best_s.topk(k, dim=1)
        best_s, best_i = ts, best_i.gather(1, pos)
    assert (best_s - D[bf]).abs().max().item() < 5e-6
    pad = torch.full((256, 1), 2.0, dtype=torch.float64, device="cuda")
    gap = torch.minimum((best_s - torch.cat([pad, best_s[:, :-1]], 1)).abs(), (best_s - torch.cat([best_s[:, 1:], -pad], 1)).abs())
    differ = best_i != I[bf]
    assert not (differ & (gap > 2e-5)).any()          # ids may only swap between scores closer than fp32 noise
    assert differ.float().mean().item() < 2e-3
    # row independence: the same queries searched on their own give the same lists
    Ds, Is = index.search_device(q[sel].contiguous(), k)
    assert torch.equal(Ds, D[sel])
    assert (Is == I[sel]).float().mean().item() > 0.9999


def _flat_grads(m):
    return {n: m.clip_model.g(n).detach().clone() for n, *_ in m.clip_model._specs}


def _step_grads(m, batch):
    if m.clip_model.grad is not None:
        for g_ in m.clip_model.grad.values():
            g_.zero_()
    loss, acc = m._loss_forward(batch)
    emb = m._train_state["emb"].clone()          # the embeddings the loss was computed from (gradient-cache phase A)
    m._loss_backward(batch, None)
    torch.cuda.synchronize()
    return loss.item(), acc.item(), _flat_grads(m), emb


def _rel_err(ga, gb):
    num = sum((ga[n].double() - gb[n].double()).pow(2).sum() for n in ga).sqrt().item()
    den = sum(ga[n].double().pow(2).sum() for n in ga).sqrt().item()
    return num / den


def test_training_step_at_configs1_size():
    sys.path.insert(0, ROOT)
    import bench
    from uniir_b200.clip_sf import CLIPScoreFusion
    gc.collect(); torch.cuda.empty_cache()
    _need_hbm(120)
    pairs = 4096
    dev = torch.device("cuda", torch.cuda.current_device())
    m = CLIPScoreFusion("ViT-L/14", device=dev, micro_batch=128, seed=11)
    m.retain_micro_batches = 1
    batch = bench.synth_batch(m.dims, pairs, 2023, dev)
    r = {}
    loss, acc, g0, emb = _step_grads(m, batch)
    r["finite"] = bool(all(torch.isfinite(v).all() for v in g0.values()) and torch.isfinite(emb).all())
    # loss / accuracy from the embeddings, in fp64 torch (clip_sf.py:88-101, 133-144)
    qn = torch.nn.functional.normalize(emb[0::2].double(), dim=-1); pn = torch.nn.functional.normalize(emb[1::2].double(), dim=-1)
    scale = m.get_logit_scale().double().reshape(())
    cosm = qn @ pn.T
    logits = cosm * scale
    tgt = torch.arange(pairs, device=dev)
    r["loss"], r["loss_ref"] = loss, torch.nn.functional.cross_entropy(logits, tgt).item()
    r["acc"], r["acc_ref"] = acc, (logits.argmax(1) == tgt).float().mean().item()
    # d loss / d logit_scale in closed form: scale * mean_i(sum_j softmax_ij cos_ij - cos_ii)
    p = logits.softmax(dim=1)
    r["dls_ref"] = (scale * ((p * cosm).sum(1) - cosm.diagonal()).mean()).item()
    r["dls"] = g0["logit_scale"].double().sum().item()
    # the eval-mode encoder gives the same embeddings as the training pass (clip_sf.py:53-63), within bf16 noise
    emb_eval = m.encode_multimodal_input(batch["txt_batched"][:512], batch["image_batched"][:512],
                                         batch["txt_mask_batched"][:512], batch["image_mask_batched"][:512])
    r["eval_cos_min"] = torch.nn.functional.cosine_similarity(emb_eval, emb[:512], dim=-1).min().item()
    del emb, emb_eval, cosm, logits, p
    # the in-batch loss does not depend on the order of the pairs; nor does the summed gradient
    perm = torch.randperm(pairs, generator=torch.Generator().manual_seed(5)).to(dev)
    items = torch.stack([2 * perm, 2 * perm + 1], 1).flatten()
    pb = dict(batch, txt_batched=batch["txt_batched"][items].contiguous(), image_batched=batch["image_batched"][items].contiguous())
    r["loss_perm"], r["acc_perm"], g1, _ = _step_grads(m, pb)
    del pb
    r["grad_rel_err_perm"] = _rel_err(g0, g1)
    del g1
    # a different tiling of the gradient cache (64-item micro-batches, none retained): same loss, same gradient
    m.micro_batch = 64
    m.retain_micro_batches = 0
    gc.collect(); torch.cuda.empty_cache()
    r["loss_mb64"], r["acc_mb64"], g2, _ = _step_grads(m, batch)
    r["grad_rel_err_mb64"] = _rel_err(g0, g2)
    print("configs[1] full-size properties:", r)
    assert r["finite"]
    assert abs(r["loss"] - r["loss_ref"]) < 2e-5 * max(1.0, abs(r["loss_ref"])), r
    assert abs(r["acc"] - r["acc_ref"]) <= 2.0 / pairs, r
    assert abs(r["dls"] - r["dls_ref"]) < 1e-4 * abs(r["dls_ref"]) + 2e-6, r
    assert r["eval_cos_min"] > 1 - 1e-3, r
    for tag in ("perm", "mb64"):
        assert abs(r["loss_" + tag] - r["loss"]) < 2e-5 * max(1.0, abs(r["loss"])), r
        assert abs(r["acc_" + tag] - r["acc"]) <= 2.0 / pairs, r
        assert r["grad_rel_err_" + tag] < 2e-3, r
