"""The epoch-level loops with the reference's signatures: CLIP `train_one_epoch` / `eval_engine`
(uniir_clip/engine.py:7-84), BLIP `train_one_epoch` with the alpha ramp and `eval_engine` with the queue snapshot
(uniir_blip/engine.py:9-114), `generate_embeds_and_ids_for_dataset_with_gather` (mbeir_embedder.py:32-120) and
`build_model_from_config` (common/utils.py:64-152)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import NO_DROP

pytestmark = pytest.mark.gpu


def _clip_batches(dims, n_batches, bs, seed=0):
    from uniir_b200.preprocess import synthetic_clip_tokens
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        n = 2 * bs
        out.append(dict(txt_batched=synthetic_clip_tokens(n, g, dims.context, dims.vocab),
                        image_batched=torch.randn(n, 3, dims.image_size, dims.image_size, generator=g),
                        txt_mask_batched=torch.ones(n, dtype=torch.long), image_mask_batched=torch.ones(n, dtype=torch.long),
                        index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]}))
    return out


def test_clip_train_one_epoch_accumulates_and_matches_manual_steps():
    """Two accumulation steps per optimizer step, cosine schedule: the epoch loop must leave the same weights as the
    same sequence of train_step calls, report the mean loss, and advance the scheduler once per optimizer step."""
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.engine import CosineLR, FlatAdamW, eval_engine, train_one_epoch, train_step
    cfg = NS(trainer_config=NS(print_freq=2, gradient_accumulation_steps=2), evaluator=NS(print_freq=2))
    res = {}
    for how in ("epoch", "manual"):
        m = CLIPScoreFusion("tiny-test", device="cuda", micro_batch=8, seed=7)
        opt = FlatAdamW(m.clip_model, lr=1e-3)
        sched = CosineLR(opt, T_max=4)
        batches = _clip_batches(m.dims, 4, 8, seed=1)
        if how == "epoch":
            stats = train_one_epoch(m, [dict(b) for b in batches], opt, 0, "cuda:0", sched, 0, None, cfg)
            assert set(stats) >= {"loss", "inbatch_accuracy", "lr"} and np.isfinite(stats["loss"])
            res["stats"] = stats
        else:
            losses = []
            for i, b in enumerate(batches):
                gb = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
                loss, _ = train_step(m, gb, opt, sched, accumulation_steps=2, do_step=(i % 2 == 1))
                losses.append(loss.item())
            res["mean_loss"] = float(np.mean(losses))
        assert opt.step_count == 2 and sched.t == 2
        res[how] = {n: m.clip_model.w(n).clone() for n, *_ in m.clip_model._specs}
        res[how + "_model"] = m
    assert abs(res["stats"]["loss"] - res["mean_loss"]) < 1e-3 * max(1.0, abs(res["mean_loss"]))
    for n, w in res["manual"].items():
        assert (w - res["epoch"][n]).norm() <= 2e-3 * w.norm() + 2e-4 * w.numel() ** 0.5, n
    # eval_engine: no parameter changes, same numbers as direct no-grad forwards
    m = res["epoch_model"]
    before = {n: m.clip_model.w(n).clone() for n, *_ in m.clip_model._specs}
    val = _clip_batches(m.dims, 3, 8, seed=5)
    stats = eval_engine(m, [dict(b) for b in val], "cuda:0", cfg)
    with torch.no_grad():
        direct = [m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}) for b in val]
    assert abs(stats["loss"] - float(np.mean([o["loss"].item() for o in direct]))) < 1e-5
    assert abs(stats["inbatch_accuracy"] - float(np.mean([o["accuracy"].item() for o in direct]))) < 1e-6
    for n, w in before.items():
        assert torch.equal(w, m.clip_model.w(n)), n


def _blip_batches(d, n_batches, seed0=40):
    from test_oracle_blip import _inputs
    out = []
    for s in range(n_batches):
        ids, att, img = _inputs(d, 16, 24, seed0 + s)
        out.append(dict(txt_batched={"input_ids": ids, "attention_mask": att}, image_batched=img, txt_mask_batched=None,
                        image_mask_batched=None, p_did_list=torch.tensor([5, 6, 5, 7, 8, 9, 10, 11 + s]),
                        index_mapping={"query": [[2 * i] for i in range(8)], "pos_cand": [[2 * i + 1] for i in range(8)]}))
    return out


def test_blip_train_one_epoch_alpha_ramp_and_eval_engine_restores_state():
    from uniir_b200 import blip_engine
    from uniir_b200.blip_ff import BLIPFeatureFusion
    from uniir_b200.engine import CosineLR, FlatAdamW, train_step_blip
    cfg = NS(trainer_config=NS(print_freq=10, gradient_accumulation_steps=1), evaluator=NS(print_freq=10), model=NS(alpha=0.4))
    assert blip_engine.epoch_alpha(cfg, 0, 0, 4) == 0.0 and blip_engine.epoch_alpha(cfg, 0, 2, 4) == 0.2
    assert blip_engine.epoch_alpha(cfg, 1, 0, 4) == 0.4
    res = {}
    for how in ("epoch", "manual"):
        m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=32, micro_batch=8, seed=11)
        m.train_micro_batch = 8
        opt = FlatAdamW(m._online, lr=1e-3, weight_decay=0.05)
        sched = CosineLR(opt, T_max=8)
        batches = _blip_batches(m.dims, 3)
        if how == "epoch":
            stats = blip_engine.train_one_epoch(m, [dict(b) for b in batches], opt, 0, "cuda:0", sched, 0, None, cfg)
            assert np.isfinite(stats["loss"]) and 0.0 <= stats["inbatch_accuracy"] <= 1.0
        else:
            for i, b in enumerate(batches):
                train_step_blip(m, dict(b), opt, cfg.model.alpha * min(1, i / 3), sched)
        torch.cuda.synchronize()
        assert opt.step_count == 3 and int(m.new_ptr_queue) == 24
        res[how] = {k: v.detach().clone() for k, v in m.state_dict().items()}
        res[how + "_model"] = m
    for k, v in res["manual"].items():
        w = res["epoch"][k]
        if v.dtype.is_floating_point:
            # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is ~0 amplify the
            # summation-order noise of the atomics, so allow a fraction of the 3-step update (3e-3 per element)
            assert (v - w).norm() <= 5e-3 * v.norm() + 3e-4 * v.numel() ** 0.5, k
        else:
            assert torch.equal(v, w), k
    # eval_engine: numbers come back, every parameter / queue / pointer is exactly as before
    m = res["epoch_model"]
    before = {k: v.detach().clone() for k, v in m.state_dict().items()}
    stats = blip_engine.eval_engine(m, m, [dict(b) for b in _blip_batches(m.dims, 2, seed0=70)], "cuda:0", cfg)
    assert np.isfinite(stats["loss"]) and 0.0 <= stats["inbatch_accuracy"] <= 1.0
    for k, v in before.items():
        assert torch.equal(v, m.state_dict()[k]), k


def test_generate_embeds_with_gather_world1_and_build_model_from_config(tmp_path):
    from uniir_b200 import config as cfg_mod
    from uniir_b200.clip_sf import CLIPScoreFusion
    from uniir_b200.mbeir_embedder import build_model_from_config, generate_embeds_and_ids_for_dataset_with_gather
    m = CLIPScoreFusion("tiny-test", device="cuda", seed=3)
    ckpt_dir = tmp_path / "checkpoint" / "X"
    os.makedirs(ckpt_dir)
    torch.save({"model": m.state_dict()}, str(ckpt_dir / "w.pth"))
    c = cfg_mod.create({"uniir_dir": str(tmp_path), "model": {"name": "CLIPScoreFusion", "clip_vision_model_name": "tiny-test",
                        "pretrained_clip_model_dir": "checkpoint/CLIP/", "ckpt_config": {"ckpt_dir": "checkpoint/X", "ckpt_name": "w.pth"}}})
    m2 = build_model_from_config(c)
    for (k, a), (k2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    c.model.name = "NoSuchModel"
    with pytest.raises(NotImplementedError):
        build_model_from_config(c)
    c.model.name = "CLIPScoreFusion"; c.model.ckpt_config.ckpt_name = "missing.pth"
    with pytest.raises(AssertionError):
        build_model_from_config(c)
    batches = _clip_batches(m.dims, 3, 5, seed=9)          # 10 items per batch
    loader = []
    for i, b in enumerate(batches):
        b = dict(b); b.pop("index_mapping"); b["did_list"] = [1000 * (i + 1) + j for j in range(10)]
        loader.append(b)
    emb, ids = generate_embeds_and_ids_for_dataset_with_gather(m2, [dict(b) for b in loader], "cuda:0", use_fp16=True)
    assert emb.dtype == np.float16 and emb.shape == (30, m.dims.embed_dim) and ids == [i for b in loader for i in b["did_list"]]
    direct = torch.cat([m(dict((k, v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()), encode_mbeir_batch=True)[0] for b in loader])
    assert np.array_equal(emb, direct.half().cpu().numpy())
    loader[1]["did_list"][0] = loader[0]["did_list"][0]       # duplicate id -> the reference's uniqueness assertion
    with pytest.raises(AssertionError):
        generate_embeds_and_ids_for_dataset_with_gather(m2, [dict(b) for b in loader], "cuda:0")


def test_blip_autograd_bridge_fills_param_grads_like_the_fused_path():
    """The reference loop's `loss.backward()` (uniir_blip/engine.py:35-56) on the BLIP arm: p.grad of every online
    parameter equals the flat gradient buffers of the fused path; momentum / frozen parameters get none."""
    from uniir_b200.blip_ff import BLIPFeatureFusion
    res = {}
    for how in ("bridge", "fused"):
        m = BLIPFeatureFusion(med_config=NO_DROP, vit="tiny-test", image_size=64, embed_dim=64, queue_size=32, micro_batch=8, seed=13)
        m.train_micro_batch = 8
        b = _blip_batches(m.dims, 1, seed0=90)[0]
        if how == "bridge":
            out = m(b, alpha=0.3)
            assert out["loss"].requires_grad and not out["accuracy"].requires_grad
            (out["loss"] * 2.0).backward()
            res[how] = (out["loss"].item(), {n: m.get_parameter(n).grad.clone() for n, *_ in m._online._specs})
            assert all(m.get_parameter(n).grad is None for n, *_ in m._mom._specs)
        else:
            loss, _ = m._loss_forward(b, 0.3)
            m.backward_contrastive()
            res[how] = (loss.item(), {n: 2.0 * m._online.g(n).clone() for n, *_ in m._online._specs})
    assert abs(res["bridge"][0] - res["fused"][0]) < 1e-5 * max(1.0, abs(res["fused"][0]))
    for n, g in res["fused"][1].items():
        assert (g - res["bridge"][1][n]).norm() <= 2e-3 * g.norm() + 1e-7, n
