"""Pin the BLIP_FF oracle restatement: (1) against the reference's own vendored modules imported through
shims (only where /root/reference exists), (2) against the committed golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

import ref_blip_shim
from conftest import GOLDEN
from oracle.blip_oracle import BLIP_CONFIGS, BLIPFeatureFusionOracle, BLIPScoreFusionOracle


def _inputs(d, B, S, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(B, S, dtype=torch.long); att = torch.zeros(B, S, dtype=torch.long)
    for i in range(B):
        L = int(torch.randint(1, S - 2, (1,), generator=g))
        ids[i, 0] = 101; ids[i, 1:1 + L] = torch.randint(1000 % d.vocab, d.vocab - 2, (L,), generator=g); ids[i, 1 + L] = 102
        att[i, :L + 2] = 1
    img = torch.randn(B, 3, d.image_size, d.image_size, generator=g)
    return ids, att, img


@pytest.mark.skipif(not ref_blip_shim.available(), reason="reference tree not present (GPU box): golden vectors cover it")
def test_restatement_matches_reference_modules(tmp_path):
    vit, med = ref_blip_shim.load()
    d = BLIP_CONFIGS["tiny-test"]
    torch.manual_seed(0)
    orc = BLIPFeatureFusionOracle("tiny-test").eval()
    from functools import partial
    rv = vit.VisionTransformer(img_size=d.image_size, patch_size=d.patch, embed_dim=d.v_width, depth=d.v_layers,
                               num_heads=d.v_heads, mlp_ratio=4, qkv_bias=True,
                               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).eval()
    cfg = dict(json.load(open("/root/reference/src/models/uniir_blip/backbone/configs/med_config.json")))
    cfg.update(hidden_size=d.t_width, intermediate_size=d.t_inter, num_attention_heads=d.t_heads,
               num_hidden_layers=d.t_layers, vocab_size=d.vocab, max_position_embeddings=d.max_pos, encoder_width=d.v_width)
    p = tmp_path / "med.json"; p.write_text(json.dumps(cfg))
    rb = med.BertModel(config=med.BertConfig.from_json_file(str(p)), add_pooling_layer=True).eval()
    rv.load_state_dict(orc.visual_encoder.state_dict(), strict=True)
    miss, unexp = rb.load_state_dict(orc.text_encoder.state_dict(), strict=False)
    assert not unexp and all("position_ids" in m or "token_type" in m for m in miss), (miss, unexp)
    ids, att, img = _inputs(d, 3, 20, 1)
    with torch.no_grad():
        v_ref = rv(img)
        v_orc = orc.visual_encoder(img)
        e_ref = rb(ids, attention_mask=att, encoder_hidden_states=v_ref,
                   encoder_attention_mask=torch.ones(v_ref.shape[:-1], dtype=torch.long), return_dict=True).pooler_output
        e_orc = orc.encode_multimodal_input(ids, att, img)
    assert torch.allclose(v_orc, v_ref, atol=1e-5, rtol=1e-4), (v_orc - v_ref).abs().max()
    assert torch.allclose(e_orc, e_ref, atol=1e-5, rtol=1e-4), (e_orc - e_ref).abs().max()


class _SeqMasks:
    """Keep decisions handed out in call order (call i draws from a generator seeded base + i): the reference's
    nn.Dropout / DropPath modules and the oracle's hook see the same bits iff they sit at the same sites."""

    def __init__(self, base):
        self.base, self.i = base, 0

    def keep(self, shape, p):
        g = torch.Generator().manual_seed(self.base + self.i)
        self.i += 1
        return (torch.rand(shape, generator=g) >= p).float()


@pytest.mark.skipif(not ref_blip_shim.available(), reason="reference tree not present (GPU box)")
def test_dropout_sites_match_reference_modules_in_train_mode(tmp_path, monkeypatch):
    """Train mode: med.py:97,217,244,325 (hidden / attention-probability dropout 0.1) and vit.py:151-152 (DropPath,
    linspace(0, 0.3, depth) here).  The reference modules and the oracle get the same keep decisions in call order;
    outputs and gradients agree only if every dropout of the reference has its hook at the same place in the oracle."""
    import oracle.blip_oracle as bo
    vit, med = ref_blip_shim.load()
    import sys
    DropPath = sys.modules["timm.models.layers"].DropPath
    d = BLIP_CONFIGS["tiny-test"]
    torch.manual_seed(0)
    orc = BLIPFeatureFusionOracle("tiny-test").train()
    from functools import partial
    rate = 0.3
    rv = vit.VisionTransformer(img_size=d.image_size, patch_size=d.patch, embed_dim=d.v_width, depth=d.v_layers,
                               num_heads=d.v_heads, mlp_ratio=4, qkv_bias=True, drop_path_rate=rate,
                               norm_layer=partial(torch.nn.LayerNorm, eps=1e-6)).train()
    cfg = dict(json.load(open("/root/reference/src/models/uniir_blip/backbone/configs/med_config.json")))
    assert cfg["hidden_dropout_prob"] == 0.1 and cfg["attention_probs_dropout_prob"] == 0.1
    cfg.update(hidden_size=d.t_width, intermediate_size=d.t_inter, num_attention_heads=d.t_heads,
               num_hidden_layers=d.t_layers, vocab_size=d.vocab, max_position_embeddings=d.max_pos, encoder_width=d.v_width)
    p = tmp_path / "med.json"; p.write_text(json.dumps(cfg))
    rb = med.BertModel(config=med.BertConfig.from_json_file(str(p)), add_pooling_layer=True).train()
    rv.load_state_dict(orc.visual_encoder.state_dict(), strict=True)
    rb.load_state_dict(orc.text_encoder.state_dict(), strict=False)
    ids, att, img = _inputs(d, 4, 20, 1)
    # ---- reference: nn.Dropout and DropPath draw from the sequence
    seq = _SeqMasks(1000)
    monkeypatch.setattr(torch.nn.Dropout, "forward",
                        lambda self, x: x if (not self.training or self.p == 0) else x * seq.keep(x.shape, self.p) / (1 - self.p))
    monkeypatch.setattr(DropPath, "mask_fn", staticmethod(lambda shape, p: seq.keep(shape, p)))
    v_ref = rv(img)
    e_ref = rb(ids, attention_mask=att, encoder_hidden_states=v_ref,
               encoder_attention_mask=torch.ones(v_ref.shape[:-1], dtype=torch.long), return_dict=True).pooler_output
    calls_ref = seq.i
    e_ref.square().sum().backward()
    # ---- oracle: the hook draws from a fresh copy of the same sequence
    seq2 = _SeqMasks(1000)
    dpr = torch.linspace(0, rate, d.v_layers).tolist()

    def hook(tag, kind, layer, x):
        if kind.startswith("path"):
            pl = dpr[layer]
            if pl == 0.0:
                return x                                   # nn.Identity in the reference (vit.py:136): no draw
            return x / (1 - pl) * seq2.keep((x.shape[0],) + (1,) * (x.ndim - 1), pl)
        return x * seq2.keep(x.shape, 0.1) / (1 - 0.1)
    monkeypatch.setattr(bo, "DROPOUT", hook)
    e_orc = orc.encode_multimodal_input(ids, att, img)
    assert seq2.i == calls_ref == 2 * (d.v_layers - 1) + 1 + 5 * d.t_layers
    e_orc.square().sum().backward()
    assert torch.allclose(e_orc, e_ref, atol=1e-5, rtol=1e-4), (e_orc - e_ref).abs().max()
    for n, pr in rb.named_parameters():
        if n in ("embeddings.word_embeddings.weight", "encoder.layer.0.attention.self.key.weight",
                 "encoder.layer.1.crossattention.output.dense.bias", "encoder.layer.0.output.dense.weight", "pooler.dense.weight"):
            assert torch.allclose(pr.grad, orc.text_encoder.get_parameter(n).grad, atol=1e-6, rtol=1e-4), n
    for n in ("blocks.0.attn.qkv.weight", "blocks.1.mlp.fc2.bias", "cls_token"):
        assert torch.allclose(rv.get_parameter(n).grad, orc.visual_encoder.get_parameter(n).grad, atol=1e-6, rtol=1e-4), n


def test_param_count_of_blip_ff_large_online_encoder():
    m = BLIPFeatureFusionOracle("large", queue_size=64)
    n = sum(p.numel() for k, p in m.named_parameters() if k.startswith(("visual_encoder.", "text_encoder.")))
    assert abs(n - 445.87e6) < 0.6e6, n   # 445.87 M online parameters


def test_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "blip_tiny.npz"))
    m = BLIPFeatureFusionOracle("tiny-test", queue_size=64).eval()
    missing, unexpected = m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}, strict=False)
    assert not unexpected and all("_m." in k or k in ("query_queue", "cand_queue", "idx_queue", "new_ptr_queue", "temp") for k in missing)
    with torch.no_grad():
        e = m.encode_multimodal_input(torch.from_numpy(g["ids"]), torch.from_numpy(g["att"]), torch.from_numpy(g["img"]))
    assert np.allclose(e.numpy(), g["emb"], atol=1e-5)


@pytest.mark.skipif(not ref_blip_shim.available(), reason="reference tree not present (GPU box): golden vectors cover it")
def test_training_loss_matches_reference_class(tmp_path):
    """Two consecutive steps of the REAL reference BLIPFeatureFusion.compute_contrastive_loss (momentum update,
    soft labels, queue) vs the oracle restatement: loss, accuracy, gradients, EMA'd weights, queue contents."""
    import types
    from oracle.blip_oracle import BLIPDims
    bf = ref_blip_shim.load_blip_ff()
    cfg = dict(json.load(open("/root/reference/src/models/uniir_blip/backbone/configs/med_config.json")))
    cfg.update(hidden_size=64, intermediate_size=256, num_attention_heads=1, num_hidden_layers=2, vocab_size=600,
               max_position_embeddings=128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    p = tmp_path / "med.json"; p.write_text(json.dumps(cfg))
    torch.manual_seed(0)
    ref = bf.BLIPFeatureFusion(med_config=str(p), image_size=64, vit="base", embed_dim=64, queue_size=16, config=None).eval()
    d = BLIPDims(image_size=64, patch=16, v_width=768, v_layers=12, v_heads=12, t_width=64, t_layers=2, t_heads=1,
                 t_inter=256, vocab=600, max_pos=128)
    orc = BLIPFeatureFusionOracle(dims=d, embed_dim=64, queue_size=16).eval()
    missing, unexpected = orc.load_state_dict(ref.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    for step in range(2):
        ids, att, img = _inputs(d, 8, 20, 10 + step)
        mk = lambda: dict(image_batched=img, txt_mask_batched=None, image_mask_batched=None, p_did_list=[7, 8, 7, 9 + step],
                          index_mapping={"query": [[0], [2], [4], [6]], "pos_cand": [[1], [3], [5], [7]]})
        br = mk(); br["txt_batched"] = types.SimpleNamespace(input_ids=ids, attention_mask=att)
        bo = mk(); bo["txt_batched"] = {"input_ids": ids, "attention_mask": att}
        ref.zero_grad(); orc.zero_grad()
        o_r = ref.compute_contrastive_loss(br, alpha=0.3); o_r["loss"].backward()
        o_o = orc.compute_contrastive_loss(bo, alpha=0.3); o_o["loss"].backward()
        assert abs(o_r["loss"].item() - o_o["loss"].item()) < 1e-5 and abs(o_r["accuracy"].item() - o_o["accuracy"].item()) < 1e-6
        for n in ("temp", "text_encoder.pooler.dense.weight", "text_encoder.encoder.layer.0.crossattention.self.key.weight",
                  "visual_encoder.blocks.3.attn.qkv.weight", "visual_encoder.patch_embed.proj.bias"):
            a, b = ref.get_parameter(n).grad, orc.get_parameter(n).grad
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-4), n
        assert torch.equal(ref.idx_queue, orc.idx_queue) and torch.equal(ref.new_ptr_queue, orc.new_ptr_queue)
        assert torch.allclose(ref.cand_queue, orc.cand_queue, atol=1e-6)
        assert torch.allclose(ref.visual_encoder_m.blocks[0].attn.qkv.weight, orc.visual_encoder_m.blocks[0].attn.qkv.weight, atol=1e-7)


@pytest.mark.skipif(not ref_blip_shim.available(), reason="reference tree not present (GPU box): golden vectors cover it")
def test_hard_negative_variant_matches_reference_class(tmp_path):
    """blip_ff.py:126-131,160-171,208-218,233-246: negatives in place of the first queue columns, un-normalised
    negative momentum features, the coin flip between enqueueing positives and first negatives (same torch seed on
    both sides) -- three steps so that both outcomes of the flip and a non-trivial queue are covered."""
    import types
    from oracle.blip_oracle import BLIPDims
    bf = ref_blip_shim.load_blip_ff()
    cfg = dict(json.load(open("/root/reference/src/models/uniir_blip/backbone/configs/med_config.json")))
    cfg.update(hidden_size=64, intermediate_size=256, num_attention_heads=1, num_hidden_layers=2, vocab_size=600,
               max_position_embeddings=128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    p = tmp_path / "med.json"; p.write_text(json.dumps(cfg))
    torch.manual_seed(0)
    ref = bf.BLIPFeatureFusion(med_config=str(p), image_size=64, vit="base", embed_dim=64, queue_size=16, config=None).eval()
    d = BLIPDims(image_size=64, patch=16, v_width=768, v_layers=12, v_heads=12, t_width=64, t_layers=2, t_heads=1,
                 t_inter=256, vocab=600, max_pos=128)
    orc = BLIPFeatureFusionOracle(dims=d, embed_dim=64, queue_size=16).eval()
    orc.load_state_dict(ref.state_dict(), strict=False)
    flips = set()
    for step in range(3):
        ids, att, img = _inputs(d, 16, 20, 30 + step)          # 4 pairs x [query, pos, neg, neg]
        mk = lambda: dict(image_batched=img, txt_mask_batched=None, image_mask_batched=None, p_did_list=[7, 8, 7, 9 + step],
                          nc_dids_list=[[20, 21], [22, 8], [24, 25], [26, 27 + step]],
                          index_mapping={"query": [[4 * i] for i in range(4)], "pos_cand": [[4 * i + 1] for i in range(4)],
                                         "neg_cand_list": [[4 * i + 2, 4 * i + 3] for i in range(4)]})
        br = mk(); br["txt_batched"] = types.SimpleNamespace(input_ids=ids, attention_mask=att)
        bo = mk(); bo["txt_batched"] = {"input_ids": ids, "attention_mask": att}
        ref.zero_grad(); orc.zero_grad()
        torch.manual_seed(102 + step); o_r = ref.compute_contrastive_loss(br, alpha=0.3); o_r["loss"].backward()
        torch.manual_seed(102 + step); o_o = orc.compute_contrastive_loss(bo, alpha=0.3); o_o["loss"].backward()
        torch.manual_seed(102 + step); flips.add(bool(torch.rand(1) < 0.5))
        assert abs(o_r["loss"].item() - o_o["loss"].item()) < 1e-5 and abs(o_r["accuracy"].item() - o_o["accuracy"].item()) < 1e-6
        for n in ("temp", "text_encoder.pooler.dense.weight", "visual_encoder.blocks.3.attn.qkv.weight"):
            assert torch.allclose(ref.get_parameter(n).grad, orc.get_parameter(n).grad, atol=1e-6, rtol=1e-4), n
        assert torch.equal(ref.idx_queue, orc.idx_queue) and torch.equal(ref.new_ptr_queue, orc.new_ptr_queue)
        assert torch.allclose(ref.cand_queue, orc.cand_queue, atol=1e-6) and torch.allclose(ref.query_queue, orc.query_queue, atol=1e-6)
    assert flips == {True, False}


@pytest.mark.skipif(not ref_blip_shim.available(), reason="reference tree not present (GPU box): golden vectors cover it")
def test_blip_sf_restatement_matches_reference_class(tmp_path):
    """The REAL reference BLIPScoreFusion (blip_sf.py) vs the oracle restatement: state-dict keys, masked
    score-fusion embeddings (online and momentum), and two consecutive training steps (loss, accuracy, gradients
    incl. the frozen cross-attention, EMA'd weights, queues)."""
    import types
    from oracle.blip_oracle import BLIPDims
    bs = ref_blip_shim.load_blip_sf()
    cfg = dict(json.load(open("/root/reference/src/models/uniir_blip/backbone/configs/med_config.json")))
    cfg.update(hidden_size=64, intermediate_size=256, num_attention_heads=1, num_hidden_layers=2, vocab_size=600,
               max_position_embeddings=128, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    p = tmp_path / "med.json"; p.write_text(json.dumps(cfg))
    torch.manual_seed(0)
    ref = bs.BLIPScoreFusion(med_config=str(p), image_size=64, vit="base", embed_dim=48, queue_size=16, config=None).eval()
    d = BLIPDims(image_size=64, patch=16, v_width=768, v_layers=12, v_heads=12, t_width=64, t_layers=2, t_heads=1,
                 t_inter=256, vocab=600, max_pos=128)
    orc = BLIPScoreFusionOracle(dims=d, embed_dim=48, queue_size=16).eval()
    missing, unexpected = orc.load_state_dict(ref.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    assert {n for n, q in ref.named_parameters() if not q.requires_grad} == {n for n, q in orc.named_parameters() if not q.requires_grad}
    ids, att, img = _inputs(d, 8, 20, 3)
    mt = torch.tensor([1, 1, 0, 1, 1, 0, 1, 1]); mv = torch.tensor([1, 0, 1, 1, 1, 1, 0, 1])
    txt_ns = types.SimpleNamespace(input_ids=ids, attention_mask=att)
    with torch.no_grad():
        for mom in (False, True):
            e_ref = ref.encode_multimodal_input(txt_ns, img, mt, mv, use_momentum=mom)
            e_orc = orc.encode_multimodal_input(ids, att, img, mt, mv, use_momentum=mom)
            assert torch.allclose(e_orc, e_ref, atol=1e-5, rtol=1e-4), (e_orc - e_ref).abs().max()
    for step in range(2):
        ids, att, img = _inputs(d, 8, 20, 10 + step)
        mk = lambda: dict(image_batched=img, txt_mask_batched=mt, image_mask_batched=mv, p_did_list=[7, 8, 7, 9 + step],
                          index_mapping={"query": [[0], [2], [4], [6]], "pos_cand": [[1], [3], [5], [7]]})
        br = mk(); br["txt_batched"] = types.SimpleNamespace(input_ids=ids, attention_mask=att)
        bo = mk(); bo["txt_batched"] = {"input_ids": ids, "attention_mask": att}
        ref.zero_grad(); orc.zero_grad()
        o_r = ref.compute_contrastive_loss(br, alpha=0.3); o_r["loss"].backward()
        o_o = orc.compute_contrastive_loss(bo, alpha=0.3); o_o["loss"].backward()
        assert abs(o_r["loss"].item() - o_o["loss"].item()) < 1e-5 and abs(o_r["accuracy"].item() - o_o["accuracy"].item()) < 1e-6
        for n in ("temp", "text_proj.weight", "vision_proj.bias", "text_encoder.encoder.layer.1.attention.self.key.weight",
                  "visual_encoder.blocks.3.attn.qkv.weight", "visual_encoder.patch_embed.proj.bias"):
            a, b = ref.get_parameter(n).grad, orc.get_parameter(n).grad
            assert torch.allclose(a, b, atol=1e-6, rtol=1e-4), n
        assert ref.get_parameter("text_encoder.encoder.layer.0.crossattention.self.key.weight").grad is None
        assert orc.get_parameter("text_encoder.encoder.layer.0.crossattention.self.key.weight").grad is None
        assert torch.equal(ref.idx_queue, orc.idx_queue) and torch.equal(ref.new_ptr_queue, orc.new_ptr_queue)
        assert torch.allclose(ref.cand_queue, orc.cand_queue, atol=1e-6)
        assert torch.allclose(ref.text_proj_m.weight, orc.text_proj_m.weight, atol=1e-7)


def test_blip_sf_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "blip_sf_tiny.npz"))
    m = BLIPScoreFusionOracle("tiny-test", embed_dim=64, queue_size=64).eval()
    missing, unexpected = m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}, strict=False)
    assert not unexpected and all("_m." in k or k in ("query_queue", "cand_queue", "idx_queue", "new_ptr_queue", "temp") for k in missing)
    with torch.no_grad():
        e = m.encode_multimodal_input(torch.from_numpy(g["ids"]), torch.from_numpy(g["att"]), torch.from_numpy(g["img"]),
                                      torch.from_numpy(g["mt"]), torch.from_numpy(g["mv"]))
    assert np.allclose(e.numpy(), g["emb"], atol=1e-5)
    assert np.abs(g["emb"][2] - g["emb"][4]).max() > 1e-3     # image-only rows differ; masks were applied, not ignored


@pytest.mark.parametrize("mode", ["ff", "sf"])
def test_product_parameter_layout_equals_the_reference_state_dict(mode):
    """uniir_b200's flat parameter specs (names, shapes; online and momentum sets) against the oracle classes, whose
    state dicts load into / from the real reference classes with strict key equality (tests above)."""
    from uniir_b200.blip_ff import BLIP_CONFIGS as PCFG, blip_param_specs
    d = PCFG["tiny-test"]
    orc = (BLIPFeatureFusionOracle("tiny-test", embed_dim=48, queue_size=16) if mode == "ff"
           else BLIPScoreFusionOracle("tiny-test", embed_dim=48, queue_size=16))
    want = {k: tuple(v.shape) for k, v in orc.named_parameters()}
    got = {}
    for sfx in ("", "_m"):
        for name, shape, decay, _ in blip_param_specs(d, sfx, mode, 48):
            assert name not in got
            got[name] = tuple(shape)
            if mode == "sf":       # frozen cross-attention -> the segment without weight decay / gradients
                assert decay == (".crossattention." not in name), name
            else:
                assert decay, name
    assert got == want, (sorted(set(got) ^ set(want))[:10])
    # the momentum set is an exact prefix of the online set in each flat segment (EMA = one kernel per segment)
    on, mo = blip_param_specs(d, "", mode, 48), blip_param_specs(d, "_m", mode, 48)
    for seg in (True, False):
        a = [(n, s) for n, s, dc, _ in on if dc == seg]
        b = [(n, s) for n, s, dc, _ in mo if dc == seg]
        assert [s for _, s in a[:len(b)]] == [s for _, s in b]
        assert len(a) - len(b) == (1 if seg else 0)      # temp, last in the decayed segment
