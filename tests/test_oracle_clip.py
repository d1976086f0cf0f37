"""Pin the CLIP oracle: (1) against an independent implementation (transformers.CLIPModel with
mapped weights -- the reference's CLIP arithmetic lives in the un-vendored openai/CLIP package,
SURVEY.md 8c), (2) against the committed golden vectors, (3) loss semantics of clip_sf.py:68-147."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle.clip_oracle import (CLIP_CONFIGS, CLIPOracle, CLIPScoreFusionOracle,
                                contrastive_loss_from_embeddings)


def _hf_from_oracle(orc: CLIPOracle):
    tr = pytest.importorskip("transformers")
    d = orc.dims
    cfg = tr.CLIPConfig(
        text_config=dict(vocab_size=d.vocab, hidden_size=d.t_width, intermediate_size=4 * d.t_width,
                         num_hidden_layers=d.t_layers, num_attention_heads=d.t_heads,
                         max_position_embeddings=d.context, hidden_act="quick_gelu", layer_norm_eps=1e-5,
                         eos_token_id=d.vocab - 1, bos_token_id=d.vocab - 2, pad_token_id=0, projection_dim=d.embed_dim),
        vision_config=dict(hidden_size=d.v_width, intermediate_size=4 * d.v_width, num_hidden_layers=d.v_layers,
                           num_attention_heads=d.v_heads, image_size=d.image_size, patch_size=d.patch,
                           hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=d.embed_dim),
        projection_dim=d.embed_dim)
    hf = tr.CLIPModel(cfg).eval()
    sd = orc.state_dict()
    new = {}

    def blocks(src, dst, n, w):
        for i in range(n):
            s, t = f"{src}.resblocks.{i}", f"{dst}.encoder.layers.{i}"
            qw, kw, vw = sd[f"{s}.attn.in_proj_weight"].split(w, 0)
            qb, kb, vb = sd[f"{s}.attn.in_proj_bias"].split(w, 0)
            for nm, W, B in (("q", qw, qb), ("k", kw, kb), ("v", vw, vb)):
                new[f"{t}.self_attn.{nm}_proj.weight"], new[f"{t}.self_attn.{nm}_proj.bias"] = W, B
            new[f"{t}.self_attn.out_proj.weight"] = sd[f"{s}.attn.out_proj.weight"]
            new[f"{t}.self_attn.out_proj.bias"] = sd[f"{s}.attn.out_proj.bias"]
            for a, b in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2")):
                new[f"{t}.{b}.weight"], new[f"{t}.{b}.bias"] = sd[f"{s}.{a}.weight"], sd[f"{s}.{a}.bias"]
            for a, b in (("c_fc", "fc1"), ("c_proj", "fc2")):
                new[f"{t}.mlp.{b}.weight"], new[f"{t}.mlp.{b}.bias"] = sd[f"{s}.mlp.{a}.weight"], sd[f"{s}.mlp.{a}.bias"]

    blocks("transformer", "text_model", d.t_layers, d.t_width)
    blocks("visual.transformer", "vision_model", d.v_layers, d.v_width)
    new["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    new["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    new["text_model.final_layer_norm.weight"], new["text_model.final_layer_norm.bias"] = sd["ln_final.weight"], sd["ln_final.bias"]
    new["text_projection.weight"] = sd["text_projection"].t()
    new["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    new["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    new["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    new["vision_model.pre_layrnorm.weight"], new["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"]
    new["vision_model.post_layernorm.weight"], new["vision_model.post_layernorm.bias"] = sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]
    new["visual_projection.weight"] = sd["visual.proj"].t()
    new["logit_scale"] = sd["logit_scale"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return hf


def _tokens(n, d, seed):
    from uniir_b200.preprocess import synthetic_clip_tokens
    return synthetic_clip_tokens(n, torch.Generator().manual_seed(seed), d.context, d.vocab)


def test_oracle_matches_hf_clip_both_towers():
    torch.manual_seed(0)
    orc = CLIPOracle("tiny-test").eval()
    hf = _hf_from_oracle(orc)
    d = orc.dims
    ids = _tokens(5, d, 1)
    img = torch.randn(5, 3, d.image_size, d.image_size)
    with torch.no_grad():
        t_o, v_o = orc.encode_text(ids), orc.encode_image(img)
        mask = (ids != 0).long()  # HF pools at eos_token_id; causal mask only (pad keys are after EOT)
        t_h = hf.get_text_features(input_ids=ids.long(), attention_mask=torch.ones_like(mask))
        v_h = hf.get_image_features(pixel_values=img)
        t_h = getattr(t_h, "pooler_output", t_h)
        v_h = getattr(v_h, "pooler_output", v_h)
    assert torch.allclose(t_o, t_h, atol=2e-5, rtol=1e-4), (t_o - t_h).abs().max()
    assert torch.allclose(v_o, v_h, atol=2e-5, rtol=1e-4), (v_o - v_h).abs().max()


def test_param_count_matches_published_clip_sizes():
    assert sum(p.numel() for p in CLIPOracle("ViT-L/14").parameters()) == 427_616_513
    assert sum(p.numel() for p in CLIPOracle("ViT-B/32").parameters()) == 151_277_313


def _golden_model():
    g = np.load(os.path.join(GOLDEN, "clip_tiny.npz"))
    m = CLIPScoreFusionOracle("tiny-test")
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}, strict=True)
    bs = g["ids"].shape[0] // 2
    batch = dict(txt_batched=torch.from_numpy(g["ids"]), image_batched=torch.from_numpy(g["img"]),
                 txt_mask_batched=torch.from_numpy(g["mt"]), image_mask_batched=torch.from_numpy(g["mv"]),
                 index_mapping={"query": [[2 * i] for i in range(bs)], "pos_cand": [[2 * i + 1] for i in range(bs)]})
    return g, m, batch


def test_oracle_reproduces_golden_vectors():
    g, m, batch = _golden_model()
    out = m(batch)
    out["loss"].backward()
    with torch.no_grad():
        emb = m.encode_multimodal_input(batch["txt_batched"], batch["image_batched"], batch["txt_mask_batched"],
                                        batch["image_mask_batched"])
    assert np.allclose(emb.numpy(), g["emb"], atol=1e-5)
    assert abs(out["loss"].item() - float(g["loss"])) < 1e-5 and out["accuracy"].item() == float(g["acc"])
    for k in g.files:
        if k.startswith("g:"):
            assert np.allclose(m.get_parameter(k[2:]).grad.numpy(), g[k], atol=1e-5, rtol=1e-4), k


def test_loss_semantics_clip_sf():
    """clip_sf.py:88-147: one-directional CE over q @ all_p.T * exp(logit_scale), accuracy by argmax."""
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(10, 16, generator=g)
    im = {"query": [[0], [2], [4], [6], [8]], "pos_cand": [[1], [3], [5], [7], [9]]}
    out = contrastive_loss_from_embeddings(emb, im, torch.tensor(50.0))
    q, p = F.normalize(emb[0::2], dim=-1), F.normalize(emb[1::2], dim=-1)
    s = q @ p.t() * 50.0
    ref = -(s.log_softmax(1).diag()).mean()
    assert torch.allclose(out["loss"], ref, atol=1e-6)
    assert out["accuracy"].item() == (s.argmax(1) == torch.arange(5)).float().mean().item()
    # gathered variant: targets shift by rank*bs (clip_sf.py:135-136)
    other = F.normalize(torch.randn(5, 16, generator=g), dim=-1)
    out2 = contrastive_loss_from_embeddings(emb, im, torch.tensor(50.0), all_p_fn=lambda x: torch.cat([other, x]), rank=1)
    s2 = q @ torch.cat([other, p]).t() * 50.0
    assert torch.allclose(out2["loss"], F.cross_entropy(s2, 5 + torch.arange(5)), atol=1e-6)


def test_hard_negative_branch_runs_and_matches_manual():
    g = torch.Generator().manual_seed(1)
    bs, neg = 3, 2
    emb = torch.randn(bs * (2 + neg), 8, generator=g)
    im = {"query": [[0], [4], [8]], "pos_cand": [[1], [5], [9]], "neg_cand_list": [[2, 3], [6, 7], [10, 11]]}
    out = contrastive_loss_from_embeddings(emb, im, torch.tensor(10.0), in_batch_neg_num=1)
    assert out["loss"].ndim == 0 and 0 <= out["accuracy"].item() <= 1


def test_encode_mbeir_batch_contract():
    m = CLIPScoreFusionOracle("tiny-test")
    d = m.clip_model.dims
    b = dict(txt_batched=_tokens(3, d, 2), image_batched=torch.randn(3, 3, d.image_size, d.image_size),
             txt_mask_batched=torch.ones(3, dtype=torch.long), image_mask_batched=torch.ones(3, dtype=torch.long),
             did_list=[10000001, 10000002, 10000003])
    emb, ids = m(b, encode_mbeir_batch=True)
    assert emb.shape == (3, d.embed_dim) and ids == b["did_list"]
    b2 = dict(b); b2.pop("did_list")
    with pytest.raises(AssertionError):
        m(b2, encode_mbeir_batch=True)
    b3 = dict(b); b3["did_list"] = ["1:1", "1:2", "1:3"]
    with pytest.raises(AssertionError):
        m(b3, encode_mbeir_batch=True)
