"""Pin the BLIP_FF oracle restatement: (1) against the reference's own vendored modules imported through
shims (only where /root/reference exists), (2) against the committed golden vectors."""
import json
import os

import numpy as np
import pytest
import torch

import ref_blip_shim
from conftest import GOLDEN
from oracle.blip_oracle import BLIP_CONFIGS, BLIPFeatureFusionOracle


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


def test_param_count_of_blip_ff_large_online_encoder():
    n = sum(p.numel() for p in BLIPFeatureFusionOracle("large").parameters())
    assert abs(n - 445.87e6) < 0.6e6, n   # 445.87 M online parameters (+temp)


def test_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "blip_tiny.npz"))
    m = BLIPFeatureFusionOracle("tiny-test").eval()
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")}, strict=True)
    with torch.no_grad():
        e = m.encode_multimodal_input(torch.from_numpy(g["ids"]), torch.from_numpy(g["att"]), torch.from_numpy(g["img"]))
    assert np.allclose(e.numpy(), g["emb"], atol=1e-5)
