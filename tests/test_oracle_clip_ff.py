"""Pin the CLIP_FF oracle (oracle/clip_ff_oracle.py; SURVEY 8f-3, no product path yet): the T5 fusion stack against
HuggingFace's own `T5Stack` -- the class the reference instantiates (clip_ff.py:14,82-98) -- with copied weights,
the relative-position bucketing against its static method, and the model-level contract of clip_ff.py."""
import pytest
import torch

from oracle.clip_ff_oracle import CLIPFeatureFusionOracle, T5StackOracle, t5_relative_position_bucket


def _hf_stack(d_model, heads):
    from transformers.models.t5 import T5Config
    from transformers.models.t5.modeling_t5 import T5Stack
    c = T5Config()                      # exactly the reference's construction (clip_ff.py:82-98)
    c.num_layers = 2; c.num_decoder_layers = 2; c.num_heads = heads; c.d_model = d_model; c.d_kv = 64
    return T5Stack(c).eval(), c


@pytest.mark.parametrize("d_model,heads,S", [(768, 12, 334), (512, 12, 127), (128, 2, 94)])
def test_t5_stack_restatement_matches_huggingface(d_model, heads, S):
    hf, cfg = _hf_stack(d_model, heads)
    assert (cfg.d_ff, cfg.relative_attention_num_buckets, cfg.relative_attention_max_distance, cfg.feed_forward_proj,
            cfg.layer_norm_epsilon) == (2048, 32, 128, "relu", 1e-6)
    torch.manual_seed(0)
    with torch.no_grad():
        for n, p in hf.named_parameters():
            p.normal_(0.0, 0.05) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    orc = T5StackOracle(d_model=d_model, heads=heads).eval()
    sd = {k: v for k, v in hf.state_dict().items() if not k.startswith("embed_tokens")}   # unused with inputs_embeds
    missing, unexpected = orc.load_state_dict(sd, strict=True)
    x = torch.randn(3, S, d_model) * 0.7
    with torch.no_grad():
        ref = hf(inputs_embeds=x, attention_mask=None, use_cache=False, return_dict=True).last_hidden_state
        out = orc(x)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (out - ref).abs().max()
    # gradients too (the training path of the next round)
    xr = x.clone().requires_grad_(True); xo = x.clone().requires_grad_(True)
    hf(inputs_embeds=xr, attention_mask=None, use_cache=False, return_dict=True).last_hidden_state.square().mean().backward()
    orc(xo).square().mean().backward()
    assert torch.allclose(xo.grad, xr.grad, atol=1e-6, rtol=1e-3)
    g_ref = dict(hf.named_parameters())["block.0.layer.0.SelfAttention.relative_attention_bias.weight"].grad
    g_out = orc.block[0].layer[0].SelfAttention.relative_attention_bias.weight.grad
    assert torch.allclose(g_out, g_ref, atol=1e-6, rtol=1e-3)


def test_relative_position_bucket_matches_huggingface():
    from transformers.models.t5.modeling_t5 import T5Attention
    pos = torch.arange(400)
    rel = pos[None, :] - pos[:, None]
    ref = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(t5_relative_position_bucket(rel), ref)


def test_clip_ff_model_contract():
    torch.manual_seed(1)
    m = CLIPFeatureFusionOracle("tiny-test").eval()
    d = m.clip_model.dims
    keys = set(m.state_dict())
    assert "clip_model.text_projection" not in keys and "clip_model.visual.proj" in keys                # clip_ff.py:110
    assert "t5_layers.block.0.layer.0.SelfAttention.relative_attention_bias.weight" in keys
    assert "t5_layers.block.1.layer.0.SelfAttention.relative_attention_bias.weight" not in keys
    B = 4
    g = torch.Generator().manual_seed(2)
    ids = torch.randint(1, d.vocab - 2, (B, d.context), generator=g); ids[:, 0] = d.vocab - 2; ids[:, 9] = d.vocab - 1
    img = torch.randn(B, 3, d.image_size, d.image_size, generator=g)
    one = torch.ones(B, dtype=torch.long)
    with torch.no_grad():
        e1 = m.encode_multimodal_input(ids, img, one, one)
        e0 = m.encode_multimodal_input(ids, img, 0 * one, 0 * one)
        t, v = m.encode_text(ids), m.encode_image(img)
    assert e1.shape == (B, d.embed_dim) and torch.equal(e1, e0)                  # masks are ignored (clip_ff.py:173-176)
    assert t.shape == (B, d.context, d.t_width) and v.shape == (B, d.v_tokens, d.embed_dim)
    b = dict(txt_batched=ids, image_batched=img, txt_mask_batched=one, image_mask_batched=one,
             index_mapping={"query": [[0], [2]], "pos_cand": [[1], [3]]})
    out = m(b)
    assert out["loss"].ndim == 0 and out["accuracy"].ndim == 0
    out["loss"].backward()
    assert m.t5_layers.block[1].layer[1].DenseReluDense.wi.weight.grad.abs().sum() > 0
    b2 = dict(b); b2.pop("index_mapping"); b2["did_list"] = [5, 6, 7, 8]
    emb, idl = m(b2, encode_mbeir_batch=True)
    assert emb.shape == (B, d.embed_dim) and idl == [5, 6, 7, 8]
    with pytest.raises(AssertionError):
        m({**b2, "did_list": ["1:2"] * B}, encode_mbeir_batch=True)


@pytest.mark.parametrize("name", ["tiny-test", "ViT-B/32"])
def test_product_parameter_layout_equals_the_oracle_state_dict(name):
    """uniir_b200.clip_ff's flat parameter specs (names, shapes) against the oracle's state dict (reference layout:
    clip_model.* without text_projection, t5_layers.* with transformers' T5Stack names), and the product's host-side
    bucketing table against transformers' own function."""
    from transformers.models.t5.modeling_t5 import T5Attention
    from uniir_b200.clip_ff import clip_ff_param_specs, t5_heads, t5_relative_position_bucket as product_bucket
    from uniir_b200.params import CLIP_CONFIGS
    orc = CLIPFeatureFusionOracle(name)
    want = {k: tuple(v.shape) for k, v in orc.state_dict().items()}
    got = {n: tuple(s) for n, s, _, _ in clip_ff_param_specs(CLIP_CONFIGS[name], t5_heads(name))}
    assert got == want, sorted(set(got) ^ set(want))[:10]
    specs = [n for n, *_ in clip_ff_param_specs(CLIP_CONFIGS[name], t5_heads(name))]
    for i in range(2):      # q, k, v adjacent: the fused projection is a view of the flat buffer
        j = specs.index(f"t5_layers.block.{i}.layer.0.SelfAttention.q.weight")
        assert specs[j + 1].endswith("SelfAttention.k.weight") and specs[j + 2].endswith("SelfAttention.v.weight")
    pos = torch.arange(334)
    ref = T5Attention._relative_position_bucket(pos[None, :] - pos[:, None], bidirectional=True, num_buckets=32, max_distance=128)
    assert torch.equal(product_bucket(334), ref)


def test_t5_dropout_sites_match_huggingface_in_train_mode(monkeypatch):
    """Train mode: transformers' T5Stack (dropout_rate 0.1, the default the reference keeps) and the oracle get the same
    keep decisions in call order -- torch.nn.functional.dropout is what both nn.Dropout and T5Attention call --; outputs
    and gradients agree only if every dropout of transformers' stack has its hook at the same place in the oracle."""
    import oracle.clip_ff_oracle as co
    d_model, heads, S = 128, 2, 41
    hf, cfg = _hf_stack(d_model, heads)
    assert cfg.dropout_rate == 0.1
    hf.train()
    torch.manual_seed(0)
    with torch.no_grad():
        for n, p in hf.named_parameters():
            p.normal_(0.0, 0.05) if "layer_norm" not in n else p.normal_(1.0, 0.1)
    orc = T5StackOracle(d_model=d_model, heads=heads).eval()
    orc.load_state_dict({k: v for k, v in hf.state_dict().items() if not k.startswith("embed_tokens")}, strict=True)
    x = torch.randn(3, S, d_model) * 0.7

    class Seq:
        def __init__(self):
            self.i, self.log = 0, []

        def keep(self, shape):
            g = torch.Generator().manual_seed(500 + self.i)
            self.i += 1
            self.log.append(tuple(shape))
            return (torch.rand(shape, generator=g) >= 0.1).float()
    a, b = Seq(), Seq()
    monkeypatch.setattr(torch.nn.functional, "dropout",
                        lambda inp, p=0.5, training=True, inplace=False: inp if (not training or p == 0) else inp * a.keep(inp.shape) / (1 - p))
    xr = x.clone().requires_grad_(True)
    ref = hf(inputs_embeds=xr, attention_mask=None, use_cache=False, return_dict=True).last_hidden_state
    ref.square().mean().backward()
    monkeypatch.undo()
    monkeypatch.setattr(co, "DROPOUT", lambda kind, layer, t: t * b.keep(t.shape) / (1 - 0.1))
    xo = x.clone().requires_grad_(True)
    out = orc(xo)
    out.square().mean().backward()
    assert a.log == b.log and a.i == 2 + 4 * 2                     # input, 2 x (probs, attn out, ff inner, ff out), final
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4), (out - ref).abs().max()
    assert torch.allclose(xo.grad, xr.grad, atol=1e-6, rtol=1e-3)
    for n in ("block.0.layer.0.SelfAttention.relative_attention_bias.weight", "block.1.layer.1.DenseReluDense.wi.weight",
              "block.0.layer.0.SelfAttention.o.weight", "final_layer_norm.weight"):
        assert torch.allclose(orc.get_parameter(n).grad, hf.get_parameter(n).grad, atol=1e-6, rtol=1e-3), n
