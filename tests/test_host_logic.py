"""Host-side logic that needs no GPU: parameter layout / decay groups / state-dict compatibility,
the contiguous row partition, LR schedule, tokeniser/preprocess contracts."""
import math

import pytest
import torch

from oracle.clip_oracle import CLIPScoreFusionOracle
from uniir_b200.dist_utils import ContiguousDistributedSampler, shard_range
from uniir_b200.params import CLIP_CONFIGS, FlatParamModule, clip_param_specs


def test_param_specs_match_upstream_layout_and_count():
    specs = clip_param_specs(CLIP_CONFIGS["ViT-L/14"])
    assert sum(math.prod(s[1]) if s[1] else 1 for s in specs) == 427_616_513
    names = {s[0] for s in specs}
    for k in ("visual.conv1.weight", "visual.proj", "text_projection", "logit_scale", "token_embedding.weight",
              "transformer.resblocks.11.attn.in_proj_weight", "visual.transformer.resblocks.23.mlp.c_proj.bias"):
        assert k in names
    shapes = {s[0]: s[1] for s in specs}
    assert shapes["visual.conv1.weight"] == (1024, 3, 14, 14) and shapes["visual.positional_embedding"] == (257, 1024)
    assert shapes["visual.proj"] == (1024, 768) and shapes["text_projection"] == (768, 768)


def test_decay_groups_follow_reference_rule():
    """clip_scorefusion/train.py:52-61: no decay for ndim<2, bn, ln, bias, logit_scale."""
    for name, shape, decay, _ in clip_param_specs(CLIP_CONFIGS["ViT-B/32"]):
        expect = not (len(shape) < 2 or "bn" in name or "ln" in name or "bias" in name or "logit_scale" in name)
        assert decay == expect, name


def test_flat_module_state_dict_roundtrip_with_oracle_keys():
    orc = CLIPScoreFusionOracle("tiny-test")
    flat = FlatParamModule(clip_param_specs(CLIP_CONFIGS["tiny-test"]), torch.device("cpu"))
    wrap = torch.nn.Module(); wrap.clip_model = flat
    wrap.load_state_dict(orc.state_dict(), strict=True)
    for k, v in orc.state_dict().items():
        assert torch.equal(wrap.state_dict()[k], v), k
    # parameters are views into the flat buffers (one optimizer kernel / one all-reduce per group)
    p = flat.get_parameter("visual.proj")
    assert p.data_ptr() == flat.w("visual.proj").data_ptr()
    flat.flat[True].zero_()
    assert p.abs().sum() == 0


@pytest.mark.parametrize("n,w", [(10, 3), (5_600_000, 8), (7, 8), (16, 4), (0, 2)])
def test_shard_range_is_the_reference_contiguous_partition(n, w):
    per = math.ceil(n / w)
    rows = []
    for r in range(w):
        lo, hi = shard_range(n, w, r)
        assert lo == min(n, r * per) and hi == min(n, (r + 1) * per)
        rows += list(range(lo, hi)) if n < 1000 else []
        s = ContiguousDistributedSampler(range(n), num_replicas=w, rank=r)
        assert len(s) == per
        if n < 1000:
            assert list(s) == list(range(lo, hi))
    if n < 1000:
        assert rows == list(range(n))


def test_cosine_lr_matches_torch():
    from uniir_b200.engine import CosineLR

    class Opt:
        param_groups = [{"lr": 1e-3}]
    o = Opt(); sch = CosineLR(o, T_max=10)
    p = torch.nn.Parameter(torch.zeros(1)); to = torch.optim.SGD([p], lr=1e-3)
    ts = torch.optim.lr_scheduler.CosineAnnealingLR(to, T_max=10)
    for _ in range(10):
        sch.step(); to.step(); ts.step()
        assert abs(o.param_groups[0]["lr"] - to.param_groups[0]["lr"]) < 1e-12


def test_synthetic_tokens_have_clip_tokenize_shape():
    from uniir_b200.preprocess import synthetic_clip_tokens
    ids = synthetic_clip_tokens(20, torch.Generator().manual_seed(0))
    assert ids.shape == (20, 77) and ids.dtype == torch.int32
    assert (ids[:, 0] == 49406).all() and (ids.max(1).values == 49407).all()
    eot = ids.argmax(1)
    for i in range(20):
        assert (ids[i, eot[i] + 1:] == 0).all()


def test_tokenizer_requires_bpe_file(monkeypatch):
    from uniir_b200.preprocess import clip_tokenizer
    monkeypatch.delenv("UNIIR_CLIP_BPE", raising=False)
    with pytest.raises(FileNotFoundError):
        clip_tokenizer()(["a photo"])


def test_image_preprocess_contract():
    from PIL import Image
    from uniir_b200.preprocess import clip_image_preprocess
    t = clip_image_preprocess(224)(Image.new("RGB", (320, 240), (0, 0, 0)))
    assert t.shape == (3, 224, 224) and t.dtype == torch.float32
    assert torch.allclose(t[:, 0, 0], torch.tensor([-0.48145466 / 0.26862954, -0.4578275 / 0.26130258, -0.40821073 / 0.27577711]), atol=1e-5)
