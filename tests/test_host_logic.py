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


def test_flat_adamw_state_dict_is_torch_adamw_layout_of_the_reference_groups():
    """Reference checkpoints carry `optimizer.state_dict()` of torch.optim.AdamW over two groups built with the
    conditions of clip_scorefusion/train.py:195-199 (train.py:52-61, 64-79).  FlatAdamW must write that layout (a torch
    optimizer built the reference's way loads it and finds every tensor in the right slot) and read it back."""
    import torch
    from oracle.clip_oracle import CLIPScoreFusionOracle
    from uniir_b200.engine import FlatAdamW
    from uniir_b200.params import CLIP_CONFIGS, FlatParamModule, clip_param_specs
    torch.manual_seed(0)
    P = FlatParamModule(clip_param_specs(CLIP_CONFIGS["tiny-test"]), "cpu")
    opt = FlatAdamW(P, lr=3e-5)
    for k in opt.m:
        opt.m[k].normal_(); opt.v[k].uniform_()
    opt.step_count = 3
    sd = opt.state_dict()
    # the reference's grouping, verbatim conditions, on a model with the reference's parameter names / order
    ref_model = CLIPScoreFusionOracle("tiny-test")
    exclude = lambda n, p: p.ndim < 2 or any(sub in n for sub in ["bn", "ln", "bias", "logit_scale"])
    named = list(ref_model.named_parameters())
    gain_or_bias = [p for n, p in named if exclude(n, p) and p.requires_grad]
    rest = [p for n, p in named if not exclude(n, p) and p.requires_grad]
    topt = torch.optim.AdamW([{"params": gain_or_bias, "weight_decay": 0.0}, {"params": rest, "weight_decay": 0.2}],
                             lr=1e-5, betas=(0.9, 0.98), eps=1.0e-6)
    topt.load_state_dict(sd)                                   # shapes / group sizes must fit
    assert topt.param_groups[0]["weight_decay"] == 0.0 and topt.param_groups[1]["weight_decay"] == 0.2
    assert topt.param_groups[0]["lr"] == 3e-5
    names0 = [n[len("clip_model."):] for n, p in named if exclude(n, p)]
    names1 = [n[len("clip_model."):] for n, p in named if not exclude(n, p)]
    for p, n in list(zip(gain_or_bias, names0)) + list(zip(rest, names1)):
        st = topt.state[p]
        assert float(st["step"]) == 3.0
        assert torch.equal(st["exp_avg"], P._view(opt.m, n)) and torch.equal(st["exp_avg_sq"], P._view(opt.v, n)), n
    # and back: a state dict written by torch loads into a fresh FlatAdamW
    opt2 = FlatAdamW(FlatParamModule(clip_param_specs(CLIP_CONFIGS["tiny-test"]), "cpu"), lr=1e-5)
    opt2.load_state_dict(topt.state_dict())
    assert opt2.step_count == 3 and opt2.param_groups[0]["lr"] == 3e-5
    for n, *_ in P._specs:      # (the alignment padding between parameters carries no state)
        assert torch.equal(opt2.P._view(opt2.m, n), P._view(opt.m, n)) and torch.equal(opt2.P._view(opt2.v, n), P._view(opt.v, n)), n
    # no step taken yet: empty state, like torch
    assert FlatAdamW(P, lr=1e-5).state_dict()["state"] == {}


def test_bpe_pretokenizer_unicode_classes():
    """clip.tokenize splits on \\p{L}+ | \\p{N} | [^\\s\\p{L}\\p{N}]+ : non-ASCII letters stay in one word, every
    digit is its own token (the ASCII-only classes split accented / CJK text differently)."""
    from uniir_b200.preprocess import bpe_pretokenizer
    pat, fallback = bpe_pretokenizer(), bpe_pretokenizer(force_stdlib=True)
    cases = {"café über straße": ["café", "über", "straße"], "東京タワー 2023年": ["東京タワー", "2", "0", "2", "3", "年"],
             "it's a dog's life!!": ["it", "'s", "a", "dog", "'s", "life", "!!"], "x_y 42": ["x", "_", "y", "4", "2"],
             "<|startoftext|>abc<|endoftext|>": ["<|startoftext|>", "abc", "<|endoftext|>"]}
    for text, want in cases.items():
        assert pat.findall(text) == want, text
        assert fallback.findall(text) == want, text


def test_clip_image_preprocess_matches_torchvision_arithmetic():
    tv = pytest.importorskip("torchvision.transforms")
    import numpy as np
    import torch
    from PIL import Image
    from uniir_b200.preprocess import CLIP_MEAN, CLIP_STD, clip_image_preprocess
    ref = tv.Compose([tv.Resize(224, interpolation=tv.InterpolationMode.BICUBIC), tv.CenterCrop(224),
                      lambda im: im.convert("RGB"), tv.ToTensor(), tv.Normalize(CLIP_MEAN, CLIP_STD)])
    mine = clip_image_preprocess(224)
    rng = np.random.default_rng(0)
    for w, h in ((640, 427), (333, 500), (224, 224), (225, 300), (301, 226), (1000, 227)):
        img = Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8))
        a, b = mine(img), ref(img)
        assert a.shape == b.shape == (3, 224, 224)
        assert torch.allclose(a, b, atol=1e-6), (w, h, (a - b).abs().max())


def _blip_test_image(seed):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from make_golden_blip_transform import test_image
    return test_image(seed)


def test_blip_image_transforms_match_golden_vectors_of_the_reference():
    """blip_image_preprocess (train: RandomResizedCrop + flip + RandomAugment(2, 5); eval: square bicubic resize)
    against tensors the reference's own get_blip_transform produced for the same seeds
    (scripts/make_golden_blip_transform.py).  Same PIL / cv2 / torch builds in this image: bit-exact."""
    import os
    import warnings
    import numpy as np
    from conftest import GOLDEN
    from uniir_b200.preprocess import blip_image_preprocess
    pytest.importorskip("cv2")
    g = np.load(os.path.join(GOLDEN, "blip_transform.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for seed in range(12):
            img = _blip_test_image(seed)
            for train in (True, False):
                torch.manual_seed(seed); np.random.seed(seed)
                out = blip_image_preprocess(32, is_train=train, min_scale=0.5)(img)
                assert out.shape == (3, 32, 32) and out.dtype == torch.float32
                assert np.array_equal(out.numpy(), g[f"{'train' if train else 'eval'}_{seed}"]), (seed, train)
    a = {k for k in g.files if k.startswith("train")}
    assert any(not np.array_equal(g[k], g[k.replace("train", "eval")]) for k in a)


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/src/models/uniir_blip/backbone/transform"),
                    reason="reference tree not present: the golden vectors cover it")
def test_blip_train_transform_and_every_augment_op_match_the_reference_code():
    """The reference's transform run here: 150 seeds of the whole train transform, and RandomAugment at other
    magnitudes (levels 2 and 9 exercise Brightness / Sharpness away from their identity at level 5)."""
    import os
    import sys
    import warnings
    import numpy as np
    pytest.importorskip("cv2")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from make_golden_blip_transform import load_reference_transform
    from uniir_b200.preprocess import BLIP_AUG_OPS, blip_image_preprocess, blip_random_augment
    bt, ra = load_reference_transform()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for seed in range(150):
            img = _blip_test_image(seed)
            torch.manual_seed(seed); np.random.seed(seed)
            ref = bt.get_blip_transform(64, 0.5, True)(img)
            torch.manual_seed(seed); np.random.seed(seed)
            assert torch.equal(blip_image_preprocess(64, is_train=True, min_scale=0.5)(img), ref), seed
        seen = set()
        for level in (2, 9):
            aug = ra.RandomAugment(2, level, isPIL=True, augs=list(BLIP_AUG_OPS))
            for seed in range(120):
                img = _blip_test_image(seed)
                np.random.seed(seed)
                seen.update(np.random.choice(list(BLIP_AUG_OPS), 2))
                np.random.seed(seed)
                ref = aug(img)
                np.random.seed(seed)
                assert np.array_equal(blip_random_augment(img, 2, level), ref), (level, seed)
        assert seen == set(BLIP_AUG_OPS)


def test_blip_tokenizer_from_a_local_vocabulary(tmp_path, monkeypatch):
    """blip.py:221-226 / blip_ff.py:70-80 with the vocabulary read from UNIIR_BERT_VOCAB (no hub download offline):
    [DEC] and [ENC] are appended after the vocabulary (30522 -> 30524 for bert-base-uncased, the model's embedding size),
    uncased WordPiece, [CLS] .. [SEP], max_length padding with a prefix attention mask."""
    from uniir_b200.preprocess import blip_tokenizer
    monkeypatch.delenv("UNIIR_BERT_VOCAB", raising=False)
    with pytest.raises(FileNotFoundError):
        blip_tokenizer(16)
    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "a", "photo", "of", "dog", "##s", "red", "car", ",", "."]
    (tmp_path / "vocab.txt").write_text("\n".join(words) + "\n")
    monkeypatch.setenv("UNIIR_BERT_VOCAB", str(tmp_path))                  # a directory holding vocab.txt
    wrap, tok = blip_tokenizer(12, model_vocab=len(words) + 2)
    assert len(tok) == len(words) + 2 and tok.bos_token == "[DEC]" and tok.enc_token_id == len(words) + 1
    out = wrap(["A photo of DOGS, a red car.", "zebra"])
    ids, att = out["input_ids"], out["attention_mask"]
    assert ids.shape == (2, 12) and att.shape == (2, 12)
    v = {w: i for i, w in enumerate(words)}
    want = [v[w] for w in ("[CLS]", "a", "photo", "of", "dog", "##s", ",", "a", "red", "car", ".", "[SEP]")]
    assert ids[0].tolist() == want and att[0].tolist() == [1] * 12
    assert ids[1].tolist() == [v["[CLS]"], v["[UNK]"], v["[SEP]"]] + [0] * 9 and att[1].tolist() == [1, 1, 1] + [0] * 9
    with pytest.raises(ValueError):
        blip_tokenizer(12, model_vocab=len(words))        # the two special tokens would not fit the embedding table


def test_counter_based_dropout_masks_behave_like_independent_bernoulli_draws():
    """Statistics of the mask function of include/uniir_b200.h (restated in tests/cpu_kernels.py, bit-exact with the
    kernel per tests/test_dropout_gpu.py): drop rate, no correlation between sites, consecutive seeds, neighbouring
    elements or the rows / columns of a [tokens, width] activation -- each within 4.5 sigma of an ideal generator."""
    import cpu_kernels as ck
    n, p = 1 << 20, 0.1
    idx = torch.arange(n, dtype=torch.int64)
    sig = (p * (1 - p) / n) ** 0.5
    masks = {}
    for seed in (0, 1, 2, 123456789, 2 ** 31 - 2):
        for site in (0, 1, 2, 61, 4096, 4143):
            k = ck.keep_mask(seed, site, idx, p).float()
            assert abs((1 - k.mean().item()) - p) < 4.5 * sig, (seed, site)
            masks[(seed, site)] = k - (1 - p)
    keys = list(masks)
    var = p * (1 - p)
    worst = 0.0
    for i in range(len(keys)):
        for j in range(i + 1, len(keys), 7):                       # a spread of (seed, site) pairs
            worst = max(worst, abs((masks[keys[i]] * masks[keys[j]]).mean().item()) / var)
    assert worst < 4.5 / n ** 0.5 * 1.3, worst                       # correlation coefficient of independent streams
    m = masks[(1, 2)]
    for lag in (1, 2, 3, 64, 256, 768, 1024, 3072, 4096, 768 * 77):      # neighbours, rows of width 768 / 1024 / 3072 / 4096, items
        assert abs((m[lag:] * m[:-lag]).mean().item()) / var < 4.5 / n ** 0.5, lag
    grid = (m + (1 - p))[:1024 * 768].view(1024, 768)                # 1024 tokens x 768 features
    row_rate, col_rate = 1 - grid.mean(1), 1 - grid.mean(0)
    assert (row_rate - p).abs().max().item() < 5 * (var / 768) ** 0.5
    assert (col_rate - p).abs().max().item() < 5 * (var / 1024) ** 0.5
    # batch position, not micro-batch position: element i of a micro-batch starting at item0 == element item0 * per_item + i
    a = ck.keep_mask(7, 3, torch.arange(1000) + 5 * 1000, p)
    b = ck.keep_mask(7, 3, torch.arange(6000), p)[5000:]
    assert torch.equal(a, b)
