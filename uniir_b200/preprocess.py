"""Host-side input preparation with the reference's contracts:
  * image: PIL -> Tensor[3,224,224], bicubic resize of the short side to 224, centre crop,
    CLIP normalisation (the transform ``clip.load`` returns; clip_sf.py:25,32-33)
  * text : ``clip.tokenize(txt, context_length=77, truncate=True)`` (clip_sf.py:35-41)
The BPE vocabulary file ships with the openai/CLIP package, which is not vendored in the
reference and not available offline; point UNIIR_CLIP_BPE at ``bpe_simple_vocab_16e6.txt.gz``
to tokenise real text.  Synthetic token ids (bench / tests) do not need it.
Known deviation: upstream ``clip.tokenize`` first runs ``ftfy.fix_text`` (mojibake repair); ftfy is not
installable here, so text is only HTML-unescaped and whitespace-normalised -- well-formed UTF-8 text tokenises
identically, text with encoding damage does not.
"""
from __future__ import annotations

import gzip
import html
import os
import re
from functools import lru_cache

import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
SOT, EOT = 49406, 49407


def clip_image_preprocess(size=224):
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(CLIP_STD).view(3, 1, 1)

    def fn(img):
        from PIL import Image
        # torchvision arithmetic: Resize(size) keeps the short side at `size` and truncates the long one
        # (int(size * long / short)); CenterCrop rounds half the margin (int(round((h - size) / 2.0)))
        w, h = img.size
        if w <= h:
            nw, nh = size, int(size * h / w)
        else:
            nw, nh = int(size * w / h), size
        if (nw, nh) != (w, h):
            img = img.resize((nw, nh), Image.BICUBIC)
        left, top = int(round((nw - size) / 2.0)), int(round((nh - size) / 2.0))
        img = img.crop((left, top, left + size, top + size)).convert("RGB")
        t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        return (t - mean) / std

    return fn


@lru_cache()
def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


_SPECIALS = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|"


def bpe_pretokenizer(force_stdlib=False):
    """Upstream pattern: ...|[\\p{L}]+|[\\p{N}]|[^\\s\\p{L}\\p{N}]+ (case-insensitive) -- Unicode letters / numbers, not
    ASCII ranges: accented and CJK text (WebQA, OVEN, InfoSeek, EDIS) must split the way the checkpoints saw it.
    Uses the `regex` module when importable; otherwise stdlib `re` with the closest classes it has
    ([^\\W\\d_] = letters, \\d = decimal digits; the other \\p{N} characters such as superscripts and vulgar
    fractions are then grouped with the punctuation run -- tests/test_host_logic.py quantifies it)."""
    if not force_stdlib:
        try:
            import regex
            return regex.compile(_SPECIALS + r"[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)
        except ImportError:
            pass
    return re.compile(_SPECIALS + r"[^\W\d_]+|\d|(?:[^\s\w]|_)+", re.IGNORECASE)


class _BPE:
    """Byte-pair encoder over the published CLIP merges file (same algorithm as GPT-2's)."""

    def __init__(self, path):
        merges = gzip.open(path).read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in merges[1:49152 - 256 - 2 + 1]]
        b2u = _bytes_to_unicode()
        vocab = list(b2u.values())
        vocab = vocab + [v + "</w>" for v in vocab]
        vocab += ["".join(m) for m in merges]
        vocab += ["<|startoftext|>", "<|endoftext|>"]
        self.enc = dict(zip(vocab, range(len(vocab))))
        self.ranks = dict(zip(merges, range(len(merges))))
        self.b2u = b2u
        self.cache = {}
        self.pat = bpe_pretokenizer()

    def _bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        self.cache[token] = word
        return word

    def encode(self, text):
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids = []
        for tok in self.pat.findall(text):
            tok = "".join(self.b2u[b] for b in tok.encode("utf-8"))
            ids.extend(self.enc[t] for t in self._bpe(tok))
        return ids


def word_hash_tokenizer(context_length=77, vocab=49408):
    """NOT the CLIP vocabulary: [SOT, 1 + crc32(word) % (vocab - 3) ..., EOT, 0 ...] per whitespace-separated word,
    truncated like clip.tokenize(truncate=True).  For pipeline runs where the BPE merges file is unavailable (no
    network): synthetic data, smoke tests, the end-to-end CLI tests.  Selected with UNIIR_CLIP_BPE=word-hash."""
    import zlib
    sot, eot = vocab - 2, vocab - 1

    def tok(txts):
        if isinstance(txts, str):
            txts = [txts]
        out = torch.zeros(len(txts), context_length, dtype=torch.int32)
        for i, t in enumerate(txts):
            ids = [sot] + [1 + zlib.crc32(w.encode()) % (vocab - 3) for w in t.lower().split()] + [eot]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = eot
            out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
        return out
    return tok


def clip_tokenizer(context_length=77, vocab=49408):
    state = {}
    if os.environ.get("UNIIR_CLIP_BPE") == "word-hash":
        return word_hash_tokenizer(context_length, vocab)

    def tokenizer_wrapper(txt):
        if isinstance(txt, str):
            txt = [txt]
        if "bpe" not in state:
            path = os.environ.get("UNIIR_CLIP_BPE")
            if not path or not os.path.exists(path):
                raise FileNotFoundError(
                    "CLIP BPE vocabulary not found: set UNIIR_CLIP_BPE=/path/to/bpe_simple_vocab_16e6.txt.gz "
                    "(it ships with openai/CLIP, which the reference installs from git)")
            state["bpe"] = _BPE(path)
        out = torch.zeros(len(txt), context_length, dtype=torch.int32)
        for i, t in enumerate(txt):
            ids = [SOT] + state["bpe"].encode(t) + [EOT]
            if len(ids) > context_length:  # truncate=True
                ids = ids[:context_length]
                ids[-1] = EOT
            out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
        return out

    return tokenizer_wrapper


def synthetic_clip_tokens(n, generator=None, context_length=77, vocab=49408):
    """[SOT, U{1..vocab-3} x L, EOT, 0...] with L ~ U{1..75}."""
    ids = torch.zeros(n, context_length, dtype=torch.int32)
    lens = torch.randint(1, context_length - 1, (n,), generator=generator)
    body = torch.randint(1, vocab - 2, (n, context_length), generator=generator, dtype=torch.int32)
    for i in range(n):
        Ln = int(lens[i])
        ids[i, 0] = vocab - 2
        ids[i, 1:1 + Ln] = body[i, :Ln]
        ids[i, 1 + Ln] = vocab - 1
    return ids


# ---- BLIP train-time augmentation (backbone/transform/blip_transform.py:11-40): RandomResizedCrop(scale=(0.5, 1),
# bicubic) -> RandomHorizontalFlip -> RandomAugment(N=2, M=5, ten ops) -> ToTensor -> Normalize.  The random streams
# are consumed in the reference's order -- torch's global RNG for crop and flip (torchvision), numpy's global RNG for
# the augment ops (backbone/transform/randaugment.py) -- so the same seeds give the same pixels.

BLIP_AUG_OPS = ("Identity", "AutoContrast", "Brightness", "Sharpness", "Equalize", "ShearX", "ShearY", "TranslateX",
                "TranslateY", "Rotate")                                            # blip_transform.py:24-35
_AUG_FILL = (128, 128, 128)                                                        # randaugment.py:267 replace_value
_AUG_MAX_LEVEL = 10


def _lut_per_channel(img, table_of):
    """Apply a 256-entry lookup table, computed per channel by `table_of(channel)` (None = leave the channel)."""
    out = img.copy()
    for c in range(img.shape[2]):
        t = table_of(img[:, :, c])
        if t is not None:
            out[:, :, c] = t[img[:, :, c]]
    return out


def _autocontrast_table(ch):
    """randaugment.py:10-40 with cutoff 0: stretch [min, max] to [0, 255].  The reference computes the offset as
    `-low * scale` with `low` a numpy uint8 scalar, whose negation wraps modulo 256; reproduced here, because a drop-in
    has to give the reference's pixels: a channel whose minimum is not 0 saturates."""
    lo, hi = int(ch.min()), int(ch.max())
    if hi <= lo:
        return None
    scale = 255 / (hi - lo)
    offset = ((-lo) & 0xFF) * scale
    return np.clip(np.clip(np.arange(256) * scale + offset, 0, 255), 0, 255).astype(np.uint8)


def _equalize_table(ch):
    """PIL.ImageOps.equalize's table as randaugment.py:43-63 builds it (integer arithmetic: the counts are exact)."""
    hist = np.bincount(ch.reshape(-1), minlength=256).astype(np.int64)
    nz = hist[hist != 0]
    step = int(nz[:-1].sum()) // 255
    if step == 0:
        return None
    shifted = np.concatenate([[step // 2], hist[:-1]])
    return np.clip(np.cumsum(shifted) // step, 0, 255).astype(np.uint8)


def _warp(img, matrix):
    import cv2       # the reference's augment ops are cv2 calls (randaugment.py:1); same library, same pixels
    h, w = img.shape[:2]
    return cv2.warpAffine(img, np.float32(matrix), (w, h), borderValue=_AUG_FILL, flags=cv2.INTER_LINEAR).astype(np.uint8)


def _sharpness(img, factor):
    """randaugment.py:127-144: blend with a 3x3 smoothing (centre weight 5, /13), interior pixels only."""
    import cv2
    if factor == 1.0:
        return img
    k = np.ones((3, 3), dtype=np.float32)
    k[1, 1] = 5
    k /= 13
    blurred = cv2.filter2D(img, -1, k)
    if factor == 0.0:
        return blurred
    out = img.astype(np.float32)
    inner = blurred.astype(np.float32)[1:-1, 1:-1, :]
    out[1:-1, 1:-1, :] = inner + factor * (out[1:-1, 1:-1, :] - inner)
    return out.astype(np.uint8)


def _signed(magnitude, negative_if):
    """One numpy draw decides the sign (randaugment.py:186-196, 250-256)."""
    return -magnitude if negative_if(np.random.random()) else magnitude


def blip_random_augment(img, n_ops=2, level=5, ops=BLIP_AUG_OPS):
    """RandomAugment(N, M, isPIL=True, augs) of randaugment.py:289-317 on an HWC uint8 array: N ops drawn with
    replacement, each applied with probability 0.5 at magnitude level / 10 of its range."""
    img = np.array(img)
    frac = level / _AUG_MAX_LEVEL
    for name in np.random.choice(list(ops), n_ops):
        if np.random.random() > 0.5:
            continue
        if name == "Identity":
            pass
        elif name == "AutoContrast":
            img = _lut_per_channel(img, _autocontrast_table)
        elif name == "Equalize":
            img = _lut_per_channel(img, _equalize_table)
        elif name == "Brightness":                                  # enhance factor 0.1 .. 1.9 (1.0 at level 5)
            table = (np.arange(256, dtype=np.float32) * (frac * 1.8 + 0.1)).clip(0, 255).astype(np.uint8)
            img = table[img]
        elif name == "Sharpness":
            img = _sharpness(img, frac * 1.8 + 0.1)
        elif name == "ShearX":
            f = _signed(frac * 0.3, lambda u: u > 0.5)
            img = _warp(img, [[1, f, 0], [0, 1, 0]])
        elif name == "ShearY":
            f = _signed(frac * 0.3, lambda u: u > 0.5)
            img = _warp(img, [[1, 0, 0], [f, 1, 0]])
        elif name == "TranslateX":
            o = _signed(frac * 10.0, lambda u: u > 0.5)
            img = _warp(img, [[1, 0, -o], [0, 1, 0]])
        elif name == "TranslateY":
            o = _signed(frac * 10.0, lambda u: u > 0.5)
            img = _warp(img, [[1, 0, 0], [0, 1, -o]])
        elif name == "Rotate":
            import cv2
            deg = _signed(frac * 30, lambda u: u < 0.5)
            h, w = img.shape[:2]
            img = cv2.warpAffine(img, cv2.getRotationMatrix2D((w / 2, h / 2), deg, 1), (w, h), borderValue=_AUG_FILL)
        else:
            raise ValueError(f"unknown augmentation op {name!r}")
    return img


def _random_resized_crop_box(width, height, scale, ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision.transforms.RandomResizedCrop.get_params on torch's global RNG: ten attempts at a box of
    scale x area and log-uniform aspect ratio, then the central fallback.  Returns (top, left, h, w)."""
    import math
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target * aspect)))
        h = int(round(math.sqrt(target / aspect)))
        if 0 < w <= width and 0 < h <= height:
            top = torch.randint(0, height - h + 1, size=(1,)).item()
            left = torch.randint(0, width - w + 1, size=(1,)).item()
            return top, left, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w, h = width, int(round(width / min(ratio)))
    elif in_ratio > max(ratio):
        h, w = height, int(round(height * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def blip_image_preprocess(size=224, is_train=False, min_scale=0.5):
    """get_blip_transform(image_size, min_scale, is_train) of backbone/transform/blip_transform.py.
    Eval (:42-48): bicubic resize to (size, size), ToTensor, Normalize with the CLIP mean / std constants (:9).
    Train (:11-40): RandomResizedCrop(size, scale=(min_scale, 1), bicubic), RandomHorizontalFlip, RandomAugment(2, 5,
    the ten ops of BLIP_AUG_OPS), ToTensor, Normalize -- random draws in the reference's order and from the same global
    generators (torch: crop, flip; numpy: augment), so equal seeds give equal tensors (tests/test_host_logic.py)."""
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(CLIP_STD).view(3, 1, 1)

    def to_tensor(arr):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint8)).permute(2, 0, 1).float() / 255.0
        return (t - mean) / std

    def eval_fn(img):
        from PIL import Image
        return to_tensor(np.asarray(img.convert("RGB").resize((size, size), Image.BICUBIC)))

    def train_fn(img):
        from PIL import Image
        w, h = img.size
        top, left, ch, cw = _random_resized_crop_box(w, h, (min_scale, 1.0))
        img = img.crop((left, top, left + cw, top + ch)).resize((size, size), Image.BICUBIC)
        if torch.rand(1) < 0.5:
            img = img.transpose(Image.FLIP_LEFT_RIGHT)
        return to_tensor(blip_random_augment(img))

    return train_fn if is_train else eval_fn


def blip_tokenizer(max_length=100, model_vocab=None, vocab_path=None):
    """The reference's BLIP tokenizer (blip.py:221-226 + blip_ff.py:70-80): BertTokenizer("bert-base-uncased") with the
    "[DEC]" (bos) and "[ENC]" special tokens appended, called with padding="max_length", truncation=True,
    max_length=config.tokenizer_max_length, return_tensors="pt".  Returns (wrapper, tokenizer).
    The vocabulary is not in the reference tree and there is no network: `vocab_path` / UNIIR_BERT_VOCAB must point at
    bert-base-uncased's vocab.txt (or a directory holding it, e.g. a HuggingFace snapshot)."""
    path = vocab_path or os.environ.get("UNIIR_BERT_VOCAB", "")
    if os.path.isdir(path):
        path = os.path.join(path, "vocab.txt")
    if not os.path.isfile(path):
        raise FileNotFoundError("the bert-base-uncased vocabulary (BertTokenizer, blip.py:221-226) is not available offline: "
                                "set UNIIR_BERT_VOCAB to its vocab.txt, or pass pre-tokenised "
                                "{'input_ids','attention_mask'} batches")
    import inspect
    from transformers import BertTokenizer
    if "vocab" in inspect.signature(BertTokenizer.__init__).parameters:     # transformers 5: token -> id mapping
        with open(path, encoding="utf-8") as f:
            vocab = {line.rstrip("\n"): i for i, line in enumerate(f)}
        tok = BertTokenizer(vocab=vocab, do_lower_case=True)               # bert-base-uncased's tokenizer_config.json
    else:                                                                   # transformers 4 (the reference pins 4.15)
        tok = BertTokenizer(vocab_file=path, do_lower_case=True)
    tok.add_special_tokens({"bos_token": "[DEC]"})
    tok.add_special_tokens({"additional_special_tokens": ["[ENC]"]})
    tok.enc_token_id = tok.convert_tokens_to_ids("[ENC]")      # == additional_special_tokens_ids[0] (attribute gone in transformers 5)
    if model_vocab is not None and len(tok) > model_vocab:
        raise ValueError(f"{path}: {len(tok)} tokens (with [DEC], [ENC]) exceed the model's vocabulary of {model_vocab}")

    def tokenizer_wrapper(txt):
        return tok(txt, padding="max_length", truncation=True, max_length=max_length, return_tensors="pt")
    return tokenizer_wrapper, tok
