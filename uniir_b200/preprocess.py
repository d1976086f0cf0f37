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


def blip_image_preprocess(size=224):
    """BLIP eval transform (backbone/transform/blip_transform.py:42-48): bicubic resize to (size,size),
    ToTensor, Normalize with the CLIP mean/std constants (blip_transform.py:9)."""
    mean = torch.tensor(CLIP_MEAN).view(3, 1, 1)
    std = torch.tensor(CLIP_STD).view(3, 1, 1)

    def fn(img):
        from PIL import Image
        img = img.convert("RGB").resize((size, size), Image.BICUBIC)
        t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        return (t - mean) / std

    return fn
