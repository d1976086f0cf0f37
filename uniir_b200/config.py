"""Minimal OmegaConf stand-in: YAML -> attribute-access config with ``${a.b.c}`` interpolation.

The reference loads every YAML with ``OmegaConf.load`` (train.py:327, mbeir_embedder.py:509,
mbeir_retriever.py:734) and relies on attribute access, ``${...}`` interpolation
(inbatch.yaml:6-7) and ``config.x = y`` assignment.  omegaconf is not installable offline, so
the CLIs use this loader; the reference YAML files are accepted unchanged.
"""
from __future__ import annotations

import re

import yaml

_PAT = re.compile(r"\$\{([^${}]+)\}")


class Config(dict):
    """dict with attribute access; nested dicts become Config, lists stay lists."""

    def __init__(self, data=None, _root=None):
        super().__init__()
        object.__setattr__(self, "_root", _root if _root is not None else self)
        for k, v in (data or {}).items():
            super().__setitem__(k, self._wrap(v))

    def _wrap(self, v):
        if isinstance(v, dict) and not isinstance(v, Config):
            return Config(v, self._root)
        if isinstance(v, list):
            return [self._wrap(x) for x in v]
        return v

    def _resolve(self, v, depth=0):
        if isinstance(v, str) and "${" in v:
            if depth > 20:
                raise ValueError(f"interpolation cycle in {v!r}")
            whole = _PAT.fullmatch(v)
            if whole:  # keep the referenced value's type
                return self._resolve(self._lookup(whole.group(1)), depth + 1)
            return _PAT.sub(lambda m: str(self._resolve(self._lookup(m.group(1)), depth + 1)), v)
        return v

    def _lookup(self, dotted):
        node = object.__getattribute__(self, "_root")
        for part in dotted.strip().split("."):
            node = dict.__getitem__(node, part) if isinstance(node, dict) else node[int(part)]
        return node

    def __getitem__(self, k):
        return self._resolve(super().__getitem__(k))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def to_dict(self):
        out = {}
        for k, v in self.items():
            out[k] = v.to_dict() if isinstance(v, Config) else ([x.to_dict() if isinstance(x, Config) else x for x in v] if isinstance(v, list) else v)
        return out


def load(path) -> Config:
    with open(path) as f:
        return Config(yaml.safe_load(f) or {})


def create(data) -> Config:
    return Config(data)
