"""Golden vectors of the reference's BLIP train / eval image transform (backbone/transform/blip_transform.py run HERE,
where /root/reference exists) for tests/test_host_logic.py on machines without the reference tree.

    python scripts/make_golden_blip_transform.py     -> tests/golden/blip_transform.npz
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/models/uniir_blip/backbone/transform"


def load_reference_transform():
    pkg = types.ModuleType("ref_blip_transform_pkg")
    pkg.__path__ = [REF]
    sys.modules[pkg.__name__] = pkg
    mods = {}
    for name in ("randaugment", "blip_transform"):
        spec = importlib.util.spec_from_file_location(f"{pkg.__name__}.{name}", os.path.join(REF, f"{name}.py"))
        mods[name] = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mods[name]
        spec.loader.exec_module(mods[name])
    return mods["blip_transform"], mods["randaugment"]


def test_image(seed):
    """Blocky, non-square synthetic picture; every third one has no black pixel (the AutoContrast offset quirk)."""
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(40, 300)), int(rng.integers(40, 300))
    base = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3), dtype=np.uint8)
    if seed % 3 == 0:
        base = (base // 2 + 40).astype(np.uint8)
    return Image.fromarray(base).resize((w, h), Image.BILINEAR)


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    bt, _ = load_reference_transform()
    out = {}
    for seed in range(12):
        img = test_image(seed)
        for train in (True, False):
            torch.manual_seed(seed); np.random.seed(seed)
            out[f"{'train' if train else 'eval'}_{seed}"] = bt.get_blip_transform(32, 0.5, train)(img).numpy()
    path = os.path.join(ROOT, "tests", "golden", "blip_transform.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
