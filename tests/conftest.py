import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# BertConfig overrides that switch the train-mode dropout off, as the reference pin tests do (tests/test_oracle_blip.py):
# the oracle and the golden vectors are the deterministic (eval-mode) arithmetic
NO_DROP = {"hidden_dropout_prob": 0.0, "attention_probs_dropout_prob": 0.0}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    """libuniir_b200.so built in-tree (nvcc cross-compiles without a GPU)."""
    from uniir_b200 import build
    return build.build(verbose=False)
