"""The C-ABI shared library builds without a GPU, loads, and exports exactly the entry points that
include/uniir_b200.h declares (no compute calls here)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "uniir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uniir_[a-z0-9_]+)\s*\(", src)))


def test_header_and_bindings_agree(built_lib):
    from uniir_b200 import lib
    hdr = _header_functions()
    assert hdr == lib.exported_symbols(), (set(hdr) ^ set(lib.exported_symbols()))


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (uniir_[a-z0-9_]+)", out))
    assert set(_header_functions()) <= exported, set(_header_functions()) - exported


def test_library_loads_and_reports_version(built_lib):
    from uniir_b200 import lib
    l = lib.load()
    assert l.uniir_abi_version() == 1
    assert lib.launch_count() == 0 or lib.launch_count() > 0
    assert l.uniir_last_error() is not None


def test_sass_contains_blackwell_tensor_and_tma_ops(built_lib):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md evidence table)."""
    out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in out and "LDTM" in out and "UTMALDG" in out


def test_product_path_has_no_oracle_import_and_fails_loudly_without_cuda():
    import torch
    pkg = os.path.join(ROOT, "uniir_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in re.sub(r'""".*?"""', "", open(os.path.join(pkg, f)).read(), flags=re.S).replace("oracle/", ""), f
    if not torch.cuda.is_available():
        from uniir_b200 import lib
        from uniir_b200.clip_sf import CLIPScoreFusion
        from uniir_b200.retrieval import FlatIPIndex
        with pytest.raises(lib.UniirError):
            CLIPScoreFusion("tiny-test", device="cpu")
        with pytest.raises(lib.UniirError):
            FlatIPIndex(8, device="cpu")
