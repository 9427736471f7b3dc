"""CPU tests of the drop-in boundary: libkaito_rag.so loads here (no GPU) and exports every
symbol include/kaito_rag.h declares; without a device krag_init fails loudly (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "kaito_rag.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(krag_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from kaito_b200 import _native
    assert sorted(_native.SYMBOLS) == _declared_symbols()


def test_library_exports_every_declared_symbol():
    from kaito_b200 import _native
    L = _native.load()
    for name in _declared_symbols():
        assert getattr(L, name) is not None, name
    assert L.krag_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device error path cannot be exercised")
    from kaito_b200 import _native
    with pytest.raises(_native.KragError) as e:
        _native.Context(device_id=0)
    assert e.value.code == _native.KRAG_E_NO_DEVICE


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "kaito_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libkrag_oracle" not in txt and not re.search(r'#include\s*"[^"]*oracle', txt), f
