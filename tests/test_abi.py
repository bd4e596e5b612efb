"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dexb200.h declares, and refuses to work without a GPU (no CPU fallback)."""
import os
import re

import pytest

import dextractor_b200 as dx
from dextractor_b200.lib import SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dexb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dx_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = dx.load_library()
    missing = [s for s in _declared() if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dx.DexError) as e:
        dx.Context(0)
    assert e.value.code == -9          # DX_E_NOGPU


def test_product_never_imports_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "dextractor_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".c", ".sh")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"oracle|dx_oracle|libdxoracle|_ref/", txt):
                    bad.append(f)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if not os.path.isfile(os.path.join(ROOT, "tools", f)):
            continue
        txt = open(os.path.join(ROOT, "tools", f), errors="replace").read()
        if re.search(r"oracle|dx_oracle|libdxoracle", txt):
            bad.append(f)
    assert not bad, bad
