"""The C-ABI shared library loads and exports every symbol include/gxsv.h declares (no GPU needed),
and the product path fails loudly without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest

from graphix_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols() -> list[str]:
    text = open(os.path.join(ROOT, "include", "gxsv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gxsv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    _build.build()
    lib = C.CDLL(_build.LIBPATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gxsv.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    assert lib.gxsv_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from graphix_b200 import B200Statevector

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200Statevector(nqubit=1)


def test_product_package_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under graphix_b200/ may import or link it."""
    pkg = os.path.join(ROOT, "graphix_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src, f
