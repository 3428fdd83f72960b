"""Import the read-only reference (graphix) in THIS container.

Test/fixture tooling only: used by ``tests/golden/make_golden.py`` and by the
optional conformance tests that run when ``/root/reference`` is present.  It is
never imported by the product package, ``bench.py`` or ``smoke()``.

The reference imports every backend eagerly (graphix/sim/__init__.py:5-9), so
three absent modules are stubbed before ``import graphix``:
``graphix._version`` (setuptools_scm artefact), ``quimb``/``quimb.tensor`` and
``matplotlib``.
"""
from __future__ import annotations

import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("GRAPHIX_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graphix"))


def import_reference():
    """Return the imported ``graphix`` package of the reference tree."""
    if "graphix" in sys.modules:
        return sys.modules["graphix"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.path.insert(0, REFERENCE_ROOT)
    v = types.ModuleType("graphix._version")
    v.__version__ = v.version = "0+ref"
    sys.modules["graphix._version"] = v
    if "quimb" not in sys.modules:
        q, qt = types.ModuleType("quimb"), types.ModuleType("quimb.tensor")
        qt.Tensor = type("Tensor", (), {})
        qt.TensorNetwork = type("TensorNetwork", (), {})
        q.tensor = qt
        sys.modules["quimb"], sys.modules["quimb.tensor"] = q, qt
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.transforms", "matplotlib.patches",
              "matplotlib.lines", "matplotlib.axes", "matplotlib.figure", "matplotlib.text",
              "matplotlib.typing", "matplotlib.collections", "matplotlib.path", "matplotlib.colors",
              "matplotlib.legend_handler"):
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    import graphix  # noqa: PLC0415

    return graphix
