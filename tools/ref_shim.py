"""Import the read-only reference (graphix) in THIS container.

Test/fixture tooling only: used by ``tests/golden/make_golden.py``, by the conformance
tests that run when the reference is present (``/root/reference`` in the build
container, the copy staged under ``oracle/_ref`` by ``oracle/stage_ref.py`` on the GPU
box) and by the reference arm of ``bench.py``.  It is never imported by the product
package or ``smoke()``.

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

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "graphix_ref.zip")


def _find_root() -> str:
    """The read-only reference tree in the build container, else the archive staged by oracle/stage_ref.py (GPU box;
    a zip on sys.path is imported through zipimport)."""
    env = os.environ.get("GRAPHIX_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/graphix") and not os.environ.get("GRAPHIX_REFERENCE_STAGED_ONLY"):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    if REFERENCE_ROOT.endswith(".zip"):
        return os.path.isfile(REFERENCE_ROOT)
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "graphix"))


def missing_dependencies() -> list[str]:
    """Third-party modules the reference needs that this interpreter lacks (numba, networkx, sympy, ...)."""
    import importlib.util

    return [m for m in ("numba", "networkx", "sympy", "scipy", "typing_extensions")
            if importlib.util.find_spec(m) is None]


def extract_tests(dest: str) -> str:
    """The reference's own test files as real files under ``dest`` (pytest cannot collect from a zip)."""
    if REFERENCE_ROOT.endswith(".zip"):
        import zipfile

        with zipfile.ZipFile(REFERENCE_ROOT) as z:
            z.extractall(dest, [n for n in z.namelist() if n.startswith("tests/")])
        return os.path.join(dest, "tests")
    return os.path.join(REFERENCE_ROOT, "tests")


def import_reference():
    """Return the imported ``graphix`` package of the reference tree."""
    if "graphix" in sys.modules:
        return sys.modules["graphix"]
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.path.insert(0, REFERENCE_ROOT)
    if not REFERENCE_ROOT.endswith(".zip") and not os.path.exists(os.path.join(REFERENCE_ROOT, "graphix", "_version.py")):
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
