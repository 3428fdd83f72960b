"""Stage the reference package for the GPU box:  python oracle/stage_ref.py

The reference (TeamGraphix/graphix) is pure Python + numba, so "building" it is packaging: the package
(`/root/reference/graphix`) and its test files are archived — by this recipe, never by hand, never into tracked paths —
into ONE zip under the git-ignored `oracle/_ref/` (`graphix_ref.zip`, importable as is through zipimport; it travels to
the GPU box with the snapshot like the built `.so` files), together with the setuptools_scm artefact
`graphix/_version.py` that an installed copy would carry (graphix/__init__.py:5, pyproject.toml:80-81).  `quimb` and
`matplotlib` stay absent; `tools/ref_shim.py` stubs them at import time.

Test infrastructure only: used by `tests/test_graphix_dropin_gpu.py` (real graphix driving the CUDA engine) and by the
reference arm of `bench.py` (`--impl reference`: the reference's own numba kernels on the host cores, kind "reference").
The product package never imports anything below `oracle/`.
"""
from __future__ import annotations

import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GRAPHIX_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "graphix_ref.zip")


def stage(verbose: bool = True) -> str | None:
    """Write oracle/_ref/graphix_ref.zip from the reference tree; returns its path (None if there is nothing)."""
    if not os.path.isdir(os.path.join(SRC, "graphix")):
        if verbose:
            print(f"stage_ref: {SRC}/graphix not present — keeping {ARCHIVE} as is", file=sys.stderr)
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(DST, exist_ok=True)
    n = 0
    tmp = ARCHIVE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for sub in ("graphix", "tests"):
            for root, dirs, files in os.walk(os.path.join(SRC, sub)):
                dirs[:] = sorted(d for d in dirs if d not in ("__pycache__", ".pytest_cache"))
                for f in sorted(files):
                    if f.endswith(".pyc"):
                        continue
                    full = os.path.join(root, f)
                    z.write(full, os.path.relpath(full, SRC))
                    n += 1
        z.writestr("graphix/_version.py", '__version__ = version = "0+ref.staged"\n')
    os.replace(tmp, ARCHIVE)
    if verbose:
        print(f"stage_ref: {n} files -> {ARCHIVE} ({os.path.getsize(ARCHIVE)} bytes)")
    return ARCHIVE


if __name__ == "__main__":
    stage()
