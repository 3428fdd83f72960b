"""Test configuration.

``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol check (no GPU needed).
``-m gpu``      : parity tests proper — the CUDA path through the C ABI vs the oracle / golden vectors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def kernels_golden():
    return dict(np.load(os.path.join(GOLDEN, "kernels.npz")))


@pytest.fixture()
def fx_rng():
    return np.random.default_rng(25)
