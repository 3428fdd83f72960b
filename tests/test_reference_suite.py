"""The reference's OWN test files, with its built-in name "statevector" routed to the product's backend class
(tests/reference_conformance.py).  Runs only where /root/reference exists (the build container).  No GPU here:
the state object under the backend is the CPU oracle, so this checks the Backend host logic — add_nodes /
entangle_nodes / measure / correct_byproduct / apply_clifford / finalize, the BranchSelector protocol, with_capacity —
against everything the reference's pattern, transpiler, simulator and branch-selector tests expect."""
from __future__ import annotations

import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if not os.path.isdir("/root/reference/graphix"):
    pytest.skip("reference tree not present", allow_module_level=True)


def test_reference_pattern_transpiler_simulator_selector_tests_pass_with_the_product_backend():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_conformance.py")],
                         capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = out.stdout[-2000:]
    m = re.search(r"(\d+) passed", out.stdout)
    assert m, tail
    assert int(m.group(1)) >= 1000, tail
    assert "failed" not in out.stdout.splitlines()[-1], tail
