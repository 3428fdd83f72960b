"""REAL graphix driving the CUDA engine on the GPU box.

The reference package travels to the GPU box as the archive staged by ``oracle/stage_ref.py`` (run by
``__graft_entry__.build()`` where ``/root/reference`` exists).  Each check runs in a fresh interpreter in which graphix
is imported before ``graphix_b200``, so that the product backend is a real ``DenseStateBackend`` subclass:

* ``tests/dropin_gpu_worker.py`` — ``pattern.simulate(backend=B200StatevectorBackend.with_capacity(...))`` against
  graphix's own ``backend="statevector"`` on config 1 and on a 16-qubit random pattern (Const selectors, graphix's default
  ``RandomBranchSelector`` with the same rng, every execution mode), and ``Circuit.simulate(backend=...)`` on the device;
* ``tests/reference_conformance.py`` with ``GXSV_CONFORMANCE_ENGINE=cuda`` — the reference's OWN test files
  (test_statevec-independent backend users: test_branch_selector, test_simulator, test_pattern, test_transpiler) with the
  name ``StatevectorBackend`` routed to ``B200StatevectorBackend`` on the real CUDA engine.
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_shim  # noqa: E402


def _skip_reason() -> str | None:
    if not ref_shim.reference_available():
        return (f"reference package not found at {ref_shim.REFERENCE_ROOT}: run oracle/stage_ref.py (or "
                "__graft_entry__.build()) where /root/reference exists")
    missing = ref_shim.missing_dependencies()
    if missing:
        return "the reference needs python modules that this box lacks: " + ", ".join(missing)
    return None


@pytest.fixture(scope="module", autouse=True)
def _need_reference():
    why = _skip_reason()
    if why:
        print("SKIP:", why)
        pytest.skip(why)


def test_real_graphix_drives_the_cuda_engine():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_gpu_worker.py")], capture_output=True,
                         text=True, timeout=900, cwd=ROOT)
    m = re.search(r"DROPIN_RESULT (\{.*\})", out.stdout)
    assert m, (out.stdout[-1500:], out.stderr[-3000:])
    res = json.loads(m.group(1))
    assert res["backend_base"] == "DenseStateBackend"
    assert res["ok"], res["failures"]
    assert res["checks"] >= 30 and res["min_fidelity"] >= 1 - 1e-10


def test_reference_test_files_pass_on_the_cuda_engine():
    env = dict(os.environ, GXSV_CONFORMANCE_ENGINE="cuda")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_conformance.py"), "test_branch_selector.py",
                          "test_simulator.py", "test_pattern.py", "test_transpiler.py"],
                         capture_output=True, text=True, timeout=1500, cwd=ROOT, env=env)
    tail = out.stdout[-2500:]
    m = re.search(r"(\d+) passed", out.stdout)
    assert m, (tail, out.stderr[-2000:])
    assert int(m.group(1)) >= 1000, tail
    assert "failed" not in out.stdout.splitlines()[-1], tail
