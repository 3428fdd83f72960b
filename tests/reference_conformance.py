"""Run the REFERENCE's own test files with its "statevector" backend name routed to B200StatevectorBackend.

Container-only tooling (needs /root/reference; no GPU here, so the amplitude engine under the backend is the CPU
oracle — what is exercised is the product's Backend host logic against graphix's real driver, commands, measurements,
branch selectors and test expectations).  INTEGRATION.md §3 describes the same substitution on a GPU machine.

    python tests/reference_conformance.py [pytest args...]
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_shim import REFERENCE_ROOT, extract_tests, import_reference  # noqa: E402

import_reference()
import pytest  # noqa: E402

import graphix.simulator  # noqa: E402
import graphix.transpiler  # noqa: E402

import graphix_b200.backend as backend_mod  # noqa: E402
from oracle.oracle import OracleStatevector  # noqa: E402


class OracleBackedBackend(backend_mod.B200StatevectorBackend):
    """B200StatevectorBackend host logic over the CPU oracle state (no GPU in this container)."""

    def __init__(self, state=None, **kwargs):
        if state is None:
            state = OracleStatevector(nqubit=0)
        super().__init__(state, **kwargs)

    @classmethod
    def with_capacity(cls, max_qubits, state=None, **kwargs):
        st = OracleStatevector(nqubit=0, max_qubits=max_qubits) if state is None else OracleStatevector(state, max_qubits=max_qubits)
        return cls(st, **kwargs)


class OracleBackedDMBackend(__import__("graphix_b200.density_matrix", fromlist=["x"]).B200DensityMatrixBackend):
    """B200DensityMatrixBackend host logic (measure / apply_noise / ...) over the numpy density-matrix oracle."""

    def __init__(self, state=None, **kwargs):
        from oracle.dm_oracle import OracleDensityMatrix

        kwargs.pop("symbolic", None)
        super().__init__(OracleDensityMatrix(nqubit=0) if state is None else state, **kwargs)


def main() -> int:
    from graphix.sim.density_matrix import DensityMatrix
    from graphix.sim.statevec import Statevector
    from oracle.dm_oracle import OracleDensityMatrix

    if os.environ.get("GXSV_CONFORMANCE_ENGINE") == "cuda":
        # GPU box: the real CUDA engine under the product backends (tests/test_graphix_dropin_gpu.py)
        import graphix_b200

        graphix.simulator.StatevectorBackend = graphix_b200.B200StatevectorBackend
        graphix.transpiler.StatevectorBackend = graphix_b200.B200StatevectorBackend
        graphix.simulator.DensityMatrixBackend = graphix_b200.B200DensityMatrixBackend
    else:
        DensityMatrix.register(OracleDensityMatrix)
        graphix.simulator.DensityMatrixBackend = OracleBackedDMBackend
        Statevector.register(OracleStatevector)   # same virtual-subclass registration graphix_b200.statevec does for B200Statevector
        graphix.simulator.StatevectorBackend = OracleBackedBackend
        graphix.transpiler.StatevectorBackend = OracleBackedBackend
    files = [a for a in sys.argv[1:] if not a.startswith("-")] or ["test_pattern.py", "test_simulator.py",
                                                                   "test_branch_selector.py", "test_transpiler.py",
                                                                   "test_noisy_density_matrix.py", "test_noise_model.py"]
    extra = [a for a in sys.argv[1:] if a.startswith("-")]
    # tensor-network tests cannot run here (quimb is a stub in this container) and are not on the path
    import tempfile

    with tempfile.TemporaryDirectory() as tmp:
        tests_dir = extract_tests(tmp)        # real files: the reference tree itself, or extracted from the staged archive
        rootdir = os.path.dirname(tests_dir)
        args = ["-p", "no:cacheprovider", "-o", "addopts=", "-c", "/dev/null", f"--rootdir={rootdir}", "-q",
                "-W", "ignore", "-k", "not tensornetwork and not _tn and not mps",
                *[os.path.join(tests_dir, f) for f in files], *extra]
        return pytest.main(args)


if __name__ == "__main__":
    raise SystemExit(main())
