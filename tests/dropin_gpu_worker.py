"""Real graphix driving the CUDA engine (run by tests/test_graphix_dropin_gpu.py in a fresh interpreter on the GPU box):

    python tests/dropin_gpu_worker.py            -> one JSON line of results

graphix is imported first (from the archive staged by oracle/stage_ref.py, or /root/reference in the build container),
so that ``graphix_b200.B200StatevectorBackend`` picks graphix's ``DenseStateBackend`` as its base class and is accepted
by graphix's own ``pattern.simulate`` / ``PatternSimulator`` (graphix/simulator.py:396-479,549) and ``Circuit.simulate``
(graphix/transpiler.py:553-642,1165).  Every object handed to the backend is a real graphix object (``Measurement``,
``Clifford``, ``command.*``, ``BranchSelector``, ``PlanarState``); the comparison is against graphix's own
``backend="statevector"`` (numba) on the same pattern, selector and rng.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np  # noqa: E402
from ref_shim import import_reference  # noqa: E402

import_reference()
from graphix import Circuit  # noqa: E402
from graphix.branch_selector import ConstBranchSelector, RandomBranchSelector  # noqa: E402
from graphix.random_objects import rand_circuit  # noqa: E402
from graphix.sim.base_backend import Backend, DenseStateBackend  # noqa: E402
from graphix.sim.statevec import Statevector  # noqa: E402
from graphix.simulator import PatternSimulator  # noqa: E402

import graphix_b200  # noqa: E402
from graphix_b200 import B200Statevector, B200StatevectorBackend  # noqa: E402


def pipeline(circ):
    p = circ.transpile().pattern
    p.standardize()
    p.minimize_space()
    return p


def fid(a, b) -> float:
    return float(abs(np.vdot(np.asarray(a.flatten()), np.asarray(b.flatten()))) ** 2)


def main() -> None:
    out = {"checks": 0, "min_fidelity": 1.0, "failures": []}

    def check(name, ok, detail=None):
        out["checks"] += 1
        if not ok:
            out["failures"].append({"check": name, "detail": detail})

    b = B200StatevectorBackend.with_capacity(3, branch_selector=ConstBranchSelector(0))
    check("isinstance Backend / DenseStateBackend", isinstance(b, Backend) and isinstance(b, DenseStateBackend))
    check("state is a virtual graphix Statevector", isinstance(b.state, Statevector))

    cases = [("config1 rand_circuit(8,5,rng=42)", pipeline(rand_circuit(8, 5, rng=np.random.default_rng(42)))),
             ("16-qubit rand_circuit(15,6,rng=7)", pipeline(rand_circuit(15, 6, rng=np.random.default_rng(7))))]
    for name, pattern in cases:
        out.setdefault("patterns", {})[name] = {"commands": len(list(pattern)), "max_space": pattern.max_space()}
        for sel in (0, 1):
            ref = pattern.simulate(backend="statevector", branch_selector=ConstBranchSelector(sel))
            for kw in ({}, {"fusion": 1}, {"strict": False, "batch": 8}):
                mine = pattern.simulate(backend=B200StatevectorBackend.with_capacity(
                    pattern.max_space(), branch_selector=ConstBranchSelector(sel), **kw))
                f = fid(ref, mine)
                out["min_fidelity"] = min(out["min_fidelity"], f)
                check(f"{name} const{sel} {kw}", f >= 1 - 1e-10 and isinstance(mine, B200Statevector), f)
                check(f"{name} const{sel} {kw} isclose both ways", ref.isclose(mine) and mine.isclose(ref))
        # graphix's default selector (RandomBranchSelector, pr_calc=True): same rng -> same outcomes, same state
        res = []
        for backend in ("statevector", B200StatevectorBackend.with_capacity(pattern.max_space(),
                                                                            branch_selector=RandomBranchSelector())):
            sim = PatternSimulator(pattern, backend=backend) if not isinstance(backend, str) else \
                PatternSimulator(pattern, backend=backend, branch_selector=RandomBranchSelector())
            sim.run(rng=np.random.default_rng(2026))
            res.append((dict(sim.measure_method.results), sim.backend.state, list(sim.backend.node_index)))
        check(f"{name} random selector: same outcomes", res[0][0] == res[1][0])
        f = fid(res[0][1], res[1][1])
        out["min_fidelity"] = min(out["min_fidelity"], f)
        check(f"{name} random selector: same state", f >= 1 - 1e-10, f)
        check(f"{name} finalize contract", res[1][2] == list(pattern.output_nodes))
        # selector kwargs together with an instantiated backend are refused by graphix (simulator.py:549-556)
        try:
            pattern.simulate(backend=B200StatevectorBackend.with_capacity(pattern.max_space()),
                             branch_selector=ConstBranchSelector(0))
            check(f"{name} kwargs refused", False)
        except ValueError:
            check(f"{name} kwargs refused", True)

    # Circuit.simulate on the device: evolve_single / evolve(k = 2, 3) / swap / entangle / measure
    rng = np.random.default_rng(5)
    circ = rand_circuit(10, 6, rng=rng, use_ccx=True, use_rzz=True)
    circ.swap(0, 5)
    circ.cz(1, 2)
    ref = circ.simulate(branch_selector=ConstBranchSelector(0)).state
    mine = circ.simulate(backend=B200StatevectorBackend.with_capacity(circ.width, branch_selector=ConstBranchSelector(0))).state
    f = fid(ref, mine)
    out["min_fidelity"] = min(out["min_fidelity"], f)
    check("Circuit.simulate on the device vs numba", f >= 1 - 1e-10, f)
    # the circuit and its pattern agree on the device (tests/test_transpiler.py of the reference)
    pat = pipeline(circ)
    mine_p = pat.simulate(backend=B200StatevectorBackend.with_capacity(pat.max_space(), branch_selector=ConstBranchSelector(0)))
    f = fid(mine, mine_p)
    out["min_fidelity"] = min(out["min_fidelity"], f)
    check("pattern == circuit, both on the device", f >= 1 - 1e-10, f)
    c2 = Circuit(3)
    c2.rx(0, 0.7)
    c2.cnot(0, 1)
    c2.m(0, __import__("graphix.fundamentals", fromlist=["Axis"]).Axis.Z)
    r_ref = c2.simulate(branch_selector=ConstBranchSelector(1))
    r_me = c2.simulate(backend=B200StatevectorBackend.with_capacity(3, branch_selector=ConstBranchSelector(1)))
    check("Circuit.m on the device", tuple(r_ref.classical_measures) == tuple(r_me.classical_measures)
          and fid(r_ref.state, r_me.state) >= 1 - 1e-10)
    out["graphix"] = getattr(sys.modules["graphix"], "__file__", "?")
    out["backend_base"] = B200StatevectorBackend.__mro__[1].__name__
    out["ok"] = not out["failures"]
    print("DROPIN_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
