"""Drop-in conformance with the REAL graphix objects (runs only where /root/reference is importable, i.e. the
build container — the GPU box has no graphix, and this container has no GPU).

``B200StatevectorBackend`` must be accepted by graphix's own ``PatternSimulator`` / ``pattern.simulate``
(``isinstance(backend, Backend)``, graphix/simulator.py:549) and its host logic must work on graphix's own
``Measurement`` / ``Clifford`` / ``command`` / ``BranchSelector`` objects.  The amplitude engine is swapped for the
CPU oracle here (the CUDA engine needs a GPU); everything above the state object is the product code.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ref_shim import import_reference, reference_available  # noqa: E402

if not reference_available():
    pytest.skip("reference tree not present", allow_module_level=True)

import_reference()
import importlib  # noqa: E402

import graphix_b200.backend as backend_mod  # noqa: E402

importlib.reload(backend_mod)   # pick up graphix's DenseStateBackend as the base class now that graphix is importable
from graphix.branch_selector import ConstBranchSelector, FixedBranchSelector, RandomBranchSelector  # noqa: E402
from graphix.random_objects import rand_circuit  # noqa: E402
from graphix.sim.base_backend import Backend, DenseStateBackend  # noqa: E402
from graphix.sim.statevec import Statevector, StatevectorBackend  # noqa: E402
from graphix.simulator import PatternSimulator  # noqa: E402
from graphix.states import BasicStates  # noqa: E402

from oracle.oracle import OracleStatevector  # noqa: E402

B200StatevectorBackend = backend_mod.B200StatevectorBackend


def make_backend(cap, **kw):
    return B200StatevectorBackend(OracleStatevector(nqubit=0, max_qubits=cap), **kw)


def test_is_a_graphix_backend():
    b = make_backend(3, branch_selector=ConstBranchSelector(0))
    assert isinstance(b, Backend) and isinstance(b, DenseStateBackend)
    assert b.nqubit == 0
    with pytest.raises(ValueError):
        B200StatevectorBackend(OracleStatevector(nqubit=0), symbolic=True)


@pytest.mark.parametrize("seed", range(4))
def test_pattern_simulate_with_graphix_objects(seed):
    rng = np.random.default_rng(seed)
    circ = rand_circuit(4, 3, rng=rng)
    pattern = circ.transpile().pattern
    pattern.standardize()
    pattern.minimize_space()
    ref = pattern.simulate(backend="statevector", branch_selector=ConstBranchSelector(seed & 1))
    mine = pattern.simulate(backend=make_backend(pattern.max_space(), branch_selector=ConstBranchSelector(seed & 1)))
    assert abs(np.vdot(ref.flatten(), mine.flatten())) ** 2 > 1 - 1e-10
    assert ref.isclose(Statevector(mine.flatten()))
    # graphix refuses selector kwargs together with an instantiated backend (simulator.py:549-556)
    with pytest.raises(ValueError):
        pattern.simulate(backend=make_backend(9), branch_selector=ConstBranchSelector(0))


def test_random_selector_same_rng_and_finalize_contract():
    circ = rand_circuit(3, 2, rng=np.random.default_rng(11))
    pattern = circ.transpile().pattern
    outs = []
    for b in (StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=RandomBranchSelector()),
              make_backend(pattern.max_space(), branch_selector=RandomBranchSelector())):
        sim = PatternSimulator(pattern, backend=b)
        sim.run(rng=np.random.default_rng(5))
        outs.append((dict(sim.measure_method.results), np.asarray(b.state.flatten()), list(b.node_index)))
    assert outs[0][0] == outs[1][0]
    assert abs(np.vdot(outs[0][1], outs[1][1])) ** 2 > 1 - 1e-10
    assert outs[1][2] == pattern.output_nodes          # tests/test_simulator.py:59-63 of the reference


def test_prepared_backend_and_input_state_rules():
    circ = rand_circuit(2, 2, rng=np.random.default_rng(3))
    pattern = circ.transpile().pattern
    b = make_backend(pattern.max_space(), branch_selector=FixedBranchSelector({}, default=ConstBranchSelector(0)))
    b.add_nodes(pattern.input_nodes, BasicStates.ZERO)
    PatternSimulator(pattern, backend=b).run(input_state=None)
    ref = pattern.simulate(backend="statevector", input_state=BasicStates.ZERO, branch_selector=ConstBranchSelector(0))
    assert abs(np.vdot(ref.flatten(), b.state.flatten())) ** 2 > 1 - 1e-10
    with pytest.raises(ValueError):
        PatternSimulator(pattern, backend=b).run()     # backend already holds qubits
