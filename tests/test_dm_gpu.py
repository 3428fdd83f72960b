"""Density-matrix path on the GPU (vectorised rho on the gxsv engine) against golden data of the reference
DensityMatrix / DensityMatrixBackend with DepolarisingNoiseModel, and against the numpy oracle."""
from __future__ import annotations

import os

import numpy as np
import pytest
from helpers import GOLDEN, check_dm_trace, load_golden, rand_state

from graphix_b200 import B200DensityMatrix, B200DensityMatrixBackend, PatternSimulator, mbqc, noise
from oracle.dm_oracle import OracleDensityMatrix
from oracle.oracle import outcome_to_operator_matrix

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "dm_kernels.npz")))


@pytest.mark.parametrize("n", range(1, 5))
def test_dm_kernels_against_reference_golden(g, n):
    rho, op = g[f"n{n}_rho"], g[f"n{n}_op"]
    np.testing.assert_allclose(B200DensityMatrix(rho).rho, rho, atol=1e-15)
    dm = B200DensityMatrix(rho)
    dm.add_nodes(1, mbqc.BasicStates.PLUS)
    np.testing.assert_allclose(dm.rho, g[f"n{n}_addplus"], atol=TOL)
    dm = B200DensityMatrix(rho)
    dm.add_nodes(1, g[f"n{n}_s1"])
    np.testing.assert_allclose(dm.rho, g[f"n{n}_add1"], atol=TOL)
    for q in range(n):
        dm = B200DensityMatrix(rho)
        dm.evolve_single(op, q)
        np.testing.assert_allclose(dm.rho, g[f"n{n}_evolve_q{q}"], atol=TOL)
        assert abs(B200DensityMatrix(rho).expectation_single(op, q) - complex(g[f"n{n}_expect_q{q}"])) < TOL
        dm = B200DensityMatrix(rho)
        dm.apply_channel(noise.depolarising_channel(0.13), [q])
        np.testing.assert_allclose(dm.rho, g[f"n{n}_depol_q{q}"], atol=TOL)
        for r in (0, 1):
            pm = g[f"n{n}_proj_q{q}_r{r}_op"]
            assert abs(B200DensityMatrix(rho).expectation_single(pm, q) - complex(g[f"n{n}_proj_q{q}_r{r}_p"])) < TOL
            dm = B200DensityMatrix(rho)
            dm.project_qubit(pm, q)
            assert dm.nqubit == n - 1
            np.testing.assert_allclose(dm.rho, g[f"n{n}_proj_q{q}_r{r}"], atol=1e-11)
        for q2 in range(n):
            if q2 == q:
                continue
            dm = B200DensityMatrix(rho)
            dm.entangle((q, q2))
            np.testing.assert_allclose(dm.rho, g[f"n{n}_cz_{q}_{q2}"], atol=TOL)
            dm = B200DensityMatrix(rho)
            dm.apply_channel(noise.two_qubit_depolarising_channel(0.21), [q, q2])
            np.testing.assert_allclose(dm.rho, g[f"n{n}_depol2_{q}_{q2}"], atol=TOL)
            dm = B200DensityMatrix(rho)
            dm.evolve(g[f"n{n}_ev2_{q}_{q2}_op"], [q, q2])
            np.testing.assert_allclose(dm.rho, g[f"n{n}_ev2_{q}_{q2}"], atol=1e-11)
    dm = B200DensityMatrix(rho)
    dm.permute([int(x) for x in g[f"n{n}_perm_p"]])
    np.testing.assert_allclose(dm.rho, g[f"n{n}_perm"], atol=TOL)


@pytest.mark.parametrize("name", ["dm_rand3x2", "dm_handbuilt", "dm_rand5x3", "dm_rand9x3"])
def test_noisy_patterns_against_reference_golden(name):
    d = load_golden(name)
    for trace in d["traces"]:
        check_dm_trace(d, trace, lambda cap, sel: B200DensityMatrixBackend.with_capacity(cap, branch_selector=sel))


def test_dm_soup_against_oracle():
    rng = np.random.default_rng(31)
    dev, orc = B200DensityMatrix(nqubit=0), OracleDensityMatrix(nqubit=0)
    states = [mbqc.BasicStates.PLUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE, mbqc.PlanarState(mbqc.Plane.YZ, 0.37)]
    for step in range(220):
        n = orc.nqubit
        kind = rng.choice(["N", "E", "M", "U", "C1", "C2"], p=[0.28, 0.25, 0.2, 0.1, 0.1, 0.07])
        if kind == "N" and n < 6:
            st = states[int(rng.integers(len(states)))]
            dev.add_nodes(1, st)
            orc.add_nodes(1, st)
        elif kind == "E" and n >= 2:
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            dev.entangle((a, b))
            orc.entangle((a, b))
        elif kind == "M" and n >= 1:
            q = int(rng.integers(n))
            vec = rng.normal(size=3)
            vec /= np.linalg.norm(vec)
            pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
            po = orc.expectation_single(pm, q)
            assert abs(dev.expectation_single(pm, q) - po) < 1e-10
            if po.real > 1e-6:
                dev.project_qubit(pm, q)
                orc.project_qubit(pm, q)
        elif kind == "U" and n >= 1:
            q = int(rng.integers(n))
            m = mbqc.Clifford(int(rng.integers(24))).matrix
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        elif kind == "C1" and n >= 1:
            q = int(rng.integers(n))
            ch = noise.depolarising_channel(float(rng.uniform(0, 0.5)))
            dev.apply_channel(ch, [q])
            orc.apply_channel(ch, [q])
        elif kind == "C2" and n >= 2:
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            ch = noise.two_qubit_depolarising_channel(float(rng.uniform(0, 0.5)))
            dev.apply_channel(ch, [a, b])
            orc.apply_channel(ch, [a, b])
        if step % 20 == 19:
            assert dev.nqubit == orc.nqubit
            np.testing.assert_allclose(dev.rho, orc.rho, atol=1e-11, err_msg=f"step {step}")
    np.testing.assert_allclose(dev.rho, orc.rho, atol=1e-11)
    assert abs(dev.trace() - 1) < 1e-10


def test_dm_remove_qubit_and_ptrace_against_oracle(g):
    for n in (2, 3, 4):
        rho = g[f"n{n}_rho"]
        for q in range(n):
            dev, orc = B200DensityMatrix(rho), OracleDensityMatrix(rho)
            dev.remove_qubit(q)
            orc.remove_qubit(q)
            np.testing.assert_allclose(dev.rho, orc.rho, atol=1e-12)
    dev, orc = B200DensityMatrix(g["n4_rho"]), OracleDensityMatrix(g["n4_rho"])
    dev.ptrace([0, 2])
    orc.ptrace(2)
    orc.ptrace(0)
    orc.normalize()
    np.testing.assert_allclose(dev.rho, orc.rho, atol=1e-12)


def test_dm_string_backend_noise_contract_and_fidelity():
    d = load_golden("dm_rand3x2")
    from graphix_b200.pattern_io import pattern_from_dict

    pattern = pattern_from_dict(d)
    sim = PatternSimulator(pattern, "densitymatrix", branch_selector=mbqc.ConstBranchSelector(0))
    sim.run()                                        # noiseless density-matrix run == statevector run
    from graphix_b200 import B200StatevectorBackend

    sv = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=mbqc.ConstBranchSelector(0))
    PatternSimulator(pattern, sv).run()
    assert abs(sim.backend.state.fidelity(sv.state) - 1) < 1e-10
    with pytest.raises(ValueError):
        PatternSimulator(pattern, "statevector", noise_model=noise.DepolarisingNoiseModel())
    with pytest.raises(NotImplementedError):
        B200DensityMatrix(nqubit=2).project_qubit(np.eye(2), 0)     # not rank 1
    with pytest.raises(ValueError):
        B200DensityMatrix(np.eye(2))                                  # trace 2
    psi = rand_state(np.random.default_rng(2), 3)
    dm = B200DensityMatrix(psi)
    np.testing.assert_allclose(dm.rho, np.outer(psi, psi.conj()), atol=1e-15)
