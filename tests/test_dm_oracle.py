"""Density-matrix oracle (oracle/dm_oracle.py) and noise mirror against golden data of the reference
DensityMatrix / DensityMatrixBackend / DepolarisingNoiseModel."""
from __future__ import annotations

import os

import numpy as np
import pytest
from helpers import GOLDEN, check_dm_trace, load_golden

from graphix_b200 import mbqc, noise
from oracle.dm_oracle import OracleDensityMatrix, OracleDMBackend

TOL = 1e-13


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN, "dm_kernels.npz")))


@pytest.mark.parametrize("n", range(1, 5))
def test_dm_kernels_against_reference(g, n):
    rho, op = g[f"n{n}_rho"], g[f"n{n}_op"]
    dm = OracleDensityMatrix(rho)
    dm.add_nodes(1, mbqc.BasicStates.PLUS)
    np.testing.assert_allclose(dm.rho, g[f"n{n}_addplus"], atol=TOL)
    dm = OracleDensityMatrix(rho)
    dm.add_nodes(1, g[f"n{n}_s1"])
    np.testing.assert_allclose(dm.rho, g[f"n{n}_add1"], atol=TOL)
    for q in range(n):
        dm = OracleDensityMatrix(rho)
        dm.evolve_single(op, q)
        np.testing.assert_allclose(dm.rho, g[f"n{n}_evolve_q{q}"], atol=TOL)
        assert abs(OracleDensityMatrix(rho).expectation_single(op, q) - complex(g[f"n{n}_expect_q{q}"])) < TOL
        dm = OracleDensityMatrix(rho)
        dm.apply_channel(noise.depolarising_channel(0.13), [q])
        np.testing.assert_allclose(dm.rho, g[f"n{n}_depol_q{q}"], atol=TOL)
        for r in (0, 1):
            pm = g[f"n{n}_proj_q{q}_r{r}_op"]
            assert abs(OracleDensityMatrix(rho).expectation_single(pm, q) - complex(g[f"n{n}_proj_q{q}_r{r}_p"])) < TOL
            dm = OracleDensityMatrix(rho)
            dm.project_qubit(pm, q)
            np.testing.assert_allclose(dm.rho, g[f"n{n}_proj_q{q}_r{r}"], atol=1e-12)
        for q2 in range(n):
            if q2 == q:
                continue
            dm = OracleDensityMatrix(rho)
            dm.entangle((q, q2))
            np.testing.assert_allclose(dm.rho, g[f"n{n}_cz_{q}_{q2}"], atol=TOL)
            dm = OracleDensityMatrix(rho)
            dm.apply_channel(noise.two_qubit_depolarising_channel(0.21), [q, q2])
            np.testing.assert_allclose(dm.rho, g[f"n{n}_depol2_{q}_{q2}"], atol=TOL)
            dm = OracleDensityMatrix(rho)
            dm.evolve(g[f"n{n}_ev2_{q}_{q2}_op"], [q, q2])
            np.testing.assert_allclose(dm.rho, g[f"n{n}_ev2_{q}_{q2}"], atol=1e-12)
    dm = OracleDensityMatrix(rho)
    dm.permute([int(x) for x in g[f"n{n}_perm_p"]])
    np.testing.assert_allclose(dm.rho, g[f"n{n}_perm"], atol=TOL)


def test_channels_are_normalised():
    for ch in (noise.depolarising_channel(0.3), noise.two_qubit_depolarising_channel(0.4)):
        tot = sum(abs(k.coef) ** 2 * (k.operator.conj().T @ k.operator) for k in ch)
        np.testing.assert_allclose(tot, np.eye(tot.shape[0]), atol=1e-14)
    assert len(noise.two_qubit_depolarising_channel(0.1)) == 16
    with pytest.raises(ValueError):
        noise.KrausChannel([noise.KrausData(1.0, mbqc.Ops.X), noise.KrausData(1.0, mbqc.Ops.Z)])


@pytest.mark.parametrize("name", ["dm_rand3x2", "dm_handbuilt", "dm_rand5x3"])
def test_noisy_patterns_against_reference(name):
    d = load_golden(name)
    for trace in d["traces"]:
        check_dm_trace(d, trace, lambda cap, sel: OracleDMBackend(branch_selector=sel))
