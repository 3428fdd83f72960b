"""Kernel-level parity of the CUDA path (through the C ABI) against golden vectors of the reference
and against the oracle.  Mirrors the reference's tests/test_statevec.py:151-294, 449-480, 648-673."""
from __future__ import annotations

import itertools

import numpy as np
import pytest
from helpers import rand_state

from graphix_b200 import B200Statevector, mbqc
from oracle.oracle import OracleStatevector, outcome_to_operator_matrix

pytestmark = pytest.mark.gpu
TOL = 1e-13


def fid(a, b):
    return abs(np.vdot(a, b)) ** 2


def _sv(psi, cap=None, **kw):
    return B200Statevector(psi, max_qubits=cap, **kw)


@pytest.mark.parametrize("n", range(1, 7))
def test_kernels_against_reference_golden(kernels_golden, n):
    g = kernels_golden
    psi, op = g[f"n{n}_psi"], g[f"n{n}_op"]
    np.testing.assert_allclose(_sv(psi).flatten(), psi, atol=1e-15)
    sv = _sv(psi, n + 2)
    sv.add_nodes(1, mbqc.BasicStates.PLUS)
    np.testing.assert_allclose(sv.flatten(), g[f"n{n}_addplus"], atol=TOL)
    sv = _sv(psi)
    sv.add_nodes(1, g[f"n{n}_s1"])
    np.testing.assert_allclose(sv.flatten(), g[f"n{n}_add1"], atol=TOL)
    sv = _sv(psi)
    sv.add_nodes(2, g[f"n{n}_s2"])
    np.testing.assert_allclose(sv.flatten(), g[f"n{n}_add2"], atol=TOL)
    for q in range(n):
        sv = _sv(psi)
        sv.evolve_single(op, q)
        np.testing.assert_allclose(sv.flatten(), g[f"n{n}_evolve_q{q}"], atol=TOL)
        assert abs(_sv(psi).expectation_single(op, q) - complex(g[f"n{n}_expect_q{q}"])) < TOL
        for r in (0, 1):
            pm = g[f"n{n}_proj_q{q}_r{r}_op"]
            sv = _sv(psi)
            assert abs(sv.expectation_single(pm, q) - complex(g[f"n{n}_proj_q{q}_r{r}_p"])) < TOL
            sv.project_qubit(pm, q)
            assert sv.nqubit == n - 1
            # same vector as the reference, phase included
            np.testing.assert_allclose(sv.flatten(), g[f"n{n}_proj_q{q}_r{r}"], atol=1e-12)
        if f"n{n}_projgen_q{q}" in g:
            sv = _sv(psi)
            sv.project_qubit(op, q)  # non-projector: literal two-pass reference algorithm
            np.testing.assert_allclose(sv.flatten(), g[f"n{n}_projgen_q{q}"], atol=1e-12)
        for q2 in range(n):
            if q2 == q:
                continue
            sv = _sv(psi)
            sv.entangle((q, q2))
            np.testing.assert_array_equal(sv.flatten(), g[f"n{n}_cz_{q}_{q2}"])
            sv = _sv(psi)
            sv.swap((q, q2))
            np.testing.assert_array_equal(sv.flatten(), g[f"n{n}_swap_{q}_{q2}"])
    for t in range(3):
        sv = _sv(psi)
        sv.permute([int(x) for x in g[f"n{n}_perm{t}_p"]])
        np.testing.assert_array_equal(sv.flatten(), g[f"n{n}_perm{t}"])


@pytest.mark.parametrize("fusion", [1, 2])
@pytest.mark.parametrize("n", range(2, 7))
def test_multi_qubit_evolve_against_reference_golden(kernels_golden, n, fusion):
    """Statevector.evolve / expectation_value for 2-, 3- and 4-qubit operators on arbitrary qubit tuples."""
    g = kernels_golden
    for k in (2, 3, 4):
        for t in range(3):
            key = f"n{n}_ev{k}_{t}"
            if key not in g:
                continue
            qs = [int(x) for x in g[key + "_q"]]
            sv = _sv(g[f"n{n}_psi"], fusion=fusion)
            assert abs(sv.expectation_value(g[key + "_op"], qs) - complex(g[key + "_val"])) < 1e-12
            sv.evolve(g[key + "_op"], qs)
            np.testing.assert_allclose(sv.flatten(), g[key], atol=1e-12)
    with pytest.raises(ValueError):
        _sv(g[f"n{n}_psi"]).evolve(np.eye(4), [0, 0])


def test_multi_qubit_evolve_large_against_oracle():
    rng = np.random.default_rng(77)
    n = 16
    psi = rand_state(rng, n)
    dev, orc = _sv(psi), OracleStatevector(psi)
    for k, qs in ((2, [3, 12]), (3, [15, 0, 7]), (4, [9, 2, 14, 5]), (2, [0, 1])):
        opk = rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k))
        dev.evolve(opk, qs)
        orc.evolve(opk, qs)
        nrm = np.linalg.norm(orc.flatten())
        np.testing.assert_allclose(dev.flatten() / nrm, orc.flatten() / nrm, atol=1e-12)


def test_kernels_17_qubits_golden(kernels_golden):
    g = kernels_golden
    n = 17
    psi = rand_state(np.random.default_rng(17), n)
    w = rand_state(np.random.default_rng(1717), n + 1)
    op = g["n17_op"]
    for q in (0, 5, 16):
        sv = _sv(psi)
        sv.evolve_single(op, q)
        assert abs(np.vdot(w[: 1 << n], sv.flatten()) - complex(g[f"n17_evolve_q{q}_dig"])) < 1e-12
        assert abs(_sv(psi).expectation_single(op, q) - complex(g[f"n17_expect_q{q}"])) < 1e-11
        sv = _sv(psi)
        sv.project_qubit(g[f"n17_proj_q{q}_op"], q)
        assert abs(np.vdot(w[: 1 << (n - 1)], sv.flatten()) - complex(g[f"n17_proj_q{q}_dig"])) < 1e-12
    sv = _sv(psi)
    sv.entangle((3, 11))
    assert abs(np.vdot(w[: 1 << n], sv.flatten()) - complex(g["n17_cz_3_11_dig"])) < 1e-12
    sv = _sv(psi, n + 1)
    sv.add_nodes(1, mbqc.BasicStates.PLUS)
    assert abs(np.vdot(w, sv.flatten()) - complex(g["n17_addplus_dig"])) < 1e-12


@pytest.mark.parametrize("fusion", [1, 2])
@pytest.mark.parametrize("n", [9, 12, 14, 18])
def test_every_axis_against_oracle(n, fusion):
    """All kernels on every axis at sizes that cross the tile (11 bits) and hole (bit 8) boundaries."""
    rng = np.random.default_rng(100 + n)
    psi = rand_state(rng, n)
    op = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    for q in range(n):
        vec = rng.normal(size=3)
        vec /= np.linalg.norm(vec)
        pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
        dev, orc = _sv(psi, fusion=fusion), OracleStatevector(psi)
        assert abs(dev.expectation_single(pm, q) - orc.expectation_single(pm, q)) < 1e-12
        dev.project_qubit(pm, q)
        orc.project_qubit(pm, q)
        np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-12)
        # keep going on the shrunk state: layout now has a hole or a shifted tile
        q2 = int(rng.integers(n - 1))
        dev.evolve_single(op, q2)
        orc.evolve_single(op, q2)
        a, b = (int(x) for x in rng.choice(n - 1, size=2, replace=False))
        dev.entangle((a, b))
        orc.entangle((a, b))
        dev.add_nodes(1, mbqc.BasicStates.PLUS)
        orc.add_nodes(1, mbqc.BasicStates.PLUS)
        dev.entangle((n - 1, a))
        orc.entangle((n - 1, a))
        np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-12)  # op is not unitary: both unnormalised alike


@pytest.mark.parametrize("fusion,seed", [(0, 2026), (1, 2026), (2, 2026), (2, 7), (2, 8), (2, 9), (-3, 7), (-8, 8), (-8, 9),
                                         (-8, 10), (-8, 11)])
def test_random_command_soup_against_oracle(fusion, seed):
    """Long random sequence of N / E-runs / M / 1q ops: exercises hole creation, hole filling,
    tile shifts, extent trimming, the pending-scale bookkeeping and (fusion = 2) virtual qubits,
    queued CZs absorbed by projections and the fused N,E,M butterfly together."""
    rng = np.random.default_rng(seed)
    n0 = 13
    psi = rand_state(rng, n0)
    if fusion < 0:      # negative: lazy, non-strict, batches of -fusion projections
        dev = _sv(psi, 16, fusion=2, strict=False, batch=-fusion)
    else:
        dev = _sv(psi, 16, fusion=fusion)
    orc = OracleStatevector(psi, max_qubits=16)
    for step in range(160):
        n = orc.nqubit
        kind = rng.choice(["N", "E", "M", "U", "Z", "P", "K"], p=[0.25, 0.28, 0.25, 0.08, 0.08, 0.03, 0.03])
        if kind == "N" and n < 16:
            st = [mbqc.BasicStates.PLUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE, mbqc.BasicStates.MINUS,
                  mbqc.PlanarState(mbqc.Plane.YZ, 0.37)][int(rng.integers(5))]
            dev.add_nodes(1, st)
            orc.add_nodes(1, st)
        elif kind == "E" and n >= 2:
            for _ in range(int(rng.integers(1, 5))):
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                dev.entangle((a, b))
                orc.entangle((a, b))
        elif kind == "M" and n > 6:
            q = int(rng.integers(n))
            vec = rng.normal(size=3)
            vec /= np.linalg.norm(vec)
            pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
            if rng.integers(2) and (fusion >= 0 or rng.integers(3) == 0):
                pd, po = dev.expectation_single(pm, q), orc.expectation_single(pm, q)
                assert abs(pd - po) < 1e-10
            dev.project_qubit(pm, q)
            orc.project_qubit(pm, q)
        elif kind == "U":
            q = int(rng.integers(n))
            m = mbqc.Clifford(int(rng.integers(24))).matrix
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        elif kind == "Z":
            q = int(rng.integers(n))
            m = [mbqc.Ops.Z, mbqc.Ops.X, mbqc.Ops.S][int(rng.integers(3))]
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        elif kind == "P":
            # relabels: permute / swap with queued work and virtual qubits in flight
            if rng.integers(2):
                perm = [int(x) for x in rng.permutation(n)]
                dev.permute(perm)
                orc.permute(perm)
            else:
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                dev.swap((a, b))
                orc.swap((a, b))
        elif kind == "K":
            # 2-qubit unitary on arbitrary (possibly still virtual) qubits
            qs = [int(x) for x in rng.choice(n, size=2, replace=False)]
            u, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
            dev.evolve(u, qs)
            orc.evolve(u, qs)
        if step % 20 == 19:
            assert dev.nqubit == orc.nqubit
            np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    assert abs(dev.norm2() - 1) < 1e-12


def test_lazy_fusion_triples_and_virtual_qubits():
    """fusion = 2: N and E are queued; M of a qubit with a virtual partner runs the fused butterfly
    (the new qubit takes the measured qubit's physical bit, no hole, state never grows)."""
    rng = np.random.default_rng(5)
    n = 12
    psi = rand_state(rng, n)
    dev, orc = _sv(psi, n, fusion=2), OracleStatevector(psi, max_qubits=n + 3)
    for step in range(40):
        m = int(rng.integers(n))
        others = [int(x) for x in rng.choice([i for i in range(n) if i != m], size=2, replace=False)]
        st = [mbqc.BasicStates.PLUS, mbqc.PlanarState(mbqc.Plane.XY, 0.3), mbqc.BasicStates.ZERO][step % 3]
        for s in (dev, orc):
            s.add_nodes(1, st)                   # new qubit, logical index n
            s.entangle((n, m))
            s.entangle((m, others[0]))           # real-real edge on the measured qubit
            if step % 2:
                s.entangle((n, others[1]))       # real partner of the new qubit
            if step % 5 == 0:
                s.add_nodes(1, mbqc.BasicStates.PLUS)   # second virtual partner of m
                s.entangle((n + 1, m))
                s.entangle((n + 1, n))                  # virtual-virtual edge
        assert dev.max_qubits <= n + 1           # virtual qubits occupy no memory
        vec = rng.normal(size=3)
        vec /= np.linalg.norm(vec)
        pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
        dev.project_qubit(pm, m)
        orc.project_qubit(pm, m)
        if step % 5 == 0:
            # drop the extra qubit again so the width stays n
            pm2 = outcome_to_operator_matrix((0.6, 0.0, 0.8), 0)
            dev.project_qubit(pm2, n)
            orc.project_qubit(pm2, n)
        assert dev.nqubit == orc.nqubit == n
        if step % 8 == 7:
            np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    st = dev.stats()
    # CZs are only ever issued by the read-outs above (flatten flushes the queue), never per command
    assert st["fused"]["launches"] >= 40 and st["cz"]["launches"] <= 6


@pytest.mark.parametrize("fusion", [0, 1, 2])
def test_small_register_soup_down_to_zero_qubits(fusion):
    """Registers of 0-4 qubits: gates on freshly prepared (still virtual) qubits, measuring the last qubit
    (scalar state, phase included).  Regression: a virtual qubit whose amplitude s0 is complex (Y|+>) used
    to lose that phase when it was measured before being materialised."""
    st = mbqc.PlanarState(mbqc.Plane.YZ, 0.37)
    rng = np.random.default_rng(0)
    for trial in range(24):
        vec = rng.normal(size=3)
        vec /= np.linalg.norm(vec)
        pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
        first = [mbqc.BasicStates.PLUS, st, mbqc.BasicStates.MINUS][trial % 3]
        res = []
        for s in (OracleStatevector(nqubit=0, max_qubits=4), B200Statevector(nqubit=0, max_qubits=4, fusion=fusion)):
            s.add_nodes(1, first)
            if trial & 1:
                s.evolve_single([mbqc.Ops.Y, mbqc.Clifford(17).matrix][(trial >> 3) & 1], 0)
            if trial & 2:
                s.add_nodes(1, st)
                if trial & 4:
                    s.entangle((0, 1))
                    s.evolve_single(mbqc.Ops.S, 1)
                s.project_qubit(outcome_to_operator_matrix((0.6, 0, 0.8), 0), 1)
            s.project_qubit(pm, 0)
            assert s.nqubit == 0
            res.append(complex(s.flatten()[0]))
        assert abs(res[0] - res[1]) < 1e-12, (trial, res)
    rng = np.random.default_rng(501)
    dev, orc = B200Statevector(nqubit=0, max_qubits=6, fusion=fusion), OracleStatevector(nqubit=0, max_qubits=6)
    states = [mbqc.BasicStates.PLUS, mbqc.BasicStates.MINUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE, st]
    for step in range(400):
        n = orc.nqubit
        kind = rng.choice(["N", "E", "M", "U", "D"], p=[0.3, 0.3, 0.24, 0.08, 0.08])
        if kind == "N" and n < 6:
            s_ = states[int(rng.integers(len(states)))]
            dev.add_nodes(1, s_)
            orc.add_nodes(1, s_)
        elif kind == "E" and n >= 2:
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            dev.entangle((a, b))
            orc.entangle((a, b))
        elif kind == "M" and n >= 1:
            q = int(rng.integers(n))
            vec = rng.normal(size=3)
            vec /= np.linalg.norm(vec)
            pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
            if orc.expectation_single(pm, q).real > 1e-6:
                dev.project_qubit(pm, q)
                orc.project_qubit(pm, q)
        elif kind in ("U", "D") and n >= 1:
            q = int(rng.integers(n))
            m = mbqc.Clifford(int(rng.integers(24))).matrix if kind == "U" else \
                [mbqc.Ops.Z, mbqc.Ops.X, mbqc.Ops.S, mbqc.Ops.Y][int(rng.integers(4))]
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        if step % 25 == 24:
            np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11, err_msg=f"step {step}")


def test_measure_all_then_regrow():
    """Shrink to 0 qubits (scalar state) and grow again; capacity growth without with_capacity."""
    sv = B200Statevector(nqubit=3, data=mbqc.BasicStates.PLUS)
    proj0 = outcome_to_operator_matrix((0, 0, 1), 0)
    for _ in range(3):
        sv.project_qubit(proj0, 0)
    assert sv.nqubit == 0
    np.testing.assert_allclose(np.abs(sv.flatten()), [1.0], atol=1e-14)
    for _ in range(5):
        sv.add_nodes(1, mbqc.BasicStates.PLUS)
    np.testing.assert_allclose(np.abs(sv.flatten()), np.full(32, 1 / np.sqrt(32)), atol=1e-14)
    assert sv.max_qubits >= 5          # lazily prepared qubits take memory only once they are materialised


def test_zero_norm_projection_raises():
    """tests/test_statevec.py:269-294 of the reference."""
    sv = B200Statevector([mbqc.BasicStates.ZERO, mbqc.BasicStates.PLUS])
    with pytest.raises(RuntimeError):
        sv.project_qubit(outcome_to_operator_matrix((0, 0, 1), 1), 0)
    sv = B200Statevector([mbqc.BasicStates.ONE, mbqc.BasicStates.PLUS])
    sv.project_qubit(outcome_to_operator_matrix((0, 0, 1), 1), 0)
    np.testing.assert_allclose(sv.flatten(), np.array([1, 1]) / np.sqrt(2), atol=1e-15)
    # XY(1e-5) measured to outcome 1 has p ~ 2.5e-10, just above the cut-off (tests/test_branch_selector.py:152-168)
    sv = B200Statevector([mbqc.BasicStates.PLUS, mbqc.BasicStates.PLUS])
    v = mbqc.Plane.XY.polar(1e-5)
    sv.project_qubit(outcome_to_operator_matrix(v, 1), 0)
    assert abs(sv.norm2() - 1) < 1e-9


def test_errors():
    sv = B200Statevector(nqubit=2, data=mbqc.BasicStates.PLUS)
    with pytest.raises(IndexError):
        sv.evolve_single(np.eye(2), 2)
    with pytest.raises(IndexError):
        sv.entangle((0, 2))
    with pytest.raises(IndexError):
        sv.project_qubit(np.eye(2) / 2, -1)
    with pytest.raises(ValueError):
        sv.permute([0, 0])
    with pytest.raises(ValueError):
        sv.permute([0])
    with pytest.raises(ValueError):
        B200Statevector([1.0, 1.0])  # not normalised (statevec.py:196-197)
    with pytest.raises(ValueError):
        B200Statevector([1.0, 0.0, 0.0])  # not a power of two
    with pytest.raises(NotImplementedError):
        sv.remove_qubit(0)
    with pytest.raises(TypeError):
        sv.evolve_single(np.array([[1, 0], [0, object()]], dtype=object), 0)


def test_permute_all_three_qubit_permutations():
    """tests/test_statevec.py:648-673 of the reference: permute == transpose, bit exact."""
    psi = rand_state(np.random.default_rng(9), 3)
    for perm in itertools.permutations(range(3)):
        sv = _sv(psi)
        sv.permute(list(perm))
        expect = np.transpose(psi.reshape(2, 2, 2), perm).reshape(8)
        np.testing.assert_array_equal(sv.flatten(), expect)


def test_all_cliffords_and_basic_states():
    """tests/test_statevec.py:539-554 of the reference."""
    for st in mbqc.BasicStates.VEC:
        for k in range(24):
            sv = B200Statevector(st)
            sv.evolve_single(mbqc.Clifford(k).matrix, 0)
            np.testing.assert_allclose(sv.flatten(), mbqc.Clifford(k).matrix @ st.to_statevector_numpy(), atol=1e-15)


def test_to_dict_encoding():
    """tests/test_statevec.py:423-447 of the reference."""
    sv = B200Statevector([mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE])
    assert list(sv.to_dict()) == ["01"]
    assert list(sv.to_dict(encoding="LSB")) == ["10"]
    assert abs(sv.to_prob_dict()["01"] - 1) < 1e-15


def test_fidelity_on_device():
    psi = rand_state(np.random.default_rng(4), 10)
    phi = rand_state(np.random.default_rng(5), 10)
    a, b = _sv(psi), _sv(phi)
    assert abs(a.fidelity(b) - fid(psi, phi)) < 1e-13
    assert a.isclose(_sv(psi * np.exp(0.3j)))


def test_large_state_roundtrip_properties():
    """Size-independent properties at a size the oracle would take long for (24 -> 25 qubits):
    N then M of the same qubit is the identity; CZ twice is the identity; norms stay 1."""
    n = 24
    sv = B200Statevector(nqubit=n, data=mbqc.BasicStates.PLUS, max_qubits=n + 1)
    rng = np.random.default_rng(0)
    for _ in range(6):
        a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
        sv.entangle((a, b))
        sv.evolve_single(mbqc.Clifford(int(rng.integers(24))).matrix, int(rng.integers(n)))
    ref = B200Statevector(sv)  # device copy through flatten
    sv.add_nodes(1, mbqc.BasicStates.PLUS)
    sv.entangle((n, 3))
    sv.entangle((3, n))
    sv.project_qubit(outcome_to_operator_matrix((1, 0, 0), 0), n)
    assert sv.nqubit == n
    assert abs(sv.norm2() - 1) < 1e-12
    assert abs(sv.fidelity(ref) - 1) < 1e-12


@pytest.mark.parametrize("tma", [0, 1, 4])
@pytest.mark.parametrize("axes,glog", [(1, 0), (2, 1), (3, 0), (3, 3), (4, 0), (4, 2), (5, 0), (5, 3), (6, 0), (6, 1)])
@pytest.mark.parametrize("n", [12, 15, 17])
def test_batched_passes_on_many_axes_against_oracle(n, axes, glog, tma):
    """Non-strict batches of up to 8 fused (N, E.., M) stages on up to `axes` distinct physical bits: more than three
    axes run the phased shared-memory tile kernel (k_fused_tile).  Random measured qubits (so the axis sets, the phase
    plans and the low / high bit positions vary), CZs to neighbours of both the measured and the new qubit, non-|+>
    preparations; compared with the oracle entry by entry (phase included)."""
    from graphix_b200 import _lib

    rng = np.random.default_rng(1000 * n + 10 * axes + glog)
    psi = rand_state(rng, n)
    dev = _sv(psi, n + 1, fusion=2, strict=False, batch=8)
    dev.set_option(_lib.OPT_BATCH_AXES, axes)
    dev.set_option(_lib.OPT_CHUNK_LOG2, glog)
    dev.set_option(_lib.OPT_TMA, 1 if tma == 1 else 0)     # 1: tiles staged by bulk copies (k_fused_tma); else register-staged
    dev.set_option(_lib.OPT_TILE_REGS, 4 if tma == 4 else 3)   # 4: 2^4 amplitudes per thread in k_fused_tile
    orc = OracleStatevector(psi, max_qubits=n + 1)
    npass0 = dev.stats()["fused"]["launches"]
    for step in range(48):
        m = int(rng.integers(n))
        others = [int(x) for x in rng.choice([i for i in range(n) if i != m], size=3, replace=False)]
        st = [mbqc.BasicStates.PLUS, mbqc.PlanarState(mbqc.Plane.XY, 0.3), mbqc.BasicStates.MINUS,
              mbqc.PlanarState(mbqc.Plane.YZ, 0.7)][int(rng.integers(4))]
        vec = rng.normal(size=3)
        vec /= np.linalg.norm(vec)
        pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
        for s in (dev, orc):
            s.add_nodes(1, st)
            s.entangle((n, m))
            s.entangle((m, others[0]))
            if step % 2:
                s.entangle((n, others[1]))
            if step % 3 == 0:
                s.entangle((m, others[2]))
            s.project_qubit(pm, m)
        if step % 16 == 15:
            np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    assert abs(dev.norm2() - 1) < 1e-12
    passes = dev.stats()["fused"]["launches"] - npass0
    if axes >= 5:
        assert passes <= 16          # 48 stages: at least three per pass on average


@pytest.mark.parametrize("defer,io,tw", [(1, 1, 0), (0, 0, 4), (1, 2, 8), (1, 2, 1), (1, 2, 4), (1, 0, 1)])
@pytest.mark.parametrize("axes,batch,glog", [(7, 12, 0), (7, 12, 2), (6, 12, 1), (4, 12, 0), (3, 12, 0)])
@pytest.mark.parametrize("n", [14, 17, 20])
def test_twelve_stage_passes_deferred_partner_cz_and_layouts(n, axes, batch, glog, defer, io, tw):
    """The bench-default pass shape (up to 12 stages on up to 7 axes, 2^4 amplitudes per thread) against the oracle, with
    the round-2 kernel variants switched both ways: partner CZs left queued (GXSV_OPT_DEFER_VCZ, the default) or applied
    as signs inside the stage, extra load / store layouts never / when an axis is low / always (io), tiles private to a
    warp or shared by 4 / 8 warps (tw; 0 = chosen by the number of axes).  Mostly |+> partners and XY-plane measurements, as in
    transpiled patterns (the 6-FMA butterfly with per-pair and thread-uniform signs), mixed with |-> / planar partners and
    arbitrary measurement axes (the general stages); CZs hang on both the measured and the new qubit, and some
    measured qubits are the partners of earlier stages of the same pass (their deferred CZs come back as that stage's
    own signs)."""
    from graphix_b200 import _lib

    rng = np.random.default_rng(7000 * n + 100 * axes + 10 * glog + 2 * defer + io + 1000000 * tw)
    psi = rand_state(rng, n)
    dev = _sv(psi, n + 1, fusion=2, strict=False, batch=batch)
    dev.set_option(_lib.OPT_BATCH_AXES, axes)
    dev.set_option(_lib.OPT_CHUNK_LOG2, glog)
    dev.set_option(_lib.OPT_TILE_REGS, 4)
    dev.set_option(_lib.OPT_DEFER_VCZ, defer)
    dev.set_option(_lib.OPT_TILE_IO, io)
    dev.set_option(_lib.OPT_TILE_WARPS, tw)
    orc = OracleStatevector(psi, max_qubits=n + 1)
    npass0 = dev.stats()["fused"]["launches"]
    last_new = None
    for step in range(60):
        # every third stage measures the qubit the previous stage created (same physical bit: axis reuse)
        if last_new is not None and step % 3 == 1:
            m = last_new
        else:
            m = int(rng.integers(n))
        others = [int(x) for x in rng.choice([i for i in range(n) if i != m], size=3, replace=False)]
        kind = int(rng.integers(8))
        st = [mbqc.BasicStates.PLUS] * 5 + [mbqc.BasicStates.MINUS, mbqc.PlanarState(mbqc.Plane.XY, 0.3),
                                            mbqc.PlanarState(mbqc.Plane.YZ, 0.7)]
        st = st[kind]
        if step % 4 == 3:
            vec = rng.normal(size=3)
            vec /= np.linalg.norm(vec)
        else:
            ang = float(rng.uniform(0, 2 * np.pi))
            vec = np.array([np.cos(ang), np.sin(ang), 0.0])
        pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
        for s in (dev, orc):
            s.add_nodes(1, st)
            s.entangle((n, m))
            if step % 2 == 0:
                s.entangle((m, others[0]))
            if step % 3 != 2:
                s.entangle((n, others[1]))
            if step % 5 == 0:
                s.entangle((n, others[2]))
            s.project_qubit(pm, m)
        # the new qubit sits at logical index n - 1 after the measured one was removed
        last_new = n - 1
        if step % 20 == 19:
            np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    np.testing.assert_allclose(dev.flatten(), orc.flatten(), atol=1e-11)
    assert abs(dev.norm2() - 1) < 1e-12
    passes = dev.stats()["fused"]["launches"] - npass0
    if axes >= 6:
        assert passes <= 20          # 60 stages: at least three per pass on average


def test_read_outs_on_the_device():
    """Sparse to_dict / to_prob_dict (statevec.py:591-678) compacted on the device, the in-place inner product behind
    fidelity / expectation_value with two DIFFERENT physical layouts, entangle(q, q), and the clone."""
    rng = np.random.default_rng(77)
    n = 18
    # a sparse state: GHZ-like superposition of 5 basis states with random amplitudes, scrambled by relabels
    idx = np.sort(rng.choice(1 << n, size=5, replace=False))
    amps = rng.normal(size=5) + 1j * rng.normal(size=5)
    amps /= np.linalg.norm(amps)
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[idx] = amps
    sv, orc = _sv(psi), OracleStatevector(psi)
    perm = [int(x) for x in rng.permutation(n)]
    sv.permute(perm)
    orc.permute(perm)
    sv.swap((0, 5))
    orc.swap((0, 5))
    flat = orc.flatten()
    keep = np.flatnonzero(np.abs(flat) > 1e-8)
    d = sv.to_dict()
    assert list(d) == [format(int(i), f"0{n}b") for i in keep]
    np.testing.assert_allclose(np.array(list(d.values())), flat[keep], atol=1e-14)
    np.testing.assert_allclose(np.array(list(sv.to_prob_dict(encoding="LSB").values())), np.abs(flat[keep]) ** 2, atol=1e-14)
    assert list(sv.to_prob_dict(encoding="LSB")) == [format(int(i), f"0{n}b")[::-1] for i in keep]
    # dense state: more entries than the sparse buffer holds -> falls back to the full read-out, same answer
    dense = rand_state(rng, 12)
    small = _sv(dense)
    small.SPARSE_CAP = 64
    assert len(small.to_dict(atol=0.0)) == 1 << 12
    # inner product of two registers with different layouts (one went through projections / relabels)
    a = rand_state(rng, 14)
    sa, sb = _sv(a, 15), _sv(a, 15)
    pm = outcome_to_operator_matrix((0.6, 0.0, 0.8), 0)
    for s in (sa, sb):
        s.add_nodes(1, mbqc.BasicStates.PLUS)
        s.entangle((14, 2))
    sa.project_qubit(pm, 2)                 # layouts now differ: sa has a hole / a moved qubit
    sb.permute([0, 1, 14, *range(3, 14), 2])
    sb.project_qubit(pm, 14)
    sb.permute([0, 1, 13, *range(2, 13)])    # bring the new qubit of sb where sa has it? (only the state matters)
    oa = OracleStatevector(a, max_qubits=15)
    oa.add_nodes(1, mbqc.BasicStates.PLUS)
    oa.entangle((14, 2))
    oa.project_qubit(pm, 2)
    np.testing.assert_allclose(sa.flatten(), oa.flatten(), atol=1e-12)
    assert abs(sa.fidelity(sb) - abs(np.vdot(sa.flatten(), sb.flatten())) ** 2) < 1e-12
    op = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    assert abs(sa.expectation_value(op, [3, 9]) - oa.expectation_value(op, [3, 9])) < 1e-11
    np.testing.assert_allclose(sa.flatten(), oa.flatten(), atol=1e-12)       # expectation_value left the state alone
    # entangle(q, q) is Z on q in the reference kernel (statevec.py:930-937)
    sa.entangle((4, 4))
    oa.evolve_single(mbqc.Ops.Z, 4)
    np.testing.assert_allclose(sa.flatten(), oa.flatten(), atol=1e-12)


def test_nonstrict_zero_norm_is_a_runtime_error_not_nan():
    """An impossible branch in non-strict mode: the stored state stays finite and the FIRST error the caller sees is the
    reference's RuntimeError (statevec.py:1192-1193), whatever is asked next."""
    for batch in (0, 4):
        sv = B200Statevector(nqubit=0, max_qubits=14, strict=False, batch=batch)
        sv.add_nodes(12, mbqc.BasicStates.ZERO)
        sv.add_nodes(1, mbqc.BasicStates.PLUS)
        sv.entangle((0, 12))
        sv.project_qubit(outcome_to_operator_matrix((0.0, 0.0, 1.0), 1), 0)      # |0> measured as 1: probability 0
        sv.add_nodes(1, mbqc.BasicStates.PLUS)
        sv.entangle((1, 12))
        sv.project_qubit(outcome_to_operator_matrix((1.0, 0.0, 0.0), 0), 1)
        p = sv.expectation_single(outcome_to_operator_matrix((1.0, 0.0, 0.0), 0), 3)
        assert np.isfinite(p.real) and abs(p.imag) <= 1e-10          # no NaN reaches f_expectation0's assertion
        with pytest.raises(RuntimeError):
            sv.sync()
