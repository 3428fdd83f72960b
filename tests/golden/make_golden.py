"""Generate the golden fixtures of tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py --no-bench # skip the large command-only bench patterns

Outputs (all committed):
  kernels.npz           known-answer vectors for every numba kernel of graphix/sim/statevec.py:841-1198,
                        produced by the reference ``Statevector`` methods on seeded random inputs
  cliffords.npz         the 24 Clifford matrices (graphix/_db.py:15-66) and Pauli ops
  measure_adapt.json    Bloch vectors of signal-adapted measurements (graphix/simulator.py:205-228)
  patterns/*.json.gz    command streams (+ per-measurement p0, outcomes and the final state from the reference
                        ``StatevectorBackend`` for the parity cases; commands only for the bench-size cases)
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from ref_shim import import_reference  # noqa: E402

import_reference()

import networkx as nx  # noqa: E402
from graphix import Circuit  # noqa: E402
from graphix.branch_selector import BranchSelector, ConstBranchSelector, RandomBranchSelector  # noqa: E402
from graphix.clifford import Clifford  # noqa: E402
from graphix.command import C as CmdC  # noqa: E402
from graphix.command import E, M, N, X, Z  # noqa: E402
from graphix.fundamentals import Plane  # noqa: E402
from graphix.measurements import Measurement  # noqa: E402
from graphix.ops import Ops  # noqa: E402
from graphix.pattern import Pattern  # noqa: E402
from graphix.random_objects import rand_circuit  # noqa: E402
from graphix.sim.base_backend import _outcome_to_operator_matrix  # noqa: E402
from graphix.sim.statevec import Statevector, StatevectorBackend  # noqa: E402
from graphix.simulator import PatternSimulator  # noqa: E402
from graphix.states import BasicStates, PlanarState  # noqa: E402

from graphix_b200.pattern_io import pattern_to_dict, save_pattern  # noqa: E402

PAT_DIR = os.path.join(HERE, "patterns")


# ---------------------------------------------------------------------------
# kernel known-answer vectors
# ---------------------------------------------------------------------------
def rand_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def rand_op(rng):
    return rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))


def rand_bloch(rng):
    v = rng.normal(size=3)
    return tuple(v / np.linalg.norm(v))


def gen_kernels() -> None:
    rng = np.random.default_rng(20261019)
    out: dict[str, np.ndarray] = {}
    for n in range(1, 7):
        psi = rand_state(rng, n)
        out[f"n{n}_psi"] = psi
        op = rand_op(rng)
        out[f"n{n}_op"] = op
        s1 = rand_state(rng, 1)
        out[f"n{n}_s1"] = s1
        s2 = rand_state(rng, 2)
        out[f"n{n}_s2"] = s2
        # N
        sv = Statevector(psi, max_qubits=n + 2)
        sv.add_nodes(1, BasicStates.PLUS)
        out[f"n{n}_addplus"] = sv.flatten().copy()
        sv = Statevector(psi)
        sv.add_nodes(1, s1)
        out[f"n{n}_add1"] = sv.flatten().copy()
        sv = Statevector(psi)
        sv.add_nodes(2, s2)
        out[f"n{n}_add2"] = sv.flatten().copy()
        for q in range(n):
            sv = Statevector(psi)
            sv.evolve_single(op, q)
            out[f"n{n}_evolve_q{q}"] = sv.flatten().copy()
            sv = Statevector(psi)
            out[f"n{n}_expect_q{q}"] = np.array(sv.expectation_single(op, q))
            for r in (0, 1):
                vec = rand_bloch(rng)
                out[f"n{n}_proj_q{q}_r{r}_vec"] = np.array(vec)
                pm = _outcome_to_operator_matrix(vec, r)
                out[f"n{n}_proj_q{q}_r{r}_op"] = pm
                sv = Statevector(psi)
                out[f"n{n}_proj_q{q}_r{r}_p"] = np.array(sv.expectation_single(pm, q))
                sv.project_qubit(pm, q)
                out[f"n{n}_proj_q{q}_r{r}"] = sv.flatten().copy()
            # general (non-projector) op through project_qubit
            sv = Statevector(psi)
            try:
                sv.project_qubit(op, q)
                out[f"n{n}_projgen_q{q}"] = sv.flatten().copy()
            except RuntimeError:
                pass
            for q2 in range(n):
                if q2 == q:
                    continue
                sv = Statevector(psi)
                sv.entangle((q, q2))
                out[f"n{n}_cz_{q}_{q2}"] = sv.flatten().copy()
                sv = Statevector(psi)
                sv.swap((q, q2))
                out[f"n{n}_swap_{q}_{q2}"] = sv.flatten().copy()
        for t in range(3):
            perm = [int(x) for x in rng.permutation(n)]
            sv = Statevector(psi)
            sv.permute(perm)
            out[f"n{n}_perm{t}_p"] = np.array(perm)
            out[f"n{n}_perm{t}"] = sv.flatten().copy()
    # projection on Pauli eigenstates incl. the zero-norm edge (tests/test_statevec.py:269-294)
    # 17 qubits: exercises the prange code path (NUM_QUBIT_PARALLEL + 2); only digests are stored
    n = 17
    psi = rand_state(np.random.default_rng(17), n)  # regenerated from the seed by the tests
    w = rand_state(np.random.default_rng(1717), n + 1)
    op = rand_op(rng)
    out["n17_op"] = op
    for q in (0, 5, 16):
        sv = Statevector(psi)
        sv.evolve_single(op, q)
        out[f"n17_evolve_q{q}_dig"] = np.array(np.vdot(w[: 1 << n], sv.flatten()))
        sv = Statevector(psi)
        out[f"n17_expect_q{q}"] = np.array(sv.expectation_single(op, q))
        vec = rand_bloch(rng)
        pm = _outcome_to_operator_matrix(vec, 1)
        out[f"n17_proj_q{q}_op"] = pm
        sv = Statevector(psi)
        sv.project_qubit(pm, q)
        out[f"n17_proj_q{q}_dig"] = np.array(np.vdot(w[: 1 << (n - 1)], sv.flatten()))
    sv = Statevector(psi)
    sv.entangle((3, 11))
    out["n17_cz_3_11_dig"] = np.array(np.vdot(w[: 1 << n], sv.flatten()))
    sv = Statevector(psi, max_qubits=n + 1)
    sv.add_nodes(1, BasicStates.PLUS)
    out["n17_addplus_dig"] = np.array(np.vdot(w, sv.flatten()))
    # k-qubit operators through Statevector.evolve / expectation_value (statevec.py:361-415); separate rng so that
    # the arrays above keep their values
    rng2 = np.random.default_rng(361395)
    for n in range(2, 7):
        psi = out[f"n{n}_psi"]
        for k in (2, 3, 4):
            if k > n:
                continue
            for t in range(3):
                qs = [int(x) for x in rng2.permutation(n)[:k]]
                opk = rng2.normal(size=(1 << k, 1 << k)) + 1j * rng2.normal(size=(1 << k, 1 << k))
                sv = Statevector(psi)
                out[f"n{n}_ev{k}_{t}_val"] = np.array(sv.expectation_value(opk, qs))
                sv.evolve(opk, qs)
                out[f"n{n}_ev{k}_{t}_q"] = np.array(qs)
                out[f"n{n}_ev{k}_{t}_op"] = opk
                out[f"n{n}_ev{k}_{t}"] = sv.flatten().copy()
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)
    print("kernels.npz:", len(out), "arrays")


def gen_cliffords() -> None:
    mats = np.stack([np.asarray(Clifford(i).matrix, dtype=np.complex128) for i in range(24)])
    np.savez_compressed(os.path.join(HERE, "cliffords.npz"), clifford=mats, X=np.asarray(Ops.X), Y=np.asarray(Ops.Y),
                        Z=np.asarray(Ops.Z), H=np.asarray(Ops.H), S=np.asarray(Ops.S))
    states = {name: np.asarray(getattr(BasicStates, name).to_statevector_numpy()).view(np.float64).tolist()
              for name in ("ZERO", "ONE", "PLUS", "MINUS", "PLUS_I", "MINUS_I")}
    rows = []
    for plane in (Plane.XY, Plane.YZ, Plane.XZ):
        for angle in (0.0, 0.1, 0.25, 0.5, 0.77, 1.0, 1.3, 1.5, 1.99, -0.4, -1.0):
            for s in (0, 1):
                for t in (0, 1):
                    m = Measurement.XY(0)
                    m = type(m)(angle, plane)
                    if s:
                        m = m.clifford(Clifford.X)
                    if t:
                        m = m.clifford(Clifford.Z)
                    b = m.to_bloch()
                    rows.append({"plane": plane.name, "angle": angle, "s": s, "t": t,
                                 "vec": [float(x) for x in b.plane.polar(b.angle)]})
            st = PlanarState(plane, angle).to_statevector_numpy()
            rows.append({"plane": plane.name, "angle": angle, "state": np.asarray(st).view(np.float64).tolist()})
    pauli = {}
    for name, m in (("X", Measurement.X), ("Y", Measurement.Y), ("Z", Measurement.Z),
                    ("-X", -Measurement.X), ("-Y", -Measurement.Y), ("-Z", -Measurement.Z)):
        b = m.to_bloch()
        pauli[name] = {"plane": b.plane.name, "angle": float(b.angle), "vec": [float(x) for x in b.plane.polar(b.angle)]}
    with open(os.path.join(HERE, "measure_adapt.json"), "w") as f:
        json.dump({"rows": rows, "basic_states": states, "pauli": pauli}, f, indent=0)
    print("cliffords.npz, measure_adapt.json")


# ---------------------------------------------------------------------------
# patterns
# ---------------------------------------------------------------------------
class Recording(BranchSelector):
    """Calls f_expectation0 for every measurement and records it (tests/test_branch_selector.py:29-44 style)."""

    def __init__(self, inner):
        self.inner = inner
        self.p0 = {}

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        p = f_expectation0()
        self.p0[qubit] = float(p)
        return self.inner.measure(qubit, lambda: p, rng, stacklevel=stacklevel + 1)


def encode_input(input_state):
    if input_state is BasicStates.PLUS:
        return {"kind": "plus"}
    if isinstance(input_state, list) and isinstance(input_state[0], PlanarState):
        return {"kind": "states", "states": [[s.plane.name, float(s.angle)] for s in input_state]}
    v = np.asarray(input_state, dtype=np.complex128)
    return {"kind": "vector", "data": v.view(np.float64).tolist()}


def trace(pattern, selector_spec, input_state=BasicStates.PLUS, rng_seed=None):
    """Run the reference StatevectorBackend under a recording selector."""
    if selector_spec[0] == "const":
        inner = ConstBranchSelector(selector_spec[1])
    else:
        inner = RandomBranchSelector(pr_calc=True)
    rec = Recording(inner)
    backend = StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=rec)
    sim = PatternSimulator(pattern, backend=backend)
    rng = None if rng_seed is None else np.random.default_rng(rng_seed)
    sim.run(input_state, rng=rng)
    return {
        "selector": list(selector_spec),
        "rng_seed": rng_seed,
        "input": encode_input(input_state),
        "p0": {str(k): v for k, v in rec.p0.items()},
        "outcomes": {str(k): int(v) for k, v in sim.measure_method.results.items()},
        "final_state": np.asarray(backend.state.flatten()).view(np.float64).tolist(),
    }


def emit(name, pattern, traces, meta=None):
    d = pattern_to_dict(pattern, name=name, meta=meta)
    d["traces"] = traces
    path = os.path.join(PAT_DIR, name + ".json.gz")
    save_pattern(path, d)
    print(f"{name}: {len(d['commands'])} commands, max_space {d['max_space']}, {os.path.getsize(path)} B")


def pipeline(circuit):
    p = circuit.transpile().pattern
    p.standardize()
    p.minimize_space()
    return p


def qaoa_circuit(n, seed=0):
    """QAOA MaxCut layer on the complete graph K_n (examples/qaoa.py:19-29 generalised)."""
    rng = np.random.default_rng(seed)
    g = nx.complete_graph(n)
    xi = rng.random(g.number_of_edges())
    theta = rng.random(n)
    c = Circuit(n)
    for i, (u, v) in enumerate(g.edges):
        c.cnot(u, v)
        c.rz(v, xi[i])
        c.cnot(u, v)
    for v in g.nodes:
        c.rx(v, theta[v])
    return c


def qft_circuit(n):
    """QFT on |+...+> (examples/qft_with_tn.py:22-56)."""
    c = Circuit(n)
    for i in range(n):
        c.h(i)

    def cp(theta, control, target):
        c.rz(control, theta / 2)
        c.rz(target, theta / 2)
        c.cnot(control, target)
        c.rz(target, -theta / 2)
        c.cnot(control, target)

    for i in range(n):
        c.h(i)
        for q in range(i + 1, n):
            cp(1.0 / 2 ** (q - i), q, i)
    for q in range(n // 2):
        c.swap(q, n - q - 1)
    return c


def gen_parity_patterns() -> None:
    os.makedirs(PAT_DIR, exist_ok=True)
    # config 1: the reference's CPU-runnable case
    circ = rand_circuit(8, 5, rng=np.random.default_rng(42))
    p = pipeline(circ)
    ref_state = circ.simulate().statevec if hasattr(circ.simulate(), "statevec") else None
    emit("config1_rand8x5", p,
         [trace(p, ("const", 0)), trace(p, ("const", 1)), trace(p, ("random",), rng_seed=7)],
         meta={"generator": "rand_circuit(8, 5, rng=default_rng(42)) -> transpile -> standardize -> minimize_space"})
    # standardized only (N* E* M* order, max_space = all nodes) on a small circuit: long runs of E
    circ = rand_circuit(3, 2, rng=np.random.default_rng(3))
    p = circ.transpile().pattern
    p.standardize()
    emit("standardized_rand3x2", p, [trace(p, ("const", 0)), trace(p, ("random",), rng_seed=11)],
         meta={"generator": "rand_circuit(3, 2, rng=default_rng(3)) -> transpile -> standardize"})
    # micro pattern with non-trivial probabilities
    p = Pattern(input_nodes=[])
    p.extend([N(0), N(1), E((0, 1)), M(0, Measurement.XY(0.1)), M(1, Measurement.YZ(0.3))])
    emit("micro_prob", p, [trace(p, ("const", 0)), trace(p, ("const", 1))])
    # hand-built: every command kind, all three planes, Pauli measurements, signals, Cliffords, non-|+> preparation
    p = Pattern(input_nodes=[0, 1])
    p.extend([
        N(2), N(3, PlanarState(Plane.XZ, 0.3)), N(4, BasicStates.ZERO), N(5, BasicStates.PLUS_I),
        E((0, 2)), E((1, 2)), E((2, 3)), E((3, 4)), E((4, 5)), E((0, 5)),
        CmdC(1, Clifford.H), CmdC(3, Clifford(13)),
        M(0, Measurement.XY(0.31)), M(2, Measurement.YZ(0.27), s_domain={0}),
        M(3, Measurement.XZ(1.4), s_domain={2}, t_domain={0}), M(4, Measurement.Y, t_domain={0, 2}),
        X(5, {3}), Z(5, {0, 4}), CmdC(5, Clifford(21)), Z(1, {2}), X(1, {4}),
    ])
    inp = [PlanarState(Plane.YZ, 0.6), BasicStates.MINUS]
    emit("handbuilt_allkinds", p,
         [trace(p, ("const", 0), inp), trace(p, ("const", 1), inp), trace(p, ("random",), inp, rng_seed=3),
          trace(p, ("random",), inp, rng_seed=4)])
    # arbitrary input vector
    circ = rand_circuit(4, 3, rng=np.random.default_rng(5))
    p = pipeline(circ)
    v = rand_state(np.random.default_rng(55), 4)
    emit("rand4x3_vector_input", p, [trace(p, ("const", 1), v), trace(p, ("random",), v, rng_seed=9)])
    # reduced-width versions of configs 2-4
    p = pipeline(rand_circuit(11, 6, rng=np.random.default_rng(42)))
    emit("rand11x6", p, [trace(p, ("const", 0)), trace(p, ("random",), rng_seed=1)],
         meta={"generator": "rand_circuit(11, 6, rng=default_rng(42)) pipeline"})
    p = pipeline(rand_circuit(14, 4, rng=np.random.default_rng(1)))
    emit("rand14x4", p, [trace(p, ("const", 1))], meta={"generator": "rand_circuit(14, 4, rng=default_rng(1)) pipeline"})
    p = pipeline(qaoa_circuit(7, seed=0))
    emit("qaoa7", p, [trace(p, ("const", 0)), trace(p, ("random",), rng_seed=2)],
         meta={"generator": "qaoa_circuit(7, seed=0) pipeline (config 3 at reduced width)"})
    p = pipeline(qft_circuit(10))
    emit("qft10", p, [trace(p, ("const", 0))], meta={"generator": "qft_circuit(10) pipeline (config 4 at reduced width)"})
    p = pipeline(qft_circuit(13))
    emit("qft13", p, [trace(p, ("const", 0))], meta={"generator": "qft_circuit(13) pipeline (config 4 at reduced width)"})


def gen_bench_patterns() -> None:
    os.makedirs(PAT_DIR, exist_ok=True)
    p = pipeline(rand_circuit(26, 20, rng=np.random.default_rng(42)))
    emit("config2_rand26x20", p, [], meta={"generator": "rand_circuit(26, 20, rng=default_rng(42)) pipeline"})
    p = pipeline(qaoa_circuit(29, seed=0))
    emit("config3_qaoa29", p, [], meta={"generator": "qaoa_circuit(29, seed=0) pipeline"})
    for n in (29, 32, 33, 34, 35):
        p = pipeline(qft_circuit(n))
        emit(f"config4_qft{n}", p, [], meta={"generator": f"qft_circuit({n}) pipeline"})
    # weak-scaling series of the N = 1 workload: one more circuit qubit per doubling of the GPU count
    for w in (27, 28, 29):
        p = pipeline(rand_circuit(w, 20, rng=np.random.default_rng(42)))
        emit(f"scale_rand{w}x20", p, [], meta={"generator": f"rand_circuit({w}, 20, rng=default_rng(42)) pipeline"})
    # mid-size cases whose reference run is still feasible on host cores (used by the CPU baseline leg)
    p = pipeline(rand_circuit(20, 5, rng=np.random.default_rng(42)))
    emit("rand20x5", p, [], meta={"generator": "rand_circuit(20, 5, rng=default_rng(42)) pipeline"})
    p = pipeline(rand_circuit(22, 3, rng=np.random.default_rng(42)))
    emit("rand22x3", p, [], meta={"generator": "rand_circuit(22, 3, rng=default_rng(42)) pipeline"})


# ---------------------------------------------------------------------------
# density-matrix path (BASELINE config 5)
# ---------------------------------------------------------------------------
def rand_dm(rng, n):
    a = rng.normal(size=(1 << n, 1 << n)) + 1j * rng.normal(size=(1 << n, 1 << n))
    rho = a @ a.conj().T
    return rho / np.trace(rho)


def gen_dm_kernels() -> None:
    from graphix.channels import depolarising_channel, two_qubit_depolarising_channel
    from graphix.sim.density_matrix import DensityMatrix

    rng = np.random.default_rng(43496)
    out: dict[str, np.ndarray] = {}
    for n in range(1, 5):
        rho = rand_dm(rng, n)
        out[f"n{n}_rho"] = rho
        op = rand_op(rng)
        out[f"n{n}_op"] = op
        s1 = rand_state(rng, 1)
        out[f"n{n}_s1"] = s1
        dm = DensityMatrix(rho)
        dm.add_nodes(1, BasicStates.PLUS)
        out[f"n{n}_addplus"] = dm.rho.copy()
        dm = DensityMatrix(rho)
        dm.add_nodes(1, s1)
        out[f"n{n}_add1"] = dm.rho.copy()
        for q in range(n):
            dm = DensityMatrix(rho)
            dm.evolve_single(op, q)
            out[f"n{n}_evolve_q{q}"] = dm.rho.copy()
            out[f"n{n}_expect_q{q}"] = np.array(DensityMatrix(rho).expectation_single(op, q))
            dm = DensityMatrix(rho)
            dm.apply_channel(depolarising_channel(0.13), [q])
            out[f"n{n}_depol_q{q}"] = dm.rho.copy()
            for r in (0, 1):
                vec = rand_bloch(rng)
                pm = _outcome_to_operator_matrix(vec, r)
                out[f"n{n}_proj_q{q}_r{r}_op"] = pm
                out[f"n{n}_proj_q{q}_r{r}_p"] = np.array(DensityMatrix(rho).expectation_single(pm, q))
                dm = DensityMatrix(rho)
                dm.project_qubit(pm, q)
                out[f"n{n}_proj_q{q}_r{r}"] = dm.rho.copy()
            for q2 in range(n):
                if q2 == q:
                    continue
                dm = DensityMatrix(rho)
                dm.entangle((q, q2))
                out[f"n{n}_cz_{q}_{q2}"] = dm.rho.copy()
                dm = DensityMatrix(rho)
                dm.apply_channel(two_qubit_depolarising_channel(0.21), [q, q2])
                out[f"n{n}_depol2_{q}_{q2}"] = dm.rho.copy()
                op2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
                out[f"n{n}_ev2_{q}_{q2}_op"] = op2
                dm = DensityMatrix(rho)
                dm.evolve(op2, [q, q2])
                out[f"n{n}_ev2_{q}_{q2}"] = dm.rho.copy()
        perm = [int(x) for x in rng.permutation(n)]
        dm = DensityMatrix(rho)
        dm.permute(perm)
        out[f"n{n}_perm_p"] = np.array(perm)
        out[f"n{n}_perm"] = dm.rho.copy()
    np.savez_compressed(os.path.join(HERE, "dm_kernels.npz"), **out)
    print("dm_kernels.npz:", len(out), "arrays")


NOISE = dict(prepare_error_prob=0.01, x_error_prob=0.01, z_error_prob=0.01, entanglement_error_prob=0.01,
             measure_channel_prob=0.01, measure_error_prob=0.0)


def dm_trace(pattern, selector_spec, noise_params, rng_seed, input_state=BasicStates.PLUS, store_rho=True):
    from graphix.noise_models.depolarising import DepolarisingNoiseModel
    from graphix.sim.density_matrix import DensityMatrixBackend

    inner = ConstBranchSelector(selector_spec[1]) if selector_spec[0] == "const" else RandomBranchSelector(pr_calc=True)
    rec = Recording(inner)
    backend = DensityMatrixBackend(branch_selector=rec)
    nm = DepolarisingNoiseModel(**noise_params) if noise_params is not None else None
    sim = PatternSimulator(pattern, backend=backend, noise_model=nm)
    sim.run(input_state, rng=np.random.default_rng(rng_seed))
    rho = np.asarray(backend.state.rho)
    t = {
        "selector": list(selector_spec), "rng_seed": rng_seed, "input": encode_input(input_state), "noise": noise_params,
        "p0": {str(k): v for k, v in rec.p0.items()},
        "outcomes": {str(k): int(v) for k, v in sim.measure_method.results.items()},
        "trace": float(np.trace(rho).real), "purity": float(np.trace(rho @ rho).real),
        "diag_head": np.real(np.diag(rho))[:64].tolist(),
    }
    if store_rho:
        t["final_rho"] = np.ascontiguousarray(rho).reshape(-1).view(np.float64).tolist()
    return t


def gen_dm_patterns(big: bool) -> None:
    os.makedirs(PAT_DIR, exist_ok=True)
    p = pipeline(rand_circuit(3, 2, rng=np.random.default_rng(42)))
    emit("dm_rand3x2", p, [dm_trace(p, ("const", 0), NOISE, 7), dm_trace(p, ("const", 1), NOISE, 8),
                           dm_trace(p, ("random",), NOISE, 9), dm_trace(p, ("random",), None, 10),
                           dm_trace(p, ("random",), dict(NOISE, measure_error_prob=0.3), 11)])
    p = Pattern(input_nodes=[0, 1])
    p.extend([N(2), N(3, PlanarState(Plane.XZ, 0.3)), E((0, 2)), E((1, 2)), E((2, 3)), CmdC(1, Clifford.H),
              M(0, Measurement.XY(0.31)), M(2, Measurement.YZ(0.27), s_domain={0}), X(3, {2}), Z(3, {0}),
              CmdC(3, Clifford(13)), Z(1, {2})])
    inp = [PlanarState(Plane.YZ, 0.6), BasicStates.MINUS]
    heavy = dict(prepare_error_prob=0.05, x_error_prob=0.1, z_error_prob=0.07, entanglement_error_prob=0.2,
                 measure_channel_prob=0.15, measure_error_prob=0.0)
    emit("dm_handbuilt", p, [dm_trace(p, ("const", 0), heavy, 1, inp), dm_trace(p, ("const", 1), heavy, 2, inp),
                             dm_trace(p, ("random",), heavy, 3, inp)])
    p = pipeline(rand_circuit(5, 3, rng=np.random.default_rng(42)))
    emit("dm_rand5x3", p, [dm_trace(p, ("const", 0), NOISE, 7), dm_trace(p, ("random",), NOISE, 5)],
         meta={"generator": "rand_circuit(5, 3, rng=default_rng(42)) pipeline (config 5 at reduced width)"})
    if big:
        import time

        p = pipeline(rand_circuit(9, 3, rng=np.random.default_rng(42)))
        t0 = time.time()
        emit("dm_rand9x3", p, [dm_trace(p, ("const", 0), NOISE, 7, store_rho=False)],
             meta={"generator": "rand_circuit(9, 3, rng=default_rng(42)) pipeline; digests only"})
        print("  reference DM run took", time.time() - t0, "s")
        p = pipeline(rand_circuit(12, 3, rng=np.random.default_rng(42)))
        emit("config5_rand12x3", p, [], meta={"generator": "rand_circuit(12, 3, rng=default_rng(42)) pipeline",
                                              "noise": NOISE, "selector": "ConstBranchSelector(0)", "rng": "default_rng(7)"})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-bench", action="store_true")
    ap.add_argument("--only-bench", action="store_true")
    ap.add_argument("--only-dm", action="store_true")
    a = ap.parse_args()
    if a.only_dm:
        gen_dm_kernels()
        gen_dm_patterns(big=True)
        raise SystemExit(0)
    if not a.only_bench:
        gen_dm_kernels()
        gen_dm_patterns(big=False)
        gen_kernels()
        gen_cliffords()
        gen_parity_patterns()
    if not a.no_bench:
        gen_bench_patterns()
