"""Host-side logic of the sharded (multi-GPU) path under gloo on CPU, world_size 2 and 4.

The local shard engine is replaced by the numpy test double (tests/cpu_shard_engine.py); everything
else — global-bit placement, queued local/global CZs, rank-bit relabels, the half-shard exchange over
``torch.distributed`` send/recv, the norm all-reduce — is the product code of graphix_b200/sharded.py.
Results are checked against the golden traces of the reference and against the oracle.
"""
from __future__ import annotations

import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, case: str, q) -> None:
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = CASES[case](rank, world)
        q.put((rank, "ok", out))
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_case(case: str, world: int):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in range(world):
            results.append(q.get(timeout=int(os.environ.get("GXSV_SOAK_TIMEOUT", "300"))))
            if results[-1][1] != "ok":
                break       # the other ranks may be waiting in a collective for this one: report now
    finally:
        failed = len(results) < world or results[-1][1] != "ok"
        for p in procs:
            if failed:
                p.terminate()
            p.join(timeout=60)
    for rank, status, payload in results:
        assert status == "ok", f"rank {rank}:\n{payload}"
    return {rank: payload for rank, _, payload in results}


# ---------------------------------------------------------------------------
# cases (run inside the workers)
# ---------------------------------------------------------------------------
def _backend(cap, selector, strict=True):
    from cpu_shard_engine import CpuShardEngine

    from graphix_b200.sharded import ShardedStatevectorBackend

    g = dist.get_world_size().bit_length() - 1
    return ShardedStatevectorBackend.with_capacity(cap, engine=CpuShardEngine(max(cap - g, 1), strict=strict),
                                                   branch_selector=selector, strict=strict)


def case_golden(rank, world, strict=True):
    from helpers import check_trace, load_golden

    used_global = 0
    for name in ("config1_rand8x5", "micro_prob", "handbuilt_allkinds", "qaoa7", "qft10", "rand11x6"):
        d = load_golden(name)
        for trace in d["traces"]:
            holder = {}

            def factory(cap, sel, holder=holder):
                holder["b"] = _backend(cap, sel, strict)
                return holder["b"]

            check_trace(d, trace, factory)
            used_global += holder["b"].state.n_swaps
    return used_global


def case_soup(rank, world, strict=True, engine_factory=None, seeds=None):
    """Random command soup against the oracle with a tight capacity so that global bits are always in use.

    ``engine_factory(local_bits)`` lets tests/test_sharded_gpu.py replay the same streams on the CUDA engine."""
    from graphix_b200 import mbqc
    from oracle.oracle import OracleStatevector, outcome_to_operator_matrix

    from cpu_shard_engine import CpuShardEngine
    from graphix_b200.sharded import ShardedStatevector

    g = world.bit_length() - 1
    swaps = 0
    first = int(os.environ.get("GXSV_SOAK_FIRST", "0"))
    explicit = seeds is not None
    if not explicit:
        seeds = list(range(first, first + int(os.environ.get("GXSV_SOAK_SEEDS", "6"))))
    if not explicit and "GXSV_SOAK_SEEDS" not in os.environ:
        # regressions found by the long soak: 13 (world 2) exchanges while a virtual local qubit carries a
        # rank-dependent factor; 16 (world 4) exchanges after a projection left exact zeros in the rank scalars
        seeds += [13, 16]
    # GXSV_SOAK_MIX=gates: more single-qubit gates (relabels of rank bits, queued Z's) and fewer new qubits
    mix = {"default": [0.3, 0.3, 0.2, 0.08, 0.08, 0.04],
           "gates": [0.22, 0.25, 0.16, 0.16, 0.16, 0.05]}[os.environ.get("GXSV_SOAK_MIX", "default")]
    for seed in seeds:
        rng = np.random.default_rng(1000 + seed)
        cap = 7
        engine = CpuShardEngine(cap - g, strict=strict) if engine_factory is None else engine_factory(cap - g)
        sh = ShardedStatevector(engine, strict=strict)
        if not strict:
            sh.WINDOW = 5
        if seed % 2 == 0:
            sh.CHUNK_MIN = 2        # several chunks per exchange: exercises the double-buffered pipeline
        orc = OracleStatevector(nqubit=0, max_qubits=cap)
        states = [mbqc.BasicStates.PLUS, mbqc.BasicStates.MINUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE,
                  mbqc.PlanarState(mbqc.Plane.YZ, 0.37), mbqc.PlanarState(mbqc.Plane.XY, 1.21)]
        for step in range(150):
            n = orc.nqubit
            kind = rng.choice(["N", "E", "M", "U", "D", "P"], p=mix)
            if kind == "N" and n < cap:
                st = states[int(rng.integers(len(states)))]
                sh.add_nodes(1, st)
                orc.add_nodes(1, st)
            elif kind == "E" and n >= 2:
                for _ in range(int(rng.integers(1, 4))):
                    a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                    sh.entangle((a, b))
                    orc.entangle((a, b))
            elif kind == "M" and n >= 1:
                q = int(rng.integers(n))
                vec = rng.normal(size=3)
                vec /= np.linalg.norm(vec)
                pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
                po = orc.expectation_single(pm, q)
                if strict or step % 4 == 0:
                    ps = sh.expectation_single(pm, q)
                    assert abs(ps - po) < 1e-10, (seed, step, ps, po)
                if po.real > 1e-6:
                    sh.project_qubit(pm, q)
                    orc.project_qubit(pm, q)
            elif kind == "U" and n >= 1:
                q = int(rng.integers(n))
                m = mbqc.Clifford(int(rng.integers(24))).matrix
                sh.evolve_single(m, q)
                orc.evolve_single(m, q)
            elif kind == "D" and n >= 1:
                q = int(rng.integers(n))
                m = [mbqc.Ops.Z, mbqc.Ops.X, mbqc.Ops.S, mbqc.Ops.Y][int(rng.integers(4))]
                sh.evolve_single(m, q)
                orc.evolve_single(m, q)
            elif kind == "P" and n >= 2:
                perm = [int(x) for x in rng.permutation(n)]
                sh.permute(perm)
                orc.permute(perm)
            if step % 15 == 14:
                assert sh.nqubit == orc.nqubit
                np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11, err_msg=f"seed {seed} step {step}")
                assert abs(sh.norm2() - 1) < 1e-10
        np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11)
        swaps += sh.n_swaps
    return swaps


def case_vector_inputs(rank, world, engine_factory=None):
    """Joint 2^k input vectors (statevec.py:186-205) tensored onto partly filled sharded states: the slices that land
    on rank bits, the ones that fill the shard, exact zeros in a rank's slice; then commands on top."""
    from graphix_b200 import mbqc
    from oracle.oracle import OracleStatevector, outcome_to_operator_matrix

    from cpu_shard_engine import CpuShardEngine
    from graphix_b200.sharded import ShardedStatevector

    g = world.bit_length() - 1
    cap = 7
    for seed in range(10):
        rng = np.random.default_rng(7000 + seed)
        engine = CpuShardEngine(cap - g) if engine_factory is None else engine_factory(cap - g)
        sh = ShardedStatevector(engine, strict=True)
        orc = OracleStatevector(nqubit=0, max_qubits=cap)
        for _ in range(int(rng.integers(0, 4))):
            st = [mbqc.BasicStates.PLUS, mbqc.BasicStates.ONE, mbqc.PlanarState(mbqc.Plane.YZ, 0.37)][int(rng.integers(3))]
            sh.add_nodes(1, st)
            orc.add_nodes(1, st)
        if orc.nqubit >= 2:
            sh.entangle((0, 1))
            orc.entangle((0, 1))
        k = int(rng.integers(2, cap - orc.nqubit + 1))
        vec = rng.normal(size=1 << k) + 1j * rng.normal(size=1 << k)
        if seed % 3 == 0:
            vec[: 1 << (k - 1)] = 0.0          # first new qubit in |1>: exact zeros in some ranks' slices
        vec /= np.linalg.norm(vec)
        if seed % 4 == 1:
            # the same joint vector handed over shard by shard (ShardedInput: no rank ever sees the full vector)
            from graphix_b200.sharded import ShardedInput

            sh.add_nodes(k, ShardedInput(k, lambda t, kg, kl, vec=vec: np.ascontiguousarray(vec.reshape(1 << kg, 1 << kl)[t])))
        else:
            sh.add_nodes(k, vec if seed % 2 else list(vec))
        orc.add_nodes(k, vec)
        assert sh.nqubit == orc.nqubit
        np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-12, err_msg=f"seed {seed} after add")
        for step in range(30):
            n = orc.nqubit
            kind = rng.choice(["E", "M", "U"], p=[0.5, 0.3, 0.2])
            if kind == "E" and n >= 2:
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                sh.entangle((a, b))
                orc.entangle((a, b))
            elif kind == "M" and n >= 2:
                q = int(rng.integers(n))
                v3 = rng.normal(size=3)
                v3 /= np.linalg.norm(v3)
                pm = outcome_to_operator_matrix(v3, int(rng.integers(2)))
                po = orc.expectation_single(pm, q)
                assert abs(sh.expectation_single(pm, q) - po) < 1e-10, (seed, step)
                if po.real > 1e-6:
                    sh.project_qubit(pm, q)
                    orc.project_qubit(pm, q)
            elif kind == "U" and n >= 1:
                q = int(rng.integers(n))
                m = mbqc.Clifford(int(rng.integers(24))).matrix
                sh.evolve_single(m, q)
                orc.evolve_single(m, q)
        np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11, err_msg=f"seed {seed} end")
        # tensor() with another state object, then the dictionary read-outs (statevec.py:508-524, 590-668)
        if orc.nqubit <= cap - 2:
            other = OracleStatevector(nqubit=0, max_qubits=2)
            w = rng.normal(size=4) + 1j * rng.normal(size=4)
            other.add_nodes(2, w / np.linalg.norm(w))
            sh.tensor(other)
            orc.tensor(other)
            np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11, err_msg=f"seed {seed} tensor")
        flat = orc.flatten()
        # shard(): this rank's slice of the oracle vector, selected by rank_bit_values()
        fixed = sh.rank_bit_values()
        idx = tuple(fixed.get(i, slice(None)) for i in range(orc.nqubit))
        np.testing.assert_allclose(sh.shard(), flat.reshape((2,) * orc.nqubit)[idx].reshape(-1), atol=1e-11,
                                   err_msg=f"seed {seed} shard")
        got = sh.to_dict()
        keep = [i for i in range(flat.size) if abs(flat[i]) > 1e-8]
        assert sorted(got) == sorted(format(i, f"0{orc.nqubit}b") for i in keep)
        for i in keep:
            assert abs(got[format(i, f"0{orc.nqubit}b")] - flat[i]) < 1e-11
        probs = sh.to_prob_dict(encoding="LSB")
        assert abs(sum(probs.values()) - 1.0) < 1e-9
        assert isinstance(sh.draw(), str)
        with pytest.raises(NotImplementedError):
            sh.remove_qubit(0)
        with pytest.raises(NotImplementedError):
            sh.evolve(np.eye(4), [0, 1])
    with pytest.raises(ValueError, match="not normalized"):
        ShardedStatevector(CpuShardEngine(cap - g) if engine_factory is None else engine_factory(cap - g)).add_nodes(
            2, np.ones(4))
    return 0


def case_random_branch(rank, world):
    """Every rank must take the same branch: probabilities are all-reduced, the rng is seeded identically."""
    from helpers import load_golden

    from graphix_b200 import PatternSimulator, mbqc
    from graphix_b200.pattern_io import pattern_from_dict
    from oracle.oracle import OracleBackend

    d = load_golden("qaoa7")
    pattern = pattern_from_dict(d)
    rec = mbqc.RecordingBranchSelector(mbqc.RandomBranchSelector(pr_calc=True))
    b = _backend(pattern.max_space(), rec)
    sim = PatternSimulator(pattern, b)
    sim.run(rng=np.random.default_rng(99))
    rec2 = mbqc.RecordingBranchSelector(mbqc.RandomBranchSelector(pr_calc=True))
    o = OracleBackend.with_capacity(pattern.max_space(), branch_selector=rec2)
    sim2 = PatternSimulator(pattern, o)
    sim2.run(rng=np.random.default_rng(99))
    assert sim.measure_method.results == sim2.measure_method.results
    assert max(abs(rec.p0[k] - rec2.p0[k]) for k in rec2.p0) < 1e-10
    assert abs(np.vdot(o.state.flatten(), b.state.flatten())) ** 2 > 1 - 1e-10
    return [int(v) for v in sim.measure_method.results.values()]


def case_golden_nonstrict(rank, world):
    """Non-blocking projections: norms all-reduced "on the stream", zero-norm check deferred to the next sync."""
    import pytest as _pytest

    from graphix_b200 import PatternSimulator, mbqc

    swaps = case_golden(rank, world, strict=False)
    # |0> measured in Z with outcome 1 has probability 0: the error surfaces at finalize, not at measure
    p = mbqc.Pattern([], [mbqc.N(0, mbqc.BasicStates.ZERO), mbqc.N(1), mbqc.M(0, mbqc.Measurement.Z)])
    b = _backend(2, mbqc.ConstBranchSelector(1), strict=False)
    with _pytest.raises(RuntimeError):
        PatternSimulator(p, b).run()
    b = _backend(2, mbqc.ConstBranchSelector(1), strict=True)
    with _pytest.raises(RuntimeError):
        PatternSimulator(p, b).run()
    return swaps


def case_schedule(rank, world):
    """with_capacity(..., pattern=p): the measurement order drives the eviction policy (Belady) -> same states,
    fewer half-shard exchanges than the newest-qubit heuristic."""
    from cpu_shard_engine import CpuShardEngine
    from helpers import check_trace, load_golden

    from graphix_b200.pattern_io import pattern_from_dict
    from graphix_b200.sharded import ShardedStatevectorBackend

    g = dist.get_world_size().bit_length() - 1
    swaps = {False: 0, True: 0}
    for name in ("rand11x6", "qft10", "qaoa7"):
        d = load_golden(name)
        pattern = pattern_from_dict(d)
        for use in (False, True):
            holder = {}

            def factory(cap, sel, holder=holder, use=use):
                holder["b"] = ShardedStatevectorBackend.with_capacity(
                    cap, engine=CpuShardEngine(max(cap - g, 1)), branch_selector=sel, pattern=pattern if use else None)
                return holder["b"]

            for trace in d["traces"]:
                check_trace(d, trace, factory)
                swaps[use] += holder["b"].state.n_swaps
    assert swaps[True] < swaps[False], swaps
    return swaps[True]


def case_soup_nonstrict(rank, world):
    return case_soup(rank, world, strict=False)


CASES = {"golden": case_golden, "soup": case_soup, "vector_inputs": case_vector_inputs, "random_branch": case_random_branch,
         "golden_nonstrict": case_golden_nonstrict, "soup_nonstrict": case_soup_nonstrict, "schedule": case_schedule}


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_golden_traces(world):
    out = run_case("golden", world)
    assert len(set(out.values())) == 1       # every rank made the same layout decisions
    assert out[0] > 0                        # and the global <-> local exchange was exercised


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_command_soup_vs_oracle(world):
    out = run_case("soup", world)
    assert out[0] > 0 and len(set(out.values())) == 1


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_joint_vector_inputs(world):
    run_case("vector_inputs", world)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_nonstrict_command_soup(world):
    out = run_case("soup_nonstrict", world)
    assert out[0] > 0 and len(set(out.values())) == 1


def test_sharded_schedule_hint_reduces_exchanges():
    out = run_case("schedule", 4)
    assert len(set(out.values())) == 1


def test_sharded_nonstrict_deferred_norms():
    out = run_case("golden_nonstrict", 2)
    assert out[0] == out[1] and out[0] > 0


def test_sharded_random_branch_consistent():
    out = run_case("random_branch", 2)
    assert out[0] == out[1]
