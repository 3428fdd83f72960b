"""Sharded path on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`).

Same cases as tests/test_sharded_cpu.py but with the CUDA shard engine and NCCL, plus a 1-GPU vs
2-GPU self-consistency check at a width the oracle does not reach."""
from __future__ import annotations

import os
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.gpu


def _worker(rank: int, world: int, port: int, case: str, q) -> None:
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        out = CASES[case](rank, world)
        q.put((rank, "ok", out))
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_case(case: str, world: int):
    from test_sharded_cpu import _free_port

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, payload in results:
        assert status == "ok", f"rank {rank}:\n{payload}"
    return {rank: payload for rank, _, payload in results}


def _backend(cap, selector, strict=True):
    from graphix_b200.sharded import ShardedStatevectorBackend

    dev = dist.get_rank() if dist.is_initialized() else 0
    return ShardedStatevectorBackend.with_capacity(cap, branch_selector=selector, device=dev, strict=strict)


def case_golden_nonstrict(rank, world):
    """Non-blocking batched projections on every shard; outcomes replayed without probability queries."""
    from helpers import check_trace, load_golden

    swaps = fused = 0
    for name in ("config1_rand8x5", "handbuilt_allkinds", "qaoa7", "qft10", "rand11x6", "rand14x4", "qft13"):
        d = load_golden(name)
        for trace in d["traces"]:
            holder = {}

            def factory(cap, sel, holder=holder):
                holder["b"] = _backend(cap, sel, strict=False)
                return holder["b"]

            check_trace(d, trace, factory, record=False)
            swaps += holder["b"].state.n_swaps
            fused += holder["b"].state.stats()["fused"]["launches"]
            check_trace(d, trace, factory, record=True)
    assert fused > 0
    return swaps


def case_golden(rank, world):
    from helpers import check_trace, load_golden

    swaps = 0
    for name in ("config1_rand8x5", "micro_prob", "handbuilt_allkinds", "qaoa7", "qft10", "rand11x6", "rand14x4", "qft13"):
        d = load_golden(name)
        for trace in d["traces"]:
            holder = {}

            def factory(cap, sel, holder=holder):
                holder["b"] = _backend(cap, sel)
                return holder["b"]

            check_trace(d, trace, factory)
            swaps += holder["b"].state.n_swaps
    return swaps


def case_soup_nonstrict(rank, world):
    return case_soup(rank, world, strict=False)


def case_soup(rank, world, strict=True):
    from graphix_b200 import mbqc
    from graphix_b200.sharded import GpuShardEngine, ShardedStatevector
    from oracle.oracle import OracleStatevector, outcome_to_operator_matrix

    g = world.bit_length() - 1
    swaps = 0
    for seed in range(3):
        rng = np.random.default_rng(500 + seed)
        cap = 14
        sh = ShardedStatevector(GpuShardEngine(cap - g, device=rank, strict=strict), strict=strict)
        if not strict:
            sh.WINDOW = 7
        orc = OracleStatevector(nqubit=0, max_qubits=cap)
        states = [mbqc.BasicStates.PLUS, mbqc.BasicStates.MINUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE,
                  mbqc.PlanarState(mbqc.Plane.YZ, 0.37)]
        for step in range(260):
            n = orc.nqubit
            kind = rng.choice(["N", "E", "M", "U", "D"], p=[0.34, 0.3, 0.2, 0.08, 0.08])
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
                if step % 3 == 0 and (strict or step % 4 == 0):
                    assert abs(sh.expectation_single(pm, q) - po) < 1e-10
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
            if step % 40 == 39:
                np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11, err_msg=f"seed {seed} step {step}")
        np.testing.assert_allclose(sh.flatten(), orc.flatten(), atol=1e-11)
        swaps += sh.n_swaps
    # the tight-capacity streams of the CPU (gloo) soak, including the seeds that once failed there, on the CUDA engine
    import test_sharded_cpu  # noqa: PLC0415

    nsoak = int(os.environ.get("GXSV_GPU_SOAK_SEEDS", "0"))     # longer replay on demand
    seeds = list(range(nsoak)) if nsoak else [0, 1, 13, 16, 22]
    swaps += test_sharded_cpu.case_soup(rank, world, strict=strict, seeds=seeds,
                                        engine_factory=lambda bits: GpuShardEngine(bits, device=rank, strict=strict))
    return swaps


def case_vector_inputs(rank, world):
    """Joint input vectors sliced over the ranks: the streams of tests/test_sharded_cpu.py on the CUDA engine."""
    import test_sharded_cpu  # noqa: PLC0415

    from graphix_b200.sharded import GpuShardEngine

    return test_sharded_cpu.case_vector_inputs(rank, world,
                                               engine_factory=lambda bits: GpuShardEngine(bits, device=rank))


def case_vs_single_gpu(rank, world):
    """rand20x5 (21 live qubits): sharded over `world` GPUs vs one GPU, same selector -> same state."""
    from helpers import load_golden

    from graphix_b200 import B200StatevectorBackend, PatternSimulator, mbqc
    from graphix_b200.pattern_io import pattern_from_dict

    pattern = pattern_from_dict(load_golden("rand20x5"))
    sel = mbqc.ConstBranchSelector(1)
    b = _backend(pattern.max_space(), sel)
    PatternSimulator(pattern, b).run()
    sharded = b.state.flatten()
    swaps = b.state.n_swaps
    single = B200StatevectorBackend.with_capacity(pattern.max_space(), branch_selector=sel, device=rank)
    PatternSimulator(pattern, single).run()
    fid = abs(np.vdot(single.state.flatten(), sharded)) ** 2
    assert fid > 1 - 1e-10, fid
    return swaps


CASES = {"golden": case_golden, "soup": case_soup, "vs_single": case_vs_single_gpu,
         "golden_nonstrict": case_golden_nonstrict, "soup_nonstrict": case_soup_nonstrict,
         "vector_inputs": case_vector_inputs}


def test_sharded_single_rank_runs_the_distributed_engine_mode():
    """world_size 1 (no process group): the sharded state object drives libgxsv in GXSV_OPT_DIST mode
    (local norms + set_norm, settle, lazy engine) on one GPU; golden traces and the soup must still match."""
    from helpers import check_trace, load_golden

    from graphix_b200.sharded import ShardedStatevectorBackend

    for name in ("config1_rand8x5", "micro_prob", "handbuilt_allkinds", "qaoa7", "rand11x6"):
        d = load_golden(name)
        for trace in d["traces"]:
            check_trace(d, trace, lambda cap, sel: ShardedStatevectorBackend.with_capacity(cap, branch_selector=sel, device=0))
    assert case_soup(0, 1) == 0
    assert case_soup(0, 1, strict=False) == 0
    assert case_vector_inputs(0, 1) == 0
    assert case_golden_nonstrict(0, 1) == 0


def _need(world: int) -> None:
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_gpu_golden(world):
    _need(world)
    out = run_case("golden", world)
    assert len(set(out.values())) == 1 and out[0] > 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gpu_soup(world):
    _need(world)
    out = run_case("soup", world)
    assert len(set(out.values())) == 1 and out[0] > 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gpu_joint_vector_inputs(world):
    _need(world)
    run_case("vector_inputs", world)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_gpu_nonstrict(world):
    _need(world)
    out = run_case("golden_nonstrict", world)
    assert len(set(out.values())) == 1 and out[0] > 0
    out = run_case("soup_nonstrict", world)
    assert len(set(out.values())) == 1 and out[0] > 0


@pytest.mark.parametrize("world", [2])
def test_sharded_gpu_vs_single(world):
    _need(world)
    run_case("vs_single", world)
