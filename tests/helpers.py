"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

from graphix_b200 import mbqc
from graphix_b200.pattern_io import load_pattern_dict, pattern_from_dict
from graphix_b200.simulator import PatternSimulator

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P_TOL = 1e-10          # north_star: per-measurement outcome probabilities within 1e-10
F_TOL = 1e-10          # north_star: final-state fidelity >= 1 - 1e-10


def golden_pattern_names(with_traces: bool = True) -> list[str]:
    names = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "patterns", "*.json.gz"))):
        name = os.path.basename(p)[: -len(".json.gz")]
        if name.startswith(("dm_", "config5")):
            continue                                     # density-matrix fixtures: tests/test_dm_*.py
        if with_traces and name.startswith(("config2", "config3", "config4", "rand20", "rand22", "scale_")):
            continue
        names.append(name)
    return names


def load_golden(name: str) -> dict:
    return load_pattern_dict(os.path.join(GOLDEN, "patterns", name + ".json.gz"))


def decode_input(enc: dict):
    if enc["kind"] == "plus":
        return mbqc.BasicStates.PLUS
    if enc["kind"] == "states":
        return [mbqc.PlanarState(mbqc.Plane[p], a) for p, a in enc["states"]]
    return np.asarray(enc["data"], dtype=np.float64).view(np.complex128)


def make_selector(trace: dict) -> mbqc.BranchSelector:
    """Replay the reference's recorded outcomes; p0 is recorded for every measurement."""
    fixed = mbqc.FixedBranchSelector({int(k): v for k, v in trace["outcomes"].items()})
    return mbqc.RecordingBranchSelector(fixed)


def run_trace(d: dict, trace: dict, backend_factory, record: bool = True):
    """Run golden pattern ``d`` under ``trace``'s outcomes; returns (p0 dict, outcomes, final flat state).

    ``record=False`` replays the outcomes without ever asking for a probability (no synchronisation point:
    this is what lets the non-strict backend batch consecutive projections)."""
    pattern = pattern_from_dict(d)
    sel = make_selector(trace) if record else mbqc.FixedBranchSelector({int(k): v for k, v in trace["outcomes"].items()})
    backend = backend_factory(pattern.max_space(), sel)
    sim = PatternSimulator(pattern, backend=backend)
    sim.run(decode_input(trace["input"]))
    return getattr(sel, "p0", {}), sim.measure_method.results, np.asarray(backend.state.flatten())


def check_trace(d: dict, trace: dict, backend_factory, record: bool = True) -> None:
    p0, outcomes, flat = run_trace(d, trace, backend_factory, record)
    for k, v in trace["p0"].items():
        if record:
            assert abs(p0[int(k)] - v) <= P_TOL, (k, p0[int(k)], v)
    assert {str(k): int(v) for k, v in outcomes.items()} == trace["outcomes"]
    ref = np.asarray(trace["final_state"], dtype=np.float64).view(np.complex128)
    fid = abs(np.vdot(ref, flat)) ** 2
    assert fid >= 1 - F_TOL, fid
    assert abs(np.vdot(flat, flat).real - 1) <= 1e-10


def rand_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


# ---------------------------------------------------------------------------
# density-matrix traces
# ---------------------------------------------------------------------------
def run_dm_trace(d: dict, trace: dict, backend_factory):
    from graphix_b200.noise import DepolarisingNoiseModel

    pattern = pattern_from_dict(d)
    # same selector kind and same seed as the reference run: the rng stream (branch draws + the one
    # `confuse_result` draw per measurement, depolarising.py:146-147) is then consumed identically
    inner = mbqc.ConstBranchSelector(trace["selector"][1]) if trace["selector"][0] == "const" else \
        mbqc.RandomBranchSelector(pr_calc=True)
    sel = mbqc.RecordingBranchSelector(inner)
    backend = backend_factory(pattern.max_space(), sel)
    nm = DepolarisingNoiseModel(**trace["noise"]) if trace["noise"] is not None else None
    sim = PatternSimulator(pattern, backend=backend, noise_model=nm)
    sim.run(decode_input(trace["input"]), rng=np.random.default_rng(trace["rng_seed"]))
    return sel.p0, sim.measure_method.results, backend.state


def check_dm_trace(d: dict, trace: dict, backend_factory, atol: float = 1e-10) -> None:
    p0, outcomes, state = run_dm_trace(d, trace, backend_factory)
    for k, v in trace["p0"].items():
        assert abs(p0[int(k)] - v) <= P_TOL, (k, p0[int(k)], v)
    assert {str(k): int(v) for k, v in outcomes.items()} == trace["outcomes"]
    rho = np.asarray(state.rho)
    assert abs(np.trace(rho).real - trace["trace"]) <= 1e-10
    assert abs(np.trace(rho @ rho).real - trace["purity"]) <= 1e-9
    np.testing.assert_allclose(np.real(np.diag(rho))[:64], trace["diag_head"], atol=atol)
    if "final_rho" in trace:
        ref = np.asarray(trace["final_rho"], dtype=np.float64).view(np.complex128).reshape(rho.shape)
        np.testing.assert_allclose(rho, ref, atol=atol)
