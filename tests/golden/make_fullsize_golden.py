"""Reference-derived parity fixtures at BASELINE sizes (run where /root/reference exists; ~30 min on 8 cores).

    python tests/golden/make_fullsize_golden.py [--only nonflow|config2]

* ``fullsize/config2_digests.json`` — the REAL reference ``StatevectorBackend`` (graphix/sim/statevec.py:765-832, numba
  kernels :841-1198) runs BASELINE config 2 (``rand_circuit(26, 20)`` pipeline, 3756 commands, 26<->27 live qubits)
  under ``ConstBranchSelector(0)`` and under seeded pseudo-random outcomes.  The 1 GiB final state cannot be committed, so
  digests are: the inner products <r_k|psi> with 4 seeded Gaussian vectors, 1024 sampled amplitudes, and the norm.
  ``digest_state`` below is imported by the GPU test so that both sides compute the same thing.
* ``patterns/nonflow22.json.gz`` — a hand-built NON-flow pattern on 22 qubits (random graph of CZs on arbitrary planar product states,
  measured at arbitrary angles in all three planes, no corrections): the per-measurement p0 recorded from the reference are far from
  1/2, and the outcomes are drawn from them (``RandomBranchSelector(pr_calc=True)``).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FULL_DIR = os.path.join(HERE, "fullsize")

PROBE_SEEDS = (101, 102, 103, 104)
SAMPLE_SEED = 999
N_SAMPLES = 1024


def probe_vector(seed: int, n: int) -> np.ndarray:
    """Seeded complex Gaussian vector of 2^n entries (one generator call per part, so that both sides agree)."""
    rng = np.random.default_rng(seed)
    re = rng.standard_normal(1 << n)
    im = rng.standard_normal(1 << n)
    return re + 1j * im


def sample_indices(n: int) -> np.ndarray:
    return np.random.default_rng(SAMPLE_SEED).integers(0, 1 << n, size=N_SAMPLES)


def digest_state(flat: np.ndarray) -> dict:
    n = int(flat.size).bit_length() - 1
    inner = []
    for s in PROBE_SEEDS:
        r = probe_vector(s, n)
        z = complex(np.vdot(r, flat))
        inner.append([z.real, z.imag])
        del r
    idx = sample_indices(n)
    return {"nqubit": n, "norm2": float(np.vdot(flat, flat).real), "inner": inner,
            "samples": np.asarray(flat[idx], dtype=np.complex128).view(np.float64).tolist()}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=[None, "nonflow", "config2"])
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, HERE)
    import make_golden as mg  # imports the reference through tools/ref_shim.py

    os.makedirs(FULL_DIR, exist_ok=True)
    if a.only in (None, "nonflow"):
        gen_nonflow(mg)
    if a.only in (None, "config2"):
        gen_config2(mg)


def gen_nonflow(mg) -> None:
    import networkx as nx

    rng = np.random.default_rng(2222)
    n = 22
    g = nx.gnm_random_graph(n, 30, seed=7)
    p = mg.Pattern(input_nodes=[])
    prep = (mg.Plane.XZ, mg.Plane.YZ, mg.Plane.XY)
    cmds = [mg.N(i) if i % 7 == 6 else mg.N(i, mg.PlanarState(prep[i % 3], float(2 * rng.random()))) for i in range(n)]
    cmds += [mg.E((int(u), int(v))) for u, v in g.edges]
    makers = (mg.Measurement.XY, mg.Measurement.YZ, mg.Measurement.XZ)
    order = [int(x) for x in rng.permutation(n)]
    for node in order[:18]:
        cmds.append(mg.M(node, makers[int(rng.integers(3))](float(2 * rng.random()))))
    p.extend(cmds)
    traces = [mg.trace(p, ("random",), rng_seed=31), mg.trace(p, ("random",), rng_seed=32), mg.trace(p, ("const", 0))]
    for t in traces:
        ps = np.array(list(t["p0"].values()))
        print("  nonflow22 p0 range", ps.min(), ps.max(), "far from 1/2:", int(np.sum(np.abs(ps - 0.5) > 0.05)))
    mg.emit("nonflow22", p, traces, meta={"generator": "tests/golden/make_fullsize_golden.py: gnm_random_graph(22, 30, seed=7) "
                                                      "graph state, 18 arbitrary-angle measurements, no corrections"})


class ReplayOutcomes:
    """Outcomes fixed in advance (never asks for a probability)."""

    def __init__(self, outcomes):
        self.outcomes = outcomes

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):
        return self.outcomes[qubit]


def gen_config2(mg) -> None:
    from graphix.branch_selector import BranchSelector

    class Replay(ReplayOutcomes, BranchSelector):
        pass

    from graphix_b200.pattern_io import load_pattern_dict

    d = load_pattern_dict(os.path.join(mg.PAT_DIR, "config2_rand26x20.json.gz"))
    p = mg.pipeline(mg.rand_circuit(26, 20, rng=np.random.default_rng(42)))
    assert mg.pattern_to_dict(p)["commands"] == d["commands"], "committed config 2 fixture differs from the generator"
    measured = [c[1] for c in d["commands"] if c[0] == "M"]
    runs = {"const0": {m: 0 for m in measured},
            "seeded5": {m: int(b) for m, b in zip(measured, np.random.default_rng(5).integers(0, 2, size=len(measured)))}}
    out = {"pattern": "config2_rand26x20", "probe_seeds": list(PROBE_SEEDS), "sample_seed": SAMPLE_SEED, "runs": {}}
    for name, outcomes in runs.items():
        backend = mg.StatevectorBackend.with_capacity(p.max_space(), branch_selector=Replay(outcomes))
        t0 = time.time()
        mg.PatternSimulator(p, backend=backend).run()
        dt = time.time() - t0
        flat = np.asarray(backend.state.flatten())
        rec = digest_state(flat)
        rec["outcomes"] = {str(k): v for k, v in outcomes.items()}
        rec["reference_seconds"] = dt
        rec["reference_commands_per_s"] = len(d["commands"]) / dt
        rec["host_cores"] = os.cpu_count()
        out["runs"][name] = rec
        print(f"  config2 {name}: {dt:.1f} s on {os.cpu_count()} cores = {len(d['commands']) / dt:.2f} commands/s, "
              f"norm2 = {rec['norm2']:.15f}", flush=True)
        del backend, flat
        with open(os.path.join(FULL_DIR, "config2_digests.json"), "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
