"""Randomised soak of libgxsv against the CPU oracle (run on a GPU box): python tests/soak_gpu.py [seconds] [first_seed]

Every execution mode, registers of 0-14 qubits, every operation of the Statevector surface in random order; states are
compared entry by entry every few steps.  Prints the first failing (mode, seed, step) or a summary."""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from graphix_b200 import B200Statevector, mbqc  # noqa: E402
from oracle.oracle import OracleStatevector, outcome_to_operator_matrix  # noqa: E402

MODES = [{"fusion": 0}, {"fusion": 1}, {"fusion": 2}, {"fusion": 2, "batch": 0}, {"fusion": 2, "strict": False, "batch": 0},
         {"fusion": 2, "strict": False, "batch": 3}, {"fusion": 2, "strict": False, "batch": 12}]
STATES = [mbqc.BasicStates.PLUS, mbqc.BasicStates.MINUS, mbqc.BasicStates.ZERO, mbqc.BasicStates.ONE,
          mbqc.BasicStates.PLUS_I, mbqc.PlanarState(mbqc.Plane.YZ, 0.37), mbqc.PlanarState(mbqc.Plane.XZ, 1.21),
          mbqc.PlanarState(mbqc.Plane.XY, 0.77)]


def rand_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def one(mode, seed, steps=240, cap=14):
    rng = np.random.default_rng(seed)
    n0 = int(rng.integers(0, 11))
    psi = rand_state(rng, n0) if n0 else None
    dev = B200Statevector(psi, max_qubits=cap, **mode) if n0 else B200Statevector(nqubit=0, max_qubits=cap, **mode)
    orc = OracleStatevector(psi, max_qubits=cap) if n0 else OracleStatevector(nqubit=0, max_qubits=cap)
    for step in range(steps):
        n = orc.nqubit
        kind = rng.choice(["N", "E", "M", "U", "D", "P", "K", "T"], p=[0.26, 0.27, 0.24, 0.07, 0.07, 0.03, 0.03, 0.03])
        if kind == "N" and n < cap:
            st = STATES[int(rng.integers(len(STATES)))]
            dev.add_nodes(1, st)
            orc.add_nodes(1, st)
        elif kind == "T" and n + 2 <= cap:
            v = rand_state(rng, 2)
            dev.add_nodes(2, v)
            orc.add_nodes(2, v)
        elif kind == "E" and n >= 2:
            for _ in range(int(rng.integers(1, 6))):
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                dev.entangle((a, b))
                orc.entangle((a, b))
        elif kind == "M" and n >= 1:
            q = int(rng.integers(n))
            if rng.integers(4) == 0:
                vec = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, -1)][int(rng.integers(4))]
            else:
                vec = rng.normal(size=3)
                vec = vec / np.linalg.norm(vec)
            pm = outcome_to_operator_matrix(vec, int(rng.integers(2)))
            po = orc.expectation_single(pm, q)
            if rng.integers(3) == 0:
                pd = dev.expectation_single(pm, q)
                assert abs(pd - po) < 1e-10, ("expect", step, pd, po)
            if po.real > 1e-6:
                dev.project_qubit(pm, q)
                orc.project_qubit(pm, q)
        elif kind == "U" and n >= 1:
            q = int(rng.integers(n))
            m = mbqc.Clifford(int(rng.integers(24))).matrix
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        elif kind == "D" and n >= 1:
            q = int(rng.integers(n))
            m = [mbqc.Ops.Z, mbqc.Ops.X, mbqc.Ops.S, mbqc.Ops.Y][int(rng.integers(4))]
            dev.evolve_single(m, q)
            orc.evolve_single(m, q)
        elif kind == "P" and n >= 2:
            perm = [int(x) for x in rng.permutation(n)]
            dev.permute(perm)
            orc.permute(perm)
        elif kind == "K" and n >= 2:
            k = int(rng.integers(2, min(n, 4) + 1))
            qs = [int(x) for x in rng.choice(n, size=k, replace=False)]
            u, _ = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))
            dev.evolve(u, qs)
            orc.evolve(u, qs)
        if step % 24 == 23 or step == steps - 1:
            assert dev.nqubit == orc.nqubit, ("nqubit", step)
            err = float(np.max(np.abs(dev.flatten() - orc.flatten())))
            assert err < 1e-10, ("state", step, err)


def main() -> None:
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    runs = 0
    while time.time() - t0 < budget:
        for mode in MODES:
            try:
                one(mode, seed)
            except Exception as e:  # noqa: BLE001
                print("FAIL", mode, "seed", seed, repr(e)[:300], flush=True)
                raise SystemExit(1)
            runs += 1
        seed += 1
    print(f"soak ok: {runs} runs, seeds up to {seed - 1}, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
