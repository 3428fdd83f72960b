"""Small run through every kernel of libgxsv for `compute-sanitizer --tool memcheck` (one tool per gpurun call,
/opt/skills/guides/B200_PROFILING.md).  Sizes are tiny on purpose (the sanitizer slows kernels by 10-100x) but cross
the layout boundaries: tile contraction (<= 11 bits), hole creation (bit >= 8), hole filling, extent trimming,
queued CZ stars and the general diagonal pass, fused and batched projections, k-qubit evolve, density-matrix traces,
pack/unpack, gather."""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from graphix_b200 import B200DensityMatrix, B200Statevector, mbqc, noise  # noqa: E402
from graphix_b200.sharded import GpuShardEngine  # noqa: E402


def proj(vec, r):
    x, y, z = vec
    s = -1.0 if r else 1.0
    return np.array([[0.5 + 0.5 * s * z, 0.5 * s * (x - 1j * y)], [0.5 * s * (x + 1j * y), 0.5 - 0.5 * s * z]])


def soup(sv, rng, steps, cap):
    for _ in range(steps):
        n = sv.nqubit
        kind = rng.choice(["N", "E", "M", "U", "D", "K"], p=[0.3, 0.3, 0.22, 0.08, 0.06, 0.04])
        if kind == "N" and n < cap:
            sv.add_nodes(1, [mbqc.BasicStates.PLUS, mbqc.BasicStates.ONE, mbqc.PlanarState(mbqc.Plane.YZ, 0.3)][int(rng.integers(3))])
        elif kind == "E" and n >= 2:
            for _ in range(int(rng.integers(1, 5))):
                a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
                sv.entangle((a, b))
        elif kind == "M" and n > 3:
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            q = int(rng.integers(n))
            pm = proj(v, int(rng.integers(2)))
            if sv.expectation_single(pm, q).real > 1e-6:
                sv.project_qubit(pm, q)
        elif kind == "U" and n >= 1:
            sv.evolve_single(mbqc.Clifford(int(rng.integers(24))).matrix, int(rng.integers(n)))
        elif kind == "D" and n >= 1:
            sv.evolve_single([mbqc.Ops.Z, mbqc.Ops.S][int(rng.integers(2))], int(rng.integers(n)))
        elif kind == "K" and n >= 3:
            k = int(rng.integers(2, 4))
            qs = [int(x) for x in rng.choice(n, size=k, replace=False)]
            u, _ = np.linalg.qr(rng.normal(size=(1 << k, 1 << k)) + 1j * rng.normal(size=(1 << k, 1 << k)))
            sv.evolve(u, qs)
    return sv


def main() -> None:
    rng = np.random.default_rng(1)
    for kw in ({"fusion": 0}, {"fusion": 1}, {"fusion": 2}, {"fusion": 2, "strict": False, "batch": 8}):
        sv = B200Statevector(nqubit=12, data=mbqc.BasicStates.PLUS, max_qubits=14, **kw)
        soup(sv, rng, 120, 14)
        assert abs(sv.norm2() - 1) < 1e-9
        sv.flatten()
    a = B200Statevector(nqubit=9, data=mbqc.BasicStates.PLUS)
    assert abs(a.fidelity(B200Statevector(a)) - 1) < 1e-12
    dm = B200DensityMatrix(nqubit=3, data=mbqc.BasicStates.PLUS)
    dm.entangle((0, 1))
    dm.apply_channel(noise.two_qubit_depolarising_channel(0.1), [0, 2])
    dm.apply_channel(noise.depolarising_channel(0.2), [1])
    dm.project_qubit(proj((0.6, 0.0, 0.8), 0), 1)
    dm.remove_qubit(0)
    assert abs(dm.trace() - 1) < 1e-10
    eng = GpuShardEngine(12, device=0)
    for _ in range(10):
        eng.add_qubit(2 ** -0.5, 2 ** -0.5)
    eng.entangle(0, 9)
    eng.settle(9)
    import torch

    buf = torch.empty(2 * eng.half_size(), dtype=torch.float64, device="cuda")
    eng.pack_half(9, 1, 0, eng.half_size(), buf)
    eng.unpack_half(9, 0, 0, eng.half_size(), buf, 1.0 + 0j)
    eng.sync()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
