"""Per-kernel device timing at a given width (run on the GPU box): python tools/microbench.py [nqubits]"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from graphix_b200 import B200Statevector, _lib, mbqc  # noqa: E402


def proj(vec, r):
    x, y, z = vec
    s = -1.0 if r else 1.0
    return np.array([[0.5 + 0.5 * s * z, 0.5 * s * (x - 1j * y)], [0.5 * s * (x + 1j * y), 0.5 - 0.5 * s * z]])


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    sv = B200Statevector(nqubit=n - 1, data=mbqc.BasicStates.PLUS, max_qubits=n)
    sv.set_option(_lib.OPT_PROFILE, 1)
    pm = proj(mbqc.Plane.XY.polar(0.3), 0)
    rows = []

    def run(label, fn, reset=True):
        sv.sync()
        sv.reset_stats()
        fn()
        sv.sync()
        st = sv.stats()
        for k, v in st.items():
            if v["launches"]:
                gbs = v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 if v["device_ms"] else 0.0
                rows.append({"case": label, "kind": k, "launches": v["launches"], "ms_per_launch": v["device_ms"] / v["launches"],
                             "algo_GBps": gbs})
                print(f"{label:34s} {k:9s} x{v['launches']:3d}  {v['device_ms'] / v['launches']:8.4f} ms  {gbs:8.1f} GB/s (algorithmic)", flush=True)

    # warm-up
    for _ in range(2):
        sv.add_nodes(1, mbqc.BasicStates.PLUS)
        sv.entangle((n - 1, 0))
        sv.project_qubit(pm, 0)

    for target in (0, 3, 7, 8, 12, n - 2):
        def triple(t=target):
            for _ in range(reps):
                sv.add_nodes(1, mbqc.BasicStates.PLUS)      # N at width n-1 -> n
                sv.entangle((n - 1, t))                      # E
                sv.project_qubit(pm, t)                      # M at width n -> n-1
        run(f"N,E,M(q={target}) width {n - 1}<->{n}", triple)
        print("   layout:", sv.layout(), flush=True)

    def ones():
        for q in (0, 5, n - 2):
            sv.evolve_single(mbqc.Ops.X, q)
            sv.evolve_single(mbqc.Ops.Z, q)
            sv.evolve_single(mbqc.Clifford(7).matrix, q)
    run(f"X,Z,C width {n - 1}", ones)

    def expect():
        for q in (0, 5, n - 2):
            sv.expectation_single(pm, q)
    run(f"expectation width {n - 1}", expect)
    print("norm2", sv.norm2())
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/microbench_{n}.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
