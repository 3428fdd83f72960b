"""Batched-projection pass timing as a function of register size, axis positions, axes per pass and chunking
(run on the GPU box):  python tools/microbench_batch.py <n materialised qubits> [reps]

Every case builds batches of fused (N, E, M) stages on chosen PHYSICAL bits through the public API and reports the
CUDA-event time per pass of the `fused` kernel kind (k_fused / k_fused_batch / k_fused_tile)."""
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


def sections_1_to_3(run, n, hi, lowmid, mid, top, spread, five_top, five_spread, six):
    # 1) one stage per pass (k_fused, two streams) vs position
    for ax in ((10,), (n // 2,), (hi,)):
        run("1 axis, 1 stage", ax, 1, 3, 0)
    # 2) three axes, three stages (k_fused_batch<3>, eight streams): position and chunking
    for name, ax in (("low", lowmid), ("mid", mid), ("top", top), ("spread", spread)):
        for glog in (3,):
            run(f"3 axes {name}", ax, 3, 3, glog)
    # 3) k_fused_tile: 4 / 5 / 6 axes, 8 stages
    for name, ax in (("4 axes spread", (9, n // 2, hi - 2, hi)), ("5 axes top", five_top), ("5 axes spread", five_spread),
                     ("6 axes", six)):
        run(name, ax, 8, len(ax), 3)
    # the same 8 stages on 5 axes as 3-axis passes (what round 1 ran)
    run("5 axes spread via A<=3", five_spread, 8, 3, 3)
    run("5 axes incl. low bits", (1, 4, n // 2, hi - 3, hi), 8, 5, 3)
    run("3 axes, 8 stages", spread, 8, 3, 3)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    sv = B200Statevector(nqubit=0, max_qubits=n, strict=False, batch=8)
    sv.add_nodes(n, mbqc.BasicStates.PLUS)
    sv.flush()
    sv.sync()
    sv.set_option(_lib.OPT_PROFILE, 1)
    pm = proj(mbqc.Plane.XY.polar(0.3), 0)
    rows = []
    hi = n - 1

    def run(label, axes, stages, batch_axes, glog, tma=int(os.environ.get("MB_TMA", "0")), io=1, tw=0):
        sv.set_option(_lib.OPT_TMA, tma)
        sv.set_option(_lib.OPT_TILE_WARPS, tw)
        sv.set_option(_lib.OPT_TILE_IO, io)
        sv.set_option(_lib.OPT_TILE_REGS, int(os.environ.get("MB_TILE_REGS", "4")))
        sv.set_option(_lib.OPT_BATCH_AXES, batch_axes)
        sv.set_option(_lib.OPT_CHUNK_LOG2, glog)
        sv.set_option(_lib.OPT_BATCH, min(12, stages))
        for timed in (False, True):
            sv.sync()
            sv.reset_stats()
            for _ in range(reps if timed else 1):
                for s in range(stages):
                    p = axes[s % len(axes)]
                    q = sv.layout().index(p)
                    sv.add_nodes(1, mbqc.BasicStates.PLUS)
                    sv.entangle((sv.nqubit - 1, q))
                    sv.project_qubit(pm, q)
                sv.flush() if False else None
            sv.sync()
        st = sv.stats()["fused"]
        ms = st["device_ms"] / max(st["launches"], 1)
        gbs = st["algo_bytes"] / (st["device_ms"] * 1e-3) / 1e9 if st["device_ms"] else 0.0
        rows.append({"n": n, "case": label, "io": io, "tile_warps": tw, "axes": list(axes), "stages": stages, "batch_axes": batch_axes, "glog": glog,
                     "launches": st["launches"], "ms_per_pass": ms, "algo_GBps": gbs})
        print(f"n={n} {label:34s} tw={tw} io={io} axes={str(list(axes)):26s} stages={stages} A<={batch_axes} glog={glog}: "
              f"{st['launches']:3d} passes {ms:9.4f} ms  {gbs:7.1f} GB/s", flush=True)

    lowmid = (9, 10, 11)
    mid = (n // 2 - 1, n // 2, n // 2 + 1)
    top = (hi - 2, hi - 1, hi)
    spread = (9, n // 2, hi)
    quick = os.environ.get("MB_QUICK", "0") == "1"
    five_top = (hi - 4, hi - 3, hi - 2, hi - 1, hi)
    five_spread = (8, 12, n // 2, hi - 3, hi)
    six = (8, 12, n // 2, hi - 5, hi - 2, hi)
    if not quick:
        sections_1_to_3(run, n, hi, lowmid, mid, top, spread, five_top, five_spread, six)
    # 4) the shapes the QAOA / QFT / random-circuit patterns produce: 7 stages on 7 axes (2 phases), rarely 12 stages
    seven = (8, 11, 12, n // 2, hi - 5, hi - 2, hi)
    seven_top = tuple(range(hi - 6, hi + 1))
    # 5) warps per tile (run length per axis combination): private warp tiles against tiles shared by 4 / 8 warps
    seven7 = dict(stages=7, batch_axes=7)
    for tw in ((4,) if quick else (1, 4, 8)):
        run("7 axes spread, 7 stages", seven, 7, 7, 3, tw=tw)
        run("7 axes top, 7 stages", seven_top, 7, 7, 3, tw=tw)
        run("7 axes spread, 12 stages", seven, 12, 7, 3, tw=tw)
        run("6 axes, 8 stages", six, 8, 6, 3, tw=tw)
        run("5 axes spread, 8 stages", five_spread, 8, 5, 3, tw=tw)
        run("4 axes spread, 12 stages", (9, n // 2, hi - 2, hi), 12, 4, 3, tw=tw)
    # 6) batch axes on the lowest physical bits (QFT patterns: 60 % of the passes): extra load / store layouts off / on
    low7 = (0, 1, 12, n // 2, hi - 5, hi - 2, hi)
    low7b = (1, 3, 12, n // 2, hi - 5, hi - 2, hi)
    low4 = (0, n // 2, hi - 2, hi)
    for io in (0, 1, 2):
        run("7 axes incl. bits 0,1, 7 stages", low7, 7, 7, 3, io=io)
        run("7 axes incl. bits 1,3, 7 stages", low7b, 7, 7, 3, io=io)
        run("7 axes incl. bits 0,1, 12 stages", low7, 12, 7, 3, io=io)
        run("4 axes incl. bit 0, 12 stages", low4, 12, 4, 3, io=io)
        run("7 axes spread, 7 stages", seven, 7, 7, 3, io=io)
        run("7 axes top, 7 stages", seven_top, 7, 7, 3, io=io)
        run("5 axes spread, 8 stages", five_spread, 8, 5, 3, io=io)
    print("norm2", sv.norm2())
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/microbench_batch_{n}.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
