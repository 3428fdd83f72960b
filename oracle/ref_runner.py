"""The REAL reference backend (graphix ``StatevectorBackend``: numba kernels graphix/sim/statevec.py:841-1198) timed on the
host cores over a bounded command window — the reference arm of ``bench.py`` (``--impl reference`` and the
``cpu_baseline`` leg), kind "reference".

Test / measurement infrastructure only (never imported by the product package).  The package is imported from
``/root/reference`` in the build container, else from the archive staged by ``oracle/stage_ref.py`` (GPU box); when
neither is importable ``available()`` says why and ``bench.py`` falls back to the C port (kind "port").
"""
from __future__ import annotations

import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))

_why = None


def available() -> tuple[bool, str]:
    """(importable, reason) — imports the reference on first use (~15 s: eager numba compilation)."""
    global _why
    if _why is not None:
        return _why == "", _why
    try:
        import ref_shim

        missing = ref_shim.missing_dependencies()
        if missing:
            _why = "missing python modules: " + ", ".join(missing)
        elif not ref_shim.reference_available():
            _why = f"reference package not found at {ref_shim.REFERENCE_ROOT} (run oracle/stage_ref.py where /root/reference exists)"
        else:
            ref_shim.import_reference()
            _why = ""
    except Exception as exc:  # noqa: BLE001
        _why = f"import of the reference failed: {type(exc).__name__}: {exc}"
    return _why == "", _why


def to_reference_pattern(d: dict):
    """A graphix ``Pattern`` from a fixture of tests/golden/patterns (graphix_b200.pattern_io encoding)."""
    from graphix.clifford import Clifford
    from graphix.command import C, E, M, N, X, Z
    from graphix.fundamentals import Plane
    from graphix.measurements import Measurement
    from graphix.pattern import Pattern
    from graphix.states import PlanarState

    makers = {"XY": Measurement.XY, "YZ": Measurement.YZ, "XZ": Measurement.XZ}
    cmds = []
    for c in d["commands"]:
        k = c[0]
        if k == "N":
            cmds.append(N(c[1]) if len(c) == 2 else N(c[1], PlanarState(Plane[c[2]], c[3])))
        elif k == "E":
            cmds.append(E((c[1], c[2])))
        elif k == "M":
            cmds.append(M(c[1], makers[c[2]](c[3]), s_domain=set(c[4]), t_domain=set(c[5])))
        elif k == "X":
            cmds.append(X(c[1], set(c[2])))
        elif k == "Z":
            cmds.append(Z(c[1], set(c[2])))
        elif k == "C":
            cmds.append(C(c[1], Clifford(c[2])))
    p = Pattern(input_nodes=list(d["input_nodes"]))
    p.extend(cmds)
    return p


def time_window(d: dict, n_cmds: int, steps: int, warmup: int, threads: int | None = None) -> tuple[list[float], int]:
    """Seconds of ``steps`` windows of ``n_cmds`` consecutive commands of fixture ``d`` on the reference backend under
    ``ConstBranchSelector(0)``.  Untimed setup as in ``bench.cpu_window``: inputs prepared, pattern advanced until the
    live width first reaches its maximum.  Returns (times, numba threads used)."""
    import numba
    from graphix.branch_selector import ConstBranchSelector
    from graphix.command import CommandKind
    from graphix.sim.statevec import StatevectorBackend
    from graphix.simulator import DefaultMeasureMethod
    from graphix.states import BasicStates

    if threads:
        numba.set_num_threads(min(threads, numba.config.NUMBA_NUM_THREADS))
    pattern = to_reference_pattern(d)
    cmds = list(pattern)
    width = d["max_space"]
    backend = StatevectorBackend.with_capacity(width, branch_selector=ConstBranchSelector(0))
    backend.add_nodes(pattern.input_nodes, BasicStates.PLUS)
    mm = DefaultMeasureMethod()

    def run(cmd) -> None:
        k = cmd.kind
        if k == CommandKind.N:
            backend.add_nodes(nodes=[cmd.node], data=cmd.state)
        elif k == CommandKind.E:
            backend.entangle_nodes(edge=cmd.nodes)
        elif k == CommandKind.M:
            mm.measure(backend, cmd)
        elif k in (CommandKind.X, CommandKind.Z):
            if mm.check_domain(cmd.domain):
                backend.correct_byproduct(cmd)
        elif k == CommandKind.C:
            backend.apply_clifford(cmd.node, cmd.clifford)

    pos = 0
    while backend.nqubit < width and pos < len(cmds):
        run(cmds[pos])
        pos += 1
    times = []
    for s in range(warmup + steps):
        if pos + n_cmds > len(cmds):
            break
        t0 = time.perf_counter()
        for c in cmds[pos:pos + n_cmds]:
            run(c)
        dt = time.perf_counter() - t0
        pos += n_cmds
        if s >= warmup:
            times.append(dt)
    return times, numba.get_num_threads()
