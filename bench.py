#!/usr/bin/env python
"""Benchmark of the B200 statevector pattern simulator (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--fusion F]

A *step* is one complete pattern simulation (``PatternSimulator.run``: prepare inputs, every
N/E/M/X/Z/C command, finalize) of the workload under ``ConstBranchSelector(0)``.

* N = 1 : BASELINE config 2, ``rand_circuit(26, 20)`` space-minimised pattern, 3756 commands, 26<->27 live
  qubits, complex128 (state 1-2 GiB, far larger than the 126 MB L2, so no flush is needed between steps).
* N > 1 : (torchrun, one rank per GPU) the state is sharded by its top log2(N) qubits
  (``graphix_b200.sharded``); weak scaling: ``rand_circuit(26 + log2 N, 20)``, i.e. the per-GPU shard stays
  1-2 GiB, every command touches N times more amplitudes, and ideal scaling keeps commands/s constant.

Prints ONE JSON line (rank 0).  ``value`` = commands/s with the inputs resident in HBM, timed with CUDA events
on the launching stream (max over ranks); ``e2e`` = the same through the public call with HOST buffers (the
input statevector is copied from pinned host memory and the final state is read back inside the timed region);
``roofline`` = dominant kernel vs the measured HBM copy rate; ``cpu_baseline`` = the oracle port of the
reference numba backend timed on this box's host cores over a bounded window of the same pattern.

``--impl reference`` times that CPU port alone (all host threads, bounded window per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "pattern_sim_commands_per_s"
UNIT = "commands/s"
NOMINAL_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# kernel kind (gxsv_stats) -> kernel name in the committed ncu captures
_NCU_KERNEL = {"fused": "k_fused_tile", "contract": "k_contract", "fill": "k_fill", "cz": "k_cz_star"}


def ncu_traffic(kind: str, workload: str, batched: bool) -> float | None:
    """DRAM bytes per launch of the dominant kernel from the ncu --set full capture committed under profiles/
    (taken on config 2; not transferable to other workloads, hence None there).  A constant read from the committed
    capture, not measured in this run: it goes stale when the kernel changes (profiles/r02_ncu_traffic.json says which
    capture each number comes from)."""
    if workload != "config2_rand26x20":
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            t = json.load(f)
        name = _NCU_KERNEL.get(kind)
        if kind == "fused" and not batched:
            name = "k_fused"
        return float(t[name]["dram_bytes_per_launch"])
    except Exception:  # noqa: BLE001
        return None


def workload_for(n_gpus: int, name: str | None) -> str:
    if name:
        return name
    # weak scaling of the N = 1 workload: one more circuit qubit per doubling of the GPU count, so every GPU
    # keeps a 1-2 GiB shard.  The 33..36-live-qubit series of BASELINE config 4 (64-128 GiB per GPU) is selected
    # with --workload config4_qft32 / 33 / 34 / 35 (use --max-commands for a bounded prefix window).
    return {1: "config2_rand26x20", 2: "scale_rand27x20", 4: "scale_rand28x20", 8: "scale_rand29x20"}[n_gpus]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self) -> None:
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU reference arm (oracle port of the numba backend)
# ---------------------------------------------------------------------------
def cpu_window(pattern, n_cmds: int, steps: int, warmup: int):
    """Time ``steps`` windows of ``n_cmds`` consecutive commands of ``pattern`` on the oracle backend.

    Setup (untimed): the input qubits are prepared and the pattern is advanced until the live width first
    reaches its maximum, so every timed command runs at the widths the GPU arm runs at."""
    from graphix_b200 import mbqc
    from graphix_b200.simulator import DefaultMeasureMethod
    from oracle.oracle import OracleBackend

    cmds = list(pattern)
    backend = OracleBackend.with_capacity(pattern.max_space(), branch_selector=mbqc.ConstBranchSelector(0))
    backend.add_nodes(pattern.input_nodes, mbqc.BasicStates.PLUS)
    mm = DefaultMeasureMethod()

    def run(cmd) -> None:
        k = cmd.kind.name
        if k == "N":
            backend.add_nodes([cmd.node], cmd.state)
        elif k == "E":
            backend.entangle_nodes(cmd.nodes)
        elif k == "M":
            mm.measure(backend, cmd)
        elif k in ("X", "Z"):
            if mm.check_domain(cmd.domain):
                backend.correct_byproduct(cmd)
        elif k == "C":
            backend.apply_clifford(cmd.node, cmd.clifford)

    pos = 0
    while backend.nqubit < pattern.max_space() and pos < len(cmds):
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
    return times


def host_window(name: str, pattern, n_cmds: int, steps: int, warmup: int, force_port: bool = False):
    """The reference's CPU implementation of the path over a bounded window: the REAL reference (graphix numba backend,
    oracle/ref_runner.py; kind "reference") when it can be imported here, else the C/OpenMP port of the oracle (kind "port").
    Returns (times, kind, cores, note)."""
    if not force_port:
        try:
            from oracle import ref_runner

            ok, why = ref_runner.available()
            if ok:
                from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern_dict

                d = load_pattern_dict(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
                times, threads = ref_runner.time_window(d, n_cmds, steps, warmup)
                return times, "reference", threads, "graphix StatevectorBackend (numba kernels), numba threads = %d" % threads
            note = why
        except Exception as exc:  # noqa: BLE001
            note = f"{type(exc).__name__}: {exc}"
    else:
        note = "forced"
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)      # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every core
    return cpu_window(pattern, n_cmds, steps, warmup), "port", cores, "C/OpenMP port of the reference kernels (real reference unavailable: %s)" % note


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    name = workload_for(args.gpus, args.workload)
    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    if pattern.max_space() > 31:
        print(json.dumps({"impl": "reference", "metric": METRIC, "unit": UNIT, "value": None, "n_gpus": args.gpus,
                          "unavailable": "reference is limited to 31 qubits (int32 kernel signatures, SURVEY.md §2.1)"}))
        return
    width = pattern.max_space()
    n_cmds = args.cpu_window or max(3, min(30, int(round(12 * 2.0 ** (27 - width)))) if width >= 24 else 300)
    times, kind, cores, note = host_window(name, pattern, n_cmds, args.steps, args.warmup, force_port=args.port)
    if not times:
        raise SystemExit("pattern too short for the requested window")
    per_step = float(np.mean(times))
    value = n_cmds / per_step
    sample = (f"{n_cmds} consecutive commands of {name} per step at {width - 1}<->{width} live qubits (bounded window); "
              f"{note}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "selector": "ConstBranchSelector(0)", "commands_per_step": n_cmds},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class SeededOutcomes:
    """Deterministic pseudo-random outcomes that never ask for a probability (flow patterns: every p = 1/2)."""

    def __init__(self, seed: int) -> None:
        self.seed = seed
        self.rng = np.random.default_rng(seed)

    def measure(self, qubit, f_expectation0, rng=None, *, stacklevel=1):  # noqa: ARG002
        return int(self.rng.integers(2))


def kernel_table(stats: dict, steps: int, peak: float) -> dict:
    return {k: {"launches_per_step": v["launches"] / steps, "ms_per_launch": v["device_ms"] / v["launches"],
                "GBps": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9,
                "frac_of_measured_peak": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 / peak}
            for k, v in stats.items() if v["launches"] and v["device_ms"] > 0}


def roofline_of(stats: dict, dev_ms: float, traffic=None) -> dict:
    """Dominant kernel (by summed CUDA-event time inside the timed region) against the measured HBM copy rate."""
    peak, peak_src = measured_peak()
    kinds = {k: v for k, v in stats.items() if v["launches"] and v["device_ms"] > 0}
    dom = max(kinds, key=lambda k: kinds[k]["device_ms"])
    d = kinds[dom]
    achieved = d["algo_bytes"] / (d["device_ms"] * 1e-3) / 1e9
    total_ms = sum(v["device_ms"] for v in kinds.values())
    total_bytes = sum(v["algo_bytes"] for v in kinds.values())
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / NOMINAL_HBM_GBS, "traffic": traffic,
            "algo_bytes_per_launch": d["algo_bytes"] / d["launches"], "avg_launch_ms": d["device_ms"] / d["launches"],
            "share_of_kernel_time": d["device_ms"] / total_ms,
            "all_kernels_achieved": total_bytes / (total_ms * 1e-3) / 1e9,
            "all_kernels_frac": total_bytes / (total_ms * 1e-3) / 1e9 / peak,
            "kernel_time_share_of_step": total_ms / dev_ms}


class Timer:
    """Barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""

    def __init__(self, world: int) -> None:
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.world = torch, dist, world

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps: int) -> float:
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        self.wall_ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def timed_pattern(timer: Timer, backend, pattern, steps: int, warmup: int, run_kwargs=None, sharded: bool = False):
    """W untimed + K timed complete runs of ``pattern`` on ``backend``; returns (device ms, engine stats)."""
    from graphix_b200 import PatternSimulator, _lib

    kw = run_kwargs or (lambda: {})

    def one_step() -> None:
        backend.state.reset()
        backend.node_index.clear()
        PatternSimulator(pattern, backend).run(**kw())

    for _ in range(warmup):
        one_step()
    timer.barrier()
    backend.state.reset_stats()
    backend.state.set_option(_lib.OPT_PROFILE, 1)
    if sharded:
        backend.state.profile = True
    ms = timer.run(one_step, steps)
    stats = backend.state.stats()
    backend.state.set_option(_lib.OPT_PROFILE, 0)
    return ms, stats


def exchange_record(st, steps: int) -> dict:
    st.collect_profile()
    rec = {"swaps_per_step": st.n_swaps / steps,
           "nvlink_bytes_per_rank_per_step": st.exchanged_bytes / steps,
           "device_ms_per_step": st.exchange_ms / steps,
           "host_ms_per_step": st.exchange_host_ms / steps,
           "host_wait_for_device_ms_per_step": st.sync_wait_ms / steps,
           "GBps_per_direction": (st.exchanged_bytes / 1e9) / (st.exchange_ms * 1e-3) if st.exchange_ms else None,
           "transport": getattr(st, "transport", "nccl send/recv")}
    st.profile = False
    return rec


def free_device_memory() -> None:
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()


def make_gpu_backend(world: int, width: int, selector, *, fusion=2, nonstrict=True, batch=12, device=0, pattern=None,
                     schedule=True):
    from graphix_b200 import B200StatevectorBackend

    if world > 1:
        from graphix_b200.sharded import ShardedStatevectorBackend

        return ShardedStatevectorBackend.with_capacity(width, branch_selector=selector, fusion=fusion,
                                                       strict=not nonstrict, batch=batch or 12,
                                                       pattern=pattern if schedule else None)
    lazy = fusion == 2 and nonstrict
    # strict + batch > 0 is the library default: only projections whose outcome probability is analytic are queued
    return B200StatevectorBackend.with_capacity(width, branch_selector=selector, fusion=fusion, device=device,
                                                strict=not lazy, batch=batch if fusion == 2 else 0)


def sharded_parity(world: int) -> dict:
    """Golden traces of the REFERENCE StatevectorBackend (tests/golden/patterns) replayed through the sharded backend
    on the real NCCL ranks, strict and non-blocking: per-measurement p0 and final-state fidelity."""
    from graphix_b200 import PatternSimulator, mbqc
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern_dict, pattern_from_dict
    from graphix_b200.sharded import ShardedStatevectorBackend

    max_dp, min_fid, cases = 0.0, 1.0, 0
    for name in ("qft13", "rand14x4", "nonflow22"):
        d = load_pattern_dict(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
        pattern = pattern_from_dict(d)
        for trace in d["traces"][:2]:
            if trace["input"]["kind"] != "plus":
                continue
            outcomes = {int(k): v for k, v in trace["outcomes"].items()}
            ref = np.asarray(trace["final_state"], dtype=np.float64).view(np.complex128)
            for strict, record in ((True, True), (False, False)):
                fixed = mbqc.FixedBranchSelector(outcomes)
                sel = mbqc.RecordingBranchSelector(fixed) if record else fixed
                b = ShardedStatevectorBackend.with_capacity(pattern.max_space(), branch_selector=sel, strict=strict,
                                                            pattern=pattern)
                PatternSimulator(pattern, b).run()
                if record:
                    max_dp = max(max_dp, max(abs(sel.p0[int(k)] - v) for k, v in trace["p0"].items()))
                flat = np.asarray(b.state.flatten())
                min_fid = min(min_fid, float(abs(np.vdot(ref, flat)) ** 2))
                cases += 1
                del b
    return {"max_dp0": max_dp, "min_fidelity": min_fid, "one_minus_min_fidelity": 1.0 - min_fid, "cases": cases,
            "patterns": ["qft13", "rand14x4", "nonflow22"], "ranks": world,
            "against": "reference StatevectorBackend traces committed under tests/golden/patterns"}


def config3_record(args, timer: Timer, local_rank: int) -> dict:
    """BASELINE config 3: QAOA MaxCut on K29 (examples/qaoa.py:19-29 generalised), 8352 commands, 29<->30 live qubits,
    16 GiB register — the workload the '>= 70 % of HBM on a >= 30-live-qubit pattern' target is stated on."""
    from graphix_b200 import mbqc
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    name = "config3_qaoa29"
    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    width = pattern.max_space()
    sel = mbqc.ConstBranchSelector(0)
    out = {"workload": name, "commands_per_step": len(pattern), "live_qubits": f"{width - 1}<->{width}",
           "state_bytes_max": 16 << width, "selector": "ConstBranchSelector(0)"}
    # bench-default mode: lazy fusion, non-blocking batched projections, the whole pattern per step
    b = make_gpu_backend(1, width, sel, device=local_rank, batch=args.batch)
    ms, stats = timed_pattern(timer, b, pattern, 2, 1)
    out["batched"] = {"value": len(pattern) * 2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 2, "steps": 2, "warmup": 1,
                      "roofline": roofline_of(stats, ms), "kernels": kernel_table(stats, 2, measured_peak()[0]),
                      "projections_per_pass": sum(1 for c in pattern if c.kind.name == "M") * 2 / max(stats["fused"]["launches"], 1)}
    del b
    free_device_memory()
    # strict mode (reference error timing) on a bounded prefix window: the library default (analytic projections share
    # passes) and batch = 0 (one k_fused pass per projection, every measure() waits for its norm)
    nwin = 2400
    sub = mbqc.Pattern(pattern.input_nodes, list(pattern)[:nwin])
    for label, bt in (("strict_lazy", args.batch), ("strict_unbatched", 0)):
        b = make_gpu_backend(1, width, sel, device=local_rank, nonstrict=False, batch=bt)
        ms, stats = timed_pattern(timer, b, sub, 2, 1)
        out[label] = {"value": nwin * 2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 2, "steps": 2, "warmup": 1,
                      "batch": bt, "window": f"first {nwin} commands", "roofline": roofline_of(stats, ms),
                      "kernels": kernel_table(stats, 2, measured_peak()[0])}
        del b
        free_device_memory()
    if not args.no_cpu:
        n_cmds = 8
        times, kind, cores, note = host_window(name, pattern, n_cmds, 1, 0, force_port=args.port)
        if times:
            out["cpu_baseline"] = {"value": n_cmds / times[0], "unit": UNIT, "cores": cores, "kind": kind,
                                   "sample": f"{n_cmds} consecutive commands of {name} at {width - 1}<->{width} live "
                                             f"qubits, {times[0]:.1f} s of CPU work; {note}"}
    return out


def config4_record(args, timer: Timer, world: int, local_rank: int) -> dict:
    """BASELINE config 4: the QFT pattern of examples/qft_with_tn.py:22-56 at 33 + log2(N) live qubits on N GPUs
    (128 GiB register per GPU at every point).  Timed on a bounded prefix window; then ONE complete run under
    pseudo-random outcomes whose output is known: H on |+...+> gives |0...0>, whose QFT is |+...+>, so P(+) = 1 on every
    output qubit whatever branch is taken, and the norm is 1."""
    from graphix_b200 import PatternSimulator, mbqc
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    g = world.bit_length() - 1
    name = f"config4_qft{32 + g}"
    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    width = pattern.max_space()
    nwin = args.config4_window
    sub = mbqc.Pattern(pattern.input_nodes, list(pattern)[:nwin])
    out = {"workload": name, "n_gpus": world, "live_qubits": f"{width - 1}<->{width}", "state_bytes_max": 16 << width,
           "bytes_per_gpu": (16 << width) // world, "commands_full_pattern": len(pattern), "max_commands": nwin,
           "selector": "ConstBranchSelector(0) (timed window); seeded pseudo-random outcomes (full run)"}
    b = make_gpu_backend(world, width, mbqc.ConstBranchSelector(0), device=local_rank, batch=args.batch, pattern=sub)
    ms, stats = timed_pattern(timer, b, sub, 2, 1, sharded=world > 1)
    out.update({"value": nwin * 2 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / 2, "steps": 2, "warmup": 1,
                "roofline": roofline_of(stats, ms), "kernels": kernel_table(stats, 2, measured_peak()[0])})
    if world > 1:
        out["exchange"] = exchange_record(b.state, 2)
    # complete run + known-answer check
    if not args.no_parity:
        sel = SeededOutcomes(11)
        object.__setattr__(b, "branch_selector", sel)
        if world > 1:
            b.use_schedule(pattern)
        b.state.reset()
        b.node_index.clear()
        timer.barrier()
        sw0 = getattr(b.state, "n_swaps", 0)
        t0 = time.perf_counter()
        sim = PatternSimulator(pattern, b)
        sim.run()
        timer.barrier()
        dt = time.perf_counter() - t0
        plus = np.array([[0.5, 0.5], [0.5, 0.5]], dtype=np.complex128)
        pmin = min(float(np.real(b.state.expectation_single(plus, q))) for q in range(b.state.nqubit))
        norm2 = float(b.state.norm2())
        ones = int(sum(sim.measure_method.results.values()))
        out["full_pattern"] = {"value": len(pattern) / dt, "unit": UNIT, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3,
                               "commands_per_step": len(pattern), "timing": "wall clock between two barrier + synchronize "
                               "pairs (one complete run: host + device)", "selector": "seeded pseudo-random outcomes",
                               "swaps": getattr(b.state, "n_swaps", 0) - sw0}
        out["parity_check"] = {"what": "complete pattern under seeded pseudo-random outcomes: QFT(H|+..+>) = |+..+>",
                               "commands": len(pattern), "seconds": dt, "commands_per_s": len(pattern) / dt,
                               "outcomes_equal_to_1": ones, "output_qubits": b.state.nqubit,
                               "min_P_plus": pmin, "norm2": norm2,
                               "ok": bool(abs(pmin - 1.0) < 1e-9 and abs(norm2 - 1.0) < 1e-9)}
    del b
    free_device_memory()
    return out


def gpu_arm(args) -> None:
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from graphix_b200 import B200StatevectorBackend, PatternSimulator, _lib, mbqc
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    timer = Timer(world)
    name = workload_for(args.gpus, args.workload)
    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    width = pattern.max_space()
    if args.max_commands:
        # bounded prefix window: the first --max-commands commands; whatever is alive then is the output
        pattern = mbqc.Pattern(pattern.input_nodes, list(pattern)[: args.max_commands])
    ncmd = len(pattern)
    if args.selector == "random":
        # graphix's default: RandomBranchSelector(pr_calc=True) — every measurement first asks for its probability
        selector = mbqc.RandomBranchSelector(pr_calc=True)
    else:
        selector = mbqc.ConstBranchSelector(0)
    run_kwargs = (lambda: {"rng": np.random.default_rng(0)}) if args.selector == "random" else (lambda: {})

    # ---- reference-derived parity of the sharded path on the real ranks (before anything is timed) -------------
    shard_par = sharded_parity(world) if (world > 1 and not args.no_parity) else None

    backend = make_gpu_backend(world, width, selector, fusion=args.fusion, nonstrict=args.nonstrict, batch=args.batch,
                               device=local_rank, pattern=pattern, schedule=not args.no_schedule)
    barrier = timer.barrier

    # ---- device-resident arm -------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof = None
    if args.cprofile and rank == 0:
        import cProfile

        prof = cProfile.Profile()
        prof.enable()
    dev_ms, stats = timed_pattern(timer, backend, pattern, args.steps, args.warmup, run_kwargs, sharded=world > 1)
    if prof is not None:
        import pstats

        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(35)
    clocks = sampler.stop() if rank == 0 else {}
    value = ncmd * args.steps / (dev_ms * 1e-3)
    exchange = exchange_record(backend.state, args.steps) if world > 1 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- secondary: the same pattern in the stricter modes and under graphix's default selector, same timing rules --
    #   per_command : fusion = 1 — one kernel per command (runs of E fused), the north_star's per-command design
    #   strict_lazy : fusion = 2, strict — the library default: lazy N/E->M fusion, every projection waits for its
    #                 norm (reference error timing); `value` itself is fusion = 2, non-strict, batched
    #   random_selector : RandomBranchSelector(pr_calc=True), graphix's default — a probability query per measurement
    per_command = strict_lazy = strict_unb = random_sel = random_sel_b = None
    peak = measured_peak()[0]
    if world == 1 and args.fusion == 2 and args.nonstrict and not args.no_per_command and args.selector != "random":
        for label, kw, sel2, rk in (("strict_lazy", {"nonstrict": False, "batch": args.batch}, selector, lambda: {}),
                                    ("strict_unbatched", {"nonstrict": False, "batch": 0}, selector, lambda: {}),
                                    ("random_selector", {"nonstrict": False, "batch": args.batch},
                                     mbqc.RandomBranchSelector(pr_calc=True), lambda: {"rng": np.random.default_rng(0)}),
                                    ("random_selector_batched", {"batch": args.batch}, mbqc.RandomBranchSelector(pr_calc=True),
                                     lambda: {"rng": np.random.default_rng(0)}),
                                    ("per_command", {"fusion": 1}, selector, lambda: {})):
            b2 = make_gpu_backend(1, width, sel2, device=local_rank, **kw)
            nst = max(2, args.steps - 2) if label != "strict_lazy" else args.steps
            ms2, st2 = timed_pattern(timer, b2, pattern, nst, 2, rk)
            rec = {"fusion": kw.get("fusion", 2), "strict": "nonstrict" in kw or "fusion" in kw,
                   "batch": kw.get("batch", 0 if "fusion" in kw else args.batch),
                   "value": ncmd * nst / (ms2 * 1e-3), "unit": UNIT,
                   "ms_per_step": ms2 / nst, "steps": nst, "warmup": 2, "kernels": kernel_table(st2, nst, peak)}
            if label == "strict_lazy":
                rec["note"] = ("library default (B200StatevectorBackend()): zero-norm RuntimeError raised by measure() itself; "
                               "projections whose outcome probability is analytically 1/2 cannot fail and share passes")
                strict_lazy = rec
            elif label == "strict_unbatched":
                rec["note"] = "strict, batch=0: one k_fused pass per projection, every measure() waits for its norm"
                strict_unb = rec
            elif label == "random_selector":
                rec["selector"] = "RandomBranchSelector(pr_calc=True), rng=default_rng(0) (graphix's default selector)"
                rec["note"] = ("everything at its default: strict library mode + graphix's default selector; the probability "
                               "of an XY-plane measurement of a qubit entangled with a fresh |+> is 1/2 analytically, so most "
                               "queries need no pass")
                random_sel = rec
            elif label == "random_selector_batched":
                rec["selector"] = "RandomBranchSelector(pr_calc=True), rng=default_rng(0)"
                rec["note"] = "same selector in the bench-default mode (strict=False, batch=12)"
                random_sel_b = rec
            else:
                per_command = rec
            del b2
            free_device_memory()

    # N > 1: the sharded library default (strict: reference error timing; analytic projections are queued, every other
    # projection waits for its all-reduced norm)
    if world > 1 and args.fusion == 2 and args.nonstrict and not args.no_per_command and args.selector != "random":
        b2 = make_gpu_backend(world, width, selector, nonstrict=False, batch=args.batch, device=local_rank, pattern=pattern,
                              schedule=not args.no_schedule)
        nst = max(2, args.steps - 2)
        ms2, st2 = timed_pattern(timer, b2, pattern, nst, 2, run_kwargs, sharded=True)
        strict_lazy = {"fusion": 2, "strict": True, "batch": args.batch, "value": ncmd * nst / (ms2 * 1e-3), "unit": UNIT,
                       "ms_per_step": ms2 / nst, "steps": nst, "warmup": 2, "kernels": kernel_table(st2, nst, peak),
                       "note": "ShardedStatevectorBackend default (strict=True): zero-norm RuntimeError raised by measure() itself"}
        del b2
        free_device_memory()

    # ---- end-to-end arm: host input vector in, host final state out ------------
    e2e = None
    if world == 1 and not args.no_e2e and len(pattern.input_nodes) <= 28:
        n_in = len(pattern.input_nodes)
        n_out = len(pattern.output_nodes)
        host_in = torch.full((1 << n_in,), 2.0 ** (-n_in / 2), dtype=torch.complex128).pin_memory()
        host_out = torch.empty(1 << n_out, dtype=torch.complex128).pin_memory()
        np_in, np_out = host_in.numpy(), host_out.numpy()

        def e2e_step() -> None:
            backend.state.reset()
            backend.node_index.clear()
            PatternSimulator(pattern, backend).run(input_state=np_in, **run_kwargs())   # H2D of the input state inside
            backend.state.flatten(out=np_out)                                            # D2H of the result inside

        e2e_step()
        e2e_dev = timer.run(e2e_step, args.steps)
        e2e_ms = max(e2e_dev, timer.wall_ms)
        e2e = {"value": ncmd * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": 16 << n_in, "d2h_bytes_per_step": 16 << n_out,
               "ms_per_step": e2e_ms / args.steps}

    if world > 1 and not args.no_e2e and width - (world.bit_length() - 1) <= 30:
        # end to end on N GPUs: every rank uploads ITS slice of the joint input state from pinned host memory (the same
        # |+..+> input as at N = 1, handed over shard by shard) and reads its shard of the final state back
        from graphix_b200.sharded import ShardedInput

        eng = backend.state.engine
        n_in = len(pattern.input_nodes)
        kl_max = min(n_in, eng.capacity)
        host_in = torch.full((1 << kl_max,), 2.0 ** (-n_in / 2), dtype=torch.complex128).pin_memory()
        np_in = host_in.numpy()
        h2d = {"bytes": 0}

        def slice_for(t, kg, kl):
            h2d["bytes"] = 16 << kl
            return np_in[: 1 << kl]

        np_out = None

        def e2e_step_sharded() -> None:
            backend.state.reset()
            backend.node_index.clear()
            PatternSimulator(pattern, backend).run(input_state=ShardedInput(n_in, slice_for))   # H2D inside
            backend.state.check_deferred()
            if np_out is not None:
                eng.sv.flatten(out=np_out)                                                       # D2H inside

        e2e_step_sharded()          # untimed: learn the size of the local shard of the output state
        host_out = torch.empty(1 << eng.nqubit, dtype=torch.complex128).pin_memory()
        np_out = host_out.numpy()
        e2e_step_sharded()
        e2e_dev = timer.run(e2e_step_sharded, args.steps)
        t = torch.tensor([max(e2e_dev, timer.wall_ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": ncmd * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": world * h2d["bytes"],
               "d2h_bytes_per_step": world * (16 << eng.nqubit), "ms_per_step": e2e_ms / args.steps,
               "note": "every rank uploads its slice of the joint |+..+> input vector from pinned host memory and reads "
                       "its shard of the final state back to pinned host memory"}

    # ---- the GPU arm on exactly the command window the CPU baseline is timed on ------------------------------
    same_window = None
    cpu_n = None
    if world == 1 and not args.no_cpu and width <= 31 and args.selector != "random":
        cpu_n = args.cpu_window or (max(6, min(240, int(round(60 * 2.0 ** (27 - width))))) if width >= 22 else 600)
        same_window = gpu_window(backend, pattern, cpu_n, timer)
    del backend
    free_device_memory()

    # ---- BASELINE configs 3, 4 and 5 as sub-records of the same line (the driver only runs the default command) ---
    extras = {}
    default_run = args.workload is None and not args.max_commands and args.fusion == 2 and args.nonstrict and \
        args.selector != "random" and not args.no_extra
    if default_run:
        if world == 1:
            for key, fn in (("config3", lambda: config3_record(args, timer, local_rank)),
                            ("config4", lambda: config4_record(args, timer, world, local_rank)),
                            ("config5", lambda: dm_record(args, "config5_rand12x3", 3, 1))):
                try:
                    extras[key] = fn()
                except Exception as exc:  # noqa: BLE001 - a sub-record must never take the headline line down
                    extras[key] = {"error": f"{type(exc).__name__}: {exc}"}
                free_device_memory()
        else:
            try:
                extras["config4"] = config4_record(args, timer, world, local_rank)
            except Exception as exc:  # noqa: BLE001
                extras["config4"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------
    kinds = {k: v for k, v in stats.items() if v["launches"]}
    roofline = roofline_of(stats, dev_ms)
    roofline["traffic"] = ncu_traffic(roofline["kernel"], name, bool(args.nonstrict and args.batch)) if world == 1 else None

    # ---- CPU baseline: the oracle port on this box's cores, bounded window --------
    cpu_baseline = None
    if cpu_n:
        times, kind, cores, note = host_window(name, pattern, cpu_n, 1, 0, force_port=args.port)
        if times:
            cpu_baseline = {"value": cpu_n / times[0], "unit": UNIT, "cores": cores, "kind": kind,
                            "sample": f"{cpu_n} consecutive commands of {name} at {width - 1}<->{width} live qubits, "
                                      f"{times[0]:.1f} s of CPU work; {note}",
                            "gpu_same_window": same_window}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "commands_per_step": ncmd, "live_qubits": f"{width - 1}<->{width}",
                   "state_bytes_max": 16 << width,
                   "selector": "ConstBranchSelector(0)" if args.selector != "random" else "RandomBranchSelector(pr_calc=True), rng=default_rng(0)",
                   "fusion": args.fusion, "max_commands": args.max_commands or None,
                   "strict_norm_check": not args.nonstrict, "batch": args.batch if args.nonstrict else 0,
                   "mode": "lazy N/E->M fusion, non-blocking batched projections (opt-in: strict=False, batch=12); the "
                           "library default (strict) is reported as `strict_lazy`" if args.nonstrict and args.fusion == 2 else "as flagged",
                   "parallelism": f"shard{world}" if world > 1 else "single",
                   "l2": "state >> 126 MB L2: inputs larger than L2, no flush needed"},
        "e2e": e2e, "gpu_launches": int(sum(v["launches"] for v in kinds.values())),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "kernels": kernel_table(stats, args.steps, peak),
        "projections_per_pass": (sum(1 for c in pattern if c.kind.name == "M") * args.steps / stats["fused"]["launches"])
        if stats.get("fused", {}).get("launches") else None,
        "per_command": per_command, "strict_lazy": strict_lazy, "strict_unbatched": strict_unb,
        "random_selector": random_sel,
        "random_selector_batched": random_sel_b,
    }
    if world > 1:
        line["exchange"] = exchange
        line["sharded_parity"] = shard_par
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def gpu_window(backend, pattern, n_cmds: int, timer: "Timer") -> dict | None:
    """The GPU backend on exactly the command window ``cpu_window`` times (same setup: inputs prepared, pattern advanced
    until the live width first reaches its maximum, then ``n_cmds`` consecutive commands)."""
    import torch

    from graphix_b200.simulator import DefaultMeasureMethod

    cmds = list(pattern)
    backend.state.reset()
    backend.node_index.clear()
    from graphix_b200 import mbqc

    backend.add_nodes(pattern.input_nodes, mbqc.BasicStates.PLUS)
    mm = DefaultMeasureMethod()

    def run(cmd) -> None:
        k = cmd.kind.name
        if k == "N":
            backend.add_nodes([cmd.node], cmd.state)
        elif k == "E":
            backend.entangle_nodes(cmd.nodes)
        elif k == "M":
            mm.measure(backend, cmd)
        elif k in ("X", "Z"):
            if mm.check_domain(cmd.domain):
                backend.correct_byproduct(cmd)
        elif k == "C":
            backend.apply_clifford(cmd.node, cmd.clifford)

    pos = 0
    while backend.nqubit < pattern.max_space() and pos < len(cmds):
        run(cmds[pos])
        pos += 1
    if pos + n_cmds > len(cmds):
        return None
    backend.state.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timer.barrier()
    t0 = time.perf_counter()
    e0.record()
    for c in cmds[pos:pos + n_cmds]:
        run(c)
    backend.state.sync()          # queued projections of the window are executed and checked inside the timed region
    e1.record()
    timer.barrier()
    ms = max(e0.elapsed_time(e1), 0.0)
    wall = (time.perf_counter() - t0) * 1e3
    return {"value": n_cmds / (max(ms, 1e-6) * 1e-3), "unit": UNIT, "ms": ms, "wall_ms": wall,
            "note": "same commands, same widths, state resident; one window (no warm-up of this window)"}


# ---------------------------------------------------------------------------
# density-matrix arm (BASELINE config 5): --workload config5_rand12x3 (or any dm_* fixture)
# ---------------------------------------------------------------------------
DM_NOISE = dict(prepare_error_prob=0.01, x_error_prob=0.01, z_error_prob=0.01, entanglement_error_prob=0.01,
                measure_channel_prob=0.01, measure_error_prob=0.0)


def dm_cpu_window(pattern, max_ops: int, budget_s: float = 30.0):
    """Oracle DM backend (numpy, out of place, one evolve per Kraus operator like the reference) on the first
    commands of the noisy sequence; stops after ``max_ops`` pattern commands or ``budget_s`` seconds."""
    from graphix_b200 import mbqc
    from graphix_b200.noise import DepolarisingNoiseModel
    from graphix_b200.simulator import DefaultMeasureMethod
    from oracle.dm_oracle import OracleDMBackend

    nm = DepolarisingNoiseModel(**DM_NOISE)
    rng = np.random.default_rng(7)
    backend = OracleDMBackend(branch_selector=mbqc.ConstBranchSelector(0))
    backend.add_nodes(pattern.input_nodes, mbqc.BasicStates.PLUS)
    mm = DefaultMeasureMethod()
    done_n = False
    # untimed setup: noise channels on the input qubits
    for cmd in nm.input_nodes(pattern.input_nodes):
        backend.apply_noise(cmd)
    seq = nm.transpile(pattern)
    done = 0
    t0 = time.perf_counter()
    for cmd in seq:
        k = cmd.kind.name
        if k == "N":
            backend.add_nodes([cmd.node], cmd.state)
        elif k == "E":
            backend.entangle_nodes(cmd.nodes)
        elif k == "M":
            mm.measure(backend, cmd, noise_model=nm, rng=rng)
        elif k in ("X", "Z"):
            if mm.check_domain(cmd.domain):
                backend.correct_byproduct(cmd)
        elif k == "C":
            backend.apply_clifford(cmd.node, cmd.clifford)
        elif k == "ApplyNoise":
            if cmd.domain is None or mm.check_domain(cmd.domain):
                backend.apply_noise(cmd)
        if k == "M" or (k in ("X", "Z", "C")):
            done += 1          # a pattern command is complete once its trailing / leading noise has been applied
        elif k == "ApplyNoise" and cmd.noise.nqubits == 2:
            done += 1          # E + its two-qubit channel
        elif k == "ApplyNoise" and done_n:
            done += 1          # N + its preparation channel
            done_n = False
        if k == "N":
            done_n = True
        if done >= max_ops or time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0, backend.nqubit


def dm_record(args, name: str, steps: int, warmup: int) -> dict:
    """BASELINE config 5 (noisy pattern on the density-matrix backend, Kraus-channel kernels) as one JSON record:
    device-resident commands/s, roofline of the dominant kernel, end-to-end (final rho read back to pinned host memory
    inside the timed region) and the numpy oracle port over a bounded window."""
    import torch

    from graphix_b200 import B200DensityMatrixBackend, PatternSimulator, _lib, mbqc
    from graphix_b200.noise import DepolarisingNoiseModel
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    ncmd = len(pattern)
    width = pattern.max_space()
    dev = torch.cuda.current_device()
    backend = B200DensityMatrixBackend.with_capacity(width, branch_selector=mbqc.ConstBranchSelector(0), device=dev)
    nm = DepolarisingNoiseModel(**DM_NOISE)

    def one_step() -> None:
        backend.state.reset()
        backend.node_index.clear()
        PatternSimulator(pattern, backend, noise_model=nm).run(rng=np.random.default_rng(7))

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize()
    backend.state.reset_stats()
    backend.state.set_option(_lib.OPT_PROFILE, 1)
    sampler = ClockSampler(dev)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    stats = {k: v for k, v in backend.state.stats().items() if v["launches"]}
    backend.state.set_option(_lib.OPT_PROFILE, 0)
    # end to end: the noisy run plus the read-back of the final density matrix to host memory
    n_out = len(pattern.output_nodes)
    one_step()
    _ = backend.state.rho
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
        rho = backend.state.rho
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tr = float(np.trace(rho).real)
    cpu = None
    if not args.no_cpu:
        done, dt, nq = dm_cpu_window(pattern, args.cpu_window or 1, budget_s=15.0)
        cpu = {"value": done / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {done} commands of {name} (+ noise channels) in {dt:.1f} s on the numpy density-matrix "
                         f"oracle, register grown to {nq} of {width} qubits"}
    rec = {
        "metric": METRIC, "value": ncmd * steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "backend": "densitymatrix", "commands_per_step": ncmd, "live_qubits": width,
                   "rho_bytes_max": 16 << (2 * width), "noise": DM_NOISE, "selector": "ConstBranchSelector(0)",
                   "rng": "default_rng(7)", "l2": "rho >> 126 MB L2 at full width"},
        "e2e": {"value": ncmd * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 << (2 * n_out),
                "ms_per_step": e2e_s * 1e3 / steps, "final_trace": tr,
                "note": "inputs are |+> product states (no upload); the final rho is read back to host memory every step"},
        "gpu_launches": int(sum(v["launches"] for v in stats.values())),
        "roofline": roofline_of(stats, dev_ms), "cpu_baseline": cpu, "clocks": clocks,
        "kernels": kernel_table(stats, steps, measured_peak()[0]),
    }
    del backend
    return rec


def dm_arm(args) -> None:
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    name = args.workload
    if args.impl == "reference":
        pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
        done, dt, nq = dm_cpu_window(pattern, args.cpu_window or 10 ** 9, budget_s=60.0)
        value = done / dt
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1,
                          "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "c128", "data": "synthetic",
                          "config": {"workload": name, "backend": "densitymatrix", "noise": DM_NOISE},
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                           "sample": f"first {done} commands of {name} (+ their noise channels), {dt:.1f} s, "
                                                     f"numpy density-matrix oracle, reached {nq} qubits"},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return
    import torch

    torch.cuda.set_device(0)
    print(json.dumps(dm_record(args, name, args.steps, args.warmup)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="pattern fixture name under tests/golden/patterns")
    ap.add_argument("--fusion", type=int, default=2,
                    help="2 = lazy N/E->M fusion (default), 1 = one kernel per command with runs of E fused, 0 = none")
    ap.add_argument("--max-commands", type=int, default=0, help="simulate only the first K commands of the pattern")
    ap.add_argument("--no-per-command", action="store_true", help="skip the secondary fusion=1 measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (it is skipped anyway above 28 input qubits)")
    ap.add_argument("--cprofile", action="store_true", help="host-side cProfile of the timed region (rank 0, stderr)")
    ap.add_argument("--strict", dest="nonstrict", action="store_false",
                    help="wait for the norm of every projection like the reference (default: non-blocking projections, "
                         "zero-norm check deferred to finalize)")
    ap.add_argument("--selector", default="const0", choices=["const0", "random"],
                    help="const0 = ConstBranchSelector(0) (BASELINE configs); random = graphix's default selector")
    ap.add_argument("--no-schedule", action="store_true",
                    help="N > 1: do not hand the pattern's measurement order to the sharded backend (eviction heuristic only)")
    ap.add_argument("--batch", type=int, default=12, help="non-strict: fused projections per pass, up to 12 (0 = one pass each)")
    ap.set_defaults(nonstrict=True)
    ap.add_argument("--cpu-window", type=int, default=0, help="commands per CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--port", action="store_true",
                    help="CPU legs: time the C/OpenMP port of the oracle even when the real reference is importable")
    ap.add_argument("--no-extra", action="store_true",
                    help="default run only: skip the config3 / config4 / config5 sub-records")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the in-run parity checks (sharded golden traces, complete QFT run of config 4)")
    ap.add_argument("--config4-window", type=int, default=900, help="commands of the timed config 4 prefix window")
    args = ap.parse_args()
    if args.workload and args.workload.startswith(("config5", "dm_")):
        dm_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
