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
_NCU_KERNEL = {"fused": "k_fused_batch", "contract": "k_contract", "fill": "k_fill", "cz": "k_cz_star"}


def ncu_traffic(kind: str, workload: str, batched: bool) -> float | None:
    """DRAM bytes per launch of the dominant kernel from the ncu --set full capture committed under profiles/
    (taken on config 2; not transferable to other workloads, hence None there)."""
    if workload != "config2_rand26x20":
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
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
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    width = pattern.max_space()
    n_cmds = args.cpu_window or max(3, min(30, int(round(12 * 2.0 ** (27 - width)))) if width >= 24 else 300)
    times = cpu_window(pattern, n_cmds, args.steps, args.warmup)
    if not times:
        raise SystemExit("pattern too short for the requested window")
    per_step = float(np.mean(times))
    value = n_cmds / per_step
    sample = f"{n_cmds} consecutive commands of {name} per step at {width - 1}<->{width} live qubits (bounded window)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "selector": "ConstBranchSelector(0)", "commands_per_step": n_cmds},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
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
    run_rng = (lambda: np.random.default_rng(0)) if args.selector == "random" else (lambda: None)

    if world > 1:
        from graphix_b200.sharded import ShardedStatevectorBackend

        def make_backend():
            # non-blocking projections: norms are all-reduced on the stream, the zero-norm check is deferred to finalize
            return ShardedStatevectorBackend.with_capacity(width, branch_selector=selector, fusion=args.fusion,
                                                           strict=not args.nonstrict,
                                                           pattern=None if args.no_schedule else pattern)
    else:
        def make_backend():
            # default: lazy fusion, non-blocking projections, up to --batch projections per pass
            lazy = args.fusion == 2 and args.nonstrict
            return B200StatevectorBackend.with_capacity(width, branch_selector=selector, fusion=args.fusion,
                                                        device=local_rank, strict=not lazy,
                                                        batch=args.batch if lazy else 0)

    backend = make_backend()

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(profile: bool = False) -> None:
        backend.state.reset()
        backend.node_index.clear()
        PatternSimulator(pattern, backend).run(rng=run_rng())

    # ---- device-resident arm -------------------------------------------------
    for _ in range(args.warmup):
        one_step()
    barrier()
    backend.state.reset_stats()
    backend.state.set_option(_lib.OPT_PROFILE, 1)
    if world > 1:
        backend.state.profile = True
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = None
    if args.cprofile and rank == 0:
        import cProfile

        prof = cProfile.Profile()
        prof.enable()
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    if prof is not None:
        import pstats

        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(35)
    clocks = sampler.stop() if rank == 0 else {}
    dev_ms = e0.elapsed_time(e1)
    stats = backend.state.stats()
    backend.state.set_option(_lib.OPT_PROFILE, 0)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = ncmd * args.steps / (dev_ms * 1e-3)
    exchange = None
    if world > 1:
        st = backend.state
        st.collect_profile()
        exchange = {"swaps_per_step": st.n_swaps / args.steps,
                    "nvlink_bytes_per_rank_per_step": st.exchanged_bytes / args.steps,
                    "device_ms_per_step": st.exchange_ms / args.steps,
                    "host_ms_per_step": st.exchange_host_ms / args.steps,
                    "host_wait_for_device_ms_per_step": st.sync_wait_ms / args.steps,
                    "GBps_per_direction": (st.exchanged_bytes / 1e9) / (st.exchange_ms * 1e-3) if st.exchange_ms else None}
        st.profile = False

    # ---- secondary: the same pattern in the two stricter modes, same timing rules ----------
    #   per_command : fusion = 1 — one kernel per command (runs of E fused), the north_star's per-command design
    #   strict_lazy : fusion = 2, strict — lazy N/E->M fusion, every projection waits for its norm (reference
    #                 error timing); `value` itself is fusion = 2, non-strict, batched
    per_command = None
    strict_lazy = None
    if world == 1 and args.fusion == 2 and args.nonstrict and not args.no_per_command:
        b2 = B200StatevectorBackend.with_capacity(width, branch_selector=selector, fusion=2, device=local_rank)
        for _ in range(2):
            b2.state.reset(); b2.node_index.clear(); PatternSimulator(pattern, b2).run()
        b2.state.reset_stats()
        b2.state.set_option(_lib.OPT_PROFILE, 1)
        barrier()
        e0.record()
        for _ in range(args.steps):
            b2.state.reset(); b2.node_index.clear(); PatternSimulator(pattern, b2).run()
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        st2 = {k: v for k, v in b2.state.stats().items() if v["launches"] and v["device_ms"] > 0}
        peak2, _ = measured_peak()
        strict_lazy = {"fusion": 2, "strict": True, "value": ncmd * args.steps / (ms2 * 1e-3), "unit": UNIT,
                       "ms_per_step": ms2 / args.steps,
                       "kernels": {k: {"launches_per_step": v["launches"] / args.steps,
                                       "ms_per_launch": v["device_ms"] / v["launches"],
                                       "frac_of_measured_peak": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 / peak2}
                                   for k, v in st2.items()}}
        del b2
    if world == 1 and args.fusion == 2 and not args.no_per_command:
        b1 = B200StatevectorBackend.with_capacity(width, branch_selector=selector, fusion=1, device=local_rank)

        def step1() -> None:
            b1.state.reset()
            b1.node_index.clear()
            PatternSimulator(pattern, b1).run()

        for _ in range(max(1, args.warmup - 1)):
            step1()
        b1.state.reset_stats()
        b1.state.set_option(_lib.OPT_PROFILE, 1)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step1()
        e1.record()
        barrier()
        ms1 = e0.elapsed_time(e1)
        st1 = {k: v for k, v in b1.state.stats().items() if v["launches"]}
        peak1, _ = measured_peak()
        per_command = {
            "fusion": 1, "value": ncmd * args.steps / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1 / args.steps,
            "kernels": {k: {"launches_per_step": v["launches"] / args.steps, "ms_per_launch": v["device_ms"] / v["launches"],
                            "GBps": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9,
                            "frac_of_measured_peak": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 / peak1}
                        for k, v in st1.items() if v["device_ms"] > 0},
        }
        del b1

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
            PatternSimulator(pattern, backend).run(input_state=np_in)   # H2D of the input state inside
            backend.state.flatten(out=np_out)                             # D2H of the result inside

        e2e_step()
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max(e0.elapsed_time(e1), wall * 1e3)
        e2e = {"value": ncmd * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": 16 << n_in, "d2h_bytes_per_step": 16 << n_out,
               "ms_per_step": e2e_ms / args.steps}

    if world > 1 and not args.no_e2e and width - (world.bit_length() - 1) <= 30:
        # end to end on N GPUs: the inputs are product states (nothing to upload), the result of every step is
        # delivered to pinned HOST memory: every rank reads back its shard of the final state
        eng = backend.state.engine
        np_out = None

        def e2e_step_sharded() -> None:
            backend.state.reset()
            backend.node_index.clear()
            PatternSimulator(pattern, backend).run()
            backend.state.check_deferred()
            if np_out is not None:
                eng.sv.flatten(out=np_out)

        e2e_step_sharded()          # untimed: learn the size of the local shard of the output state
        host_out = torch.empty(1 << eng.nqubit, dtype=torch.complex128).pin_memory()
        np_out = host_out.numpy()
        e2e_step_sharded()
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step_sharded()
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
        e2e = {"value": ncmd * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": world * (16 << eng.nqubit), "ms_per_step": e2e_ms / args.steps,
               "note": "inputs are |+> product states (no upload); every rank reads its shard of the final state back"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------
    peak, peak_src = measured_peak()
    kinds = {k: v for k, v in stats.items() if v["launches"]}
    dom = max(kinds, key=lambda k: kinds[k]["device_ms"])
    d = kinds[dom]
    achieved = d["algo_bytes"] / (d["device_ms"] * 1e-3) / 1e9
    total_kernel_ms = sum(v["device_ms"] for v in kinds.values())
    total_bytes = sum(v["algo_bytes"] for v in kinds.values())
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "frac_of_nominal_8TBps": achieved / NOMINAL_HBM_GBS,
        "traffic": ncu_traffic(dom, name, bool(args.nonstrict and args.batch)) if world == 1 else None,
        "algo_bytes_per_launch": d["algo_bytes"] / d["launches"], "avg_launch_ms": d["device_ms"] / d["launches"],
        "share_of_kernel_time": d["device_ms"] / total_kernel_ms,
        "all_kernels_achieved": total_bytes / (total_kernel_ms * 1e-3) / 1e9,
        "all_kernels_frac": total_bytes / (total_kernel_ms * 1e-3) / 1e9 / peak,
        "kernel_time_share_of_step": total_kernel_ms / dev_ms,
    }
    per_kind = {k: {"launches_per_step": v["launches"] / args.steps, "ms_per_launch": v["device_ms"] / v["launches"],
                    "GBps": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 if v["device_ms"] else None}
                for k, v in kinds.items()}

    # ---- CPU baseline: the oracle port on this box's cores, bounded window --------
    cpu_baseline = None
    if world == 1 and not args.no_cpu and width <= 31:
        cores = os.cpu_count() or 1
        n_cmds = args.cpu_window or (max(6, min(240, int(round(60 * 2.0 ** (27 - width))))) if width >= 22 else 600)
        times = cpu_window(pattern, n_cmds, 1, 0)
        if times:
            cpu_baseline = {"value": n_cmds / times[0], "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n_cmds} consecutive commands of {name} at {width - 1}<->{width} live qubits, "
                                      f"{times[0]:.1f} s of CPU work"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "commands_per_step": ncmd, "live_qubits": f"{width - 1}<->{width}",
                   "state_bytes_max": 16 << (width if world == 1 else width),
                   "selector": "ConstBranchSelector(0)" if args.selector != "random" else "RandomBranchSelector(pr_calc=True), rng=default_rng(0)",
                   "fusion": args.fusion, "max_commands": args.max_commands or None,
                   "strict_norm_check": not args.nonstrict, "batch": args.batch if args.nonstrict else 0,
                   "parallelism": f"shard{world}" if world > 1 else "single",
                   "l2": "state >> 126 MB L2: inputs larger than L2, no flush needed"},
        "e2e": e2e, "gpu_launches": int(sum(v["launches"] for v in kinds.values())),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "kernels": per_kind,
        "per_command": per_command, "strict_lazy": strict_lazy,
    }
    if world > 1:
        line["exchange"] = exchange
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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


def dm_arm(args) -> None:
    import torch

    from graphix_b200 import B200DensityMatrixBackend, PatternSimulator, _lib, mbqc
    from graphix_b200.noise import DepolarisingNoiseModel
    from graphix_b200.pattern_io import GOLDEN_DIR, load_pattern

    name = args.workload
    pattern = load_pattern(os.path.join(GOLDEN_DIR, "patterns", name + ".json.gz"))
    ncmd = len(pattern)
    width = pattern.max_space()
    if args.impl == "reference":
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
    torch.cuda.set_device(0)
    backend = B200DensityMatrixBackend.with_capacity(width, branch_selector=mbqc.ConstBranchSelector(0), device=0)
    nm = DepolarisingNoiseModel(**DM_NOISE)

    def one_step() -> None:
        backend.state.reset()
        backend.node_index.clear()
        PatternSimulator(pattern, backend, noise_model=nm).run(rng=np.random.default_rng(7))

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    backend.state.reset_stats()
    backend.state.set_option(_lib.OPT_PROFILE, 1)
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    stats = {k: v for k, v in backend.state.stats().items() if v["launches"]}
    peak, peak_src = measured_peak()
    dom = max(stats, key=lambda k: stats[k]["device_ms"])
    d = stats[dom]
    total_ms = sum(v["device_ms"] for v in stats.values())
    total_bytes = sum(v["algo_bytes"] for v in stats.values())
    cpu = None
    if not args.no_cpu:
        done, dt, nq = dm_cpu_window(pattern, args.cpu_window or 10 ** 9, budget_s=30.0)
        cpu = {"value": done / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {done} commands of {name} (+ noise channels) in {dt:.1f} s on the numpy density-matrix "
                         f"oracle, register grown to {nq} of {width} qubits"}
    line = {
        "metric": METRIC, "value": ncmd * args.steps / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": name, "backend": "densitymatrix", "commands_per_step": ncmd, "live_qubits": width,
                   "rho_bytes_max": 16 << (2 * width), "noise": DM_NOISE, "selector": "ConstBranchSelector(0)",
                   "rng": "default_rng(7)", "l2": "rho >> 126 MB L2 at full width"},
        "e2e": None, "gpu_launches": int(sum(v["launches"] for v in stats.values())),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": d["algo_bytes"] / (d["device_ms"] * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": d["algo_bytes"] / (d["device_ms"] * 1e-3) / 1e9 / peak,
                     "peak_source": peak_src, "traffic": None, "share_of_kernel_time": d["device_ms"] / total_ms,
                     "all_kernels_frac": total_bytes / (total_ms * 1e-3) / 1e9 / peak,
                     "kernel_time_share_of_step": total_ms / dev_ms},
        "cpu_baseline": cpu, "clocks": clocks,
        "kernels": {k: {"launches_per_step": v["launches"] / args.steps, "ms_per_launch": v["device_ms"] / v["launches"],
                        "GBps": v["algo_bytes"] / (v["device_ms"] * 1e-3) / 1e9 if v["device_ms"] else None}
                    for k, v in stats.items()},
    }
    print(json.dumps(line))


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
    ap.add_argument("--batch", type=int, default=8, help="N = 1, non-strict: fused projections per pass (0 = one pass each)")
    ap.set_defaults(nonstrict=True)
    ap.add_argument("--cpu-window", type=int, default=0, help="commands per CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.workload and args.workload.startswith(("config5", "dm_")):
        dm_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
