"""`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on small workloads: JSON contract."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args: str) -> dict:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_statevector_json_contract():
    d = _run("--workload", "rand11x6", "--steps", "2", "--warmup", "1", "--cpu-window", "40")
    assert d["impl"] == "reference" and d["metric"] == "pattern_sim_commands_per_s" and d["unit"] == "commands/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    # the real reference (graphix numba backend) when importable here, else the C port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "commands/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"] == "rand11x6"


def test_reference_arm_port_and_real_reference_agree_on_kind_labels():
    port = _run("--workload", "rand11x6", "--steps", "1", "--warmup", "0", "--cpu-window", "40", "--port")
    assert port["cpu_baseline"]["kind"] == "port"
    if os.path.isdir("/root/reference/graphix") or os.path.exists(os.path.join(ROOT, "oracle", "_ref", "graphix_ref.zip")):
        real = _run("--workload", "rand11x6", "--steps", "1", "--warmup", "0", "--cpu-window", "40")
        assert real["cpu_baseline"]["kind"] == "reference" and "numba" in real["cpu_baseline"]["sample"]


def test_reference_arm_density_matrix():
    d = _run("--workload", "dm_rand5x3", "--cpu-window", "30")
    assert d["impl"] == "reference" and d["value"] > 0 and d["config"]["backend"] == "densitymatrix"


def test_reference_arm_refuses_widths_beyond_the_reference_limit():
    d = _run("--workload", "config4_qft35")
    assert d["impl"] == "reference" and "unavailable" in d and "31" in d["unavailable"]
