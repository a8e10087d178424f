"""bench.py prints ONE JSON line with the keys the driver reads: the CPU reference arm (runs
anywhere) and, on a GPU, this repo's arm at a reduced size (same code path, extras included)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # stdout carries the JSON line and nothing else
    return json.loads(lines[0])


def test_reference_arm_contract():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--horizon", "20")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "trajectory-iterations/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("C4 quadruped")


@pytest.mark.gpu
def test_b200_arm_contract_small():
    d = run_bench("--steps", "3", "--warmup", "3", "--batch", "64", "--horizon", "40")
    assert BASE <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["active_per_step"] == 64.0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["strong"]["global_batch"] == 1024 and set(d["other_configs"]) >= {"C2_acrobot_N40_B50"}
    assert d["cost_vs_oracle"]["rel_err"] <= 1e-5          # north-star tolerance on the converged cost
