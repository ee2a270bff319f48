"""bench.py prints ONE JSON line with the keys the driver reads (task contract): the reference arm runs on the host
cores (no GPU); our arm is run small on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "cpu_baseline", "e2e"}


def _bench(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"].startswith("simulated fish-passages/sec") and d["unit"] == "fish-passages/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "Stryke", "stryke.py"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line_small_workload():
    d = _bench("--steps", "2", "--warmup", "3", "--fish", "200000", "--iterations", "64", "--days", "2", "--no-cpu", "--no-api",
               "--also", "")
    assert COMMON | {"roofline", "gpu_launches", "clocks"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["fish_passages_per_step"] == 200000 * 64 * 2
    assert abs(d["value"] - d["config"]["fish_passages_per_step"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "issue" and r["unit"] == "Tlane-op/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.0 < r["frac"] < 1.0
    assert 150 < r["op_budget_per_fish"] < 260 and r["op_budget_terms"]["philox_blocks"] == 1      # C2: one block per fish
    assert 0.0 < r["kernel_share_of_step"] <= 1.0 and r["kernel_ms_per_step"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 2 * 2        # count + passage per step
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


@pytest.mark.gpu
def test_our_arm_reports_population_mode_and_cascade_in_the_same_line():
    d = _bench("--steps", "2", "--warmup", "3", "--fish", "100000", "--iterations", "32", "--days", "1", "--no-cpu", "--no-api")
    w = d["workloads"]
    assert set(w) == {"c3", "c5"}
    assert w["c3"]["scaling"] == "strong" and w["c5"]["scaling"] == "weak"
    assert w["c3"]["config"]["cells_per_gpu"] == 20 * 365 * 1000 * 5
    assert 1.5e9 < w["c3"]["config"]["fish_passages_per_step"] < 1.0e10
    assert w["c5"]["config"]["fish_passages_per_step"] == 1_250_000 * 1000
    for k in ("c3", "c5"):
        assert w[k]["value"] > 0 and w[k]["e2e"]["value"] > 0 and 0 < w[k]["roofline"]["frac"] < 1
