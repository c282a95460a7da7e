"""bench.py prints ONE JSON line on stdout with the keys the driver reads: the reference arm on CPU, the product arm on a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
             "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must hold exactly one line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "Mtris/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2") and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_product_arm_line():
    d = _run(["--steps", "5", "--warmup", "3", "--no-ref-kernels"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "stage_ms"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 1000 and d["gpu_launches"] >= 4 * 5
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 28048032 and d["e2e"]["d2h_bytes_per_step"] == 8294400 and d["e2e"]["frame_nonzero"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert "sm_mhz" in d["clocks"] and "workload" in d["config"] and "binning" in d["notes"]
    # value is the metric SURVEY.md 8(d) defines: T / (sum of the four stage intervals), one frame in flight
    assert abs(d["value"] - 1.0 / d["device_frame_ms_median"] * 1e3) / d["value"] < 1e-6 and d["value_unbroken_chain"] > 0 and d["value_two_in_flight"] > 0
    assert d["enqueue_ms_per_step"] > 0
    # both arms describe the workload with the same `config`
    ref = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert ref["config"] == d["config"]
