"""bench.py prints ONE JSON line with the keys the driver reads, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = run_bench("--impl", "reference", "--constraints", "300", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is False and d["unit"] == "ms"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and "workload" in d["config"]


@pytest.mark.gpu
def test_product_arm_json_line():
    d = run_bench("--constraints", "20000", "--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert (BASE_KEYS | {"clocks", "roofline", "phases_ms"}) <= set(d)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 20002 * 32 and d["e2e"]["d2h_bytes_per_step"] == 576
    assert d["gpu_launches"] > 20 and d["higher_is_better"] is False and d["scaling"] == "strong" and d["n_gpus"] == 1
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf) and rf["bound"] == "int_pipe" and 0 < rf["frac"] < 1.2 and 0 < rf["hbm"]["frac"] < 1.2
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["verified"]["ok"] is True and d["verified"]["library_pairing"] is True
