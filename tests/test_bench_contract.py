"""bench.py prints exactly one JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def _one_json_line(cmd):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _one_json_line(["--impl", "reference", "--cpu-duals", "256", "--cpu-obs", "1024", "--rv", "16", "--steps", "2", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0 and "workload" in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line_reduced_size():
    d = _one_json_line(["--duals", "4096", "--obs-per-gpu", "16384", "--rv", "32", "--steps", "5", "--warmup", "3", "--cpu-duals", "256", "--cpu-obs", "1024",
                        "--sd-iterations", "40"])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "sd_iterations_ssn"} <= set(d)
    it = d["sd_iterations_ssn"]
    assert "unavailable" not in it, it
    assert it["iterations"] == 40 and it["gpu_tables_it_per_s"] > 0 and it["cpu_tables_it_per_s"] > 0 and it["same_incumbent_estimate"] is True
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["achieved"] > 0
    assert d["gpu_launches"] == 3 * d["steps"]                       # prep, sweep, merge per cut
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.05
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["dtype"] == "f64" and d["scaling"] == "weak"
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
