"""bench.py's output contract: the JSON line of the B200 arm (checked on the line recorded from the last GPU run,
profiles/) and of the reference arm (run here, on the CPU, on the small workload)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _check_common(line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["metric"] == "icp_registrations_per_s" and line["unit"] == "registrations/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["data"] == "synthetic"
    assert line["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    assert "workload" in line["config"] and "model" not in line["config"]
    e = line["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port")


def test_recorded_b200_line_has_every_contract_key():
    line = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_fullres_warp.json")))
    _check_common(line)
    assert line["dtype"] == "f32" and line["n_gpus"] == 1 and line["warmup"] >= 3
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.05
    assert line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    c = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    assert "l2" in line["config"]


def test_reference_arm_runs_on_the_cpu_and_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "10k",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    _check_common(line)
    assert line["impl"] == "reference" and line["gpu_launches"] == 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"] > 0
    assert line["cpu_baseline"]["cores"] >= 1
