"""bench.py's JSON contract (both arms).  The reference arm runs here on CPU with a shrunken sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"SIFT_BENCH_CPU_TILE": "48"})
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "Mpixel/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


@pytest.mark.gpu
def test_own_arm_line():
    d = _run("--steps", "2", "--warmup", "3", env={"SIFT_BENCH_CPU_TILE": "64"})
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["gpu_launches"] > 0 and d["value"] > 100 and d["e2e"]["value"] > 100
    assert d["e2e"]["h2d_bytes_per_step"] == d["config"]["frames_per_gpu_per_step"] * 1920 * 1080
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and set(r["kernels"]) == {"blur_octave0", "blur_octave1", "blur_high_octaves", "scan", "refine"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["clocks"]["sm_mhz"] and d["clocks"]["samples"] >= 1
