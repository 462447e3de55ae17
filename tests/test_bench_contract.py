"""CPU: the bench line's JSON contract - on the committed line of the round (profiles/r02_bench.json, written by bench.py on a B200)
and on a live run of the reference arm (``bench.py --impl reference``, the CPU port on the smallest rung of its ladder)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config")


def _check_common(d):
    for k in BASE_KEYS:
        assert k in d, k
    assert d["unit"] == "frequencies/s" and d["higher_is_better"] is True and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["scaling"] == "strong" and d["vs_baseline"] is None          # BASELINE.md publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"] or base["metric"] in d["metric"] or d["metric"] in base["metric"]
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k


def test_committed_bench_line_keeps_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench.json")))
    _check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 1e8 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or 0.5 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.5
    dom = max(r["iteration_kernels"], key=lambda k: k["us_per_launch"])
    assert dom["id"] == r["id"]                                          # `roofline` is the dominant kernel of an iteration
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1
    cl = d["clocks"]
    assert cl["sm_mhz"] > 0 and cl["sm_max_mhz"] >= cl["sm_mhz"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    s = d["solver"]
    assert s["unconverged"] == 0 and s["max_relres"] <= 1e-8


def test_reference_arm_runs_and_keeps_the_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sizes", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] == d["value"] and "sample" in c
