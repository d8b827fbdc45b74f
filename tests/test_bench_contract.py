"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for our arm and for the reference arm.
Run at a reduced size here (the contract run is `python bench.py` with its defaults)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


@pytest.mark.gpu
def test_our_arm_prints_the_contract_line():
    d = run_bench("--ion-pairs", "30000", "--steps", "6", "--warmup", "3", "--cpu-seconds", "1", "--e2e-steps", "2")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["metric"] == "integrator_particle_updates_per_s" and d["unit"] == "particle-updates/s"
    assert d["n_gpus"] == 1 and d["steps"] == 6 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_launch"] in (88 * 30000 * 37, 128 * 30000 * 37)
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                       # host buffers: transfers are inside the timed region
    assert d["gpu_launches"] == 2 * 6                    # two kernels per step at this size, nothing else
    assert abs(d["value"] - 30000 * 37 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_arm_prints_the_contract_line():
    """runs on the host cores only (no GPU needed): the reference's kernel sources compiled for the CPU"""
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-seconds", "3")
    assert d["impl"] == "reference" and d["metric"] == "integrator_particle_updates_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["higher_is_better"] is True
