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
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "mixed(f32 pos / f64 vel)" and d["data"] == "synthetic"
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
    # round-2 keys: the other flows, config 3, both force widths, parity against the reference's CUDA kernels, the fixed
    # boxes of the strong-scaling block, the resident end-to-end number and the (blocked) full step
    assert d["parity"]["max_rel"] <= 1e-6 and d["parity"]["particles"] == 30000 * 37
    assert [f["bytes_per_particle"] for f in d["flows"]] == [336, 440, 32, 64] and all(0 < f["frac"] < 1 for f in d["flows"])
    assert d["config3_edl"]["particles"] == 40310 and d["config3_edl"]["launches_per_step"] == 1
    assert 0 <= d["force_sigma_1000"]["hardwall_fire_fraction_per_step_at_sigma_1"] <= d["force_sigma_1000"]["hardwall_fire_fraction_per_step"] <= 1
    assert d["strong"] and all(s["value"] > 0 for s in d["strong"])
    assert d["e2e_resident"]["value"] > d["e2e"]["value"]
    assert d["full_step"]["status"].startswith("blocked") and d["full_step"]["toy_force_full_step"]["ns_per_day"] > 0


def test_both_arms_describe_the_same_config():
    """the driver compares the two arms' `config` dicts key by key"""
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 8):
        for scaling in ("weak", "strong"):
            args = bench.argparse.Namespace(ion_pairs=442368, total_ion_pairs=1769472, scaling=scaling, precision="mixed")
            a, b = bench.config_dict(args, world), bench.config_dict(args, world)
            assert a == b and a["particles_total"] == a["particles_per_gpu"] * world and "model" not in a


def test_reference_arm_prints_the_contract_line():
    """runs on the host cores only (no GPU needed): the reference's kernel sources compiled for the CPU"""
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-seconds", "3")
    assert d["dtype"] == "mixed(f32 pos / f64 vel)" and d["scaling"] == "weak"
    assert set(d["config"]) == {"workload", "precision", "force_sigma", "particles_per_gpu", "particles_total", "l2", "parallelism"}
    assert d["impl"] == "reference" and d["metric"] == "integrator_particle_updates_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["higher_is_better"] is True
