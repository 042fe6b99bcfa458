"""bench.py's JSON contract, checked on CPU through the reference arm (the product arm needs a GPU) and through the
pure helpers the product arm is made of."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "64", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GStencil/s" and d["unit"] == "GStencil/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GStencil/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_object_and_peak_source():
    import bench
    peak, src = bench.measured_peak()
    assert peak > 1000 and ("measured" in src or "fallback" in src)
    pts = 512 ** 3
    r = bench.roofline_of(pts, 0.46e-3, 2, peak, src, 2175366000)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == peak
    assert abs(r["achieved"] - 16.0 * pts / 0.46e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / peak) < 1e-12
    assert abs(r["frac_of_single_sweep_roofline"] - 2 * r["frac"]) < 1e-12 and r["steps_per_launch"] == 2
    assert r["traffic"] == 2175366000


def test_clock_sampler_degrades_without_a_gpu():
    import bench
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
