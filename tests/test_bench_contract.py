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
    # 16 algorithmic bytes per point per time step x the point-steps of one launch / its duration
    assert abs(r["achieved"] - 16.0 * pts * 2 / 0.46e-3 / 1e9) < 1e-6 and abs(r["frac"] - r["achieved"] / peak) < 1e-12
    # ... and the physical view: a launch moves every point through HBM once whatever it fuses
    assert abs(r["hbm_frac"] * 2 - r["frac"]) < 1e-12 and r["steps_per_launch"] == 2
    assert r["traffic"] == 2175366000
    one = bench.roofline_of(pts, 0.35e-3, 1, peak, src, None)
    assert one["frac"] == one["hbm_frac"]


def test_reference_arm_runs_the_whole_job_with_every_host_core():
    """under torchrun the workers inherit OMP_NUM_THREADS=1: the reference arm must override it, run all N subdomains of
    the job on rank 0 and say how many cores it used (round-1 ADVICE)"""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "64",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][0])
    cores = len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2 and d["config"]["process_grid"] == "2x1x1" and d["cpu_baseline"]["cores"] == cores
    assert d["config"]["workload"] == bench_workload("mpi7pt", 64, 8)
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--size", "64",
                            "--steps", "1"], capture_output=True, text=True, timeout=60, cwd=ROOT, env=dict(env, RANK="1"))
    assert other.returncode == 0 and other.stdout.strip() == ""


def bench_workload(stencil, size, it):
    import bench
    return bench.workload_name(stencil, size, it)


def test_clock_sampler_degrades_without_a_gpu():
    import bench
    s = bench.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_watchdog_prints_the_last_snapshot_when_an_extra_hangs():
    """a hung optional leg must not take the JSON line with it: the watchdog prints the last complete snapshot (rank 0)
    and every rank exits 0"""
    code = ("import sys, time; sys.path.insert(0, %r); import bench\n"
            "rank = int(sys.argv[1]); wd = bench.Watchdog(rank, 0.6)\n"
            "wd.at('headline done', {'metric': 'GStencil/s', 'value': 1.0})\n"
            "wd.at('e2e')\n"
            "time.sleep(30)\n" % ROOT)
    r0 = subprocess.run([sys.executable, "-c", code, "0"], capture_output=True, text=True, timeout=20)
    assert r0.returncode == 0
    d = json.loads(r0.stdout.strip().splitlines()[-1])
    assert d["value"] == 1.0 and d["extras_truncated"]["stage"] == "e2e"
    r1 = subprocess.run([sys.executable, "-c", code, "1"], capture_output=True, text=True, timeout=20)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    # a finished run is left alone
    code2 = ("import sys, time; sys.path.insert(0, %r); import bench\n"
             "wd = bench.Watchdog(0, 0.3); wd.at('x', {'value': 2}); wd.finish(); time.sleep(0.8); print('end')\n" % ROOT)
    r2 = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=20)
    assert r2.returncode == 0 and r2.stdout.strip() == "end"
    import bench
    assert bench.leg_timeout(None, 150) == 150
    wd = bench.Watchdog(1, 1000.0)
    assert 100 < bench.leg_timeout(wd, 150) <= 150
    wd.seconds = 10.0
    assert bench.leg_timeout(wd, 150) == 0
    wd.finish()


class _FakeDomain:
    dom, stencil = (512, 512, 512), 1

    def __init__(self, steps=2):
        self.steps, self.filled = steps, 0

    def steps_per_pass(self):
        return self.steps

    def fill_synthetic(self, seed):
        self.filled += 1


class _FakeBk:
    STENCILS = {"7pt": 0, "mpi7pt": 1}

    def __init__(self):
        self.variant = 0

    def fused_variant(self, v=None):
        before = self.variant
        if v is not None:
            self.variant = v
        return before

    def device_sync(self):
        pass


def test_fused_kernel_selection_policy(monkeypatch):
    """bench.select_fused_kernel: a composed kernel is used only when its child trial passed, it is exact on the device and
    it is the fastest candidate; whatever goes wrong leaves the proven staged kernel in place (no GPU needed: mocked)"""
    import bench

    class R:
        def __init__(self, rc, out):
            self.returncode, self.stdout, self.stderr = rc, out, ""

    def scenario(child, times, parity=None, want="auto", steps=2):
        """times / parity: by variant number 0 staged, 1 composed, 2 wide"""
        fake, dom = _FakeBk(), _FakeDomain(steps)
        parity = parity or {}
        monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: child() if callable(child) else child)
        monkeypatch.setattr(bench, "time_sweeps", lambda bk, d, reps: (times[bk.variant], 2))
        monkeypatch.setattr(bench, "fused_vs_two_sweeps", lambda bk, d: parity.get(bk.variant, (0, 1e-16, 10)))
        info = bench.select_fused_kernel(fake, dom, None, 0, want)
        assert fake.variant == bench.FUSED_NAMES[info["selected"]]
        return info

    both = R(0, 'noise\n{"ok": true, "composed": {"ok": true}, "wide": {"ok": true}}\n')
    only_wide = R(0, '{"ok": true, "composed": {"ok": false}, "wide": {"ok": true}}')
    T = {0: 0.46e-3, 1: 0.39e-3, 2: 0.41e-3}
    assert scenario(both, T)["selected"] == "composed"
    assert scenario(both, {0: 0.46e-3, 1: 0.42e-3, 2: 0.40e-3})["selected"] == "wide"
    assert scenario(both, {0: 0.46e-3, 1: 0.50e-3, 2: 0.47e-3})["selected"] == "staged"
    assert scenario(both, T, parity={1: (3, 0.2, 10)})["selected"] == "wide"            # composed wrong, wide exact and faster
    assert scenario(both, T, parity={1: (0, 1e-9, 10), 2: (1, 1.0, 10)})["selected"] == "staged"
    assert scenario(only_wide, T)["selected"] == "wide"
    assert "composed" not in scenario(only_wide, T)["launch_ms"]                        # never run in this process
    assert scenario(R(-11, ""), T)["selected"] == "staged"                              # the child crashed
    assert scenario(R(0, '{"ok": false}'), T)["selected"] == "staged"                   # the child found mismatches

    def hang():
        raise bench.subprocess.TimeoutExpired("trial", 180)
    assert scenario(hang, T)["selected"] == "staged"                                    # the child hung
    assert scenario(both, T, want="staged")["why"] == "forced"
    assert scenario(both, {0: 0.46e-3, 1: 0.50e-3, 2: 0.40e-3}, want="composed")["selected"] == "composed"   # forced, exact
    assert scenario(both, T, parity={2: (1, 1.0, 10)}, want="wide")["selected"] == "staged"                   # forced, wrong
    assert scenario(both, T, steps=1)["selected"] == "staged"                           # nothing to select


def test_driver_legs_fall_back_to_the_staged_kernel_when_validation_fails(monkeypatch):
    import bench

    class R:
        def __init__(self, rc, out):
            self.returncode, self.stdout, self.stderr = rc, out, ""

    monkeypatch.delenv("BK_FUSED_VARIANT", raising=False)
    assert bench.drivers_accept_fused_variant(2) is None
    monkeypatch.setenv("BK_FUSED_VARIANT", "composed")
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: R(0, "perf 1\nresult match (worst relative difference 4e-16 after 24 steps)\n"))
    ok = bench.drivers_accept_fused_variant(2)
    assert ok["ok"] and ok["strong"] == ok["weak"] == "result match" and os.environ["BK_FUSED_VARIANT"] == "composed"
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: R(2, "result mismatch! (12 cells)\n"))
    bad = bench.drivers_accept_fused_variant(2)
    assert not bad["ok"] and "fallback" in bad and "BK_FUSED_VARIANT" not in os.environ
