"""bench.py's product arm, start to finish, WITHOUT a GPU: the device-touching pieces (WeakDomain, events, streams, pinned
memory, the C++ driver binaries, the composed kernel's child trial) are replaced by stand-ins, everything else -- the
argument handling, the fused-kernel selection, the order of the legs, the watchdog stages, the keys of the JSON line -- is
bench.py's own code.  The orchestration is the one part of the measurement that cannot be allowed to fail on the box:
a NameError in main() would cost the whole record.  Covers N = 1 and, with a stand-in for torch.distributed, both ranks of
N = 2."""
import ctypes as C
import json
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeBuf:
    ptr, nbytes = 0x10000000, 1177587712

    def zero(self):
        pass

    def free(self):
        pass


class FakeStorage:
    def __init__(self):
        self.dat = FakeBuf()


class FakeEvent:
    clock = [0.0]

    def __init__(self):
        self.t, self.h = 0.0, C.c_void_p(1)

    def record(self, stream=None):
        FakeEvent.clock[0] += 0.5
        self.t = FakeEvent.clock[0]

    def sync(self):
        pass

    def elapsed_ms(self, later):
        return max(later.t - self.t, 0.25)


class FakeDomain:
    made = 0

    def __init__(self, dom, st, cart=(1, 1, 1), coo=(0, 0, 0), rank=0, kernel=0):
        import bricklib_b200 as bk
        FakeDomain.made += 1
        self.dom, self.stencil, self.cart, self.coo, self.rank, self.kernel = tuple(dom), st, cart, coo, rank, kernel
        self.st_iter = bk._lib.load().bk_stencil_st_iter(st)
        self.storage = [FakeStorage(), FakeStorage()]
        self.bricks = [object(), object()]
        self.grid = types.SimpleNamespace(dims=(66, 66, 66))
        self.info = types.SimpleNamespace(allocate=lambda step: FakeStorage())
        self.decomp = types.SimpleNamespace(sep_pos=[238329, 262145, 287497])
        self.view = types.SimpleNamespace(bytes=103841792)
        self.peers = [r for r in range(cart[0] * cart[1] * cart[2]) if r != rank]
        self.fuse, self.thin, self.transport, self.periods, self.comm_stream = 2, None, "kernel", 0, None

    def connect(self, ptrs=None, hs=None):
        self.connected = True

    def enable_overlap(self):
        self.comm_stream = 1

    def set_pull_shape(self, *a):
        pass

    def fill_synthetic(self, seed, which=0, stream=None):
        pass

    def steps_per_pass(self):
        return 2 if self.stencil in (0, 1) and self.fuse == 2 else 1

    def period(self, stream=None):
        self.periods += 1
        return 6

    def _sweep(self, *a):
        pass

    def _thin(self):
        return bool(self.peers)

    def _remote(self):
        return False


class FakeLib:
    """metadata from the real library (it loads without a device), stubs for everything that needs one"""

    def __init__(self, real):
        self.real, self.launches = real, 0

    def __getattr__(self, name):
        if name in ("bk_stencil_st_iter", "bk_stencil_points", "bk_stencil_radius", "bk_stencil_fused_steps", "bk_last_error",
                    "bk_stencil_fused_variant_set", "bk_stencil_fused_variant_get"):
            return getattr(self.real, name)

        def stub(*a):
            if name == "bk_host_alloc":
                a[0]._obj.value = 0x20000000
            if name == "bk_stream_create":
                a[0]._obj.value = 0x30000000
            return 0
        return stub


class R:
    def __init__(self, out, rc=0):
        self.returncode, self.stdout, self.stderr = rc, out, ""


def fake_run(cmd, **kw):
    exe = os.path.basename(cmd[0]) if not cmd[0].endswith("python") and "python" not in os.path.basename(cmd[0]) else os.path.basename(cmd[1])
    if exe == "-m":
        exe = cmd[2]
    if exe == "composed_trial.py":
        return R(json.dumps({"ok": True, "composed": {"ok": True, "launch_ms": 0.39}, "wide": {"ok": True, "launch_ms": 0.4}}) + "\n")
    if exe == "direct_exchange_trial.py" or (exe == "torch.distributed.run" or "direct_exchange_trial.py" in " ".join(cmd)):
        return R(json.dumps({"ok": True, "n_gpus": 1, "mpi25pt": {"pull_ms": 0.95, "direct_ms": 0.85, "mismatches": 0, "ok": True}}) + "\n")
    if exe == "strong":
        tail = "result match (worst relative difference 4e-16 after 24 steps)\n" if "-v" in cmd else ""
        return R("calc : [0.001, 0.001, 0.001] (s: 0)\ncall : [1e-05, 1e-05, 1e-05] (s: 0)\nwait : [1e-06, 1e-06, 1e-06] (s: 0)\nperf 1086.7 GStencil/s\n" + tail)
    if exe == "weak":
        tail = "result match (worst relative difference 4e-16 after 24 steps)\n" if "-v" in cmd else ""
        return R("Arr: 0.0005\nperf 533.5 GStencil/s\nBri: 0.0003\nperf 929.7 GStencil/s\nArr == Bri: result match\n" + tail)
    if exe == "single":
        return R("Arr: 0.63\nTrans: 0.000357\nperf 375.5 GStencil/s 6008 GB/s\nresult match\n")
    raise AssertionError(f"unexpected subprocess {cmd}")


@pytest.fixture
def dry(monkeypatch):
    import bench
    import bricklib_b200 as bk
    fake_lib = FakeLib(bk._lib.load())
    FakeDomain.made = 0
    from bricklib_b200 import core, weak
    for mod in (bk, core, weak):            # every module binds `load` by name
        monkeypatch.setattr(mod, "load", lambda: fake_lib)
    monkeypatch.setattr(bk, "WeakDomain", FakeDomain)
    for mod in (bk, core):
        monkeypatch.setattr(mod, "Event", FakeEvent)
        monkeypatch.setattr(mod, "device_sync", lambda: None)
    monkeypatch.setattr(bk, "stencil_advance", lambda *a, **k: None)
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setattr(bench, "sampled_parity", lambda bk_, d, box_bricks=4: (4.3e-16, 131072))
    monkeypatch.setattr(bench, "fused_vs_two_sweeps", lambda bk_, d: (0, 0.0, 512 ** 3) if d.steps_per_pass() == 2 else None)
    monkeypatch.setattr(bench, "reference_period_seconds", lambda *a, **k: (0.26, "reference", 16, "avx512"))
    monkeypatch.delenv("BK_FUSED_VARIANT", raising=False)
    before = bk.fused_variant()
    yield bench
    bk.fused_variant(before)
    os.environ.pop("BK_FUSED_VARIANT", None)


def run_main(bench, monkeypatch, capsys, argv):
    monkeypatch.setattr(sys, "argv", ["bench.py", *argv])
    bench.main()
    out = [x for x in capsys.readouterr().out.splitlines() if x.startswith("{")]
    return [json.loads(x) for x in out]


def test_product_arm_runs_every_leg_and_prints_one_complete_line(dry, monkeypatch, capsys):
    lines = run_main(dry, monkeypatch, capsys, ["--steps", "4", "--warmup", "3"])
    assert len(lines) == 1
    d = lines[0]
    assert d["metric"] == "GStencil/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3 and d["value"] > 0
    assert d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None and "workload" in d["config"]
    assert "extras_truncated" not in d
    assert d["gpu_launches"] == 4 * 6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["steps_per_launch"] == 2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["fused_kernel"]["selected"] in ("staged", "composed", "wide") and "launch_ms" in d["fused_kernel"]
    assert d["loop_options"]["selected"]["first"] in ("pull", "ready") and len(d["loop_options"]["ms_per_step"]) == 2
    assert d["config"]["ready_first"] == (d["loop_options"]["selected"]["first"] == "ready")
    assert all("loop_options" in d["others"][k] for k in ("mpi13pt", "mpi25pt", "mpi125pt"))
    assert d["parity"]["ok"] and d["parity"]["fused_vs_two_sweeps"]["mismatches"] == 0
    o = d["others"]
    assert set(o) >= {"mpi13pt", "mpi25pt", "mpi125pt", "strong", "array_layout_baseline", "single_7pt_512"}
    assert all(o[k]["parity"]["ok"] and o[k]["steps_per_launch"] == 1 for k in ("mpi13pt", "mpi25pt", "mpi125pt"))
    assert o["strong"]["global_1024_sub_64"]["GStencil/s"] == 1086.7 and "share_512_stitched" in o["strong"]
    assert o["array_layout_baseline"]["mpi25pt"]["arr_equals_bri"] is True
    assert o["single_7pt_512"]["GStencil/s"] == 375.5 and o["single_7pt_512"]["validation"] == "result match"
    assert o["exchange_inside_the_sweep"]["ok"] and o["exchange_inside_the_sweep"]["mpi25pt"]["mismatches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == d["e2e"]["d2h_bytes_per_step"] == (262145 - 1) * 4096 and d["e2e"]["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 16
    assert d["baseline_configs"]["configs[4] strong 1024^3 in 64^3 subdomains"] == 1086.7
    assert FakeDomain.made == 3        # the timed domain + two more fields in flight for the end-to-end leg


def test_product_arm_with_a_forced_composed_kernel_hands_it_to_the_drivers_after_validation(dry, monkeypatch, capsys):
    d = run_main(dry, monkeypatch, capsys, ["--steps", "2", "--fused", "composed"])[0]
    assert d["fused_kernel"]["selected"] == "composed" and d["fused_kernel"]["drivers"]["ok"]
    assert "diamond" in d["roofline"]["kernel"] and d["roofline"]["traffic"] is None
    assert os.environ.get("BK_FUSED_VARIANT") == "composed"


def test_no_extras_prints_the_headline_only(dry, monkeypatch, capsys):
    d = run_main(dry, monkeypatch, capsys, ["--steps", "2", "--no-extras", "--fused", "staged"])[0]
    assert "others" not in d and "e2e" not in d and d["fused_kernel"]["why"] == "forced"
    assert d["roofline"]["traffic"] == 2175366000


class FakeDist:
    """torch.distributed for one of two ranks: collectives are identities, the log keeps the order of the calls"""

    def __init__(self):
        self.log = []

    def barrier(self, group=None):
        self.log.append("host barrier" if group is not None else "barrier")

    def destroy_process_group(self):
        self.log.append("destroy")

    def all_gather_object(self, out, obj):
        for i in range(len(out)):
            out[i] = obj


@pytest.mark.parametrize("rank", [0, 1])
def test_two_rank_orchestration(dry, monkeypatch, capsys, rank):
    bench = dry
    fd = FakeDist()
    monkeypatch.setattr(bench, "dist_setup", lambda n: (rank, 2, fd, "gloo-group"))
    monkeypatch.setattr(bench, "max_over_ranks", lambda dist, v: v)
    monkeypatch.setattr(bench, "sum_over_ranks", lambda dist, v: v)
    monkeypatch.setattr(bench, "barrier", lambda dist: dist.log.append("barrier") if dist is not None else None)
    monkeypatch.setattr(bench, "wire_peers", lambda bk, dom, dist, r, w: dom.connect())
    monkeypatch.setenv("LOCAL_RANK", str(rank))
    lines = run_main(bench, monkeypatch, capsys, ["--gpus", "2", "--steps", "4"])
    assert "host barrier" in fd.log and fd.log[-1] == "destroy"
    if rank == 1:
        assert lines == []
        return
    d = lines[0]
    assert d["n_gpus"] == 2 and d["config"]["process_grid"] == "2x1x1" and d["config"]["thin_split"] is True
    assert "cpu_baseline" not in d and "single_7pt_512" not in d["others"] and "share_512_stitched" not in d["others"]["strong"]
    assert d["others"]["strong"]["global_1024_sub_64"]["cmd"].endswith("-g 2 -S mpi7pt")
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * (262145 - 1) * 4096
