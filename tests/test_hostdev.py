"""The Python above the C ABI that only ever runs on a GPU box -- the weak loop with its device-side flag handshake, the
end-to-end pipeline of bench.py, tests/handshake_case.py, tools/composed_trial.py, bench.py's main() -- executed here on
tests/hostdev.py, a CPU stand-in for the device side of the library (streams as queues that run when the host blocks,
events, kernels that wait for a flag, hardware queues shared by a process's streams, seeded schedules, deadlock
DETECTION).  What this checks is orchestration: that every rank's operations are issued in an order that can complete
under any schedule, that the scripts are right as Python, that the numbers they move end up where the oracle says.  The
arithmetic of the stand-in is the oracle's own, so numerical agreement here says nothing about the CUDA kernels."""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bricklib_b200 as bk  # noqa: E402
import hostdev  # noqa: E402
import oracle  # noqa: E402
from oracle import schedule as S  # noqa: E402


def load_script(name):
    """a script of tools/ or tests/ as a module; what it exports to the environment at import time is undone (the tests say
    what they want set)"""
    saved = dict(os.environ)
    where = "tools" if os.path.exists(os.path.join(ROOT, "tools", name + ".py")) else "tests"
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, where, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for k in set(os.environ) - set(saved):
        del os.environ[k]
    os.environ.update(saved)
    return m


@pytest.mark.parametrize("st", [1, 2, 3, 4])
@pytest.mark.parametrize("mode", ["pull", "direct"])
@pytest.mark.parametrize("overlap,ready_first", [(False, False), (True, False), (True, True)])
def test_weak_loop_runs_on_the_stand_in_device(st, overlap, ready_first, mode):
    rng = np.random.default_rng(st)
    dom = (32, 24, 16)
    field = rng.random(dom[::-1])
    before = bk.fused_variant(bk.FUSED_COMPOSED if mode == "direct" else bk.FUSED_STAGED)
    with hostdev.installed(policy="random", seed=st) as dev:
        d = bk.WeakDomain(dom, st)
        d.connect()
        d.exchange_mode = mode
        assert d.direct_active() == (mode == "direct")
        d.ready_first = ready_first
        if overlap:
            d.enable_overlap()
        d.load_interior(field)
        launches = sum(d.period() for _ in range(2))
        bk.device_sync()
        got = d.read_interior(0)
        assert launches > 0 and dev.launches >= launches
        if overlap:     # the READY half of pass 0 is submitted before / after the pull -- or there is no pull at all
            order = [x for x in dev.executed if x.startswith("pull") or x.endswith("READY")]
            assert len(order) == (4 if mode == "pull" else 2)
        assert ("ghost bricks through the address table" in dev.executed) == (mode == "direct")
    bk.fused_variant(before)
    want = S.periodic_steps(st, field, 2 * oracle.ST_ITER[st])
    assert float((np.abs(got - want) / (np.abs(got) + np.abs(want))).max()) < 1e-12


@pytest.mark.parametrize("ranks,extra,kw", [
    (2, [], {}), (4, [], {}), (2, ["--no-overlap"], {}),
    (8, [], {"policy": "random", "seed": 3}),
    (4, [], {"hw_queues": 1}),                                  # ONE hardware queue per rank: the most adversarial valid model
    (8, [], {"hw_queues": 1, "policy": "random", "seed": 5}),
    (8, ["--no-overlap"], {"hw_queues": 2, "policy": "random", "seed": 9}),
])
@pytest.mark.parametrize("ready_first,mode", [("0", "pull"), ("1", "pull"), ("0", "direct")])
def test_handshake_case_script_completes_under_every_schedule(monkeypatch, capsys, ranks, extra, kw, ready_first, mode):
    """tests/handshake_case.py as it will run on the box (tests/test_zx_handshake_gpu.py), here on the stand-in: ranks
    ordered by the device-side flags alone; within a rank the submission order is a valid serial order, so even one
    hardware queue per rank cannot deadlock it"""
    hs = load_script("handshake_case")
    monkeypatch.setenv("BK_SKIP_ADJ_CHECK", "1")        # what the script sets for itself when it is run
    monkeypatch.setenv("BK_READY_FIRST", ready_first)
    monkeypatch.setenv("BK_EXCHANGE_MODE", mode)            # "direct": the exchange inside the sweep (no pull at all)
    before = bk.fused_variant(bk.FUSED_COMPOSED if mode == "direct" else bk.FUSED_STAGED)
    monkeypatch.setattr(sys, "argv", ["handshake_case.py", "--ranks", str(ranks), "--size", "16", "--periods", "3", "--stencils",
                                      "mpi7pt,mpi25pt", *extra])
    with hostdev.installed(**kw) as dev:
        hs.main()
    bk.fused_variant(before)
    assert ("ghost bricks through the address table" in dev.executed) == (mode == "direct")
    assert f"handshake ok: {ranks} ranks" in capsys.readouterr().out


def test_one_host_thread_feeding_several_ranks_must_not_block_in_the_adjacency_check(monkeypatch):
    """why tests/handshake_case.py switches the one-time grid-vs-adjacency check off: it synchronises the stream at the first
    launch over a new box, and a host that blocks inside rank 0's period can never enqueue the signal rank 0 waits for"""
    hs = load_script("handshake_case")
    monkeypatch.setattr(sys, "argv", ["handshake_case.py", "--ranks", "2", "--size", "16", "--periods", "2", "--stencils", "mpi7pt"])
    monkeypatch.delenv("BK_SKIP_ADJ_CHECK", raising=False)      # (the script set it at import)
    with pytest.raises(hostdev.Deadlock):
        with hostdev.installed():
            hs.main()
    monkeypatch.setenv("BK_SKIP_ADJ_CHECK", "1")
    with hostdev.installed():
        hs.main()


def test_composed_trial_script_reports_every_variant(monkeypatch, capsys):
    """tools/composed_trial.py (the child process bench.py and the GPU tests run before the composed kernel is allowed
    near anything else) as Python: every launch shape, the three variants, the JSON keys the callers read"""
    ct = load_script("composed_trial")
    monkeypatch.setattr(sys, "argv", ["composed_trial.py", "--size", "32"])
    before = bk.fused_variant()
    with hostdev.installed():
        ct.main()
    bk.fused_variant(before)
    out = json.loads([x for x in capsys.readouterr().out.splitlines() if x.startswith("{")][-1])
    assert out["ok"] and out["composed"]["ok"] and out["wide"]["ok"]
    for name in ("staged", "composed", "wide"):
        assert out[name]["mismatches"] == 0 and out[name]["steps_per_launch"] == 2 and out[name]["points"] == 32 ** 3
    assert out["composed"]["small_boxes_max_rel"] < 1e-13


def test_bench_main_on_the_stand_in_device(monkeypatch, capsys):
    """bench.py's product arm with its REAL WeakDomain / FieldPipeline / parity code on the stand-in (only the C++ driver
    binaries, the child trial and the CPU reference timing are replaced): one complete line, parity green"""
    import bench
    from test_bench_dryrun import fake_run
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setattr(bench, "reference_period_seconds", lambda *a, **k: (0.26, "reference", 16, "avx512"))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--size", "32", "--steps", "2"])
    monkeypatch.delenv("BK_FUSED_VARIANT", raising=False)
    before = bk.fused_variant()
    with hostdev.installed() as dev:
        bench.main()
    bk.fused_variant(before)
    os.environ.pop("BK_FUSED_VARIANT", None)
    d = json.loads([x for x in capsys.readouterr().out.splitlines() if x.startswith("{")][-1])
    assert "extras_truncated" not in d and d["steps"] == 2 and d["warmup"] == 3
    assert d["gpu_launches"] == 2 * 6                       # per period: pull, READY + REST of pass 0, three more passes
    assert d["parity"]["ok"] and d["parity"]["fused_vs_two_sweeps"]["mismatches"] == 0
    assert all(d["others"][k]["parity"]["ok"] for k in ("mpi13pt", "mpi25pt", "mpi125pt"))
    assert d["e2e"]["h2d_bytes_per_step"] == d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["fused_kernel"]["vs_two_sweeps"]["composed"]["mismatches"] == 0
    assert d["loop_options"]["selected"]["first"] in ("pull", "ready")
    assert dev.launches > 100


def _two_rank_pipelines(dev, ranks, one_period_in_flight, steps, issue_seed):
    """`ranks` emulated ranks, each a FieldPipeline of three fields in flight (bench.py's end-to-end leg), issuing their
    steps independently (a seeded interleaving, each rank in step order)"""
    import bench
    from bricklib_b200.weak import Handshake
    os.environ["BK_SKIP_ADJ_CHECK"] = "1"   # ONE host thread feeds all ranks here (see the handshake script); undone by the caller
    cart = {2: (2, 1, 1), 4: (2, 2, 1)}[ranks]
    coords = S.cart_coords(cart)
    slots, dom = 3, (16, 16, 16)
    doms = [[bk.WeakDomain(dom, 1, cart, coords[r], r) for r in range(ranks)] for _ in range(slots)]
    for s in range(slots):      # slot s of every rank forms one distributed field: own storages, flags, epochs
        ptrs = {r: doms[s][r].storage[0].dat.ptr for r in range(ranks)}
        hs = [Handshake(ranks) for _ in range(ranks)]
        for r in range(ranks):
            dev.process = r
            for q in range(ranks):
                hs[r].peer[q] = hs[q].buf.ptr
            doms[s][r].connect(ptrs, hs[r])
            doms[s][r].enable_overlap()
            doms[s][r].fill_synthetic(7 + s)
    bk.device_sync()
    pipes = []
    for r in range(ranks):
        dev.process = r
        pipes.append(bk.FieldPipeline([doms[s][r] for s in range(slots)], one_period_in_flight=one_period_in_flight))
    rng = np.random.default_rng(issue_seed)
    nxt = [0] * ranks
    while any(n < steps for n in nxt):
        r = int(rng.choice([q for q in range(ranks) if nxt[q] < steps]))
        pipes[r].step(nxt[r])
        nxt[r] += 1
    for p in pipes:
        p.sync()
    return doms


@pytest.mark.parametrize("kw", [{}, {"hw_queues": 1}, {"hw_queues": 2, "policy": "random", "seed": 1}, {"policy": "random", "seed": 4},
                                {"wide_pull_spin": True, "policy": "random", "seed": 2}])
@pytest.mark.parametrize("ranks", [2, 4])
def test_end_to_end_pipelines_of_several_ranks_cannot_deadlock(kw, ranks):
    """as shipped (one period in flight per GPU): completes under shared hardware queues, random schedules, random issue
    orders -- even if the pull still spun in every CTA"""
    try:
        for issue_seed in range(3):
            with hostdev.installed(**kw) as dev:
                _two_rank_pipelines(dev, ranks, True, 7, issue_seed)
    finally:
        os.environ.pop("BK_SKIP_ADJ_CHECK", None)


def test_the_hang_of_the_round_2_bench_reproduced():
    """round 2's bench.py kept the periods of all three fields in flight AND its pull spun in every CTA while it waited
    for the peers' flags.  Its only 8-GPU run hung.  On the stand-in that combination deadlocks as soon as the schedule
    is not strictly oldest-first: on each GPU one field's spinning pull holds the SMs and waits for a flag whose k_signal
    sits, on the peer, behind ANOTHER field's spinning pull.  Either change made since -- the one-CTA wait kernel, or one
    period in flight -- removes it."""
    deadlocks = 0
    try:
        for issue_seed in range(4):
            try:
                with hostdev.installed(wide_pull_spin=True, policy="random", seed=2) as dev:
                    _two_rank_pipelines(dev, 2, False, 7, issue_seed)
            except hostdev.Deadlock as e:
                deadlocks += 1
                assert "pull that spins in every CTA" in str(e) and "k_signal" in str(e)
        assert deadlocks == 4
        for issue_seed in range(4):         # the narrow wait kernel alone is enough ...
            with hostdev.installed(policy="random", seed=2) as dev:
                _two_rank_pipelines(dev, 2, False, 7, issue_seed)
        for issue_seed in range(4):         # ... and so is one period in flight alone
            with hostdev.installed(wide_pull_spin=True, policy="random", seed=2) as dev:
                _two_rank_pipelines(dev, 2, True, 7, issue_seed)
    finally:
        os.environ.pop("BK_SKIP_ADJ_CHECK", None)


@pytest.mark.parametrize("world,size,kw", [(2, 32, {}), (2, 16, {"hw_queues": 1, "policy": "random", "seed": 7}),
                                           (4, 16, {"policy": "random", "seed": 11}), (8, 16, {"hw_queues": 2, "policy": "random", "seed": 13}),
                                           (8, 16, {"wide_pull_spin": True, "policy": "random", "seed": 3})])
def test_bench_main_for_several_ranks_on_the_stand_in_device(monkeypatch, capsys, world, size, kw):
    """bench.py's main() for EVERY rank of an N-rank job at once -- one host thread per rank on the stand-in device, the
    collectives of torch.distributed replaced by ones that really wait for every rank: rendezvous over (stand-in) IPC
    handles, the fused-kernel and submission-order selections, the timed periods with their device-side handshake, parity
    against the oracle on every rank, the other stencils, rank 0's driver legs behind the host barrier, the end-to-end
    pipelines.  A protocol error anywhere in that sequence is a detected deadlock here instead of a hung 8-GPU box."""
    import bench
    from test_bench_dryrun import fake_run
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(world), "--size", str(size), "--steps", "4"])
    monkeypatch.delenv("BK_FUSED_VARIANT", raising=False)
    monkeypatch.delenv("BK_SKIP_ADJ_CHECK", raising=False)      # one host thread per rank: the check's stream synchronise is on
    before = bk.fused_variant()
    with hostdev.installed(**kw) as dev:
        dist = hostdev.RankDist(dev, world)
        monkeypatch.setattr(bench, "dist_setup", lambda n: (dev.process, world, dist.for_rank(dev.process), "gloo"))
        monkeypatch.setattr(bench, "max_over_ranks", lambda d, v: v if d is None else d.allreduce(v, "max"))
        monkeypatch.setattr(bench, "sum_over_ranks", lambda d, v: v if d is None else d.allreduce(v, "sum"))
        monkeypatch.setattr(bench, "barrier", lambda d: d.barrier() if d is not None else None)
        hostdev.run_ranks(dev, world, lambda r: bench.main())
    bk.fused_variant(before)
    os.environ.pop("BK_FUSED_VARIANT", None)
    lines = [json.loads(x) for x in capsys.readouterr().out.splitlines() if x.startswith("{")]
    assert len(lines) == 1                          # rank 0 alone prints
    d = lines[0]
    assert d["n_gpus"] == world and "extras_truncated" not in d
    assert d["parity"]["ok"] and d["parity"]["checked_points"] > 0
    assert all(d["others"][k]["parity"]["ok"] for k in ("mpi13pt", "mpi25pt", "mpi125pt"))
    assert d["others"]["strong"]["global_1024_sub_64"]["cmd"].endswith(f"-g {world} -S mpi7pt")
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] % world == 0
    assert d["loop_options"]["selected"]["first"] in ("pull", "ready") and len(d["loop_options"]["ms_per_step"]) == 4
    assert d["config"]["thin_split"] == d["loop_options"]["selected"]["thin"]


@pytest.mark.parametrize("transport", ["kernel", "ce"])
@pytest.mark.parametrize("world,kw", [(2, {}), (4, {"policy": "random", "seed": 2}), (8, {"hw_queues": 1, "policy": "random", "seed": 4})])
def test_multi_process_parity_script_with_rank_threads(monkeypatch, capsys, world, kw, transport):
    """tests/mgpu_weak_check.py (what tests/test_multi_gpu.py launches under torchrun on a multi-GPU box) for all its ranks at
    once on the stand-in: rendezvous, four stencils on one domain, overlap and fused passes, parity with the oracle"""
    import bench
    spec = importlib.util.spec_from_file_location("mgpu_weak_check", os.path.join(ROOT, "tests", "mgpu_weak_check.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    monkeypatch.setattr(sys, "argv", ["mgpu_weak_check.py", "--size", "16", "--periods", "2", "--transport", transport])
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.delenv("BK_SKIP_ADJ_CHECK", raising=False)

    def rank_main(r):
        try:
            m.main()
        except SystemExit as e:
            return e.code

    with hostdev.installed(**kw) as dev:
        dist = hostdev.RankDist(dev, world)
        monkeypatch.setattr(bench, "dist_setup", lambda n: (dev.process, world, dist.for_rank(dev.process), "gloo"))
        monkeypatch.setattr(bench, "max_over_ranks", lambda d, v: v if d is None else d.allreduce(v, "max"))
        monkeypatch.setattr(bench, "barrier", lambda d: d.barrier() if d is not None else None)
        codes = hostdev.run_ranks(dev, world, rank_main)
    assert codes == [0] * world
    res = json.loads([x for x in capsys.readouterr().out.splitlines() if x.startswith("{")][-1])
    assert res["ok"] and res["world"] == world and res["max_rel"] < 1e-12


def test_field_pipeline_streams_host_fields_through_the_device():
    """the public streaming API behind bench.py's end-to-end leg: what goes in through the pinned input buffers comes out of
    the result buffers advanced by one exchange period (checked against the oracle's periodic sweep)"""
    rng = np.random.default_rng(21)
    dom = (16, 24, 16)
    with hostdev.installed(policy="random", seed=21):
        doms = [bk.WeakDomain(dom, 1) for _ in range(3)]
        for d in doms:
            d.connect()
            d.enable_overlap()
        pipe = bk.FieldPipeline(doms)
        fields, want = [], []
        for step in range(5):
            f = rng.random(dom[::-1])
            fields.append(f)
            want.append(S.periodic_steps(1, f, oracle.ST_ITER[1]))
        got = []
        for step, f in enumerate(fields):
            slot = step % 3
            if step >= 3:                       # the slot's previous result must have left before its buffers are reused
                pipe.sync()
                got.append((step - 3, pipe.host_out(slot).copy()))
            doms[slot].load_interior(f)         # stage the field in brick order, then take it back as the slot's host input
            bk.device_sync()
            n = pipe.nbytes // 8
            pipe.host_in(slot)[:] = doms[slot].storage[0].to_host()[512:512 + n]
            doms[slot].storage[0].dat.zero()
            bk.device_sync()                    # (the null stream does not order itself against the slot's upload stream)
            pipe.step(step)
        pipe.sync()
        for step in range(2, 5):
            got.append((step, pipe.host_out(step % 3).copy()))
        got = dict(got)
        for step in (2, 3, 4):
            doms[0].storage[0].dat.zero()
            bk.device_sync()
            full = doms[0].storage[0].to_host()
            full[512:512 + got[step].size] = got[step]
            doms[0].storage[0].from_host(full)
            res = doms[0].read_interior(0)
            assert float((np.abs(res - want[step]) / (np.abs(res) + np.abs(want[step]))).max()) < 1e-12, step
        pipe.close()


@pytest.mark.parametrize("world,kw", [(1, {}), (2, {"policy": "random", "seed": 1}), (4, {"hw_queues": 1, "policy": "random", "seed": 6})])
def test_direct_exchange_trial_script_on_the_stand_in_device(monkeypatch, capsys, world, kw):
    """tools/direct_exchange_trial.py (the child job behind bench.py's `exchange_inside_the_sweep` leg) for all its ranks:
    a pull domain and a direct domain per rank on the same field, the same periods through both, the same numbers out"""
    import bench
    m = load_script("direct_exchange_trial")
    monkeypatch.setattr(sys, "argv", ["direct_exchange_trial.py", "--size", "16", "--periods", "2"])
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.delenv("BK_SKIP_ADJ_CHECK", raising=False)
    before = bk.fused_variant()
    with hostdev.installed(**kw) as dev:
        if world == 1:
            m.main()
        else:
            dist = hostdev.RankDist(dev, world)
            monkeypatch.setattr(bench, "dist_setup", lambda n: (dev.process, world, dist.for_rank(dev.process), "gloo"))
            monkeypatch.setattr(bench, "max_over_ranks", lambda d, v: v if d is None else d.allreduce(v, "max"))
            monkeypatch.setattr(bench, "sum_over_ranks", lambda d, v: v if d is None else d.allreduce(v, "sum"))
            monkeypatch.setattr(bench, "barrier", lambda d: d.barrier() if d is not None else None)
            hostdev.run_ranks(dev, world, lambda r: m.main())
        assert "ghost bricks through the address table" in dev.executed and any(x.startswith("pull") for x in dev.executed)
    bk.fused_variant(before)
    out = json.loads([x for x in capsys.readouterr().out.splitlines() if x.startswith("{")][-1])
    assert out["ok"] and out["n_gpus"] == world
    for name in ("mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"):
        assert out[name]["ok"] and out[name]["mismatches"] == 0 and out[name]["pull_ms"] > 0 and out[name]["direct_ms"] > 0
