#!/usr/bin/env python3
"""bench.py -- headline benchmark of the brick hot path on B200 (contract: see the build brief / DESIGN.md section 6).

Workload (BASELINE.json configs[2], the configuration the metric's weak-scaling claim is quoted on):
weak scaling, 7-point star stencil (stencils/mpi7pt.py), 512^3 FP64 cells per GPU in 8^3 bricks, periodic Cartesian
process grid (1 / 2x1x1 / 2x2x1 / 2x2x2), ghost depth 8.  One "step" = one exchange period of the reference's time
loop (weak/main.cu:246-287): ghost-zone exchange + ST_ITER(=8) sweeps.  GStencil/s counts interior points only
(weak/main.cu:315-317).  The other stencils (13/25/125-point) are timed in the same run and reported under "others".

  python bench.py [--gpus N --steps K --warmup W] [--stencil mpi7pt] [--size 512]       # product arm
  python bench.py --impl reference ...                                                  # reference CPU arm
Under torchrun (N>1) every rank drives one GPU; rank 0 prints the single JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# OpenMP set-up of the CPU legs, BEFORE libgomp loads.  The reference arm runs on rank 0 alone with every host core:
# torchrun exports OMP_NUM_THREADS=1 to its workers, which would time the reference on ONE thread (round-1 ADVICE).
def _host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# One hardware work queue per stream (default: 8 shared by all streams).  The multi-GPU legs keep kernels that wait for a
# peer's flag in flight on several streams; with shared queues a finite kernel of ANOTHER stream can sit behind such a
# waiter (a false dependency) -- harmless for one domain, a possible cross-process deadlock for the three independent
# domains of the end-to-end leg.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

if "reference" in sys.argv[1:]:
    os.environ["OMP_NUM_THREADS"] = str(_host_cores())
    os.environ["OMP_PROC_BIND"] = "close"
elif int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CART = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}  # MPI_Dims_create order (weak/args.cpp:101)


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread (every 5 ms) when pynvml works,
    else `nvidia-smi -lms 100` as a subprocess"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    MASKS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self.stop_flag, self.thread = None, [], False, None

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except AttributeError:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((sm, int(why)))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
            self.nvml = nv
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = [s for s, _ in self.samples]
            reasons = sorted(n for n, m in self.MASKS.items() if any(w & m for _, w in self.samples))
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.sm_max),
                    "samples": len(sm), "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


class Watchdog:
    """The headline measurement comes first; parity, the other stencils, the end-to-end leg and the C++ driver legs are
    EXTRAS.  If an extra hangs (a wedged GPU, a peer that died) the run must still end with its JSON line: a daemon
    thread waits for the deadline, rank 0 prints the last complete snapshot of the line -- marked `extras_truncated` --
    and every rank leaves with os._exit(0) (all ranks run the same clock, so torchrun sees N clean exits).  ctypes and
    torch release the GIL while they block, so the thread runs even when the main thread sits in a synchronize."""

    def __init__(self, rank, seconds):
        self.rank, self.seconds, self.t0 = rank, seconds, time.time()
        self.snapshot, self.stage, self.done = None, "start", False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def left(self):
        return self.seconds - (time.time() - self.t0)

    def at(self, stage, line=None):
        self.stage = stage
        if line is not None:
            self.snapshot = json.loads(json.dumps(line))

    def finish(self):
        self.done = True

    def _run(self):
        while not self.done and self.left() > 0:
            time.sleep(0.25)
        if self.done:
            return
        if self.rank == 0 and self.snapshot is not None:
            snap = dict(self.snapshot)
            snap["extras_truncated"] = {"stage": self.stage, "after_s": round(time.time() - self.t0, 1),
                                        "why": "an optional leg did not finish inside the bench's own deadline; the headline "
                                               "measurement above is complete"}
            sys.stdout.write(json.dumps(snap) + "\n")
            sys.stdout.flush()
        os._exit(0 if self.snapshot is not None or self.rank != 0 else 3)


def dist_setup(n_gpus):
    """one process per GPU under torchrun; returns (rank, world, torch.distributed or None, host-side gloo group)"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1, None, None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a host-side barrier (no kernel on any GPU) for the legs that rank 0 runs through a driver binary on ALL GPUs
    host_group = dist.new_group(backend="gloo")
    return rank, world, dist, host_group


def max_over_ranks(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def wire_peers(bk, dom_obj, dist, rank, world):
    """exchange CUDA-IPC handles of storage[0] and of the flag buffer, map every neighbour's memory"""
    from bricklib_b200.weak import Handshake
    import ctypes as C
    L = bk.load()
    if dist is None:
        dom_obj.connect()
        return
    hs = Handshake(world)
    h_store, h_flag = C.create_string_buffer(64), C.create_string_buffer(64)
    bk._lib.check(L.bk_ipc_export(dom_obj.storage[0].dat.ptr, h_store))
    bk._lib.check(L.bk_ipc_export(hs.buf.ptr, h_flag))
    gathered = [None] * world
    dist.all_gather_object(gathered, (h_store.raw, h_flag.raw))
    ptrs = {}
    hs.peer[rank] = hs.buf.ptr
    for p in dom_obj.peers:
        sp, fp = C.c_void_p(), C.c_void_p()
        bk._lib.check(L.bk_ipc_open(gathered[p][0], C.byref(sp)))
        bk._lib.check(L.bk_ipc_open(gathered[p][1], C.byref(fp)))
        ptrs[p] = sp.value
        hs.peer[p] = fp.value
    dom_obj.connect(ptrs, hs)


def time_periods(bk, d, steps, warmup, dist):
    """W warm-up periods, then K timed periods between barriers; returns (seconds max over ranks, launches)"""
    for _ in range(warmup):
        d.period()
    bk.device_sync()
    barrier(dist)
    ev0, ev1 = bk.Event(), bk.Event()
    launches = 0
    ev0.record()
    for _ in range(steps):
        launches += d.period()
    ev1.record()
    ev1.sync()
    bk.device_sync()
    barrier(dist)
    sec = ev0.elapsed_ms(ev1) / 1e3
    return max_over_ranks(dist, sec), launches


def time_sweeps(bk, d, reps):
    """the dominant kernel alone: `reps` launches over the interior between two events on the launching stream.
    Returns (seconds per launch, time steps one launch advances)."""
    t = d.grid.dims
    lo, hi = (1, 1, 1), tuple(x - 1 for x in t)
    steps = d.steps_per_pass()

    def one(s):
        if steps == 1:
            d._sweep(s % 2, 1 - s % 2, lo, hi, None)
        else:
            bk.stencil_advance(d.stencil, steps, d.grid, d.bricks[s % 2], d.bricks[1 - s % 2], lo, hi)

    for s in range(3):
        one(s)
    bk.device_sync()
    ev0, ev1 = bk.Event(), bk.Event()
    ev0.record()
    for s in range(reps):
        one(s)
    ev1.record()
    ev1.sync()
    return ev0.elapsed_ms(ev1) / 1e3 / reps, steps


def roofline_of(pts, launch_s, steps, peak, peak_src, traffic):
    """HBM roofline of one launch of the dominant sweep kernel, in SURVEY 8(d)'s unit: 16 algorithmic bytes per point per
    TIME STEP x the point-steps one launch processes (points x steps fused in the launch) / its CUDA-event duration.
    For the two-steps-per-pass kernel `frac` can exceed 1: the kernel moves each byte through HBM once per TWO steps
    (temporal blocking), which no one-sweep-per-pass kernel can do; `hbm_frac` is the physical view of the same launch
    (16 B x points actually read+written / duration / peak), and `traffic` the DRAM bytes ncu counted for it."""
    achieved = 16.0 * pts * steps / launch_s / 1e9
    physical = 16.0 * pts / launch_s / 1e9
    kern = "k_star2 (two time steps per pass)" if steps == 2 else "k_star (one sweep)"
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": f"{kern}: one launch = {pts} interior points x {steps} step(s) x 16 B",
            "kernel_ms": launch_s * 1e3, "steps_per_launch": steps, "peak_source": peak_src,
            "GStencil/s_kernel": pts * steps / launch_s / 1e9,
            "hbm_GB/s": physical, "hbm_frac": physical / peak,
            "note": "frac = algorithmic bytes (16 B per point per step, SURVEY 8d) / time / measured copy peak; "
                    "hbm_frac = compulsory bytes that cross HBM (16 B per point per LAUNCH) / time / peak"}


# ---- parity inside the bench: what was timed is also checked ---------------------------------------------------------
PARITY_TOL = 1e-12
PARITY_SEED = 0xB200


def sampled_parity(bk, d, box_bricks=4):
    """One exchange period of the CURRENT stencil on the position-addressable synthetic field, then four sampled boxes of
    this rank's interior (low corner, high corner, centre, one edge: every kind of ghost dependency) against the oracle
    (the C port's array sweeps on the same global field, evaluated independently from the hash).  The oracle is the
    checker here, never the thing measured.  Returns (max relative difference, points checked)."""
    import oracle
    P = oracle.port()
    st, it = d.stencil, d.st_iter
    r = oracle.RADIUS[st]
    d.fill_synthetic(PARITY_SEED)
    d.storage[1].dat.zero()
    bk.device_sync()
    d.period()
    bk.device_sync()
    nb = tuple(x // 8 for x in d.dom)
    B = tuple(min(box_bricks, n) for n in nb)
    where = [(0, 0, 0), tuple(n - b for n, b in zip(nb, B)), tuple((n - b) // 2 for n, b in zip(nb, B)),
             (0, (nb[1] - B[1]) // 2, nb[2] - B[2])]
    org, glob = d.global_origin(), d.global_cells()
    worst, pts = 0.0, 0
    for p in sorted(set(where)):
        got = d.read_bricks(p, tuple(a + b for a, b in zip(p, B)))
        lo = tuple(org[a] + 8 * p[a] - 8 for a in range(3))
        hi = tuple(org[a] + 8 * (p[a] + B[a]) + 8 for a in range(3))
        cur = bk.synthetic_field(PARITY_SEED, glob, lo, hi)
        for _ in range(it):    # the halo of 8 cells is exactly what ST_ITER steps of radius 8/ST_ITER consume
            n = cur.shape
            cur = P.sweep_array(st, np.ascontiguousarray(cur), (r, r, r), (n[2] - r, n[1] - r, n[0] - r))
        want = cur[8:-8, 8:-8, 8:-8]
        worst = max(worst, float((np.abs(got - want) / (np.abs(got) + np.abs(want) + 1e-300)).max()))
        pts += got.size
    return worst, pts


def fused_vs_two_sweeps(bk, d):
    """device-side, over the FULL interior: one two-steps-per-pass launch against two plain sweeps of the same input.
    Returns None when the stencil has no fused kernel in use, else (mismatching cells at 1e-12, max rel diff, points)."""
    if d.steps_per_pass() != 2:
        return None
    t = d.grid.dims
    lo, hi = (1, 1, 1), tuple(x - 1 for x in t)
    extra = [d.info.allocate(bk.BRICK), d.info.allocate(bk.BRICK)]
    fused, plain = bk.Brick(d.info, extra[0], 0), bk.Brick(d.info, extra[1], 0)
    d.fill_synthetic(PARITY_SEED + 1)
    bk.stencil_advance(d.stencil, 2, d.grid, d.bricks[0], fused, lo, hi)
    d._sweep(0, 1, (0, 0, 0), t, None)
    bk.stencil(d.stencil, d.grid, d.bricks[1], plain, lo, hi, None, d.kernel)
    ok, bad, rel = bk.compare_storage(d.grid, lo, hi, fused, plain, PARITY_TOL)
    for e in extra:
        e.dat.free()
    return bad, rel, int(np.prod([h - l for l, h in zip(lo, hi)])) * 512


def parity_of(bk, d, dist):
    worst, pts = sampled_parity(bk, d)
    out = {"max_rel": max_over_ranks(dist, worst), "checked_points": int(sum_over_ranks(dist, pts)), "tolerance": PARITY_TOL,
           "what": "one exchange period on the synthetic field; 4 sampled 32^3 boxes per rank (corners, centre, edge) vs the "
                   "oracle's array sweeps of the same global periodic field"}
    f = fused_vs_two_sweeps(bk, d)
    if f is not None:
        out["fused_vs_two_sweeps"] = {"mismatches": int(sum_over_ranks(dist, f[0])), "max_rel": max_over_ranks(dist, f[1]),
                                      "points": int(sum_over_ranks(dist, f[2])),
                                      "what": "k_star2 (one launch) vs two k_star sweeps, whole interior, compared on the device"}
    out["ok"] = bool(out["max_rel"] < PARITY_TOL and out.get("fused_vs_two_sweeps", {}).get("mismatches", 0) == 0)
    return out


def e2e_periods(bk, doms, steps):
    """`steps` end-to-end steps through a bricklib_b200.FieldPipeline (the public API for streaming host-resident fields
    through the GPU) after one warm-up step per slot, all inside the timed region:
    H2D of the inputs, the period, D2H of the result.  Returns (seconds per step, h2d bytes, d2h bytes)."""
    pipe = bk.FieldPipeline(doms)
    for i in range(len(doms)):                  # warm-up: one step per slot
        pipe.step(i)
    bk.device_sync()
    t0 = time.perf_counter()
    for i in range(steps):
        pipe.step(i)
    pipe.sync()
    sec = time.perf_counter() - t0
    nbytes = pipe.nbytes
    pipe.close()
    return sec / steps, nbytes, nbytes


FUSED_NAMES = {"staged": 0, "composed": 1, "wide": 2}     # bricklib_b200.FUSED_STAGED / _COMPOSED / _COMPOSED_WIDE


def select_fused_kernel(bk, d, dist, rank, want):
    """Which kernel advances two time steps per pass for the radius-1 stars: the staged one (k_star2, intermediate plane
    in shared memory) or the composed one (one 25-point diamond update, bricklib_b200/csrc/bk_diamond.h; on 4x4-brick
    tiles = "composed", on 8x4-brick tiles = "wide")?  Measure, don't guess: (1) rank 0 lets a CHILD process run the
    composed kernels first (tools/composed_trial.py: every launch shape on a small decomposition, parity against two
    plain sweeps over the whole interior, a few timed launches) -- a fault or hang there costs the child, not this run;
    (2) every rank checks each candidate that survived against two plain sweeps on the device and times it on its own
    domain; (3) a composed kernel is used only if every rank found it exact (< 1e-12, 0 mismatching cells) AND it is the
    fastest candidate (max over ranks).  Returns the record that goes into the JSON line."""
    info = {"policy": want, "selected": "staged"}
    if d.steps_per_pass() != 2:
        info["why"] = "one sweep per pass for this stencil / these options"
        return info
    if want == "staged":
        bk.fused_variant(FUSED_NAMES["staged"])
        info["why"] = "forced"
        return info
    try:
        alive = {"composed": 1.0, "wide": 1.0}
        if rank == 0:
            alive = {"composed": 0.0, "wide": 0.0}
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "composed_trial.py"), "--device",
                                    os.environ.get("LOCAL_RANK", "0"), "--size", str(d.dom[0]), "--stencil",
                                    {v: k for k, v in bk.STENCILS.items()}[d.stencil]],
                                   capture_output=True, text=True, timeout=180, cwd=ROOT)
                lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
                if r.returncode == 0 and lines:
                    info["child_trial"] = json.loads(lines[-1])
                    for name in alive:
                        alive[name] = 1.0 if info["child_trial"].get(name, {}).get("ok") else 0.0
                else:
                    info["child_trial"] = {"ok": False, "rc": r.returncode, "tail": (r.stdout + r.stderr)[-400:]}
            except Exception as exc:
                info["child_trial"] = {"ok": False, "error": str(exc)[:300]}
        for name in sorted(alive):                              # rank 0's verdict, known to all
            alive[name] = 0.0 if max_over_ranks(dist, 1.0 - alive[name]) > 0.0 else 1.0
        if want in alive:
            alive = {k: (v if k == want else 0.0) for k, v in alive.items()}
        bk.fused_variant(FUSED_NAMES["staged"])
        if not any(alive.values()):
            info["why"] = "no composed kernel passed its trial in a child process"
            return info
        times = {"staged": max_over_ranks(dist, time_sweeps(bk, d, 10)[0])}
        info["launch_ms"] = {"staged": times["staged"] * 1e3}
        info["vs_two_sweeps"] = {}
        for name in sorted(k for k, v in alive.items() if v):
            bk.fused_variant(FUSED_NAMES[name])
            bad, worst, pts = fused_vs_two_sweeps(bk, d)
            bad, worst = sum_over_ranks(dist, bad), max_over_ranks(dist, worst)
            t = max_over_ranks(dist, time_sweeps(bk, d, 10)[0])
            info["launch_ms"][name] = t * 1e3
            info["vs_two_sweeps"][name] = {"mismatches": int(bad), "max_rel": worst, "points_per_rank": int(pts)}
            if bad == 0 and worst < PARITY_TOL:
                times[name] = t
        best = min(times, key=times.get)
        if want in ("composed", "wide"):
            best = want if want in times else "staged"
        info["selected"] = best
        info["why"] = ("forced (and exact)" if best == want else "fastest of the exact candidates" if best != "staged" else
                       "no composed kernel was both exact and faster")
        bk.fused_variant(FUSED_NAMES[best])
        d.fill_synthetic(0x5EED)
        bk.device_sync()
    except Exception as exc:
        bk.fused_variant(FUSED_NAMES["staged"])
        info["selected"], info["why"] = "staged", f"selection failed: {str(exc)[:200]}"
    return info


def select_loop_options(bk, d, dist, periods):
    """Two run-time choices of the weak loop that change no result, only the schedule, settled by measurement on this
    run's own domain before anything is timed for the record (max over ranks, `periods` periods after 2 warm-ups each):
      * which of the two independent streams of a period is fed first -- the pull and the READY half of pass 0 behind it
        (the order of rounds 1 and 2), or READY first: its CTAs are then on the SMs when the wide, high-priority pull
        arrives, which trickles in as they retire instead of taking the machine for itself;
      * with neighbours on other GPUs: thin k segments for the ghost-dependent layers of the split pass (BK_PART_THIN) or
        uniform ones (round 1 fixed this per radius from one N = 8 sweep).
    Returns the record for the JSON line; leaves the fastest combination set on `d`."""
    info = {"selected": {"first": "pull", "thin": bool(d._thin())}}
    if d.comm_stream is None:
        info["why"] = "no overlap"
        return info
    forced_thin = d.thin
    thins = [False, True] if (d.peers and forced_thin is None) else [bool(d._thin())]
    try:
        t = {}
        for first in ("pull", "ready"):
            for thin in thins:
                d.ready_first, d.thin = first == "ready", thin
                t[(first, thin)] = time_periods(bk, d, periods, 2, dist)[0] / periods
        best = min(t, key=t.get)
        info["ms_per_step"] = {f"{f} first, {'thin' if th else 'uniform'} segments": v * 1e3 for (f, th), v in t.items()}
        info["selected"] = {"first": best[0], "thin": best[1]}
    except Exception as exc:
        info["why"] = f"selection failed: {str(exc)[:200]}"
    d.ready_first = info["selected"]["first"] == "ready"
    d.thin = info["selected"]["thin"] if len(thins) > 1 else forced_thin
    return info


def reference_period_seconds(stencil_id, size, periods, warm=1, cart=(1, 1, 1)):
    """the reference's own CPU implementation of the path (oracle/_ref: its BrickDecomp, its exchange() over the
    in-process MPI stand-in, its generated AVX brick code) on all host threads; else the C port.  `cart` ranks are run
    in this one process, each with its own subdomain, exchanging through the stand-in -- the whole weak-scaling job on
    the host's cores.  Returns (seconds per period, kind, cores, isa)."""
    import oracle
    from oracle import schedule as S
    R = oracle.ref()
    if R is not None:
        be, kind, isa, cores = S.RefBackend(), "reference", R.isa, R.threads
    else:
        be, kind, isa, cores = S.PortBackend(), "port", "scalar-c", _host_cores()
    dom = (size,) * 3
    be.setup(dom, tuple(cart))
    rng = np.random.default_rng(0x5EED)
    for r in range(len(be.store)):
        be.store[r][0][512:] = rng.random(be.store[r][0].size - 512)
    it = oracle.ST_ITER[stencil_id]

    def period():
        be.exchange()
        for s in range(it):
            be.sweep(stencil_id, s % 2, 1 - s % 2, skip=1 if s == it - 1 else 0)

    for _ in range(warm):
        period()
    t0 = time.perf_counter()
    for _ in range(periods):
        period()
    return (time.perf_counter() - t0) / periods, kind, cores, isa


def workload_name(stencil, size, it):
    return f"weak {stencil} {size}^3 per GPU, 8^3 bricks, 1 exchange + {it} sweeps per step"


def driver_env():
    """environment of the C++ drivers (one host thread per GPU).  OMP_PROC_BIND must not leak into them: with binding on,
    libgomp pins the initial thread to the first place at start-up and every rank thread created later inherits that
    single core -- measured: 595 instead of 1090 GStencil/s for the strong leg at N = 2"""
    env = dict(os.environ, OMP_NUM_THREADS=str(max(4, _host_cores() // 2)), OMP_PROC_BIND="false")
    env.pop("OMP_PLACES", None)
    return env


def leg_timeout(wd, cap):
    """seconds a driver leg may take: `cap`, but never more than what is left of the extras' deadline minus a reserve for the
    legs that follow; 0 = skip the leg"""
    if wd is None:
        return cap
    left = wd.left() - 75.0      # reserve: the end-to-end leg, the CPU baseline and printing the line
    return 0 if left < 20.0 else min(cap, left)


def drivers_accept_fused_variant(n):
    """The C++ driver legs inherit the fused kernel this run selected (BK_FUSED_VARIANT).  Before any of them is timed that
    way, both drivers run a small SELF-VALIDATING case with it (-v: against a CPU sweep of the global periodic array; the
    strong one on the stitched super grid, whose shell aliases other subdomains' bricks).  On any failure the legs fall back
    to the staged kernel.  Returns the record for the JSON line."""
    variant = os.environ.get("BK_FUSED_VARIANT")
    if not variant:
        return None
    out = {"variant": variant, "ok": True}
    cases = {"strong": ["-d", "128", "-s", "32", "-I", "2", "-g", str(n), "-S", "mpi7pt", "-v"],
             "weak": ["-s", "32,32,32", "-I", "2", "-g", str(n), "-S", "mpi7pt", "-v"]}
    for name, args in cases.items():
        try:
            r = subprocess.run([os.path.join(ROOT, "drivers", name), *args], capture_output=True, text=True, timeout=120,
                               env=driver_env())
            good = r.returncode == 0 and "result match (worst relative difference" in r.stdout
            out[name] = "result match" if good else (r.stdout + r.stderr)[-300:]
        except Exception as exc:
            good, out[name] = False, str(exc)[:300]
        out["ok"] = out["ok"] and good
    if not out["ok"]:
        os.environ.pop("BK_FUSED_VARIANT", None)
        out["fallback"] = "driver legs run with the staged kernel"
    return out


def strong_leg(n, wd=None):
    """BASELINE.json configs[4]: 1024^3 global, 64^3 subdomains, Z-Morton sections over n GPUs -- the C++ strong driver
    on ALL n GPUs (one host thread per GPU), stitched super grid; at n = 1 also one GPU's 1/8 share (512^3) both
    stitched and with per-subdomain launches (-M, the reference CUDA driver's structure)"""
    exe = os.path.join(ROOT, "drivers", "strong")
    out = {}

    def run(label, args, timeout):
        timeout = leg_timeout(wd, timeout)
        if not timeout:
            out[label] = {"skipped": "no time left inside the bench's deadline for extras"}
            return
        try:
            r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=timeout, env=driver_env())
            perf = [ln for ln in r.stdout.splitlines() if ln.startswith("perf ")]
            if perf:
                out[label] = {"GStencil/s": float(perf[-1].split()[1]), "cmd": "drivers/strong " + " ".join(args)}
                for key in ("calc ", "call ", "wait "):      # seconds per time step: [min, avg, max] over the ranks
                    ln = [x for x in r.stdout.splitlines() if x.startswith(key)]
                    if ln:
                        out[label][key.strip()] = ln[-1].split(":", 1)[1].split("(")[0].strip()
            else:
                out[label] = {"error": (r.stdout + r.stderr)[-300:]}
        except Exception as exc:
            out[label] = {"error": str(exc)[:300]}

    run("global_1024_sub_64", ["-d", "1024", "-s", "64", "-I", "10", "-g", str(n), "-S", "mpi7pt"], 150)
    run("global_1024_sub_64_mpi25pt", ["-d", "1024", "-s", "64", "-I", "10", "-g", str(n), "-S", "mpi25pt"], 150)
    if n == 1:
        run("share_512_stitched", ["-d", "512", "-s", "64", "-I", "20", "-g", "1", "-S", "mpi7pt"], 90)
        run("share_512_per_subdomain", ["-d", "512", "-s", "64", "-I", "20", "-g", "1", "-S", "mpi7pt", "-M"], 90)
    out["what"] = f"strong scaling: fixed 1024^3 global domain in 64^3 subdomains on {n} GPU(s); 10 exchange periods after 1 warm-up"
    return out


def array_baseline_leg(n, wd=None):
    """the reference's `Arr:` vs `Bri:` comparison through the C++ weak driver on all n GPUs -- the same
    512^3-per-GPU job on a plain array layout (arr_kernel + exchangeArr as one strided-box pull) and on bricks"""
    exe = os.path.join(ROOT, "drivers", "weak")
    out = {"what": f"drivers/weak -s 512,512,512 -I 10 -g {n} -S <stencil>: array-layout loop, then the brick loop, same input"}
    for name in ("mpi7pt", "mpi25pt"):
        timeout = leg_timeout(wd, 150)
        if not timeout:
            out[name] = {"skipped": "no time left inside the bench's deadline for extras"}
            continue
        try:
            r = subprocess.run([exe, "-s", "512,512,512", "-I", "10", "-g", str(n), "-S", name], capture_output=True, text=True,
                               timeout=timeout, env=driver_env())
            perf = [float(ln.split()[1]) for ln in r.stdout.splitlines() if ln.startswith("perf ")]
            if len(perf) == 2:
                out[name] = {"array_GStencil/s": perf[0], "brick_GStencil/s": perf[1],
                             "arr_equals_bri": "Arr == Bri: result match" in r.stdout}
            else:
                out[name] = {"error": (r.stdout + r.stderr)[-300:]}
        except Exception as exc:
            out[name] = {"error": str(exc)[:300]}
    return out


def direct_exchange_leg(n, wd=None):
    """THE EXCHANGE INSIDE THE SWEEP against the pull on this run's workload and all n GPUs (tools/direct_exchange_trial.py:
    two domains per rank on the same field, the same periods through both, results compared over the whole interior on the
    device, both timed).  A child job of n processes: the in-place kernels were written in a round without GPU time."""
    timeout = leg_timeout(wd, 200)
    if not timeout:
        return {"skipped": "no time left inside the bench's deadline for extras"}
    script = os.path.join(ROOT, "tools", "direct_exchange_trial.py")
    if n == 1:
        cmd = [sys.executable, script]
    else:
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), script]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT",
                                                            "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE",
                                                            "TORCHELASTIC_RUN_ID", "BK_FUSED_VARIANT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
        lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
        if r.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"ok": False, "rc": r.returncode, "tail": (r.stdout + r.stderr)[-500:]}
    except Exception as exc:
        return {"ok": False, "error": str(exc)[:300]}


def single_leg(wd=None):
    """BASELINE.json configs[0]: the single-GPU 7-point case of single/cuda.cpp (coeff[] stencil, in/out interleaved in one
    storage, step 1024), through the C++ single driver: kernel-only sweep rate + the host array sweep it validates against"""
    exe = os.path.join(ROOT, "drivers", "single")
    timeout = leg_timeout(wd, 120)
    if not timeout:
        return {"skipped": "no time left inside the bench's deadline for extras"}
    try:
        r = subprocess.run([exe, "-n", "512", "-s", "7pt", "-r", "50"], capture_output=True, text=True, timeout=timeout, env=driver_env())
        out = {"what": "drivers/single -n 512 -s 7pt -r 50 (one sweep per launch, interleaved storage)"}
        for ln in r.stdout.splitlines():
            f = ln.split()
            if ln.startswith("perf "):
                out["GStencil/s"], out["GB/s_algorithmic"] = float(f[1]), float(f[3])
            elif ln.startswith("Trans:"):
                out["sweep_ms"] = float(f[1]) * 1e3
            elif ln.startswith("Arr:"):
                out["host_array_sweep_s"] = float(f[1])
            elif ln.startswith("result"):
                out["validation"] = ln.strip()
        if "GStencil/s" not in out:
            out["error"] = (r.stdout + r.stderr)[-200:]
        return out
    except Exception as exc:
        return {"error": str(exc)[:200]}


def run_reference_arm(args):
    """the reference's CPU implementation of the SAME workload (N subdomains, periodic process grid) on rank 0 with
    every host core; the other ranks of a torchrun launch exit without work"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    st = oracle.STENCILS[args.stencil]
    size, n = args.size, args.gpus
    cart = CART.get(n)
    if cart is None:
        raise SystemExit("supported GPU counts: 1, 2, 4, 8")
    sec, kind, cores, isa = reference_period_seconds(st, size, args.steps, max(1, min(args.warmup, 1)), cart)
    it = oracle.ST_ITER[st]
    gst = size ** 3 * it * n / sec / 1e9
    line = {
        "impl": "reference", "metric": "GStencil/s", "value": gst, "unit": "GStencil/s", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.stencil, size, it), "process_grid": "x".join(map(str, cart)),
                   "note": f"reference CPU path (OpenMP + generated {isa} code): the {n} subdomain(s) of the job run in one "
                           f"process on {cores} host threads, exchange through its own BrickDecomp::exchange"},
        "cpu_baseline": {"value": gst, "unit": "GStencil/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} periods of {n} x {size}^3 after 1 warm-up"},
        "e2e": {"value": gst, "unit": "GStencil/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--stencil", default="mpi7pt", choices=["mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"])
    ap.add_argument("--size", type=int, default=512, help="cells per axis per GPU")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="one sweep per pass (no temporal blocking)")
    ap.add_argument("--fused", default="auto", choices=["auto", "staged", "composed", "wide"],
                    help="kernel behind the two-steps-per-pass launches of the radius-1 stars: k_star2 (staged), the composed "
                         "25-point diamond on 4x4- (composed) or 8x4-brick tiles (wide), or whichever is exact and fastest on "
                         "this box (auto)")
    ap.add_argument("--no-extras", action="store_true", help="skip other stencils / strong / e2e / cpu baseline / parity")
    ap.add_argument("--kernel", default="auto", choices=["auto", "brick", "tiled"])
    ap.add_argument("--transport", default="kernel", choices=["kernel", "ce"],
                    help="ghost exchange as one pull kernel over NVLink peer mappings (default, faster) or on the copy engines")
    ap.add_argument("--pull-shape", default="", metavar="CTAS,THREADS",
                    help="launch shape of the pull kernel, e.g. 32,1024 (narrow: leaves the other SMs to the overlapped sweep)")
    ap.add_argument("--thin", default="auto", choices=["auto", "on", "off"],
                    help="split sweeps with thin ghost-dependent k segments (BK_PART_THIN); auto = N>1 and radius <= 2")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import bricklib_b200 as bk
    rank, world, dist, host_group = dist_setup(args.gpus)
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
    L = bk.load()
    L.bk_bind_host_to_device()      # NUMA: pin this process next to its GPU before any pinned allocation (best effort)
    n = world
    cart = CART.get(n)
    if cart is None:
        raise SystemExit("supported GPU counts: 1, 2, 4, 8")
    coo = [(a, b, c) for a in range(cart[0]) for b in range(cart[1]) for c in range(cart[2])][rank]
    kernel = {"auto": bk.KERNEL_AUTO, "brick": bk.KERNEL_BRICK, "tiled": bk.KERNEL_TILED}[args.kernel]
    st = bk.STENCILS[args.stencil]
    size = args.size
    dom = (size,) * 3
    pts = size ** 3

    def make_domain():
        dm = bk.WeakDomain(dom, st, cart, coo, rank, kernel)
        wire_peers(bk, dm, dist, rank, world)
        if not args.no_overlap:
            dm.enable_overlap()
        if args.no_fuse:
            dm.fuse = 1
        dm.transport, dm.thin = args.transport, {"auto": None, "on": True, "off": False}[args.thin]
        if args.pull_shape:
            dm.set_pull_shape(*[int(x) for x in args.pull_shape.split(",")])
        dm.ready_first = ready_first
        if headline_thin is not None:
            dm.thin = headline_thin
        dm.fill_synthetic(0x5EED)          # synthetic U[0,1) field, written on the device
        bk.device_sync()
        return dm

    ready_first, headline_thin = False, None
    d = make_domain()
    fused_info = select_fused_kernel(bk, d, dist, rank, args.fused)
    if fused_info["selected"] != "staged":
        os.environ["BK_FUSED_VARIANT"] = fused_info["selected"]     # the C++ driver legs inherit the choice
    loop_info = select_loop_options(bk, d, dist, max(3, args.steps // 4))
    ready_first, headline_thin = d.ready_first, d.thin

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    sec, launches = time_periods(bk, d, args.steps, args.warmup, dist)
    clocks = sampler.stop() if rank == 0 else None
    it = d.st_iter
    value = pts * it * n * args.steps / sec / 1e9

    peak, peak_src = measured_peak()
    sweep_s, sweep_steps = time_sweeps(bk, d, 20)
    traffic = None
    composed = sweep_steps == 2 and fused_info["selected"] != "staged"
    try:
        if size == 512:     # the captures were taken at this size
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
                args.stencil + ("_composed2" if composed else "_fused2" if sweep_steps == 2 else ""))
    except Exception:
        pass

    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.stencil, size, it),
                   "process_grid": "x".join(map(str, cart)), "exchange_MB_per_gpu_per_step": d.view.bytes / 1e6,
                   "overlap": not args.no_overlap, "kernel": args.kernel, "thin_split": d._thin(),
                   "exchange_transport": "copy engines (faces) + narrow pull kernel (edges, corners) over NVLink peer mappings"
                   if d._remote() else "one pull kernel over NVLink peer mappings (CUDA IPC)", "steps_per_pass": d.steps_per_pass(),
                   "l2": f"inputs larger than L2: {2 * d.storage[0].dat.nbytes / 1e9:.2f} GB streamed per sweep"},
        "gpu_launches": launches,
        "roofline": roofline_of(pts, sweep_s, sweep_steps, peak, peak_src, traffic),
        "clocks": clocks,
    }
    line["roofline"]["traffic_source"] = ("ncu --set full capture of this kernel at this size (profiles/traffic.json), per launch"
                                          if traffic is not None else "no ncu capture of this kernel yet")
    if composed:
        line["roofline"]["kernel"] = (f"march_body<diamond, {fused_info['selected']}> (two time steps as ONE composed 25-point update): "
                                      f"one launch = {pts} interior points x 2 step(s) x 16 B")
    line["fused_kernel"] = fused_info
    line["loop_options"] = loop_info
    line["config"]["ready_first"] = bool(d.ready_first)

    wd = Watchdog(rank, float(os.environ.get("BENCH_EXTRAS_DEADLINE_S", "420")))
    wd.at("headline done", line)
    if not args.no_extras:
        # what was timed is also checked: sampled boxes vs the oracle, fused pass vs two sweeps on the device
        wd.at("parity")
        line["parity"] = parity_of(bk, d, dist)
        wd.at("others", line)
        others = {}
        line["others"] = others
        k_other = max(3, args.steps // 4)
        for name, sid in bk.STENCILS.items():
            if name in ("7pt", args.stencil):
                continue
            wd.at("others." + name)
            d.stencil, d.st_iter = sid, L.bk_stencil_st_iter(sid)
            d.fill_synthetic(0x5EED)
            d.thin = {"auto": None, "on": True, "off": False}[args.thin]
            opts = select_loop_options(bk, d, dist, k_other)
            s2, _ = time_periods(bk, d, k_other, 3, dist)
            k2, ks = time_sweeps(bk, d, 10)
            k2 = max_over_ranks(dist, k2)
            flops = 2 * L.bk_stencil_points(sid) - 1
            others[name] = {"GStencil/s": pts * d.st_iter * n * k_other / s2 / 1e9, "ms_per_step": s2 / k_other * 1e3,
                            "sweep_ms": k2 * 1e3, "steps_per_launch": ks, "hbm_GB/s": 16.0 * pts * ks / k2 / 1e9,
                            "frac_of_hbm_peak": 16.0 * pts * ks / k2 / 1e9 / peak,
                            "GFLOP/s_nominal": flops * pts * ks / k2 / 1e9,
                            "flops_note": f"nominal {flops} flop per point (FMA form, SURVEY 8d)" +
                            ("; the kernel EXECUTES ~50 flop per point (sign/permutation folds: 16 adds + 18 FMA), i.e. "
                             f"{50 * pts * ks / k2 / 1e12:.1f} TFLOP/s executed" if name == "mpi125pt" else ""),
                            "loop_options": opts, "parity": parity_of(bk, d, dist)}
            wd.at("others." + name + " done", line)
        d.stencil, d.st_iter = st, it
        d.ready_first, d.thin = ready_first, headline_thin

        # strong scaling (configs[4]), the array-layout baseline (8f#4) and the single driver (configs[0]) run through the
        # C++ drivers, one host thread per GPU, on ALL n GPUs: rank 0 runs them while the other ranks sit in a HOST
        # barrier (gloo: no kernel on any GPU; their contexts are idle).  They come before the end-to-end leg, which is
        # the one extra that keeps several domains with device-side handshakes in flight: whatever happens there, these
        # numbers are already in the snapshot.  Each leg gets what is left of the extras' deadline minus a reserve for the
        # legs that follow (and is skipped when that is too little).
        if rank == 0:
            wd.at("drivers: validation with the selected fused kernel", line)
            accepted = drivers_accept_fused_variant(n)
            if accepted is not None:
                line["fused_kernel"]["drivers"] = accepted
            wd.at("strong", line)
            others["strong"] = strong_leg(n, wd)
            wd.at("array baseline", line)
            others["array_layout_baseline"] = array_baseline_leg(n, wd)
            wd.at("exchange inside the sweep", line)
            others["exchange_inside_the_sweep"] = direct_exchange_leg(n, wd)
            if n == 1:
                wd.at("single", line)
                others["single_7pt_512"] = single_leg(wd)
            line["baseline_configs"] = {
                "configs[0] single 7pt 512^3 (N=1 only)": others.get("single_7pt_512", {}).get("GStencil/s"),
                "configs[1] 125pt 512^3 per GPU": others.get("mpi125pt", {}).get("GStencil/s"),
                "configs[2] weak 7pt / 13pt 512^3 per GPU": [value, others.get("mpi13pt", {}).get("GStencil/s")],
                "configs[3] weak 25pt 512^3 per GPU": others.get("mpi25pt", {}).get("GStencil/s"),
                "configs[4] strong 1024^3 in 64^3 subdomains": others["strong"].get("global_1024_sub_64", {}).get("GStencil/s"),
                "unit": f"GStencil/s, whole job on {n} GPU(s)",
            }
        wd.at("host barrier after the driver legs", line)
        if dist is not None:
            dist.barrier(group=host_group)

        wd.at("e2e", line)
        e2e_s, bi, bo = e2e_periods(bk, [d, make_domain(), make_domain()], 9)
        e2e_s = max_over_ranks(dist, e2e_s)
        line["e2e"] = {"value": pts * it * n / e2e_s / 1e9, "unit": "GStencil/s", "h2d_bytes_per_step": bi * n,
                       "d2h_bytes_per_step": bo * n, "ms_per_step": e2e_s * 1e3,
                       "h2d_GB/s_per_gpu": bi / e2e_s / 1e9, "d2h_GB/s_per_gpu": bo / e2e_s / 1e9,
                       "what": "per rank and step: H2D of the interior bricks from pinned host memory, exchange + sweeps, D2H "
                               "of the result bricks; three independent fields in flight, upload / compute / download on "
                               "separate streams; max over ranks"}
        wd.at("e2e done", line)
        if n == 1:
            wd.at("cpu_baseline")
            try:
                cs, kind, cores, isa = reference_period_seconds(st, size, 2, 1)
                line["cpu_baseline"] = {"value": pts * it / cs / 1e9, "unit": "GStencil/s", "cores": cores, "kind": kind,
                                        "sample": f"2 periods (exchange + {it} sweeps) of {size}^3 after 1 warm-up, {isa}"}
            except Exception as exc:  # the checker is optional for the product arm
                line["cpu_baseline"] = {"value": None, "unit": "GStencil/s", "cores": 0, "kind": "port",
                                        "sample": f"unavailable: {exc}"}
            wd.at("cpu_baseline done", line)

    if dist is not None:
        wd.at("leaving the process group", line)
        barrier(dist)
        dist.destroy_process_group()
    if rank != 0:
        wd.finish()
        return
    wd.finish()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
