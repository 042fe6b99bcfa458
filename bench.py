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

# bind OpenMP threads of the CPU reference leg -- only where a single process runs it
if int(os.environ.get("WORLD_SIZE", "1")) == 1:
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


def dist_setup(n_gpus):
    """one process per GPU under torchrun; returns (rank, world, torch.distributed or None)"""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1, None
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, dist


def max_over_ranks(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
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
    """W warm-up periods, then K timed periods between barriers; returns (seconds max over ranks, sweep seconds, launches)"""
    L = bk.load()
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
    """HBM roofline of one launch of the sweep kernel.  A launch reads and writes every interior point once whatever
    the number of time steps it fuses, so its compulsory traffic is 16 B x points; the single-sweep algorithmic figure
    of SURVEY 8(d) (16 B per point PER STEP) is reported next to it as `frac_of_single_sweep_roofline`."""
    achieved = 16.0 * pts / launch_s / 1e9
    kern = "k_star2 (two time steps per pass)" if steps == 2 else "k_star (one sweep)"
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": f"{kern}: one launch = 512^3 interior points x 16 B compulsory, {steps} step(s)",
            "kernel_ms": launch_s * 1e3, "steps_per_launch": steps, "peak_source": peak_src,
            "GStencil/s_kernel": pts * steps / launch_s / 1e9,
            "frac_of_single_sweep_roofline": 16.0 * pts * steps / launch_s / 1e9 / peak}


def e2e_periods(bk, doms, steps):
    """End to end through the public API with HOST buffers: every step uploads the field's interior bricks from pinned
    host memory (H2D), runs one period (exchange + ST_ITER sweeps) and downloads the result bricks (D2H) -- all inside
    the timed region.  Steps are independent fields, so two are kept in flight (double buffering on two streams): the
    upload of step i+1 and the download of step i-1 overlap the sweeps of step i, as a user streaming fields through
    the GPU would do.  Returns (seconds per step, h2d bytes, d2h bytes)."""
    import ctypes as C
    L = bk.load()
    d0 = doms[0]
    lo, hi = 1, d0.decomp.sep_pos[1]            # inner + skin bricks = the interior; ghosts come from the exchange
    off, nbytes = lo * 512 * 8, (hi - lo) * 512 * 8
    slots = []
    for d in doms:
        hin, hout, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
        bk._lib.check(L.bk_host_alloc(C.byref(hin), nbytes))
        bk._lib.check(L.bk_host_alloc(C.byref(hout), nbytes))
        bk._lib.check(L.bk_stream_create(C.byref(st)))
        bk._lib.check(L.bk_memcpy_d2h(hin, d.storage[0].dat.ptr + off, nbytes, None))
        slots.append((d, hin, hout, st))
    bk.device_sync()

    def step(i):
        d, hin, hout, st = slots[i % len(slots)]
        bk._lib.check(L.bk_memcpy_h2d(d.storage[0].dat.ptr + off, hin, nbytes, st))
        d.period(st)
        bk._lib.check(L.bk_memcpy_d2h(hout, d.storage[0].dat.ptr + off, nbytes, st))

    for i in range(len(slots)):                 # warm-up: one step per slot
        step(i)
    for _, _, _, st in slots:
        bk._lib.check(L.bk_stream_sync(st))
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    for _, _, _, st in slots:
        bk._lib.check(L.bk_stream_sync(st))
    sec = time.perf_counter() - t0
    for _, hin, hout, st in slots:
        L.bk_host_free(hin)
        L.bk_host_free(hout)
        L.bk_stream_destroy(st)
    return sec / steps, nbytes, nbytes


def reference_period_seconds(stencil_id, size, periods, warm=1):
    """the reference's own CPU implementation of the path (oracle/_ref: its BrickDecomp, its exchange() over the
    in-process MPI stand-in, its generated AVX brick code) on all host threads; else the C port.  Returns
    (seconds per period, kind, cores, isa)."""
    import oracle
    from oracle import schedule as S
    R = oracle.ref()
    if R is not None:
        be, kind, isa, cores = S.RefBackend(), "reference", R.isa, R.threads
    else:
        be, kind, isa, cores = S.PortBackend(), "port", "scalar-c", os.cpu_count()
    dom = (size,) * 3
    be.setup(dom, (1, 1, 1))
    rng = np.random.default_rng(0x5EED)
    be.store[0][0][512:] = rng.random(be.store[0][0].size - 512)
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


def strong_leg():
    """BASELINE.json configs[4] at one GPU's share (512 subdomains of 64^3 = the per-GPU work of 1024^3 on 8 GPUs):
    the C++ strong driver, stitched super grid (default) and per-subdomain launches (-M)"""
    exe = os.path.join(ROOT, "drivers", "strong")
    out = {}
    for label, extra in (("stitched", []), ("per_subdomain", ["-M"])):
        try:
            r = subprocess.run([exe, "-d", "512", "-s", "64", "-I", "20", "-g", "1", "-S", "mpi7pt", *extra],
                               capture_output=True, text=True, timeout=120)
            perf = [ln for ln in r.stdout.splitlines() if ln.startswith("perf ")]
            out[label] = {"GStencil/s": float(perf[-1].split()[1])} if perf else {"error": (r.stdout + r.stderr)[-200:]}
        except Exception as exc:
            out[label] = {"error": str(exc)[:200]}
    out["what"] = "drivers/strong -d 512 -s 64 -I 20 -g 1 -S mpi7pt (20 exchange periods of 8 steps after 1 warm-up)"
    return out


def single_leg():
    """BASELINE.json configs[0]: the single-GPU 7-point case of single/cuda.cpp (coeff[] stencil, in/out interleaved in one
    storage, step 1024), through the C++ single driver: kernel-only sweep rate + the host array sweep it validates against"""
    exe = os.path.join(ROOT, "drivers", "single")
    try:
        r = subprocess.run([exe, "-n", "512", "-s", "7pt", "-r", "50"], capture_output=True, text=True, timeout=180)
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    st = oracle.STENCILS[args.stencil]
    size = args.size
    sec, kind, cores, isa = reference_period_seconds(st, size, args.steps, max(1, min(args.warmup, 1)))
    it = oracle.ST_ITER[st]
    gst = size ** 3 * it / sec / 1e9
    line = {
        "impl": "reference", "metric": "GStencil/s", "value": gst, "unit": "GStencil/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"weak {args.stencil} {size}^3 per rank, 8^3 bricks, 1 exchange + {it} sweeps per step",
                   "ranks": 1, "note": "reference CPU path (OpenMP + generated %s code), single process" % isa},
        "cpu_baseline": {"value": gst, "unit": "GStencil/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} periods of {size}^3 after 1 warm-up"},
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
    ap.add_argument("--no-extras", action="store_true", help="skip other stencils / e2e / cpu baseline")
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
    rank, world, dist = dist_setup(args.gpus)
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
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
        rng = np.random.default_rng(0x5EED + rank)
        host = rng.random(dm.decomp.nbricks * 512)
        host[:512] = 0.0
        dm.storage[0].from_host(host)
        return dm

    d = make_domain()

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
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            args.stencil + ("_fused2" if sweep_steps == 2 else ""))
    except Exception:
        pass

    line = {
        "metric": "GStencil/s", "value": value, "unit": "GStencil/s", "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"weak {args.stencil} {size}^3 per GPU, 8^3 bricks, 1 exchange + {it} sweeps per step",
                   "process_grid": "x".join(map(str, cart)), "exchange_MB_per_gpu_per_step": d.view.bytes / 1e6,
                   "overlap": not args.no_overlap, "kernel": args.kernel, "thin_split": d._thin(),
                   "exchange_transport": "copy engines (faces) + narrow pull kernel (edges, corners) over NVLink peer mappings"
                   if d._remote() else "one pull kernel over NVLink peer mappings (CUDA IPC)", "steps_per_pass": d.steps_per_pass(),
                   "l2": f"inputs larger than L2: {2 * d.storage[0].dat.nbytes / 1e9:.2f} GB streamed per sweep"},
        "gpu_launches": launches,
        "roofline": roofline_of(pts, sweep_s, sweep_steps, peak, peak_src, traffic),
        "clocks": clocks,
    }

    if n == 1 and not args.no_extras:
        others = {}
        for name, sid in bk.STENCILS.items():
            if name in ("7pt", args.stencil):
                continue
            d.stencil, d.st_iter = sid, bk.load().bk_stencil_st_iter(sid)
            s2, _ = time_periods(bk, d, max(3, args.steps // 4), 3, None)
            k2, ks = time_sweeps(bk, d, 10)
            others[name] = {"GStencil/s": pts * d.st_iter * max(3, args.steps // 4) / s2 / 1e9,
                            "sweep_ms": k2 * 1e3, "steps_per_launch": ks, "hbm_GB/s": 16.0 * pts / k2 / 1e9,
                            "frac_of_hbm_peak": 16.0 * pts / k2 / 1e9 / peak,
                            "GFLOP/s": (2 * bk.load().bk_stencil_points(sid) - 1) * pts * ks / k2 / 1e9}
        d.stencil, d.st_iter = st, it
        others["strong_512_in_64_subdomains"] = strong_leg()
        others["single_7pt_512"] = single_leg()
        line["others"] = others
        # the five BASELINE.json configs, as measured in THIS run (one GPU's share of the multi-GPU ones)
        line["baseline_configs"] = {
            "configs[0] single 7pt 512^3": others["single_7pt_512"].get("GStencil/s"),
            "configs[1] single-GPU 125pt 512^3": others.get("mpi125pt", {}).get("GStencil/s"),
            "configs[2] weak 7pt / 13pt 512^3 per GPU": [value, others.get("mpi13pt", {}).get("GStencil/s")],
            "configs[3] weak 25pt 512^3 per GPU": others.get("mpi25pt", {}).get("GStencil/s"),
            "configs[4] strong 64^3 subdomains, one GPU's 512^3 share": others["strong_512_in_64_subdomains"].get(
                "stitched", {}).get("GStencil/s"),
            "unit": "GStencil/s at N=1; N=2/4/8: rerun with --gpus N (--stencil ...), drivers/strong -g 8",
        }
        e2e_s, bi, bo = e2e_periods(bk, [d, make_domain()], 8)
        line["e2e"] = {"value": pts * it / e2e_s / 1e9, "unit": "GStencil/s", "h2d_bytes_per_step": bi,
                       "d2h_bytes_per_step": bo, "ms_per_step": e2e_s * 1e3,
                       "what": "per step: H2D of the interior bricks from pinned host memory, exchange + sweeps, D2H of "
                               "the result bricks; two independent fields in flight (double buffered on two streams)"}
        try:
            cs, kind, cores, isa = reference_period_seconds(st, size, 2, 1)
            line["cpu_baseline"] = {"value": pts * it / cs / 1e9, "unit": "GStencil/s", "cores": cores, "kind": kind,
                                    "sample": f"2 periods (exchange + {it} sweeps) of {size}^3 after 1 warm-up, {isa}"}
        except Exception as exc:  # the checker is optional for the product arm
            line["cpu_baseline"] = {"value": None, "unit": "GStencil/s", "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {exc}"}
    elif n > 1 and not args.no_extras:
        e2e_s, bi, bo = e2e_periods(bk, [d, make_domain()], 6)
        e2e_s = max_over_ranks(dist, e2e_s)
        line["e2e"] = {"value": pts * it * n / e2e_s / 1e9, "unit": "GStencil/s", "h2d_bytes_per_step": bi * n,
                       "d2h_bytes_per_step": bo * n, "ms_per_step": e2e_s * 1e3,
                       "what": "per rank and step: H2D of the interior bricks from pinned host memory, exchange + sweeps, "
                               "D2H of the result bricks; two fields in flight; max over ranks"}

    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        barrier(dist)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
