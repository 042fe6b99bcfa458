#!/usr/bin/env python3
"""The device-side flag handshake of the one-process-per-GPU weak loop (bricklib_b200/weak.py: bk_flags_signal, the
k_wait + pull of bk_xplan_run_gate / _run_sync, bk_flags_wait) exercised on ONE GPU: `--ranks` emulated ranks, each with
its own compute and exchange stream and its own flag buffer, run whole exchange periods concurrently -- nothing but the
flags orders rank r's pull against its neighbours' sweeps, exactly as between processes (there the pointers are CUDA-IPC
mappings; here they are plain device pointers).  The result must equal the lock-step loop (exchange everybody, host
sync, sweep everybody) to rounding (1e-14), and the oracle's periodic global sweep to 1e-12.  Prints `handshake ok`.
Runs in a process of its own under a timeout (tests/test_zx_handshake_gpu.py): a protocol bug is a hang.  Test
infrastructure (it checks against the oracle), hence under tests/."""
import argparse
import ctypes as C
import os
import sys

# every emulated rank has two streams of its own; with the default of 8 hardware queues shared by all streams a kernel that
# waits for a flag could sit in FRONT of the very kernel that raises it (submission order is rank by rank).  Between
# processes this cannot happen -- each process owns its queues.  Before the CUDA context exists:
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# The first marching launch over a new (adjacency, grid, box) triple checks on the device that grid and adjacency agree and
# SYNCHRONISES its stream (bk_stencil.cu: marching_matches_adjacency).  Between processes that is harmless; here ONE host
# thread feeds all the ranks, and blocking inside rank 0's period -- whose stream waits for a flag rank 1 raises -- before
# rank 1's period has been enqueued would never return (tests/hostdev.py models exactly this).  The grids are BrickDecomp's.
os.environ["BK_SKIP_ADJ_CHECK"] = "1"

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", type=int, default=2, choices=[2, 4, 8])
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--periods", type=int, default=3)
    ap.add_argument("--stencils", default="mpi7pt,mpi25pt,mpi125pt")
    ap.add_argument("--no-overlap", action="store_true", help="bk_xplan_run_sync on the compute stream instead of the gated pull")
    a = ap.parse_args()
    import bricklib_b200 as bk
    from bricklib_b200.weak import Handshake
    import oracle
    from oracle import schedule as S
    L = bk.load()
    bk._lib.check(L.bk_set_device(0))
    cart = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[a.ranks]
    coords = S.cart_coords(cart)
    dom = (a.size,) * 3
    rng = np.random.default_rng(0xB200)
    glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
    blocks = S.split_global(glob, cart, dom)

    def build(st):
        doms = [bk.WeakDomain(dom, st, cart, coords[r], r) for r in range(a.ranks)]
        ptrs = {r: x.storage[0].dat.ptr for r, x in enumerate(doms)}
        for r, x in enumerate(doms):
            x.load_interior(blocks[r])
        return doms, ptrs

    worst = 0.0
    for name in a.stencils.split(","):
        st = bk.STENCILS[name]
        # (1) lock step, no flags: the order is imposed from the host
        ref, ptrs = build(st)
        for x in ref:
            x.connect(ptrs)
        for _ in range(a.periods):
            for x in ref:
                x.view.exchange()
            bk.device_sync()
            fuse = ref[0].steps_per_pass()
            for p in range(ref[0].st_iter // fuse):
                last = p == ref[0].st_iter // fuse - 1
                for x in ref:
                    t = x.grid.dims
                    lo, hi = ((1, 1, 1), tuple(v - 1 for v in t)) if last else ((0, 0, 0), t)
                    x._advance(fuse, p % 2, 1 - p % 2, lo, hi, None, bk.PART_ALL, None)
                bk.device_sync()
        want = [x.read_interior(0) for x in ref]
        # (2) every rank on its own streams, ordered by the device-side flags only
        doms, ptrs = build(st)
        hs = [Handshake(a.ranks) for _ in range(a.ranks)]
        streams = []
        for r, x in enumerate(doms):
            if hasattr(L, "process"):
                L.process = r        # tests/hostdev.py (the CPU stand-in for the device): whose streams these are
            for q in range(a.ranks):
                hs[r].peer[q] = hs[q].buf.ptr
            x.connect(ptrs, hs[r])
            if not a.no_overlap:
                x.enable_overlap()
            s = C.c_void_p()
            bk._lib.check(L.bk_stream_create(C.byref(s)))
            streams.append(s)
        bk.device_sync()
        for _ in range(a.periods):          # the host never waits: all periods of all ranks are enqueued up front
            for x, s in zip(doms, streams):
                x.period(s)
        bk.device_sync()
        for r, x in enumerate(doms):
            got = x.read_interior(0)
            same = float((np.abs(got - want[r]) / (np.abs(got) + np.abs(want[r]) + 1e-300)).max())
            assert same < 1e-14, f"{name}: rank {r} differs from the lock-step loop by {same}"
            o = S.global_origin(cart, coords[r], dom)
            gold = S.periodic_steps(st, glob, oracle.ST_ITER[st] * a.periods)[o[2]:o[2] + dom[2], o[1]:o[1] + dom[1], o[0]:o[0] + dom[0]]
            err = float((np.abs(got - gold) / (np.abs(got) + np.abs(gold) + 1e-300)).max())
            assert err < 1e-12, f"{name}: rank {r} vs oracle {err}"
            worst = max(worst, err)
        for s in streams:
            L.bk_stream_destroy(s)
    print(f"handshake ok: {a.ranks} ranks on one GPU, {a.periods} periods, max rel vs oracle {worst:.2e}")


if __name__ == "__main__":
    main()
