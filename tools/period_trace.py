#!/usr/bin/env python3
"""Timeline of one exchange period (developer aid): CUDA-event marks on the compute and exchange streams, averaged over
`--reps` periods, relative to the period start.  Works alone (N=1) or under torchrun like bench.py.

  python tools/period_trace.py --stencil mpi25pt [--transport ce|kernel] [--no-overlap]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stencil", default="mpi25pt")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--transport", default="kernel")
    ap.add_argument("--thin", action="store_true")
    ap.add_argument("--pull-shape", default="")
    ap.add_argument("--no-overlap", action="store_true")
    args = ap.parse_args()
    import bench
    import bricklib_b200 as bk
    rank, world, dist, _ = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
    cart = bench.CART[world]
    coo = [(a, b, c) for a in range(cart[0]) for b in range(cart[1]) for c in range(cart[2])][rank]
    d = bk.WeakDomain((args.size,) * 3, bk.STENCILS[args.stencil], cart, coo, rank)
    bench.wire_peers(bk, d, dist, rank, world)
    if not args.no_overlap:
        d.enable_overlap()
    d.transport, d.thin = args.transport, (True if args.thin else None)
    if args.pull_shape:
        d.set_pull_shape(*[int(x) for x in args.pull_shape.split(",")])
    host = np.random.default_rng(rank).random(d.decomp.nbricks * 512)
    host[:512] = 0
    d.storage[0].from_host(host)
    for _ in range(5):
        d.period()
    bk.device_sync()
    bench.barrier(dist)
    acc = {}
    for _ in range(args.reps):
        d.trace = []
        d.period()
        d.period()          # marks of the second period of a back-to-back pair
        bk.device_sync()
        marks = d.trace
        half = len(marks) // 2
        t0 = marks[half][1]
        for label, ev in marks[half:]:
            acc.setdefault(label, []).append(t0.elapsed_ms(ev))
        end_prev = marks[half][1]
        acc.setdefault("(previous period, start to start)", []).append(marks[0][1].elapsed_ms(end_prev))
    d.trace = None
    bench.barrier(dist)
    if rank == 0:
        print(json.dumps({"stencil": args.stencil, "world": world, "transport": args.transport,
                          "overlap": not args.no_overlap,
                          "ms": {k: round(float(np.mean(v)), 4) for k, v in acc.items()}}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
