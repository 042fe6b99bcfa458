#!/usr/bin/env python3
"""One launch, many exchange variants (developer aid; run under torchrun like bench.py, or alone at N=1).

Builds ONE weak domain per rank and times the exchange period of every stencil under several pull-kernel launch shapes,
with and without thin ghost-dependent k segments -- one process-group start-up for the whole table.

  python -m torch.distributed.run --nproc-per-node 8 ... tools/exchange_sweep.py --steps 40
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
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--stencils", default="mpi7pt,mpi25pt,mpi13pt,mpi125pt")
    ap.add_argument("--shapes", default="0x0,32x1024,16x1024,64x1024,64x512")
    args = ap.parse_args()
    import bench
    import bricklib_b200 as bk
    rank, world, dist, _ = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
    cart = bench.CART[world]
    coo = [(a, b, c) for a in range(cart[0]) for b in range(cart[1]) for c in range(cart[2])][rank]
    d = bk.WeakDomain((args.size,) * 3, bk.STENCILS["mpi7pt"], cart, coo, rank)
    bench.wire_peers(bk, d, dist, rank, world)
    d.enable_overlap()
    host = np.random.default_rng(rank).random(d.decomp.nbricks * 512)
    host[:512] = 0
    d.storage[0].from_host(host)
    pts = args.size ** 3
    rows = []
    for name in args.stencils.split(","):
        d.stencil = bk.STENCILS[name]
        d.st_iter = bk.load().bk_stencil_st_iter(d.stencil)
        for shape in args.shapes.split(","):
            ctas, thr = [int(x) for x in shape.split("x")]
            for thin in (False, True):
                d.set_pull_shape(ctas, thr)
                d.thin = thin
                sec, _ = bench.time_periods(bk, d, args.steps, 3, dist)
                rows.append({"stencil": name, "shape": shape, "thin": thin, "ms_per_period": sec / args.steps * 1e3,
                             "GStencil/s": pts * d.st_iter * world * args.steps / sec / 1e9})
                if rank == 0:
                    print(json.dumps(rows[-1]), flush=True)
    if dist is not None:
        bench.barrier(dist)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
