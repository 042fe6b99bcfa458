#!/usr/bin/env python3
"""Parity of the ONE-PROCESS-PER-GPU weak time loop (the path `bench.py --gpus N` times under torchrun) with the
oracle.  Launched by tests/test_multi_gpu.py as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/mgpu_weak_check.py --size 32 --periods 2

Every rank builds the same periodic global field (seeded), loads its own block (BrickDecomp + populate on the
MPI_Dims_create process grid, weak/args.cpp:101-111), wires its neighbours' storage through CUDA IPC exactly as
bench.py does, runs `periods` exchange periods with overlap and fused passes enabled, and compares its block of the
result with the oracle's periodic global sweep (oracle/schedule.py: periodic_steps).  Tolerance 1e-12 relative.
Rank 0 prints one JSON line {"ok": true, "max_rel": ...}; a mismatch on any rank makes every rank exit non-zero.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOL = 1e-12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--periods", type=int, default=2)
    ap.add_argument("--stencils", default="mpi7pt,mpi13pt,mpi25pt,mpi125pt")
    ap.add_argument("--transport", default="kernel", choices=["kernel", "ce"])
    args = ap.parse_args()

    import bench
    import bricklib_b200 as bk
    import oracle
    from oracle import schedule as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, world, dist, _ = bench.dist_setup(world)
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
    cart = bench.CART[world]
    coo = S.cart_coords(cart)[rank]
    dom = (args.size,) * 3
    rng = np.random.default_rng(0xB200)
    glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
    mine = S.split_global(glob, cart, dom)[rank]
    o = S.global_origin(cart, coo, dom)
    worst, bad = 0.0, []
    # ONE domain for all stencils: peers keep its storage mapped through CUDA IPC for the life of the process
    d = bk.WeakDomain(dom, bk.STENCILS[args.stencils.split(",")[0]], cart, coo, rank)
    bench.wire_peers(bk, d, dist, rank, world)
    d.enable_overlap()
    d.transport = args.transport
    for name in args.stencils.split(","):
        st = bk.STENCILS[name]
        d.stencil, d.st_iter = st, bk.load().bk_stencil_st_iter(st)
        d.storage[0].dat.zero()
        d.storage[1].dat.zero()
        d.load_interior(mine)
        bench.barrier(dist)
        for _ in range(args.periods):
            d.period()
        bk.device_sync()
        bench.barrier(dist)
        got = d.read_interior(0)
        want = S.periodic_steps(st, glob, oracle.ST_ITER[st] * args.periods)
        want = want[o[2]:o[2] + dom[2], o[1]:o[1] + dom[1], o[0]:o[0] + dom[0]]
        err = float((np.abs(got - want) / (np.abs(got) + np.abs(want) + 1e-300)).max())
        err = bench.max_over_ranks(dist, err)
        worst = max(worst, err)
        if not err < TOL:
            bad.append((name, err))
        bench.barrier(dist)
    if rank == 0:
        print(json.dumps({"ok": not bad, "max_rel": worst, "world": world, "cart": cart, "size": args.size,
                          "periods": args.periods, "bad": bad}))
    if dist is not None:
        dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
