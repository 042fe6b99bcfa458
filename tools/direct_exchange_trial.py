#!/usr/bin/env python3
"""The exchange INSIDE the sweep against the pull, on the weak-scaling workload (512^3 per GPU): every rank keeps two
domains on the same synthetic field -- one whose period pulls the neighbours' skins into its ghost bricks and then sweeps
(k_xplan + the marching kernels), one whose first pass reads the ghost bricks in place from the neighbours' storages
(bk_stencil_advance_remote: no pull, no ghost write, no re-read) -- runs the same periods on both, compares the two results
over the whole interior on the device and times both (CUDA events, max over ranks).  One JSON line on rank 0.

  python tools/direct_exchange_trial.py                                   # one GPU: the neighbours are the domain itself
  python -m torch.distributed.run --nproc-per-node N ... tools/direct_exchange_trial.py      # N GPUs over CUDA IPC / NVLink

bench.py runs it as a CHILD (rank 0, while the other ranks sit in a host barrier): the kernels behind it were written in a
round without GPU time, so a fault or a hang must cost this process, not the benchmark."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--periods", type=int, default=5)
    ap.add_argument("--stencils", default="mpi25pt,mpi13pt,mpi125pt,mpi7pt")
    a = ap.parse_args()
    import bench
    import bricklib_b200 as bk
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, world, dist, _ = bench.dist_setup(world)
    if world == 1:
        bk._lib.check(bk.load().bk_set_device(0))
    cart = bench.CART[world]
    coo = [(x, y, z) for x in range(cart[0]) for y in range(cart[1]) for z in range(cart[2])][rank]
    dom = (a.size,) * 3
    out = {"what": f"exchange inside the sweep vs pull, {a.size}^3 per GPU on {world} GPU(s), {a.periods} periods each, ms per period "
                   "(max over ranks)", "n_gpus": world}
    doms = {}
    for mode in ("pull", "direct"):         # ONE pair of domains for all stencils: peers keep the storages mapped
        d = bk.WeakDomain(dom, bk.STENCILS["mpi25pt"], cart, coo, rank)
        bench.wire_peers(bk, d, dist, rank, world)
        d.enable_overlap()
        d.exchange_mode = mode
        doms[mode] = d
    for name in a.stencils.split(","):
        st = bk.STENCILS[name]
        res = {}
        if name == "mpi7pt":                # two steps per pass: only the composed kernel can read ghosts in place
            bk.fused_variant(bk.FUSED_COMPOSED)
        try:
            for mode, d in doms.items():
                d.stencil, d.st_iter = st, bk.load().bk_stencil_st_iter(st)
                d.fill_synthetic(0xD1CE)
                d.storage[1].dat.zero()
                bk.device_sync()
                if mode == "direct" and not d.direct_active():
                    raise RuntimeError("no in-place kernel for this stencil / fused variant")
                sec, launches = bench.time_periods(bk, d, a.periods, 2, dist)
                res[mode + "_ms"] = sec / a.periods * 1e3
                res[mode + "_launches_per_period"] = launches // a.periods
            # same field, same number of periods on both: the results must be the same numbers
            p, q = doms["pull"], doms["direct"]
            t = p.grid.dims
            lo, hi = (1, 1, 1), tuple(x - 1 for x in t)
            # (the two domains have identical grids; compare through the pull domain's)
            ok, bad, rel = bk.compare_storage(p.grid, lo, hi, p.bricks[0], bk.Brick(p.info, q.storage[0], 0), 1e-13)
            res["mismatches"] = int(bench.sum_over_ranks(dist, bad))
            res["max_rel"] = bench.max_over_ranks(dist, rel)
            res["speedup"] = res["pull_ms"] / res["direct_ms"]
            res["ok"] = res["mismatches"] == 0
        except Exception as exc:
            res["error"] = f"{type(exc).__name__}: {str(exc)[:200]}"
            res["ok"] = False
        out[name] = res
        bench.barrier(dist)
    out["ok"] = all(v.get("ok") for k, v in out.items() if isinstance(v, dict))
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
