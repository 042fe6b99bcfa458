#!/usr/bin/env python3
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Writes tests/golden/* from the UNMODIFIED reference build
(oracle/_ref, see oracle/Makefile + ref_harness.cpp).  Run in the build container, where /root/reference exists:

    make -C oracle && python oracle/gen_golden.py

The reference ships no golden vectors of its own, so these are outputs OF the reference:
  decomp_tables.json  BrickDecomp<3,8,8,8> numbering for several subdomain shapes: nbricks, sep_pos, the 42 ghost/skin
                      region rows, skin_size, sha256 of the grid and of adj[1:] (brick-mpi.h:304-460)
  rank_maps.json      populate() neighbour-set -> rank maps for periodic Cartesian grids (brick-mpi.h:730-753)
  init_grid.json      sha256 of init_grid<3> grid/adjacency (bricksetup.h:73-90)
  single_sweep.npz    one sweep of the reference's generated brick code for all five stencils on a seeded 16^3 field
                      (init_grid layout, interleaved storage step 1024 -- the single/cpu.cpp configuration)
  weak_steps.npz      the weak/ time loop (reference exchange() + generated code) for 1 and 2 ranks, 24x16x32 cells
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import schedule as S  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = oracle.ref()
    assert R is not None, "build oracle/_ref first"
    os.makedirs(OUT, exist_ok=True)

    tables = {}
    for dom in [(16, 16, 16), (24, 16, 32), (64, 64, 64), (128, 64, 32), (512, 512, 512)]:
        d = R.decomp(dom)
        tables["x".join(map(str, dom))] = {
            "nbricks": d["nbricks"], "sep_pos": d["sep_pos"], "tdims": d["tdims"],
            "ghost": d["ghost"], "skin": d["skin"], "skin_size": d["skin_size"],
            "ghost_pad": d["ghost_pad"], "grid_sha256": sha(d["grid"]), "adj1_sha256": sha(d["adj"][1:]),
        }
    json.dump(tables, open(os.path.join(OUT, "decomp_tables.json"), "w"), indent=0)

    maps = {}
    for cart in [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 1, 2), (4, 3, 2)]:
        for coo in S.cart_coords(cart):
            d = R.decomp((16, 16, 16), 8, cart, coo)
            m = {}
            for (s, *_), p in list(zip(d["ghost"], d["ghost_peer"])) + list(zip(d["skin"], d["skin_peer"])):
                m[str(s)] = p
            maps["%d,%d,%d@%d,%d,%d" % (cart + coo)] = m
    json.dump(maps, open(os.path.join(OUT, "rank_maps.json"), "w"), indent=0)

    ig = {}
    for dims in [(4, 4, 4), (6, 5, 4), (10, 10, 10)]:
        g, a = R.init_grid(dims)
        ig["x".join(map(str, dims))] = {"grid_sha256": sha(g), "adj_sha256": sha(a)}
    json.dump(ig, open(os.path.join(OUT, "init_grid.json"), "w"), indent=0)

    # ---- single sweep, single/cpu.cpp configuration at N=16 -------------------------------------------------
    rng = np.random.default_rng(20261017)
    N, PAD, GZ = 16, 8, 8
    Sx = N + 2 * (PAD + GZ)
    arr = rng.random((Sx, Sx, Sx))
    coeff = rng.random(129)
    NB = (N + 2 * GZ) // 8
    grid, adj = R.init_grid((NB, NB, NB))
    single = {"input": arr, "coeff": coeff}
    for name, st in oracle.STENCILS.items():
        dat = oracle.aligned_zeros(NB ** 3 * 1024)
        R.copy_to_brick((N + 2 * GZ,) * 3, (PAD,) * 3, (0,) * 3, arr, grid, adj, dat, 1024, 0)
        R.sweep_brick(st, grid, (1, 1, 1), (NB - 1,) * 3, adj, dat, 1024, 0, dat, 1024, 512, coeff)
        out = np.zeros_like(arr)
        R.copy_from_brick((N,) * 3, (PAD,) * 3, (GZ,) * 3, out, grid, adj, dat, 1024, 512)
        single["out_" + name] = np.ascontiguousarray(out[PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ])
        if st > 0:  # the reference's scalar form must agree with its generated code
            ref_arr = R.sweep_array(st, arr, (PAD + GZ,) * 3, (PAD + GZ + N,) * 3)
            assert np.allclose(ref_arr[PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ],
                               single["out_" + name], rtol=1e-13, atol=0)
    np.savez_compressed(os.path.join(OUT, "single_sweep.npz"), **single)

    # ---- weak time loop ---------------------------------------------------------------------------------------
    weak = {}
    dom = (24, 16, 32)
    for cart in [(1, 1, 1), (2, 1, 1)]:
        glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
        tag = "c%d%d%d" % cart
        weak["in_" + tag] = glob
        for name, st in oracle.STENCILS.items():
            if st == 0:
                continue
            res = S.weak_run(S.RefBackend(), st, dom, cart, 2, S.split_global(glob, cart, dom))
            g = S.join_global(res, cart, dom)
            chk = S.periodic_steps(st, glob, 2 * oracle.ST_ITER[st])
            assert np.allclose(g, chk, rtol=1e-13, atol=0), (cart, name)
            weak["out_%s_%s" % (tag, name)] = g
    np.savez_compressed(os.path.join(OUT, "weak_steps.npz"), **weak)
    print("golden written to", OUT, {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})


if __name__ == "__main__":
    main()
