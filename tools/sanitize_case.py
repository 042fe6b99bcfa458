#!/usr/bin/env python3
"""A small end-to-end case for compute-sanitizer (racecheck / memcheck / initcheck / synccheck): one overlapped exchange
period of every MPI stencil on a 32^3 subdomain (marching kernels: mbarrier ring, bulk copies, setmaxnreg, named barriers,
the fused two-step kernel, split READY/REST launches), a generated kernel, the array baseline, and a two-rank lock-step
peer-pointer exchange.  Prints `case ok` when the numbers also match the plain per-brick family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bricklib_b200 as bk

bk.load().bk_set_device(0)
dom = (32, 32, 32)
rng = np.random.default_rng(0)
field = rng.random(dom[::-1])
for name in ("mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"):
    st = bk.STENCILS[name]
    res = []
    for kernel in (bk.KERNEL_AUTO, bk.KERNEL_BRICK):
        d = bk.WeakDomain(dom, st, kernel=kernel)
        d.connect()
        d.enable_overlap()
        d.load_interior(field)
        d.period()
        bk.device_sync()
        res.append(d.read_interior(0))
    assert np.abs(res[0] - res[1]).max() < 1e-13, name
# generated kernel + array baseline
cs = bk.compile_stencil(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "stencil_scripts", "box27_skewed.py"))
d = bk.WeakDomain(dom, 1)
d.connect()
d.load_interior(field)
cs.apply(d.grid, d.bricks[0], d.bricks[1])
a = bk.ArrayDomain(dom, 1)
a.connect()
a.load_interior(field)
a.period()
# two ranks on one GPU in lock step: each pulls its ghost ranges out of the other's storage (the peer-pointer exchange;
# the spin-flag handshake is left out on purpose -- the sanitizer serialises kernels, a spinning kernel would never end)
doms = [bk.WeakDomain((16, 16, 16), 1, (2, 1, 1), (r, 0, 0), r) for r in range(2)]
ptrs = {r: x.storage[0].dat.ptr for r, x in enumerate(doms)}
for x in doms:
    x.connect(ptrs)
    x.fill_synthetic(7)
for _ in range(2):
    for x in doms:
        x.view.exchange()
    bk.device_sync()
    for s in range(8):
        for x in doms:
            x._sweep(s % 2, 1 - s % 2, (0, 0, 0), x.grid.dims, None)
    bk.device_sync()
print("case ok")
