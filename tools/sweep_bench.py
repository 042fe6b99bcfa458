#!/usr/bin/env python3
"""Developer micro-benchmark: time one full-interior sweep (512^3 by default) per stencil and kernel family.
   python tools/sweep_bench.py [--size 512] [--stencils mpi7pt,mpi13pt] [--kernels auto,brick] [--reps 20] [--full]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bricklib_b200 as bk

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--stencils", default="mpi7pt,mpi13pt,mpi25pt,mpi125pt")
ap.add_argument("--kernels", default="auto")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--full", action="store_true", help="sweep the ghost shell too (66^3 bricks)")
ap.add_argument("--peak", type=float, default=6650.0)
ap.add_argument("--variants", default="")
ap.add_argument("--kls", default="")
ap.add_argument("--steps", type=int, default=1, help="time steps per pass (2 = fused kernel)")
args = ap.parse_args()
K = {"auto": bk.KERNEL_AUTO, "brick": bk.KERNEL_BRICK, "tiled": bk.KERNEL_TILED}
bk.load().bk_set_device(0)
d = bk.WeakDomain((args.size,) * 3, 1)
d.connect()
rng = np.random.default_rng(1)
h = rng.random(d.decomp.nbricks * 512); h[:512] = 0
d.storage[0].from_host(h); d.storage[1].from_host(h)
t = d.grid.dims
lo, hi = ((0, 0, 0), t) if args.full else ((1, 1, 1), tuple(x - 1 for x in t))
pts = args.size ** 3
import itertools
for name in args.stencils.split(","):
  for var, kl in itertools.product(args.variants.split(","), args.kls.split(",")):
    if var:
        os.environ["BK_STAR_VARIANT"] = var
    if kl:
        os.environ["BK_STAR_KL"] = kl
    for kn in args.kernels.split(","):
        d.stencil, d.kernel = bk.STENCILS[name], K[kn]
        def sweep(a, b):
            if args.steps == 1:
                d._sweep(a, b, lo, hi, None)
            else:
                bk.stencil_advance(d.stencil, args.steps, d.grid, d.bricks[a], d.bricks[b], lo, hi)
        for s in range(3):
            sweep(s % 2, 1 - s % 2)
        bk.device_sync()
        e0, e1 = bk.Event(), bk.Event()
        e0.record()
        for s in range(args.reps):
            sweep(s % 2, 1 - s % 2)
        e1.record(); e1.sync()
        ms = e0.elapsed_ms(e1) / args.reps
        gbs = 16.0 * pts / ms / 1e6
        ms /= args.steps
        print(f"{name:9s} v{var or '0'} kl{kl or '16'} {kn:6s} {'full' if args.full else 'inner'} {ms:8.4f} ms  {pts/ms/1e6:8.1f} GStencil/s  {gbs:8.1f} GB/s alg  {gbs/args.peak*100:5.1f}% of {args.peak:.0f}", flush=True)
