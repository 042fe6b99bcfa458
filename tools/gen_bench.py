#!/usr/bin/env python3
"""Developer micro-benchmark of script-compiled stencils at 512^3: the generated marching kernel (BK_KIND_GENERATED)
against the per-brick tap-table kernel, one full-interior sweep each.
   python tools/gen_bench.py [--scripts tests/stencil_scripts/box27_skewed.py,...] [--reps 10]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bricklib_b200 as bk

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--scripts", default="tests/stencil_scripts/box27_skewed.py,tests/stencil_scripts/upwind.py,cond")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--peak", type=float, default=6545.3)
ap.add_argument("--brick", action="store_true", help="also time the per-brick tap-table kernel")
args = ap.parse_args()
bk.load().bk_set_device(0)
d = bk.WeakDomain((args.size,) * 3, 1)
d.connect()
d.fill_synthetic(1, 0)
t = d.grid.dims
lo, hi = (1, 1, 1), tuple(x - 1 for x in t)
pts = args.size ** 3
CONSTS = {"upwind.py": {"W": 0.3}, "cond": {"coeff": [0.5, -0.1, 0.2, -0.3, 0.4, 0.05, -0.6]},
          "box27.py": {"w0": 0.2, "w1": 0.05, "w2": 0.02, "w3": 0.0325}}
for script in args.scripts.split(","):
    cs = bk.compile_stencil(script, CONSTS.get(os.path.basename(script)))
    for label, kern in (("auto", bk.KERNEL_AUTO),) + ((("brick", bk.KERNEL_BRICK),) if args.brick else ()):
        for s in range(2):
            cs.apply(d.grid, d.bricks[s % 2], d.bricks[1 - s % 2], lo, hi, kern)
        bk.device_sync()
        e0, e1 = bk.Event(), bk.Event()
        e0.record()
        for s in range(args.reps):
            cs.apply(d.grid, d.bricks[s % 2], d.bricks[1 - s % 2], lo, hi, kern)
        e1.record(); e1.sync()
        ms = e0.elapsed_ms(e1) / args.reps
        print(f"{os.path.basename(script):18s} {cs.kind:9s} {cs.ntaps:4d} taps {label:5s} {ms:8.4f} ms {pts/ms/1e6:8.1f} GStencil/s "
              f"{16.0*pts/ms/1e6:8.1f} GB/s = {16.0*pts/ms/1e6/args.peak*100:5.1f}% of {args.peak:.0f}", flush=True)
