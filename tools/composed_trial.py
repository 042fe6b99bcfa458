#!/usr/bin/env python3
"""First contact of the composed two-step kernel (BK_FUSED_COMPOSED, bricklib_b200/csrc/bk_diamond.h) with a device, in a
process of its own: bench.py runs this before it lets the kernel near the timed region, so that a fault or a hang of a
kernel that has not met this hardware yet costs one child process, not the benchmark.  Prints one JSON line.
   python tools/composed_trial.py [--device 0] [--size 512] [--stencil mpi7pt]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def small_boxes(bk):
    """every launch shape the library issues, on a small decomposition: whole grid (all six grid faces), interior, an
    off-centre box, and the READY / REST halves of a split launch with and without thin segments -- the current fused
    variant against two plain sweeps; returns the worst relative difference"""
    import numpy as np
    rng = np.random.default_rng(3)
    d = bk.BrickDecomp((40, 24, 32), 8)
    info = d.getBrickInfo()
    grid = bk.DeviceGrid(d.grid)
    s_in, s_tmp, s_ref, s_got = (info.allocate(512) for _ in range(4))
    h = rng.random(d.nbricks * 512)
    h[:512] = 0.0
    s_in.from_host(h)
    b_in, b_tmp, b_ref, b_got = (bk.Brick(info, s) for s in (s_in, s_tmp, s_ref, s_got))
    t = grid.dims
    own = ((1, 1, 1), tuple(x - 1 for x in t))
    worst = 0.0

    def rel(a, b):
        r = float((np.abs(a - b) / (np.abs(a) + np.abs(b) + 1e-300)).max())
        return r if r == r else float("inf")     # NaN anywhere is a failure, not "no difference"

    for lo, hi in [((0, 0, 0), t), own, ((1, 0, 1), (t[0], t[1] - 1, t[2]))]:
        bk.stencil(1, grid, b_in, b_tmp, kernel=bk.KERNEL_TILED)
        s_ref.dat.zero()
        bk.stencil(1, grid, b_tmp, b_ref, lo, hi, kernel=bk.KERNEL_TILED)
        s_got.dat.zero()
        bk.stencil_advance(1, 2, grid, b_in, b_got, lo, hi)
        bk.device_sync()
        want = s_ref.to_host()
        worst = max(worst, rel(s_got.to_host(), want))
        for thin in (0, bk.PART_THIN):
            parts = []
            for part in (bk.PART_READY, bk.PART_REST):
                s_got.dat.zero()
                bk.stencil_advance(1, 2, grid, b_in, b_got, lo, hi, own, part | thin)
                bk.device_sync()
                parts.append(s_got.to_host())
            worst = max(worst, rel(parts[0] + parts[1], want))
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--stencil", default="mpi7pt")
    a = ap.parse_args()
    import bench
    import bricklib_b200 as bk
    bk._lib.check(bk.load().bk_set_device(a.device))
    small = {}
    for name, variant in (("composed", bk.FUSED_COMPOSED), ("wide", bk.FUSED_COMPOSED_WIDE)):
        bk.fused_variant(variant)
        small[name] = small_boxes(bk)
    d = bk.WeakDomain((a.size,) * 3, bk.STENCILS[a.stencil])
    d.connect()
    out = {}
    for name, variant in (("staged", bk.FUSED_STAGED), ("composed", bk.FUSED_COMPOSED), ("wide", bk.FUSED_COMPOSED_WIDE)):
        bk.fused_variant(variant)
        bad, worst, pts = bench.fused_vs_two_sweeps(bk, d)       # vs two plain sweeps, whole interior, on the device
        sec, steps = bench.time_sweeps(bk, d, 5)
        out[name] = {"mismatches": int(bad), "max_rel": float(worst), "points": int(pts), "launch_ms": sec * 1e3,
                     "steps_per_launch": steps}
        if name in small:
            out[name]["small_boxes_max_rel"] = small[name]
            out[name]["ok"] = bool(bad == 0 and worst < 1e-12 and steps == 2 and small[name] < 1e-13)
    out["ok"] = bool(out["composed"]["ok"] or out["wide"]["ok"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
