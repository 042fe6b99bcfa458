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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--stencil", default="mpi7pt")
    a = ap.parse_args()
    import bench
    import bricklib_b200 as bk
    bk._lib.check(bk.load().bk_set_device(a.device))
    d = bk.WeakDomain((a.size,) * 3, bk.STENCILS[a.stencil])
    d.connect()
    out = {}
    for name, variant in (("staged", bk.FUSED_STAGED), ("composed", bk.FUSED_COMPOSED)):
        bk.fused_variant(variant)
        bad, worst, pts = bench.fused_vs_two_sweeps(bk, d)       # vs two plain sweeps, whole interior, on the device
        sec, steps = bench.time_sweeps(bk, d, 5)
        out[name] = {"mismatches": int(bad), "max_rel": float(worst), "points": int(pts), "launch_ms": sec * 1e3,
                     "steps_per_launch": steps}
    c = out["composed"]
    out["ok"] = bool(c["mismatches"] == 0 and c["max_rel"] < 1e-12 and c["steps_per_launch"] == 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
