#!/usr/bin/env python3
"""oracle/gen_golden_cond.py -- TEST INFRASTRUCTURE.  tests/golden/cond_sweep.npz from the UNMODIFIED reference build.

One sweep of the code the reference's generator (codegen/vecscatter, AVX backends) emits for stencils/cond.py --
    calc = sum_t coeff[t] * max(bIn(. + d_t), 0.0);   bOut = If(calc > 0, calc, -calc)
-- over a seeded 16^3 field with values of both signs and coefficients of both signs (so both clamps matter), in the
single/cpu.cpp configuration (init_grid layout, interleaved storage, step 1024).  Also asserts that the numpy
restatement the GPU tests use (oracle/schedule.py: taps_sweep with pre=("max",0), post=("abs",0)) reproduces it.
Run where /root/reference exists:  make -C oracle && python oracle/gen_golden_cond.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from oracle import schedule as S  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
COND = 5  # ref_sweep_brick id of cond.py (oracle/ref_harness.cpp)
TAPS = [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]  # coeff[0..6], stencils/cond.py:21-27


def main():
    R = oracle.ref()
    assert R is not None, "build oracle/_ref first"
    rng = np.random.default_rng(20261018)
    N, PAD, GZ = 16, 8, 8
    Sx = N + 2 * (PAD + GZ)
    arr = rng.random((Sx, Sx, Sx)) * 2.0 - 1.0
    coeff = rng.random(7) * 2.0 - 1.0
    NB = (N + 2 * GZ) // 8
    grid, adj = R.init_grid((NB, NB, NB))
    dat = oracle.aligned_zeros(NB ** 3 * 1024)
    R.copy_to_brick((N + 2 * GZ,) * 3, (PAD,) * 3, (0,) * 3, arr, grid, adj, dat, 1024, 0)
    R.sweep_brick(COND, grid, (1, 1, 1), (NB - 1,) * 3, adj, dat, 1024, 0, dat, 1024, 512, coeff)
    out = np.zeros_like(arr)
    R.copy_from_brick((N,) * 3, (PAD,) * 3, (GZ,) * 3, out, grid, adj, dat, 1024, 512)
    o = PAD + GZ
    got = np.ascontiguousarray(out[o:-o, o:-o, o:-o])
    want = S.taps_sweep(arr, list(zip(TAPS, coeff)), (o, o, o), (o + N,) * 3, ("max", 0.0), ("abs", 0.0))[o:-o, o:-o, o:-o]
    err = np.abs(got - want).max()
    assert err < 1e-14, err
    assert (got >= 0).all() and (arr < 0).any() and (coeff < 0).any()
    np.savez_compressed(os.path.join(OUT, "cond_sweep.npz"), input=arr, coeff=coeff, out=got)
    print("cond_sweep.npz written; numpy restatement vs reference generated code: max abs diff", err)


if __name__ == "__main__":
    main()
