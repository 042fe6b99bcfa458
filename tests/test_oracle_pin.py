"""Pins the C port (oracle/oracle.c) to the reference: against the committed fixtures that oracle/gen_golden.py
produced from the unmodified reference build, and -- when oracle/_ref is present -- against that build directly."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle
from oracle import schedule as S

TOL = 1e-12  # north_star: FP64 relative tolerance


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rel(a, b):
    d = np.abs(a - b)
    return float((d / (np.abs(a) + np.abs(b) + 1e-300)).max())


@pytest.fixture(scope="module")
def P():
    return oracle.port()


def test_decomp_tables_match_reference_fixture(P, golden_dir):
    tables = json.load(open(os.path.join(golden_dir, "decomp_tables.json")))
    for key, t in tables.items():
        dom = tuple(int(x) for x in key.split("x"))
        d = P.decomp(dom)
        assert d["nbricks"] == t["nbricks"] and list(d["sep_pos"]) == t["sep_pos"], key
        assert [list(r) for r in d["ghost"]] == t["ghost"], key
        assert [list(r) for r in d["skin"]] == t["skin"], key
        assert d["skin_size"] == t["skin_size"], key
        assert sha(d["grid"]) == t["grid_sha256"], key
        assert sha(d["adj"][1:]) == t["adj1_sha256"], key


def test_survey_appendix_a_numbers(P):
    """the rows SURVEY.md Appendix A quotes for 512^3 (measured with the reference's BrickDecomp)"""
    d = P.decomp((512, 512, 512))
    assert d["nbricks"] == 287497 and d["sep_pos"] == (238329, 262145, 287497)
    assert len(d["ghost"]) == 42 and sum(g[4] for g in d["ghost"]) == 25352
    assert (d["ghost"][0][3], d["ghost"][0][4], d["skin"][0][3]) == (262145, 1, 254143)
    assert (d["ghost"][5][3], d["ghost"][5][4], d["skin"][5][3]) == (262275, 3970, 246267)
    assert (d["ghost"][41][3], d["ghost"][41][4], d["skin"][41][3]) == (287496, 1, 242298)
    assert d["skin_size"][:7] == [3844, 62, 1, 62, 1, 62, 3844]


def test_rank_maps_match_reference_fixture(P, golden_dir):
    maps = json.load(open(os.path.join(golden_dir, "rank_maps.json")))
    for key, m in maps.items():
        cart, coo = [tuple(int(x) for x in part.split(",")) for part in key.split("@")]
        mine = P.rank_map(cart, coo)
        for s, r in m.items():
            assert mine[int(s)] == r, (key, s)


def test_init_grid_matches_reference_fixture(P, golden_dir):
    ig = json.load(open(os.path.join(golden_dir, "init_grid.json")))
    for key, t in ig.items():
        dims = tuple(int(x) for x in key.split("x"))
        g, a = P.init_grid(dims)
        assert sha(g) == t["grid_sha256"] and sha(a) == t["adj_sha256"], key


def test_single_sweep_matches_reference_output(P, golden_dir):
    z = np.load(os.path.join(golden_dir, "single_sweep.npz"))
    arr, coeff = z["input"], z["coeff"]
    N, PAD, GZ = 16, 8, 8
    NB = (N + 2 * GZ) // 8
    grid, adj = P.init_grid((NB, NB, NB))
    o = PAD + GZ
    for name, st in oracle.STENCILS.items():
        dat = np.zeros(NB ** 3 * 1024)
        P.copy_brick(0, (N + 2 * GZ,) * 3, (PAD,) * 3, (0,) * 3, arr, grid, dat, 1024, 0)
        P.sweep_brick(st, grid, (1, 1, 1), (NB - 1,) * 3, adj, dat, 1024, 0, dat, 1024, 512, coeff)
        out = np.zeros_like(arr)
        P.copy_brick(1, (N,) * 3, (PAD,) * 3, (GZ,) * 3, out, grid, dat, 1024, 512)
        assert rel(out[o:-o, o:-o, o:-o], z["out_" + name]) < TOL, name
        arr_form = P.sweep_array(st, arr, (o,) * 3, (o + N,) * 3, coeff)
        assert rel(arr_form[o:-o, o:-o, o:-o], z["out_" + name]) < TOL, name


@pytest.mark.parametrize("cart", [(1, 1, 1), (2, 1, 1)])
def test_weak_time_loop_matches_reference_output(P, golden_dir, cart):
    z = np.load(os.path.join(golden_dir, "weak_steps.npz"))
    dom = (24, 16, 32)
    tag = "c%d%d%d" % cart
    glob = z["in_" + tag]
    for name, st in oracle.STENCILS.items():
        if st == 0:
            continue
        res = S.weak_run(S.PortBackend(), st, dom, cart, 2, S.split_global(glob, cart, dom), skip_last=True)
        assert rel(S.join_global(res, cart, dom), z["out_%s_%s" % (tag, name)]) < TOL, name


# ---- direct comparison with the compiled reference (skipped where oracle/_ref is absent) ------------------------
needs_ref = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built for this CPU")


@needs_ref
def test_port_vs_compiled_reference_decomp(P):
    R = oracle.ref()
    for dom in [(16, 24, 32), (64, 64, 64), (40, 16, 72)]:
        a, b = P.decomp(dom), R.decomp(dom)
        assert a["nbricks"] == b["nbricks"] and a["sep_pos"] == b["sep_pos"]
        assert np.array_equal(a["grid"], b["grid"]) and np.array_equal(a["adj"][1:], b["adj"][1:])
        assert a["ghost"] == b["ghost"] and a["skin"] == b["skin"] and a["skin_size"] == b["skin_size"]
        assert all(p == (0, 0) for p in b["ghost_pad"])


@needs_ref
def test_port_vs_compiled_reference_sweeps(P):
    R = oracle.ref()
    rng = np.random.default_rng(3)
    N, PAD, GZ = 24, 8, 8
    Sx = N + 2 * (PAD + GZ)
    arr = rng.random((Sx, Sx, Sx))
    coeff = rng.random(7)
    NB = (N + 2 * GZ) // 8
    grid, adj = P.init_grid((NB, NB, NB))
    o = PAD + GZ
    for st in range(5):
        dp_, dr = oracle.aligned_zeros(NB ** 3 * 512), oracle.aligned_zeros(NB ** 3 * 512)
        op, orr = oracle.aligned_zeros(NB ** 3 * 512), oracle.aligned_zeros(NB ** 3 * 512)
        P.copy_brick(0, (N + 2 * GZ,) * 3, (PAD,) * 3, (0,) * 3, arr, grid, dp_, 512, 0)
        R.copy_to_brick((N + 2 * GZ,) * 3, (PAD,) * 3, (0,) * 3, arr, grid, adj, dr, 512, 0)
        P.sweep_brick(st, grid, (1, 1, 1), (NB - 1,) * 3, adj, dp_, 512, 0, op, 512, 0, coeff)
        R.sweep_brick(st, grid, (1, 1, 1), (NB - 1,) * 3, adj, dr, 512, 0, orr, 512, 0, coeff)
        a, b = np.zeros_like(arr), np.zeros_like(arr)
        P.copy_brick(1, (N,) * 3, (PAD,) * 3, (GZ,) * 3, a, grid, op, 512, 0)
        R.copy_from_brick((N,) * 3, (PAD,) * 3, (GZ,) * 3, b, grid, adj, orr, 512, 0)
        assert rel(a[o:-o, o:-o, o:-o], b[o:-o, o:-o, o:-o]) < TOL, st
        assert R.compare_brick((N,) * 3, (PAD,) * 3, (GZ,) * 3, a, grid, adj, orr, 512, 0)


@needs_ref
def test_port_vs_compiled_reference_weak_loop_mirrored_cart(P):
    """3 ranks along k exposes the c-1 <-> +axis pairing of populate() (brick-mpi.h:740-751)"""
    rng = np.random.default_rng(11)
    cart, dom = (3, 1, 2), (16, 16, 16)
    glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
    fields = S.split_global(glob, cart, dom)
    a = S.join_global(S.weak_run(S.RefBackend(), 2, dom, cart, 1, fields), cart, dom)
    b = S.join_global(S.weak_run(S.PortBackend(), 2, dom, cart, 1, fields), cart, dom)
    assert rel(a, b) < TOL
    assert rel(b, S.periodic_steps(2, glob, 4)) < TOL


def test_cond_restatement_matches_reference_generated_code(golden_dir):
    """the numpy meaning of a script with pointwise clamps (oracle/schedule.py: taps_sweep, pre/post) against one sweep
    of the code the reference's generator emits for stencils/cond.py: committed fixture, and the build itself if present"""
    z = np.load(os.path.join(golden_dir, "cond_sweep.npz"))
    taps = [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
    N, o = 16, 16
    want = S.taps_sweep(z["input"], list(zip(taps, z["coeff"])), (o, o, o), (o + N,) * 3, ("max", 0.0), ("abs", 0.0))
    assert np.abs(want[o:-o, o:-o, o:-o] - z["out"]).max() < 1e-14
    R = oracle.ref()
    if R is None:
        return
    NB = (N + 16) // 8
    grid, adj = R.init_grid((NB, NB, NB))
    dat = oracle.aligned_zeros(NB ** 3 * 1024)
    R.copy_to_brick((N + 16,) * 3, (8,) * 3, (0,) * 3, z["input"], grid, adj, dat, 1024, 0)
    R.sweep_brick(5, grid, (1, 1, 1), (NB - 1,) * 3, adj, dat, 1024, 0, dat, 1024, 512, z["coeff"])
    out = np.zeros_like(z["input"])
    R.copy_from_brick((N,) * 3, (8,) * 3, (8,) * 3, out, grid, adj, dat, 1024, 512)
    assert np.array_equal(out[o:-o, o:-o, o:-o], z["out"])
