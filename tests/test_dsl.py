"""The stencil-script front end (bricklib_b200/st, dsl.py): the reference's stencils/*.py expressions lowered to tap lists,
pinned against the tap lists the REFERENCE's own DSL builds (tests/golden/stencil_taps.json, oracle/gen_golden_taps.py),
and -- on the GPU -- the compiled stencils against the oracle."""
import json
import os

import numpy as np
import pytest

import bricklib_b200 as bk
from bricklib_b200 import dsl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "stencil_scripts")
REF_STENCILS = "/root/reference/stencils"
COEFF7 = [0.31, 0.11, 0.12, 0.13, 0.14, 0.15, 0.04]


def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "stencil_taps.json")))


def value_of(text, consts):
    look = dsl._resolver(consts)
    neg = text.startswith("-")
    body = text[1:] if neg else text
    try:
        v = float(body)
    except ValueError:
        v = look(body)
    return -v if neg else v


def golden_taps(entry, consts):
    acc = {}
    for *offs, c in entry["taps"]:
        acc[tuple(offs)] = acc.get(tuple(offs), 0.0) + value_of(c, consts)
    return acc


@pytest.mark.parametrize("name", ["7pt", "mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"])
def test_shipped_scripts_lower_to_the_reference_tap_lists(name):
    consts = {"coeff": COEFF7}
    taps, sc = dsl.lower(name, consts)
    g = golden()[name + ".py"]
    assert sc.dims == g["dims"] and sc.in_grids == [g["in"]] and sc.out_grid == g["out"]
    assert dict(taps) == golden_taps(g, consts)


@pytest.mark.skipif(not os.path.isdir(REF_STENCILS), reason="reference tree not present")
def test_reference_scripts_run_unmodified_on_this_dsl():
    consts = {"coeff": COEFF7}
    for name, g in golden().items():
        path = os.path.join(REF_STENCILS, name)
        taps, sc = dsl.lower(path, consts)
        if not g["linear"]:      # cond.py: the reference's AST is not a plain sum; here it lowers with its two clamps
            assert name == "cond.py" and (sc.pre, sc.post) == (("max", 0.0), ("abs", 0.0))
            assert dict(taps) == dict(dsl.lower("7pt", consts)[0])
            continue
        assert sc.dims == g["dims"] and sc.pre is None and sc.post is None
        assert dict(taps) == golden_taps(g, consts), name


def test_cond_script_lowers_to_taps_with_two_pointwise_clamps():
    """stencils/cond.py: coeff[t] * max(in, 0.0) summed, then If(calc > 0, calc, -calc)"""
    consts = {"coeff": [0.5, -0.1, 0.2, -0.3, 0.4, 0.05, -0.6]}
    taps, sc = dsl.lower("cond", consts)
    assert (sc.pre, sc.post) == (("max", 0.0), ("abs", 0.0))
    assert dict(taps) == {(0, 0, 0): 0.5, (1, 0, 0): -0.1, (-1, 0, 0): 0.2, (0, 1, 0): -0.3, (0, -1, 0): 0.4,
                          (0, 0, 1): 0.05, (0, 0, -1): -0.6}


def test_linear_form_algebra():
    taps, sc = dsl.lower(os.path.join(HERE, "upwind.py"), {"W": 0.3})
    assert dict(taps) == {(0, 0, 0): 0.5, (-1, 0, 0): -0.3, (-2, 1, 0): 0.6, (1, -1, 3): 0.25, (0, 0, -3): -0.125,
                          (0, 2, 0): 0.25, (0, -2, 0): -0.25}
    assert sc.symbols == ["W"] and sc.in_grids == ["u"] and sc.out_grid == "v"
    with pytest.raises(dsl.LoweringError):
        dsl.lower(os.path.join(HERE, "upwind.py"))          # W unbound


def test_read_of_an_assigned_grid_is_shifted_to_the_point_it_is_read_at(tmp_path):
    """tmp(i+1,j,k) after tmp(...).assign(...) = tmp's definition moved by (+1,0,0) (reference: codegen/st/grid.py inlines
    the producer at the shifted index)"""
    head = "from st.expr import Index, ConstRef\nfrom st.grid import Grid\n" \
           "i, j, k = Index(0), Index(1), Index(2)\na, t, b = Grid('a', 3), Grid('t', 3), Grid('b', 3)\n"
    p = tmp_path / "two_stage.py"
    p.write_text(head + "t(i, j, k).assign(a(i + 1, j, k) - a(i - 1, j, k))\n"
                        "b(i, j, k).assign(0.5 * t(i + 1, j, k) + 0.25 * t(i, j - 2, k) + t(i, j, k))\nSTENCIL = [b]\n")
    taps, sc = dsl.lower(str(p))
    assert dict(taps) == {(2, 0, 0): 0.5, (0, 0, 0): -0.5, (1, -2, 0): 0.25, (-1, -2, 0): -0.25, (1, 0, 0): 1.0,
                          (-1, 0, 0): -1.0}
    q = tmp_path / "bad_arity.py"
    q.write_text(head + "t(i, j, k).assign(a(i, j, k))\nb(i, j, k).assign(t(i, j))\nSTENCIL = [b]\n")
    with pytest.raises(ValueError):
        dsl.lower(str(q))


def test_c_table_emission_for_cxx_callers():
    src = dsl.emit_c("mpi13pt", "mpi13")
    assert "static const bk_tap_t mpi13[] = {" in src and "mpi13_count = 13" in src
    assert "{0, 0, 0, 0.4}" in src and "{-2, 0, 0, 0.03}" in src


def test_nonlinear_and_malformed_scripts_are_refused(tmp_path):
    head = "from st.expr import Index, ConstRef, If\nfrom st.grid import Grid\nfrom st.func import Func\n" \
           "i, j, k = Index(0), Index(1), Index(2)\na, b = Grid('a', 3), Grid('b', 3)\n"
    cases = {
        "square": "b(i, j, k).assign(a(i, j, k) * a(i + 1, j, k))\nSTENCIL = [b]\n",
        "call": "b(i, j, k).assign(Func('sqrt', 1)(a(i, j, k)))\nSTENCIL = [b]\n",
        "select": "b(i, j, k).assign(If(a(i, j, k) > 0, a(i, j, k), 2 * a(i, j, k)))\nSTENCIL = [b]\n",
        "mixed_clamps": "b(i, j, k).assign(Func('max', 2)(a(i, j, k), 0.0) + a(i + 1, j, k))\nSTENCIL = [b]\n",
        "clamp_then_add": "b(i, j, k).assign(Func('fabs', 1)(a(i, j, k) + a(i + 1, j, k)) + a(i, j, k))\nSTENCIL = [b]\n",
        "constant": "b(i, j, k).assign(a(i, j, k) + 1.0)\nSTENCIL = [b]\n",
        "two_inputs": "c = Grid('c', 3)\nb(i, j, k).assign(a(i, j, k) + c(i, j, k))\nSTENCIL = [b]\n",
        "unassigned": "STENCIL = [b]\n",
    }
    for name, body in cases.items():
        p = tmp_path / (name + ".py")
        p.write_text(head + body)
        with pytest.raises(dsl.LoweringError):
            dsl.lower(str(p))
    p = tmp_path / "arity.py"
    p.write_text(head + "b(i, j, k).assign(a(i, j))\nSTENCIL = [b]\n")
    with pytest.raises(ValueError):
        dsl.lower(str(p))


def test_generator_emits_and_compiles_a_specialised_kernel_without_a_gpu(monkeypatch):
    """a script that is neither a star nor the symmetric cube becomes CUDA source specialised to its taps, compiled for
    sm_100a by NVRTC at bk_stencil_compile time (no device needed); BK_NO_CODEGEN keeps the tap-table kernel"""
    cs = bk.compile_stencil(os.path.join(HERE, "box27_skewed.py"))
    assert cs.kind == "generated" and cs.ntaps == 27
    src = cs.source()
    assert 'extern "C" __global__' in src and "bk_gen" in src and "cp.async.bulk" in src
    assert src.count("fma(cf.c[") == 27 * 2 * 2 * 3          # taps x cells of an x-pair x rows per thread x k slots
    assert "cf.c[26]" in src and "cf.c[27]" not in src and "#define WS 3" in src
    up = bk.compile_stencil(os.path.join(HERE, "upwind.py"), {"W": 0.3})
    assert up.kind == "generated" and "#define WS 7" in up.source() and "#define RY 2" in up.source()
    cond = bk.compile_stencil("cond", {"coeff": COEFF7})
    assert cond.kind == "generated" and "fmax(" in cond.source() and "fabs(" in cond.source()
    assert bk.compile_stencil("mpi13pt").source() is None      # built-in star kernel: nothing generated
    monkeypatch.setenv("BK_NO_CODEGEN", "1")
    assert bk.compile_stencil(os.path.join(HERE, "box27_skewed.py")).kind == "taps"


# ---- GPU: compiled stencils against the oracle ---------------------------------------------------------------------
PAD = GZ = 8


def run_compiled(cs, arr, n):
    nb = tuple((x + 2 * GZ) // 8 for x in n)
    grid_h, adj = bk.init_grid(nb)
    info = bk.BrickInfo(adj)
    grid = bk.DeviceGrid(grid_h)
    b_in, b_out = bk.Brick(info, info.allocate(512), 0), bk.Brick(info, info.allocate(512), 0)
    dev = bk.DeviceBuffer.from_numpy(arr)
    bk.copyToBrick(tuple(x + 2 * GZ for x in n), (PAD,) * 3, (0,) * 3, dev, grid, b_in)
    out = {}
    for label, kernel in (("auto", bk.KERNEL_AUTO), ("brick", bk.KERNEL_BRICK)):
        b_out.storage.dat.zero()
        cs.apply(grid, b_in, b_out, (1, 1, 1), tuple(x - 1 for x in nb), kernel)
        res = bk.DeviceBuffer(arr.nbytes)
        res.zero()
        bk.copyFromBrick(n, (PAD,) * 3, (GZ,) * 3, res, grid, b_out)
        out[label] = res.download(np.float64).reshape(arr.shape)
    return out


def rel(a, b):
    return float((np.abs(a - b) / (np.abs(a) + np.abs(b) + 1e-300)).max())


@pytest.mark.gpu
@pytest.mark.parametrize("script,consts,kind,radius", [
    ("7pt", {"coeff": COEFF7}, "star", 1), ("mpi7pt", None, "star", 1), ("mpi13pt", None, "star", 2),
    ("mpi25pt", None, "star", 4), ("mpi125pt", None, "cube", 2),
    (os.path.join(HERE, "star3.py"), {"c": [0.05 * (n + 1) for n in range(19)]}, "star", 3),
    (os.path.join(HERE, "box27.py"), {"w0": 0.2, "w1": 0.05, "w2": 0.02, "w3": 0.0325}, "cube", 1),
    (os.path.join(HERE, "box27_skewed.py"), None, "generated", 1),
    (os.path.join(HERE, "upwind.py"), {"W": 0.3}, "generated", 3),
])
def test_compiled_stencils_against_the_oracle(script, consts, kind, radius):
    from oracle import schedule as S
    cs = bk.compile_stencil(script, consts)
    assert (cs.kind, cs.radius) == (kind, radius)
    n = (40, 24, 32)
    rng = np.random.default_rng(17)
    arr = rng.random(tuple(x + 2 * (PAD + GZ) for x in n[::-1]))
    o = PAD + GZ
    lo, hi = (o, o, o), tuple(o + x for x in n)
    want = S.taps_sweep(arr, cs.taps, lo, hi)
    got = run_compiled(cs, arr, n)
    for label, res in got.items():
        assert rel(res[o:-o, o:-o, o:-o], want[o:-o, o:-o, o:-o]) < 1e-12, label


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_generated_kernels_for_random_tap_sets(seed):
    """random offsets (asymmetric k ranges, plane-only stencils, corners or not) and coefficients: generated marching
    kernel == per-brick tap-table kernel == numpy, including a split (READY + REST) launch of the generated kernel"""
    from oracle import schedule as S
    rng = np.random.default_rng(100 + seed)
    rx, ry = rng.integers(0, 5), rng.integers(0, 5)
    zlo, zhi = sorted(rng.integers(-4, 5, size=2))
    if seed == 0:
        zlo = zhi = 0                      # a purely in-plane stencil
    if seed == 1:
        rx, ry, zlo, zhi = 1, 1, -1, 1     # the dense 27-point box
    if seed == 1:
        offs = [(i, j, k) for k in (-1, 0, 1) for j in (-1, 0, 1) for i in (-1, 0, 1)]
    else:
        n = int(rng.integers(3, 30))
        offs = {(int(rng.integers(-rx, rx + 1)), int(rng.integers(-ry, ry + 1)), int(rng.integers(zlo, zhi + 1)))
                for _ in range(n)}
        offs = sorted(offs | {(int(rx), 0, int(zhi)), (0, -int(ry), int(zlo))})
    taps = [(o, float(rng.uniform(0.05, 1))) for o in offs]   # positive weights: no cancellation, the relative metric is meaningful
    from bricklib_b200 import _lib
    import ctypes as C
    arr_t = (_lib.Tap * len(taps))(*[_lib.Tap(o[0], o[1], o[2], c) for o, c in taps])
    h = C.c_void_p()
    _lib.check(bk.load().bk_stencil_compile(C.byref(h), arr_t, len(taps)))
    kind = C.c_int()
    bk.load().bk_stencil_def_info(h, C.byref(kind), None, None, None, None)
    star = all(sum(x != 0 for x in o) <= 1 for o, _ in taps)
    assert kind.value == (0 if star else 3)
    dims = (40, 24, 32)
    d = bk.BrickDecomp(dims, 8)
    info, grid = d.getBrickInfo(), bk.DeviceGrid(d.grid)
    s_in, s_a, s_b, s_c = (info.allocate(512) for _ in range(4))
    host = rng.random(d.nbricks * 512)
    host[:512] = 0
    s_in.from_host(host)
    b_in, b_a, b_b, b_c = (bk.Brick(info, s) for s in (s_in, s_a, s_b, s_c))
    f_ab = bk.core._field(b_in, b_a)
    t = grid.dims
    lo, hi = (0, 0, 0), t
    own = ((1, 1, 1), tuple(x - 1 for x in t))
    u3 = bk.core._u3
    L = bk.load()
    _lib.check(L.bk_stencil_def_advance(h, 1, C.byref(f_ab), grid.dev.ptr, u3(t), u3(lo), u3(hi), None, None, 0, bk.KERNEL_AUTO, None))
    f_b = bk.core._field(b_in, b_b)
    _lib.check(L.bk_stencil_def_advance(h, 1, C.byref(f_b), grid.dev.ptr, u3(t), u3(lo), u3(hi), None, None, 0, bk.KERNEL_BRICK, None))
    f_c = bk.core._field(b_in, b_c)
    for part in (bk.PART_READY | bk.PART_THIN, bk.PART_REST | bk.PART_THIN):
        _lib.check(L.bk_stencil_def_advance(h, 1, C.byref(f_c), grid.dev.ptr, u3(t), u3(lo), u3(hi), u3(own[0]), u3(own[1]), part,
                                            bk.KERNEL_AUTO, None))
    bk.device_sync()
    a, b, c = s_a.to_host(), s_b.to_host(), s_c.to_host()
    assert np.array_equal(a, c)                                       # split launch == whole launch, bit for bit
    assert float(np.abs(a - b).max()) < 1e-13                          # generated == tap-table kernel
    # against numpy on the interior bricks (the shell reads the null brick where the reference reads the same zeros)
    g = d.grid
    full = host.reshape(-1, 8, 8, 8)[g].transpose(0, 3, 1, 4, 2, 5).reshape(g.shape[0] * 8, g.shape[1] * 8, g.shape[2] * 8)
    got = a.reshape(-1, 8, 8, 8)[g].transpose(0, 3, 1, 4, 2, 5).reshape(full.shape)
    o = 8
    want = S.taps_sweep(full, taps, (o, o, o), tuple(x - o for x in full.shape[::-1]))
    assert rel(got[o:-o, o:-o, o:-o], want[o:-o, o:-o, o:-o]) < 1e-12
    L.bk_stencil_def_destroy(h)


@pytest.mark.gpu
def test_compiled_cond_stencil_against_the_reference_fixture():
    """stencils/cond.py on the GPU against one sweep of the reference's own generated code (tests/golden/cond_sweep.npz,
    oracle/gen_golden_cond.py), and against the numpy restatement on a larger field"""
    from oracle import schedule as S
    z = np.load(os.path.join(ROOT, "tests", "golden", "cond_sweep.npz"))
    cs = bk.compile_stencil("cond", {"coeff": z["coeff"]})
    assert cs.kind == "generated" and cs.pre == ("max", 0.0) and cs.post == ("abs", 0.0)
    o = PAD + GZ
    got = run_compiled(cs, z["input"], (16, 16, 16))
    for label, res in got.items():
        assert np.abs(res[o:-o, o:-o, o:-o] - z["out"]).max() < 1e-14, label
    n = (40, 24, 32)
    arr = np.random.default_rng(23).random(tuple(x + 2 * (PAD + GZ) for x in n[::-1])) * 2 - 1
    want = S.taps_sweep(arr, cs.taps, (o, o, o), tuple(o + x for x in n), cs.pre, cs.post)
    for label, res in run_compiled(cs, arr, n).items():
        assert np.abs(res[o:-o, o:-o, o:-o] - want[o:-o, o:-o, o:-o]).max() < 1e-14, label


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"])
def test_compiled_reference_stencils_equal_the_builtin_ids(name):
    """a script-compiled stencil and the BK_ST_* id of the same spec run the same kernel with the same coefficients"""
    cs = bk.compile_stencil(name)
    st = bk.STENCILS[name]
    d = bk.BrickDecomp((32, 40, 24), 8)
    info = d.getBrickInfo()
    grid = bk.DeviceGrid(d.grid)
    s_in, s_a, s_b = (info.allocate(512) for _ in range(3))
    h = np.random.default_rng(3).random(d.nbricks * 512)
    h[:512] = 0
    s_in.from_host(h)
    b_in, b_a, b_b = (bk.Brick(info, s) for s in (s_in, s_a, s_b))
    bk.stencil(st, grid, b_in, b_a)
    cs.apply(grid, b_in, b_b)
    bk.device_sync()
    assert np.array_equal(s_a.to_host(), s_b.to_host())
    if cs.fused_steps == 2:
        s_a.dat.zero(), s_b.dat.zero()
        bk.stencil_advance(st, 2, grid, b_in, b_a)
        cs.advance(2, grid, b_in, b_b)
        bk.device_sync()
        assert np.array_equal(s_a.to_host(), s_b.to_host())


# ---- property test: random linear expressions in varied surface syntax lower to the expected taps --------------------
from hypothesis import given, settings, strategies as hst  # noqa: E402

_off = hst.integers(min_value=-3, max_value=3)
_term = hst.tuples(_off, _off, _off, hst.integers(min_value=-8, max_value=8).filter(lambda v: v != 0),
                   hst.sampled_from(["c*u", "u*c", "neg", "div", "sym", "sub"]))


def _shift(name, d):
    return name if d == 0 else f"{name} {'+' if d > 0 else '-'} {abs(d)}"


@settings(max_examples=40, deadline=None)
@given(hst.lists(_term, min_size=1, max_size=10))
def test_random_linear_expressions_lower_to_their_taps(tmp_path_factory, terms):
    want = {}
    pieces = []
    for di, dj, dk, num, form in terms:
        ref = f"u({_shift('i', di)}, {_shift('j', dj)}, {_shift('k', dk)})"
        c = num / 4.0
        if form == "c*u":
            pieces.append(f"+ {c!r} * {ref}")
        elif form == "u*c":
            pieces.append(f"+ {ref} * {c!r}")
        elif form == "neg":
            pieces.append(f"+ (-({(-c)!r} * {ref}))")
        elif form == "div":
            pieces.append(f"+ {ref} / {(1.0 / c)!r}")
            c = 1.0 / (1.0 / c)
        elif form == "sym":          # a named constant times a literal
            pieces.append(f"+ S * {ref} * {num}")
            c = 0.25 * num
        else:                        # subtraction of a scaled reference
            pieces.append(f"- {(-c)!r} * {ref}")
        want[(di, dj, dk)] = want.get((di, dj, dk), 0.0) + c
    body = ("from st.expr import Index, ConstRef\nfrom st.grid import Grid\ni, j, k = Index(0), Index(1), Index(2)\n"
            "u, v = Grid('u', 3), Grid('v', 3)\nS = ConstRef('S')\nrhs = 0 " + " ".join(pieces) +
            "\nv(i, j, k).assign(rhs)\nSTENCIL = [v]\n")
    path = tmp_path_factory.mktemp("dsl") / "random_stencil.py"
    path.write_text(body)
    taps, sc = dsl.lower(str(path), {"S": 0.25})
    got = dict(taps)
    for key in set(want) | set(got):
        assert abs(got.get(key, 0.0) - want.get(key, 0.0)) < 1e-12, (key, body)
