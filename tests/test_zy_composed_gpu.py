"""bk_stencil_advance(steps=2) through the COMPOSED kernel (BK_FUSED_COMPOSED: two radius-1 steps as one 25-point diamond
update, bricklib_b200/csrc/bk_diamond.h) against two plain sweeps and against the oracle.  The file sorts last on purpose:
the kernel was written in a round without GPU time (its arithmetic and addressing are covered on the CPU by
tests/test_composed_emulation.py), so its first run on hardware must not hide the rest of the suite behind `-x`."""
import numpy as np
import pytest

import bricklib_b200 as bk

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((np.abs(a - b) / (np.abs(a) + np.abs(b) + 1e-300)).max())


@pytest.fixture(scope="module")
def first_contact():
    """the kernel's first run happens in a CHILD process (tools/composed_trial.py: every launch shape on a small
    decomposition, then parity and timing at 128^3): a fault or a hang there must not take this pytest process -- and the
    record of the whole suite -- with it.  The in-process tests below run only when the child came back clean."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "composed_trial.py"), "--size", "128"], capture_output=True,
                           text=True, timeout=240, cwd=root)
    except subprocess.TimeoutExpired:
        return {"ok": False, "why": "the child trial did not finish in 240 s"}
    lines = [x for x in r.stdout.splitlines() if x.startswith("{")]
    if r.returncode != 0 or not lines:
        return {"ok": False, "why": f"rc {r.returncode}: {(r.stdout + r.stderr)[-1500:]}"}
    return json.loads(lines[-1])


def test_first_contact_in_a_child_process(first_contact):
    assert first_contact["ok"], first_contact
    assert first_contact["composed"]["ok"] and first_contact["wide"]["ok"], first_contact


@pytest.fixture(params=["composed", "wide"])
def composed(first_contact, request):
    """both shapes of the composed kernel (4x4-brick tiles / 8x4-brick tiles), each only if its child trial came back clean"""
    if not first_contact.get(request.param, {}).get("ok"):
        pytest.skip(f"the {request.param} kernel failed its first contact (see test_first_contact_in_a_child_process)")
    variant = {"composed": bk.FUSED_COMPOSED, "wide": bk.FUSED_COMPOSED_WIDE}[request.param]
    before = bk.fused_variant(variant)
    yield variant
    bk.fused_variant(before)


COEFF7 = np.array([0.31, 0.11, 0.19, 0.23, 0.05, 0.07, 0.29])   # stencils/7pt.py: centre, i+1, i-1, j+1, j-1, k+1, k-1


@pytest.mark.parametrize("dom", [(40, 24, 32), (64, 64, 64), (16, 16, 16), (104, 40, 16)])
@pytest.mark.parametrize("st", [0, 1])
def test_composed_two_steps_equal_two_sweeps(composed, dom, st):
    """whole grid (every grid face: the zero intermediate outside the grid), interior, an off-centre box; split launches"""
    rng = np.random.default_rng(11)
    coeff = COEFF7 if st == 0 else None
    d = bk.BrickDecomp(dom, 8)
    info = d.getBrickInfo()
    grid = bk.DeviceGrid(d.grid)
    s_in, s_tmp, s_ref, s_got = (info.allocate(512) for _ in range(4))
    h = rng.random(d.nbricks * 512)
    h[:512] = 0.0
    s_in.from_host(h)
    b_in, b_tmp, b_ref, b_got = (bk.Brick(info, s) for s in (s_in, s_tmp, s_ref, s_got))
    t = grid.dims
    boxes = [((0, 0, 0), t), ((1, 1, 1), tuple(x - 1 for x in t)), ((1, 0, 1), (t[0], t[1] - 1, t[2]))]
    for lo, hi in boxes:
        if any(a >= b for a, b in zip(lo, hi)):
            continue
        bk.stencil(st, grid, b_in, b_tmp, coeff=coeff, kernel=bk.KERNEL_TILED)
        s_ref.dat.zero()
        bk.stencil(st, grid, b_tmp, b_ref, lo, hi, coeff=coeff, kernel=bk.KERNEL_TILED)
        s_got.dat.zero()
        bk.stencil_advance(st, 2, grid, b_in, b_got, lo, hi, coeff=coeff)
        bk.device_sync()
        want, got = s_ref.to_host(), s_got.to_host()
        assert rel(got, want) < 1e-14, (lo, hi)
        assert np.array_equal(got == 0.0, want == 0.0), "bricks outside the box must stay untouched"
        own = ((1, 1, 1), tuple(x - 1 for x in t))
        for thin in (0, bk.PART_THIN):
            s_got.dat.zero()
            bk.stencil_advance(st, 2, grid, b_in, b_got, lo, hi, own, bk.PART_READY | thin, coeff=coeff)
            bk.device_sync()
            part1 = s_got.to_host()
            s_got.dat.zero()
            bk.stencil_advance(st, 2, grid, b_in, b_got, lo, hi, own, bk.PART_REST | thin, coeff=coeff)
            bk.device_sync()
            part2 = s_got.to_host()
            assert not np.any((part1 != 0.0) & (part2 != 0.0)), "READY and REST overlap"
            assert rel(part1 + part2, want) < 1e-14


def test_composed_and_staged_kernels_agree(composed):
    """the two implementations of steps=2 on the same input, whole grid incl. the ghost shell"""
    rng = np.random.default_rng(12)
    d = bk.BrickDecomp((48, 40, 56), 8)
    info = d.getBrickInfo()
    grid = bk.DeviceGrid(d.grid)
    s_in, s_a, s_b = (info.allocate(512) for _ in range(3))
    h = rng.random(d.nbricks * 512)
    h[:512] = 0.0
    s_in.from_host(h)
    b_in, b_a, b_b = (bk.Brick(info, s) for s in (s_in, s_a, s_b))
    bk.stencil_advance(1, 2, grid, b_in, b_a)
    bk.fused_variant(bk.FUSED_STAGED)
    bk.stencil_advance(1, 2, grid, b_in, b_b)
    bk.fused_variant(composed)
    bk.device_sync()
    assert rel(s_a.to_host(), s_b.to_host()) < 1e-14


@pytest.mark.parametrize("overlap", [False, True])
def test_weak_period_through_the_composed_kernel_against_the_oracle(composed, overlap):
    """one rank, periodic: two exchange periods of mpi7pt (4 composed passes each) == 16 periodic steps of the oracle"""
    import oracle
    from oracle import schedule as S
    rng = np.random.default_rng(13)
    dom = (32, 24, 40)
    field = rng.random(dom[::-1])
    d = bk.WeakDomain(dom, 1)
    d.connect()
    if overlap:
        d.enable_overlap()
    d.load_interior(field)
    assert d.steps_per_pass() == 2
    for _ in range(2):
        d.period()
    bk.device_sync()
    want = S.periodic_steps(1, field, 2 * oracle.ST_ITER[1])
    assert rel(d.read_interior(0), want) < 1e-12


def test_full_size_composed_pass_equals_two_sweeps_on_a_random_field(composed):
    """512^3, what bench.py times when it selects the composed kernel: compared over the whole interior on the device"""
    import bench
    d = bk.WeakDomain((512, 512, 512), 1)
    d.connect()
    bad, worst, pts = bench.fused_vs_two_sweeps(bk, d)
    assert bad == 0 and worst < 1e-12 and pts == 512 ** 3
    got, n = bench.sampled_parity(bk, d)
    assert got < 1e-12 and n > 0


def _driver_composed(variant, *args):
    """a C++ driver in a process of its own with BK_FUSED_VARIANT=composed: self-validating against a CPU sweep of the
    global periodic array, like the reference's drivers"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "drivers", args[0])
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is not built: run __graft_entry__.build()")
    r = subprocess.run([exe, *args[1:]], capture_output=True, text=True, timeout=300, env=dict(os.environ, BK_FUSED_VARIANT=variant))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("variant", ["composed", "wide"])
@pytest.mark.parametrize("ranks,dom", [(1, "32,32,32"), (4, "32,24,40")])
def test_cpp_weak_driver_through_the_composed_kernel(first_contact, ranks, dom, variant):
    if not first_contact.get(variant, {}).get("ok"):
        pytest.skip(f"the {variant} kernel failed its first contact")
    out = _driver_composed(variant, "weak", "-s", dom, "-I", "2", "-g", str(ranks), "-S", "mpi7pt", "-v")
    assert "result match (worst relative difference" in out and "Arr == Bri: result match" in out


@pytest.mark.parametrize("variant", ["composed", "wide"])
@pytest.mark.parametrize("ranks,d,s", [(1, 64, 32), (8, 128, 32), (2, 128, 64)])
def test_cpp_strong_driver_stitched_grid_through_the_composed_kernel(first_contact, ranks, d, s, variant):
    """the stitched super grid aliases shell positions onto other subdomains' bricks: the composed kernel resolves every
    neighbour through grid POSITIONS exactly like the staged one, so the periodic result must come out the same"""
    if not first_contact.get(variant, {}).get("ok"):
        pytest.skip(f"the {variant} kernel failed its first contact")
    out = _driver_composed(variant, "strong", "-d", str(d), "-s", str(s), "-I", "2", "-g", str(ranks), "-S", "mpi7pt", "-v")
    assert "result match (worst relative difference" in out and f"stitched ranks {ranks} of {ranks}" in out
