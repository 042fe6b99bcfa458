"""The array-layout baseline (arr_kernel weak/main.cu:27-33 + exchangeArr array-mpi.h:146-213) and
interleaved fields (BrickDecomp numfield, brick-mpi.h:304-350) on the GPU: array path == brick path == oracle, the
equality every reference driver asserts between its Arr and Bri runs."""
import os

import numpy as np
import pytest

import bricklib_b200 as bk
import oracle
from oracle import schedule as S

pytestmark = pytest.mark.gpu
TOL = 1e-12
PAD = GZ = 8


def rel(a, b):
    return float((np.abs(a - b) / (np.abs(a) + np.abs(b) + 1e-300)).max())


@pytest.mark.parametrize("shape", [(32, 32, 32), (72, 24, 40), (8, 8, 8), (130, 6, 19)])
def test_array_sweep_against_oracle_port(shape):
    """one sweep of every stencil over a cell box that is not a multiple of anything"""
    P = oracle.port()
    rng = np.random.default_rng(sum(shape))
    o = 5   # radius 4 + 1: the box may sit anywhere the stencil still fits
    ext = tuple(x + 2 * o for x in shape)
    arr = rng.random(ext[::-1])
    coeff = rng.random(7)
    a_in, a_out = bk.DeviceBuffer.from_numpy(arr), bk.DeviceBuffer(arr.nbytes)
    for st in range(5):
        a_out.zero()
        lo, hi = (o,) * 3, tuple(o + x for x in shape)
        bk.array_stencil(st, a_in, a_out, ext, lo, hi, coeff)
        got = a_out.download(np.float64).reshape(arr.shape)
        want = P.sweep_array(st, arr, lo, hi, coeff)
        assert rel(got[o:-o, o:-o, o:-o], want[o:-o, o:-o, o:-o]) < TOL, st
        assert not got[:o].any() and not got[:, :o].any() and not got[:, :, :o].any()   # nothing outside the box


def test_array_sweep_refuses_a_box_without_room_for_the_radius():
    a = bk.DeviceBuffer(16 ** 3 * 8)
    b = bk.DeviceBuffer(16 ** 3 * 8)
    with pytest.raises(bk.BrickError):
        bk.array_stencil(3, a, b, (16, 16, 16), (3, 4, 4), (12, 12, 12))    # radius 4 needs lo >= 4
    with pytest.raises(bk.BrickError):
        bk.array_stencil(1, a, a, (16, 16, 16), (1, 1, 1), (15, 15, 15))    # in place


@pytest.mark.parametrize("name", ["mpi7pt", "mpi13pt", "mpi25pt", "mpi125pt"])
def test_array_loop_equals_brick_loop_equals_oracle(name):
    """the reference's in-driver equality: the Arr: time loop and the Bri: time loop produce the same interior"""
    st = bk.STENCILS[name]
    dom = (40, 24, 32)
    rng = np.random.default_rng(11)
    field = rng.random(dom[::-1])
    a = bk.ArrayDomain(dom, st)
    a.connect()
    a.load_interior(field)
    b = bk.WeakDomain(dom, st)
    b.connect()
    b.load_interior(field)
    for _ in range(2):
        a.period()
        b.period()
    bk.device_sync()
    want = S.periodic_steps(st, field, 2 * oracle.ST_ITER[st])
    got_a, got_b = a.read_interior(0), b.read_interior(0)
    assert rel(got_a, want) < TOL and rel(got_b, want) < TOL and rel(got_a, got_b) < TOL


@pytest.mark.parametrize("cart,dom", [((2, 1, 1), (24, 16, 32)), ((2, 2, 2), (16, 24, 16)), ((3, 1, 2), (16, 16, 16))])
def test_array_exchange_many_ranks_on_one_gpu(cart, dom):
    """N emulated ranks on one GPU in lock step: every ghost box is pulled from the right neighbour's interior"""
    rng = np.random.default_rng(31)
    glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
    fields = S.split_global(glob, cart, dom)
    coos = S.cart_coords(cart)
    for st in (1, 3, 4):
        doms = [bk.ArrayDomain(dom, st, cart, coo, r) for r, coo in enumerate(coos)]
        ptrs = {r: d.arr[0].ptr for r, d in enumerate(doms)}
        for d, f in zip(doms, fields):
            d.connect(ptrs)
            d.load_interior(f)
        for d in doms:
            d.view.exchange()
        bk.device_sync()
        for d in doms:
            d.sweeps()
        bk.device_sync()
        res = S.join_global([d.read_interior(0) for d in doms], cart, dom)
        assert rel(res, S.periodic_steps(st, glob, oracle.ST_ITER[st])) < TOL, st


def test_array_exchange_moves_exactly_the_brick_exchange_volume():
    """26 boxes, and as many bytes as the 42 brick ranges of the same subdomain (weak/main.cu prints both sizes)"""
    plan, ext = bk.ArrayExchangeView.boxes((512, 512, 512), (PAD,) * 3, (GZ,) * 3)
    assert len(plan) == 26 and ext == (544, 544, 544)
    assert sum(n[0] * n[1] * n[2] for _, _, _, n in plan) * 8 == 25352 * 4096


def test_two_interleaved_fields_share_one_storage_and_one_exchange():
    """BrickDecomp(dims, depth, numfield = 2) (brick-mpi.h:304-350): the chunk of a brick id holds both fields (step 1024),
    ONE exchange of the storage moves the ghost zones of both, and each field steps like a storage of its own"""
    dom, st = (32, 24, 40), bk.STENCILS["mpi13pt"]
    rng = np.random.default_rng(8)
    f0, f1 = rng.random(dom[::-1]), rng.random(dom[::-1])
    d = bk.BrickDecomp(dom, 8)
    d.populate((1, 1, 1), (0, 0, 0))
    info, grid = d.getBrickInfo(), bk.DeviceGrid(d.grid)
    sa, sb = info.allocate(2 * bk.BRICK), info.allocate(2 * bk.BRICK)
    fields = [(bk.Brick(info, sa, 0), bk.Brick(info, sb, 0)), (bk.Brick(info, sa, bk.BRICK), bk.Brick(info, sb, bk.BRICK))]
    ext = tuple(n + 2 * (PAD + GZ) for n in dom[::-1])
    o = PAD + GZ
    for (b_in, _), f in zip(fields, (f0, f1)):
        host = np.zeros(ext)
        host[o:-o, o:-o, o:-o] = f
        bk.copyToBrick(tuple(n + 2 * GZ for n in dom), (PAD,) * 3, (0,) * 3, bk.DeviceBuffer.from_numpy(host), grid, b_in)
    view = bk.ExchangeView(d, sa, {0: sa.dat.ptr}, 0)
    assert view.bytes == 2 * d.exchange_bytes()
    it = oracle.ST_ITER[st]
    for _ in range(2):
        view.exchange()
        for s in range(it):
            for pair in fields:
                src, dst = pair[s % 2], pair[1 - s % 2]
                bk.stencil(st, grid, src, dst)
    bk.device_sync()
    for (b_in, _), f in zip(fields, (f0, f1)):
        out = bk.DeviceBuffer(int(np.prod(ext)) * 8)
        out.zero()
        bk.copyFromBrick(dom, (PAD,) * 3, (GZ,) * 3, out, grid, b_in)
        got = out.download(np.float64).reshape(ext)[o:-o, o:-o, o:-o]
        assert rel(got, S.periodic_steps(st, f, 2 * it)) < TOL


@pytest.mark.parametrize("name,ranks,dom", [("mpi7pt", 1, "32,24,40"), ("mpi13pt", 4, "32,32,32"), ("mpi25pt", 2, "32,32,32"),
                                            ("mpi125pt", 8, "16,24,16")])
def test_cpp_weak_driver_prints_the_arr_block_and_arr_equals_bri(name, ranks, dom):
    """drivers/weak keeps the reference's two timed loops: `Arr:` (array layout) then `Bri:` (bricks), and closes with the
    reference's check that the two layouts hold the same field (weak/main.cu:161-213, :325-327)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "drivers", "weak"), "-s", dom, "-I", "2", "-g", str(ranks), "-S", name, "-v"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert out.index("Arr: ") < out.index("Bri: ")
    assert out.count("perf ") == 2 and "Arr == Bri: result match" in out and "result match (worst" in out
    for key in ("calc ", "pack ", "move ", "call ", "wait ", "  | MPI size (MB):", "  | MPI speed (GB/s):"):
        assert key in out.split("Bri: ")[0]
