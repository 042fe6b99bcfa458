"""N>1 host logic on CPU: one process per rank over torch.distributed `gloo` (world_size 2 and 4).

What runs in every rank is the PRODUCT's host logic -- BrickDecomp numbering, populate()'s neighbour->rank map and
ExchangeView.plan (the list of (peer, src offset, dst offset, bytes) pulls a GPU rank would issue over NVLink,
bricklib_b200/core.py) -- through the C ABI.  No GPU here, so the transport is stood in by gloo (every rank publishes
its storage, the plan is applied with plain copies) and the sweeps by the oracle; the point is that plan + rank map,
built independently in separate processes, deliver exactly the ghost data the reference's exchange delivers
(brick-mpi.h:466-495), which is checked against the periodic global sweep after two exchange periods.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, cart, dom, stencil, periods, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import bricklib_b200 as bk
    import oracle
    from oracle import schedule as S

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coo = S.cart_coords(cart)[rank]
        # ---- product host logic (C ABI, CPU only) ----
        dec = bk.BrickDecomp(dom, 8)
        dec.populate(cart, coo)
        plan = bk.ExchangeView.plan(dec, 512)
        assert len(plan) == 42 and sum(p[3] for p in plan) == dec.exchange_bytes()
        # every rank must agree on who pulls what from whom: my pulls from p == what p expects to serve me
        allplans = [None] * world
        dist.all_gather_object(allplans, plan)
        for r, pl in enumerate(allplans):
            for peer, src, dst, nb in pl:
                assert 0 <= peer < world
                assert src + nb <= dec.sep_pos[1] * 4096 and src >= dec.sep_pos[0] * 4096      # skin range of the peer
                assert dst >= dec.sep_pos[1] * 4096 and dst + nb <= dec.nbricks * 4096       # my ghost range

        # ---- checker side: storage, sweeps (oracle), transport (gloo) ----
        P = oracle.port()
        rng = np.random.default_rng(99)
        glob = rng.random((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
        mine = S.split_global(glob, cart, dom)[rank]
        store = [oracle.aligned_zeros(dec.nbricks * 512), oracle.aligned_zeros(dec.nbricks * 512)]
        ext = tuple(n + 16 for n in dom)
        arr = np.zeros(tuple(n + 32 for n in dom[::-1]))
        arr[16:-16, 16:-16, 16:-16] = mine
        P.copy_brick(0, ext, (8,) * 3, (0,) * 3, arr, dec.grid, store[0], 512)
        it = oracle.ST_ITER[stencil]
        t = dec.tdims
        for _ in range(periods):
            pub = [torch.empty(dec.nbricks * 512, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(pub, torch.from_numpy(store[0].copy()))
            for peer, src, dst, nb in plan:
                store[0][dst // 8:(dst + nb) // 8] = pub[peer].numpy()[src // 8:(src + nb) // 8]
            for s in range(it):
                P.sweep_brick(stencil, dec.grid, (0, 0, 0), t, dec.adj, store[s % 2], 512, 0, store[1 - s % 2], 512, 0)
        out = np.zeros_like(arr)
        P.copy_brick(1, dom, (8,) * 3, (8,) * 3, out, dec.grid, store[0], 512)
        res = torch.from_numpy(np.ascontiguousarray(out[16:-16, 16:-16, 16:-16]))
        parts = [torch.empty_like(res) for _ in range(world)]
        dist.all_gather(parts, res)
        if rank == 0:
            got = S.join_global([p.numpy() for p in parts], cart, dom)
            want = S.periodic_steps(stencil, glob, it * periods)
            q.put(float((np.abs(got - want) / (np.abs(got) + np.abs(want))).max()))
    except Exception as exc:  # surface the failure in the parent
        if rank == 0:
            q.put(repr(exc))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("cart,stencil", [((2, 1, 1), 1), ((1, 2, 1), 3), ((1, 1, 2), 4), ((2, 2, 1), 2)])
def test_plan_and_rank_map_across_processes(cart, stencil):
    import torch.multiprocessing as mp
    world = cart[0] * cart[1] * cart[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    dom = (16, 16, 16)
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, cart, dom, stencil, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err = q.get(timeout=5)
    assert isinstance(err, float) and err < 1e-12, err


def _strong_rank_main(rank, world, port, dom, subdim, stencil, periods, q):
    """strong scaling, stitched: every process owns a Z-Morton section of subdomains, builds ITS stitched grid and pull
    plan through the C ABI (bk_stitch_*, bricklib_b200.strong_pull_plan) and runs the period schedule of
    drivers/strong.cpp -- exchange of the box-surface regions, then ST_ITER sweeps over the stitched grid, the shell swept
    only where it is real ghost storage -- with the oracle as the sweep and gloo as the transport"""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import itertools
    import torch
    import torch.distributed as dist
    import bricklib_b200 as bk
    import oracle
    from oracle import schedule as S

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dec = bk.BrickDecomp(dom, 8)
        nb, B = dec.nbricks, tuple(t - 2 for t in dec.tdims)
        lo, hi = bk.section_range(rank, subdim ** 3, world)
        nsub = hi - lo
        sg = bk.StitchedGrid(dec, lo, nsub, subdim)
        assert sg.is_box
        plan = bk.strong_pull_plan(dec, rank, world, subdim, sg)
        # adjacency of the stitched grid, from the positions that are swept (an aliased shell entry repeats an interior id)
        g = sg.grid
        adj = np.zeros((nsub * nb, 27), dtype=np.uint32)
        slo, shi = sg.sweep_box()
        for K, J, I in itertools.product(range(slo[2], shi[2]), range(slo[1], shi[1]), range(slo[0], shi[0])):
            for dk, dj, di in itertools.product((-1, 0, 1), repeat=3):
                k, j, i = K + dk, J + dj, I + di
                inside = 0 <= k < g.shape[0] and 0 <= j < g.shape[1] and 0 <= i < g.shape[2]
                adj[g[K, J, I], (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] = g[k, j, i] if inside else 0
        # field: the global periodic array, my subdomains' interiors loaded brick by brick
        P = oracle.port()
        G = tuple(subdim * n for n in dom)
        glob = np.random.default_rng(7).random(G[::-1])
        store = [oracle.aligned_zeros(nsub * nb * 512), oracle.aligned_zeros(nsub * nb * 512)]
        ext = tuple(n + 16 for n in dom)
        for qi in range(nsub):
            c = bk.zmort_decode(lo + qi)
            arr = np.zeros(tuple(n + 32 for n in dom[::-1]))
            arr[16:-16, 16:-16, 16:-16] = glob[c[2] * dom[2]:(c[2] + 1) * dom[2], c[1] * dom[1]:(c[1] + 1) * dom[1],
                                               c[0] * dom[0]:(c[0] + 1) * dom[0]]
            P.copy_brick(0, ext, (8,) * 3, (0,) * 3, arr, dec.grid, store[0][qi * nb * 512:(qi + 1) * nb * 512], 512)
        it = oracle.ST_ITER[stencil]
        sizes = [None] * world
        dist.all_gather_object(sizes, nsub * nb * 512)
        for _ in range(periods):
            pub = [torch.empty(n, dtype=torch.float64) for n in sizes]
            dist.all_gather(pub, torch.from_numpy(store[0].copy()))
            for owner, sub, spos, qi, gpos, n in plan:
                store[0][(qi * nb + gpos) * 512:(qi * nb + gpos + n) * 512] = \
                    pub[owner].numpy()[(sub * nb + spos) * 512:(sub * nb + spos + n) * 512]
            for s_ in range(it):
                blo, bhi = sg.sweep_box(last=s_ == it - 1)
                P.sweep_brick(stencil, g, blo, bhi, adj, store[s_ % 2], 512, 0, store[1 - s_ % 2], 512, 0)
        worst = 0.0
        want = S.periodic_steps(stencil, glob, it * periods)
        for qi in range(nsub):
            c = bk.zmort_decode(lo + qi)
            out = np.zeros(tuple(n + 32 for n in dom[::-1]))
            P.copy_brick(1, dom, (8,) * 3, (8,) * 3, out, dec.grid, store[0][qi * nb * 512:(qi + 1) * nb * 512], 512)
            got = out[16:-16, 16:-16, 16:-16]
            ref = want[c[2] * dom[2]:(c[2] + 1) * dom[2], c[1] * dom[1]:(c[1] + 1) * dom[1], c[0] * dom[0]:(c[0] + 1) * dom[0]]
            worst = max(worst, float((np.abs(got - ref) / (np.abs(got) + np.abs(ref))).max()))
        errs = [None] * world
        dist.all_gather_object(errs, worst)
        if rank == 0:
            q.put(max(errs))
    except Exception as exc:
        if rank == 0:
            q.put(repr(exc))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,stencil", [(2, 1), (4, 2), (2, 4)])
def test_stitched_strong_schedule_across_processes(world, stencil):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_strong_rank_main, args=(r, world, port, (16, 16, 16), 2, stencil, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err = q.get(timeout=5)
    assert isinstance(err, float) and err < 1e-12, err


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under torchrun: ranks != 0 exit 0 without work or output"""
    import subprocess
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
