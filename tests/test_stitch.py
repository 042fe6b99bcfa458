"""Host logic of the stitched strong-scaling grid (bk_stitch_*): pure integer work, checked on CPU.

Reference anchors: Z-Morton ids include/zmort.h:18-105, section split strong/args.cpp:104-113, ghost aliasing
strong/main.cpp:205-262 (CPU, mmap) / ghost links strong/main.cu:188-247 (CUDA, cudaCopy)."""
import itertools

import numpy as np
import pytest

import bricklib_b200 as bk


def neighbour_sign(region):
    """(di, dj, dk) in {-1,0,1} of a BitSet (include/bitset.h: bit a = +axis a, bit 31+a = -axis a, axis 1 = i)"""
    s = int(region.neighbor)
    return tuple(1 if (s >> (a + 1)) & 1 else -1 if (s >> (31 + a + 1)) & 1 else 0 for a in range(3))


@pytest.mark.parametrize("subdim,size", [(2, 1), (4, 1), (4, 2), (4, 4), (4, 8), (8, 8), (4, 3), (2, 8), (8, 64)])
def test_sections_of_power_of_two_rank_counts_are_boxes(subdim, size):
    allsubs = subdim ** 3
    d = bk.BrickDecomp((32, 32, 32), 8)
    covered = 0
    for rank in range(size):
        lo, hi = bk.section_range(rank, allsubs, size)
        covered += hi - lo
        sg = bk.StitchedGrid(d, lo, hi - lo, subdim)
        power_of_two = size & (size - 1) == 0
        assert sg.is_box == power_of_two or (hi - lo == 1), (rank, sg.n)
        if sg.is_box:
            assert int(np.prod(sg.n)) == hi - lo
            assert sg.wrap == tuple(n == subdim for n in sg.n)
    assert covered == allsubs


@pytest.mark.parametrize("dom,subdim,size", [((32, 32, 32), 2, 1), ((32, 32, 32), 4, 2), ((16, 24, 32), 4, 8),
                                             ((64, 64, 64), 2, 8), ((32, 16, 16), 4, 4)])
def test_stitched_grid_names_every_brick_once_and_aliases_periodic_shells(dom, subdim, size):
    d = bk.BrickDecomp(dom, 8)
    nb = d.nbricks
    B = tuple(t - 2 for t in d.tdims)
    for rank in range(size):
        lo, hi = bk.section_range(rank, subdim ** 3, size)
        sg = bk.StitchedGrid(d, lo, hi - lo, subdim)
        assert sg.is_box
        g = sg.grid
        assert g.shape == tuple(n * b + 2 for n, b in zip(sg.n, B))[::-1]
        interior = g[1:-1, 1:-1, 1:-1]
        # every interior position names a distinct inner/skin brick of one of my subdomains, and all of them are named
        assert len(np.unique(interior)) == interior.size
        q, local = interior // nb, interior % nb
        assert q.min() == 0 and q.max() == hi - lo - 1
        assert local.min() >= 1 and local.max() < d.sep_pos[1]
        assert interior.size == (hi - lo) * (d.sep_pos[1] - 1)
        # position -> subdomain: Z-Morton of (box origin + position // B)
        for K, J, I in itertools.product(*[range(0, s, 3) for s in interior.shape]):
            c = (sg.lo[0] + I // B[0], sg.lo[1] + J // B[1], sg.lo[2] + K // B[2])
            assert q[K, J, I] == bk.zmort_encode(c) - lo
            assert local[K, J, I] == d.grid[1 + K % B[2], 1 + J % B[1], 1 + I % B[0]]
        # shells: an aliased axis repeats the far side's interior layer, a real shell holds ghost bricks
        for axis in range(3):
            first = np.take(g, 0, axis=2 - axis)
            last = np.take(g, g.shape[2 - axis] - 1, axis=2 - axis)
            if sg.wrap[axis]:
                assert np.array_equal(first, np.take(g, g.shape[2 - axis] - 2, axis=2 - axis))
                assert np.array_equal(last, np.take(g, 1, axis=2 - axis))
            else:
                assert (first % nb >= d.sep_pos[1]).all() and (last % nb >= d.sep_pos[1]).all()


@pytest.mark.parametrize("dom,subdim,size", [((32, 32, 32), 4, 2), ((16, 24, 32), 4, 8), ((64, 64, 64), 2, 8),
                                             ((32, 32, 32), 2, 1), ((32, 16, 16), 4, 4)])
def test_needed_regions_are_exactly_the_ghost_bricks_the_grid_reads(dom, subdim, size):
    """the two halves of the contract agree: the union of the ghost ranges bk_stitch_region_needed keeps equals the set
    of ghost bricks that appear in the stitched grid (so nothing read is left unexchanged, nothing exchanged is unread)"""
    d = bk.BrickDecomp(dom, 8)
    nb = d.nbricks
    for rank in range(size):
        lo, hi = bk.section_range(rank, subdim ** 3, size)
        sg = bk.StitchedGrid(d, lo, hi - lo, subdim)
        in_grid = np.unique(sg.grid)
        in_grid = set(int(x) for x in in_grid[(in_grid % nb) >= d.sep_pos[1]])
        kept = set()
        for q in range(hi - lo):
            for r, reg in enumerate(d.ghost):
                if sg.region_needed(lo + q, r):
                    kept.update(range(q * nb + reg.pos, q * nb + reg.pos + reg.len))
        assert kept == in_grid
        if all(sg.wrap):
            assert not kept


def test_sweep_boxes_and_errors():
    d = bk.BrickDecomp((32, 32, 32), 8)
    sg = bk.StitchedGrid(d, 0, 32, 4)          # 2 ranks of a 4^3 arrangement: box 4 x 4 x 2
    assert sg.n == (4, 4, 2) and sg.wrap == (True, True, False)
    assert sg.sweep_box() == ((1, 1, 0), (sg.dims[0] - 1, sg.dims[1] - 1, sg.dims[2]))
    assert sg.sweep_box(last=True) == ((1, 1, 1), tuple(x - 1 for x in sg.dims))
    with pytest.raises(ValueError):
        sg.region_needed(40, 0)
    odd = bk.StitchedGrid(d, 0, 22, 4)          # 3 ranks: not a box
    assert not odd.is_box and odd.grid is None


def _simulate_exchange(dom, subdim, size, stitched):
    """every rank's storage holds, per brick, the code of the GLOBAL brick position whose data it carries (-1 = none);
    apply every rank's pull plan as plain range copies, as k_xplan does"""
    d = bk.BrickDecomp(dom, 8)
    nb, B = d.nbricks, tuple(t - 2 for t in d.tdims)
    G = tuple(subdim * b for b in B)
    stores, grids = [], []
    for rank in range(size):
        lo, hi = bk.section_range(rank, subdim ** 3, size)
        st = np.full((hi - lo) * nb, -1, dtype=np.int64)
        for q in range(hi - lo):
            c = bk.zmort_decode(lo + q)
            for bk_, bj, bi in itertools.product(range(B[2]), range(B[1]), range(B[0])):
                code = (c[0] * B[0] + bi) + G[0] * ((c[1] * B[1] + bj) + G[1] * (c[2] * B[2] + bk_))
                st[q * nb + d.grid[1 + bk_, 1 + bj, 1 + bi]] = code
        stores.append(st)
        grids.append(bk.StitchedGrid(d, lo, hi - lo, subdim) if stitched else None)
    before = [s.copy() for s in stores]
    moved = 0
    for rank in range(size):
        for owner, sub, spos, q, gpos, n in bk.strong_pull_plan(d, rank, size, subdim, grids[rank]):
            stores[rank][q * nb + gpos:q * nb + gpos + n] = before[owner][sub * nb + spos:sub * nb + spos + n]
            moved += n
    return d, B, G, stores, grids, moved


@pytest.mark.parametrize("dom,subdim,size", [((16, 16, 16), 2, 1), ((16, 16, 16), 4, 2), ((16, 24, 32), 2, 8),
                                             ((16, 16, 16), 4, 8), ((24, 16, 16), 4, 4)])
def test_stitched_exchange_delivers_what_the_grid_names(dom, subdim, size):
    """after one exchange of the kept regions, EVERY entry of every rank's stitched grid -- interior, real shell, aliased
    shell -- carries the data of the periodic global position it stands for"""
    d, B, G, stores, grids, moved = _simulate_exchange(dom, subdim, size, stitched=True)
    for rank, sg in enumerate(grids):
        g = sg.grid
        for K, J, I in itertools.product(*[range(s) for s in g.shape]):
            pos = [(sg.lo[a] * B[a] + p - 1) % G[a] for a, p in enumerate((I, J, K))]
            want = pos[0] + G[0] * (pos[1] + G[1] * pos[2])
            assert stores[rank][g[K, J, I]] == want, (rank, (I, J, K))
    _, _, _, _, _, moved_all = _simulate_exchange(dom, subdim, size, stitched=False)
    assert moved < moved_all or size == subdim ** 3     # stitching exchanges the box surface only


@pytest.mark.parametrize("dom,subdim,size", [((16, 16, 16), 2, 1), ((16, 16, 16), 2, 3), ((16, 24, 32), 2, 8)])
def test_per_subdomain_exchange_fills_every_ghost_brick(dom, subdim, size):
    """the unstitched plan (drivers/strong -M): every ghost brick of every subdomain receives the periodic neighbour's data"""
    d, B, G, stores, _, _ = _simulate_exchange(dom, subdim, size, stitched=False)
    nb = d.nbricks
    for rank in range(size):
        lo, hi = bk.section_range(rank, subdim ** 3, size)
        for q in range(hi - lo):
            c = bk.zmort_decode(lo + q)
            for lk, lj, li in itertools.product(*[range(t) for t in d.tdims[::-1]]):
                pos = [(c[a] * B[a] + p - 1) % G[a] for a, p in enumerate((li, lj, lk))]
                want = pos[0] + G[0] * (pos[1] + G[1] * pos[2])
                assert stores[rank][q * nb + d.grid[lk, lj, li]] == want
    for sub_id in range(subdim ** 3):
        r, idx = bk.section_owner(sub_id, subdim ** 3, size)
        lo, hi = bk.section_range(r, subdim ** 3, size)
        assert lo + idx == sub_id and idx < hi - lo
