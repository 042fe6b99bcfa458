"""Product host logic (bk_decomp.cpp through the C ABI) against the reference fixtures and the oracle port."""
import hashlib
import json
import os

import numpy as np
import pytest

import bricklib_b200 as bk
from bricklib_b200 import _lib
import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def region_rows(regs):
    return [[r.neighbor, r.skin_st, r.skin_ed, r.pos, r.len] for r in regs]


def test_decomp_matches_reference_fixture(golden_dir):
    tables = json.load(open(os.path.join(golden_dir, "decomp_tables.json")))
    for key, t in tables.items():
        dom = tuple(int(x) for x in key.split("x"))
        d = bk.BrickDecomp(dom, 8)
        assert d.nbricks == t["nbricks"] and list(d.sep_pos) == t["sep_pos"] and list(d.tdims) == t["tdims"], key
        assert region_rows(d.ghost) == t["ghost"] and region_rows(d.skin) == t["skin"], key
        assert d.skin_size == t["skin_size"], key
        assert sha(d.grid) == t["grid_sha256"] and sha(d.adj[1:]) == t["adj1_sha256"], key
        assert all(r.first_pad == 0 and r.last_pad == 0 for r in d.ghost + d.skin)


def test_decomp_matches_oracle_port_on_odd_shapes():
    P = oracle.port()
    for dom in [(16, 16, 16), (40, 16, 72), (24, 64, 16), (96, 32, 48)]:
        d, o = bk.BrickDecomp(dom, 8), P.decomp(dom)
        assert np.array_equal(d.grid, o["grid"]) and np.array_equal(d.adj, o["adj"])
        assert region_rows(d.ghost) == [list(r) for r in o["ghost"]]
        assert region_rows(d.skin) == [list(r) for r in o["skin"]]


def test_adjacency_symmetry():
    """the self-check every reference driver runs (weak/main.cu:102-109)"""
    d = bk.BrickDecomp((32, 24, 40), 8)
    g, adj = d.grid, d.adj
    for k in range(1, g.shape[0] - 1):
        for j in range(1, g.shape[1] - 1):
            for i in range(1, g.shape[2] - 1):
                b = g[k, j, i]
                for s in range(27):
                    assert adj[adj[b, s], 26 - s] == b


def test_ghost_and_skin_cover_their_ranges():
    d = bk.BrickDecomp((64, 64, 64), 8)
    ghost_ids = np.concatenate([np.arange(r.pos, r.pos + r.len) for r in d.ghost])
    assert len(ghost_ids) == len(set(ghost_ids)) == d.sep_pos[2] - d.sep_pos[1]   # ghost ranges never overlap
    assert ghost_ids.min() == d.sep_pos[1] and ghost_ids.max() == d.sep_pos[2] - 1
    for g, s in zip(d.ghost, d.skin):
        assert g.len == s.len and d.sep_pos[0] <= s.pos and s.pos + s.len <= d.sep_pos[1]
    assert len(d.id_list(0)) == 6 ** 3 and len(d.id_list(1)) == 8 ** 3 - 6 ** 3 and len(d.id_list(2)) == 10 ** 3 - 8 ** 3


def test_bad_arguments_are_rejected():
    with pytest.raises(bk.BrickError):
        bk.BrickDecomp((60, 64, 64), 8)      # not a multiple of the brick edge
    with pytest.raises(bk.BrickError):
        bk.BrickDecomp((64, 64, 8), 8)       # thinner than two ghost layers
    with pytest.raises(bk.BrickError):
        bk.BrickDecomp((64, 64, 64), 4)      # ghost depth must be whole bricks (brick-mpi.h:312)


def test_rank_map_matches_reference_fixture(golden_dir):
    maps = json.load(open(os.path.join(golden_dir, "rank_maps.json")))
    d = bk.BrickDecomp((16, 16, 16), 8)
    for key, m in maps.items():
        cart, coo = [tuple(int(x) for x in part.split(",")) for part in key.split("@")]
        mine = d.populate(cart, coo)
        for s, r in m.items():
            assert mine[int(s)] == r, (key, s)


def test_init_grid_matches_reference_fixture(golden_dir):
    ig = json.load(open(os.path.join(golden_dir, "init_grid.json")))
    for key, t in ig.items():
        dims = tuple(int(x) for x in key.split("x"))
        g, a = bk.init_grid(dims)
        assert sha(g) == t["grid_sha256"] and sha(a) == t["adj_sha256"], key


def test_zmorton_roundtrip():
    import ctypes as C
    L = _lib.load()
    assert L.bk_zmort_encode((C.c_ulong * 3)(1, 0, 0)) == 1
    assert L.bk_zmort_encode((C.c_ulong * 3)(0, 1, 0)) == 2
    assert L.bk_zmort_encode((C.c_ulong * 3)(3, 5, 6)) == 0b110_101_011
    for xyz in [(0, 0, 0), (7, 3, 5), (15, 15, 15), (100, 3, 77)]:
        z = L.bk_zmort_encode((C.c_ulong * 3)(*xyz))
        out = (C.c_ulong * 3)()
        L.bk_zmort_decode(z, out)
        assert tuple(out) == xyz


def test_shell_boxes_partition_the_boundary():
    t = (10, 9, 8)
    lo, hi, ilo, ihi = (0, 0, 0), t, (2, 2, 2), tuple(x - 2 for x in t)
    seen = np.zeros(t[::-1], dtype=int)
    for a, b in bk.shell_boxes(lo, hi, ilo, ihi):
        seen[a[2]:b[2], a[1]:b[1], a[0]:b[0]] += 1
    seen[ilo[2]:ihi[2], ilo[1]:ihi[1], ilo[0]:ihi[0]] += 1
    assert (seen == 1).all()
