"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers of the CUDA path:

* ``port()``  -> liboracle.so, our plain-C restatement (oracle/oracle.c), always buildable.
* ``ref()``   -> oracle/_ref/libbrickref_<isa>.so, the UNMODIFIED reference compiled from /root/reference by
                 oracle/Makefile (prebuilt files travel to the GPU box; absent => ``ref()`` returns None).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (bricklib_b200) never does.
"""
import ctypes as C
import os
import subprocess

# The reference's CPU path is timed with bound OpenMP threads (OMP_PROC_BIND=close); unbound
# threads scale negatively on virtualised hosts.  Must be in the environment before libgomp initialises.
os.environ.setdefault("OMP_PROC_BIND", "close")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STENCILS = {"7pt": 0, "mpi7pt": 1, "mpi13pt": 2, "mpi25pt": 3, "mpi125pt": 4}
RADIUS = {0: 1, 1: 1, 2: 2, 3: 4, 4: 2}
ST_ITER = {0: 8, 1: 8, 2: 4, 3: 2, 4: 4}
_ARRAY_FN = {1: "ref_sweep_array_7pt", 2: "ref_sweep_array_13pt", 3: "ref_sweep_array_25pt", 4: "ref_sweep_array_125pt"}

c_long_p = C.POINTER(C.c_long)
c_uint_p = C.POINTER(C.c_uint)
c_dbl_p = C.POINTER(C.c_double)


def _longs(v):
    return (C.c_long * 3)(*[int(x) for x in v])


def _ptr(a, ty):
    return a.ctypes.data_as(ty)


def aligned_zeros(n, align=4096):
    """float64 zeros whose base address is `align`-aligned (the reference's generated AVX code uses aligned loads;
    BrickStorage::allocate aligns to 2048 B, brick.h:15,:68-75)."""
    raw = np.zeros(n * 8 + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * 8].view(np.float64)


def build(verbose=False):
    """make liboracle.so (+ _ref when /root/reference is present).  Building the checker is not using it."""
    r = subprocess.run(["make", "-j4", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-2000:], r.stderr[-2000:])
    if r.returncode:
        raise RuntimeError("oracle build failed")


class RegionStruct(C.Structure):
    _fields_ = [("neighbor", C.c_uint64), ("skin_st", C.c_uint), ("skin_ed", C.c_uint), ("pos", C.c_uint),
                ("len", C.c_uint)]


class DecompStruct(C.Structure):
    _fields_ = [("dims", C.c_uint * 3), ("gdepth", C.c_uint * 3), ("tdims", C.c_uint * 3), ("nbricks", C.c_uint),
                ("sep_pos", C.c_uint * 3), ("grid", c_uint_p), ("adj", c_uint_p), ("nregions", C.c_int),
                ("ghost", RegionStruct * 64), ("skin", RegionStruct * 64), ("skin_size", C.c_long * 26)]


class Port:
    """liboracle.so (oracle/oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.orc_decomp_new.restype = C.POINTER(DecompStruct)
        L.orc_bitset_of.restype = C.c_uint64
        L.orc_bitset_neg.restype = C.c_uint64
        L.orc_bitset_neg.argtypes = [C.c_uint64]

    def sweep_array(self, stencil, inp, lo, hi, coeff=None, out=None):
        """inp: padded array indexed [k][j][i]; sweeps cells lo<=(i,j,k)<hi."""
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        out = np.zeros_like(inp) if out is None else out
        ext = inp.shape[::-1]
        cf = None if coeff is None else _ptr(np.ascontiguousarray(coeff, dtype=np.float64), c_dbl_p)
        rc = self.L.orc_sweep_array(stencil, _longs(ext), _longs(lo), _longs(hi), _ptr(inp, c_dbl_p),
                                    _ptr(out, c_dbl_p), cf)
        assert rc == 0, rc
        return out

    def init_grid(self, dims):
        n = int(np.prod(dims))
        grid = np.zeros(n, dtype=np.uint32)
        adj = np.zeros((n, 27), dtype=np.uint32)
        self.L.orc_init_grid(_longs(dims), _ptr(grid, c_uint_p), _ptr(adj, c_uint_p))
        return grid.reshape(dims[::-1]), adj

    def copy_brick(self, direction, dimlist, padding, ghost, arr, grid, dat, step, off=0):
        assert arr.dtype == np.float64 and dat.dtype == np.float64 and grid.dtype == np.uint32
        self.L.orc_copy_brick(direction, _longs(dimlist), _longs(padding), _longs(ghost), _ptr(arr, c_dbl_p),
                              _ptr(grid, c_uint_p), _ptr(dat, c_dbl_p), C.c_size_t(step), C.c_size_t(off))

    def sweep_brick(self, stencil, grid, lo, hi, adj, din, step_in, off_in, dout, step_out, off_out, coeff=None):
        sb = grid.shape[::-1]
        cf = None if coeff is None else _ptr(np.ascontiguousarray(coeff, dtype=np.float64), c_dbl_p)
        rc = self.L.orc_sweep_brick(stencil, _ptr(grid, c_uint_p), _longs(sb), _longs(lo), _longs(hi),
                                    _ptr(adj, c_uint_p), _ptr(din, c_dbl_p), C.c_size_t(step_in), C.c_size_t(off_in),
                                    _ptr(dout, c_dbl_p), C.c_size_t(step_out), C.c_size_t(off_out), cf)
        assert rc == 0, rc

    def decomp(self, dom, depth=8):
        p = self.L.orc_decomp_new((C.c_uint * 3)(*dom), C.c_uint(depth))
        if not p:
            raise ValueError("bad decomposition")
        d = p.contents
        t = tuple(d.tdims)
        out = {
            "dims": tuple(d.dims), "tdims": t, "nbricks": d.nbricks, "sep_pos": tuple(d.sep_pos),
            "grid": np.ctypeslib.as_array(d.grid, shape=(t[2], t[1], t[0])).copy(),
            "adj": np.ctypeslib.as_array(d.adj, shape=(d.nbricks, 27)).copy(),
            "ghost": [(r.neighbor, r.skin_st, r.skin_ed, r.pos, r.len) for r in d.ghost[:d.nregions]],
            "skin": [(r.neighbor, r.skin_st, r.skin_ed, r.pos, r.len) for r in d.skin[:d.nregions]],
            "skin_size": list(d.skin_size),
        }
        self.L.orc_decomp_free(p)
        return out

    def rank_map(self, cart, coo):
        sets = (C.c_uint64 * 27)()
        ranks = (C.c_int * 27)()
        self.L.orc_rank_map((C.c_int * 3)(*cart), (C.c_int * 3)(*coo), sets, ranks)
        return {int(s): int(r) for s, r in zip(sets, ranks)}

    def bitset(self, *elems):
        e = list(elems) + [0, 0, 0]
        return int(self.L.orc_bitset_of(e[0], e[1], e[2]))

    def bitset_neg(self, s):
        return int(self.L.orc_bitset_neg(C.c_uint64(s)))


def _cpu_flags():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return set(line.split(":")[1].split())
    except OSError:
        pass
    return set()


class Ref:
    """oracle/_ref/libbrickref_<isa>.so: the unmodified reference (see oracle/ref_harness.cpp)."""

    def __init__(self, path, isa):
        self.isa = isa
        L = self.L = C.CDLL(path)
        L.ref_decomp_new.restype = C.c_void_p
        L.ref_isa.restype = C.c_char_p
        for fn in ("ref_decomp_free", "ref_decomp_nbricks", "ref_decomp_sep_pos", "ref_decomp_grid", "ref_decomp_adj",
                   "ref_decomp_nregions", "ref_decomp_regions", "ref_decomp_skin_size"):
            getattr(L, fn).argtypes = None
        L.ref_decomp_nbricks.restype = C.c_uint
        self.threads = L.ref_threads()

    def init_grid(self, dims):
        n = int(np.prod(dims))
        grid = np.zeros(n, dtype=np.uint32)
        adj = np.zeros((n, 27), dtype=np.uint32)
        self.L.ref_init_grid(_longs(dims), _ptr(grid, c_uint_p), _ptr(adj, c_uint_p))
        return grid.reshape(dims[::-1]), adj

    def decomp(self, dom, depth=8, cart=(1, 1, 1), coo=(0, 0, 0), keep=False):
        h = C.c_void_p(self.L.ref_decomp_new((C.c_uint * 3)(*dom), C.c_uint(depth), (C.c_int * 3)(*cart),
                                             (C.c_int * 3)(*coo)))
        nb = self.L.ref_decomp_nbricks(h)
        t = tuple(d // 8 + 2 * (depth // 8) for d in dom)
        sep = (C.c_uint * 3)()
        self.L.ref_decomp_sep_pos(h, sep)
        grid = np.zeros((t[2], t[1], t[0]), dtype=np.uint32)
        self.L.ref_decomp_grid(h, _ptr(grid, c_uint_p))
        adj = np.zeros((nb, 27), dtype=np.uint32)
        self.L.ref_decomp_adj(h, _ptr(adj, c_uint_p))
        nr = self.L.ref_decomp_nregions(h)
        out = {"tdims": t, "nbricks": nb, "sep_pos": tuple(sep), "grid": grid, "adj": adj}
        for which, name in ((0, "ghost"), (1, "skin")):
            tab = np.zeros((nr, 6), dtype=np.uint32)
            sets = np.zeros(nr, dtype=np.uint64)
            peer = np.zeros(nr, dtype=np.int32)
            self.L.ref_decomp_regions(h, which, _ptr(tab, c_uint_p), sets.ctypes.data_as(C.POINTER(C.c_uint64)),
                                      peer.ctypes.data_as(C.POINTER(C.c_int)))
            out[name] = [(int(sets[i]), int(tab[i, 2]), int(tab[i, 3]), int(tab[i, 0]), int(tab[i, 1]))
                         for i in range(nr)]
            out[name + "_peer"] = [int(p) for p in peer]
            out[name + "_pad"] = [(int(tab[i, 4]), int(tab[i, 5])) for i in range(nr)]
        ss = (C.c_long * 26)()
        self.L.ref_decomp_skin_size(h, ss)
        out["skin_size"] = list(ss)
        if keep:
            out["handle"] = h
        else:
            self.L.ref_decomp_free(h)
        return out

    def decomp_free(self, h):
        self.L.ref_decomp_free(h)

    def exchange_post(self, handle, rank, dat, chunks, step):
        self.L.ref_decomp_exchange(handle, C.c_int(rank), _ptr(dat, c_dbl_p), C.c_long(chunks), C.c_size_t(step))

    def deliver(self):
        self.L.ref_mpi_deliver()

    def _copy(self, fn, dimlist, padding, ghost, arr, grid, adj, dat, step, off):
        assert arr.dtype == np.float64 and dat.dtype == np.float64 and grid.dtype == np.uint32
        return fn(_longs(dimlist), _longs(padding), _longs(ghost), _ptr(arr, c_dbl_p), _ptr(grid, c_uint_p),
                  _ptr(adj, c_uint_p), C.c_uint(adj.shape[0]), _ptr(dat, c_dbl_p), C.c_size_t(step), C.c_uint(off))

    def copy_to_brick(self, dimlist, padding, ghost, arr, grid, adj, dat, step, off=0):
        self._copy(self.L.ref_copy_to_brick, dimlist, padding, ghost, arr, grid, adj, dat, step, off)

    def copy_from_brick(self, dimlist, padding, ghost, arr, grid, adj, dat, step, off=0):
        self._copy(self.L.ref_copy_from_brick, dimlist, padding, ghost, arr, grid, adj, dat, step, off)

    def compare_brick(self, dimlist, padding, ghost, arr, grid, adj, dat, step, off=0):
        return bool(self._copy(self.L.ref_compare_brick, dimlist, padding, ghost, arr, grid, adj, dat, step, off))

    def sweep_brick(self, stencil, grid, lo, hi, adj, din, step_in, off_in, dout, step_out, off_out, coeff=None):
        sb = grid.shape[::-1]
        cf = None if coeff is None else _ptr(np.ascontiguousarray(coeff, dtype=np.float64), c_dbl_p)
        rc = self.L.ref_sweep_brick(stencil, _ptr(grid, c_uint_p), _longs(sb), _longs(lo), _longs(hi),
                                    _ptr(adj, c_uint_p), C.c_uint(adj.shape[0]), _ptr(din, c_dbl_p),
                                    C.c_size_t(step_in), C.c_uint(off_in), _ptr(dout, c_dbl_p), C.c_size_t(step_out),
                                    C.c_uint(off_out), cf)
        assert rc == 0, rc

    def sweep_array(self, stencil, inp, lo, hi, out=None):
        """reference scalar form ST_CPU (fake.h); stencil in 1..4 (the coeff[] 7pt has no ST_CPU macro)."""
        inp = np.ascontiguousarray(inp, dtype=np.float64)
        out = np.zeros_like(inp) if out is None else out
        fn = getattr(self.L, _ARRAY_FN[stencil])
        fn(_longs(inp.shape[::-1]), _longs(lo), _longs(hi), _ptr(inp, c_dbl_p), _ptr(out, c_dbl_p))
        return out


_PORT = None
_REF = False


def port():
    global _PORT
    if _PORT is None:
        _PORT = Port()
    return _PORT


def ref():
    """the compiled reference, or None when oracle/_ref holds no loadable build for this CPU."""
    global _REF
    if _REF is False:
        _REF = None
        flags = _cpu_flags()
        order = []
        if {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd"} <= flags:
            order.append("avx512")
        if {"avx2", "fma", "bmi2"} <= flags:
            order.append("avx2")
        for isa in order:
            path = os.path.join(HERE, "_ref", f"libbrickref_{isa}.so")
            if os.path.exists(path):
                try:
                    _REF = Ref(path, isa)
                    break
                except OSError:
                    continue
    return _REF
