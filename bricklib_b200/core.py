"""Host-side mirror of the reference's brick interface, over the C ABI (include/bricklib_b200.h).

Same names and argument meaning as the reference so the parity tests read like its drivers:
  BrickStorage / BrickInfo / Brick      include/brick.h:53-395        (device resident here)
  init_grid                             include/bricksetup.h:73-90
  copyToBrick / copyFromBrick           include/bricksetup.h:172-221
  compareBrick                          include/brickcompare.h:30-57  (explicit tolerance)
  BrickDecomp, ExchangeView             include/brick-mpi.h:82-124, :178-513
  brick_kernel launch -> stencil()      weak/main.cu:35-43, :277-282
Python is plumbing for tests/bench; the C++ drivers in drivers/ use the same C ABI through include/*.h.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Field, Region, Seg, check, load

BRICK = 512  # elements of an 8x8x8 brick


def _l3(v):
    return (C.c_long * 3)(*[int(x) for x in v])


def _u3(v):
    return (C.c_uint * 3)(*[int(x) for x in v])


class DeviceBuffer:
    """cudaMalloc'ed bytes (IPC-exportable, unlike a slice of a caching allocator)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(load().bk_dev_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        self._owned = True

    @classmethod
    def from_numpy(cls, a, stream=None):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a, stream)
        return b

    def upload(self, a, stream=None):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(load().bk_memcpy_h2d(self.ptr, a.ctypes.data, a.nbytes, stream))
        check(load().bk_stream_sync(stream))

    def download(self, dtype, count=None, offset_bytes=0, stream=None):
        dtype = np.dtype(dtype)
        n = (self.nbytes - offset_bytes) // dtype.itemsize if count is None else int(count)
        out = np.empty(n, dtype=dtype)
        check(load().bk_memcpy_d2h(out.ctypes.data, self.ptr + offset_bytes, out.nbytes, stream))
        check(load().bk_stream_sync(stream))
        return out

    def zero(self, stream=None):
        check(load().bk_dev_memset(self.ptr, 0, self.nbytes, stream))

    def free(self):
        if self._owned and self.ptr:
            load().bk_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BrickInfo:
    """BrickInfo<3>: the adjacency list, mirrored to the device (replaces movBrickInfo, brick-gpu.h:43-57)."""

    def __init__(self, adj_host):
        adj_host = np.ascontiguousarray(adj_host, dtype=np.uint32).reshape(-1, 27)
        self.nbricks = adj_host.shape[0]
        self.adj_host = adj_host
        self.adj = DeviceBuffer.from_numpy(adj_host)

    def allocate(self, step):
        return BrickStorage.allocate(self.nbricks, step)


class BrickStorage:
    """BrickStorage on the device: `chunks` bricks, `step` elements apart (brick.h:53-82)."""

    def __init__(self, chunks, step, dat):
        self.chunks, self.step, self.dat = int(chunks), int(step), dat

    @classmethod
    def allocate(cls, chunks, step):
        buf = DeviceBuffer(int(chunks) * int(step) * 8)
        buf.zero()  # the null brick (id 0) and padding read as 0.0 instead of garbage
        return cls(chunks, step, buf)

    def brick_ptr(self, b, offset=0):
        return self.dat.ptr + (int(b) * self.step + offset) * 8

    def to_host(self):
        return self.dat.download(np.float64)

    def from_host(self, a):
        self.dat.upload(np.ascontiguousarray(a, dtype=np.float64))


class Brick:
    """Brick<Dim<8,8,8>,Dim<4,8>>(bInfo, bStorage, offset): a view (brick.h:389-393)."""

    def __init__(self, info, storage, offset=0):
        self.info, self.storage, self.offset = info, storage, int(offset)
        self.step = storage.step
        self.dat = storage.dat.ptr + self.offset * 8


def init_grid(dimlist):
    """init_grid<3>(grid_ptr, dimlist): returns (grid[k][j][i] host array, BrickInfo)."""
    n = int(np.prod(dimlist))
    grid = np.zeros(n, dtype=np.uint32)
    adj = np.zeros((n, 27), dtype=np.uint32)
    check(load().bk_init_grid(_l3(dimlist), grid.ctypes.data_as(_lib.up), adj.ctypes.data_as(_lib.up)))
    return grid.reshape(tuple(dimlist)[::-1]), adj


class DeviceGrid:
    """dense brick-id array on the device + its extents (i first)."""

    def __init__(self, grid_host):
        self.host = np.ascontiguousarray(grid_host, dtype=np.uint32)
        self.dims = tuple(self.host.shape[::-1])
        self.dev = DeviceBuffer.from_numpy(self.host)


def copyToBrick(dimlist, padding, ghost, arr_dev, grid, brick, stream=None):
    check(load().bk_copy_to_brick(_l3(dimlist), _l3(padding), _l3(ghost), arr_dev.ptr, grid.dev.ptr, brick.dat,
                                  brick.step, stream))


def copyFromBrick(dimlist, padding, ghost, arr_dev, grid, brick, stream=None):
    check(load().bk_copy_from_brick(_l3(dimlist), _l3(padding), _l3(ghost), arr_dev.ptr, grid.dev.ptr, brick.dat,
                                    brick.step, stream))


def compareBrick(dimlist, padding, ghost, arr_dev, grid, brick, tol=1e-12, stream=None):
    """returns (ok, mismatching cells, max relative difference)"""
    bad = C.c_ulonglong()
    rel = C.c_double()
    check(load().bk_compare_brick(_l3(dimlist), _l3(padding), _l3(ghost), arr_dev.ptr, grid.dev.ptr, brick.dat,
                                  brick.step, tol, C.byref(bad), C.byref(rel), stream))
    return bad.value == 0, bad.value, rel.value


def fill_synthetic(grid, brick, origin_cells, global_cells, seed, stream=None):
    """bk_fill_synthetic: every non-null brick of `grid` gets the counter-based synthetic field (a hash of the global
    periodic cell coordinate); origin_cells = global coordinate of cell 0 of grid position (0,0,0)"""
    check(load().bk_fill_synthetic(grid.dev.ptr, _u3(grid.dims), _l3(origin_cells), _l3(global_cells), int(seed),
                                   brick.dat, brick.step, stream))


def synthetic_field(seed, global_cells, lo, hi):
    """the same field on the host (numpy): cells lo <= (i,j,k) < hi of the global periodic array, returned as [k][j][i].
    Bit-identical to bk_fill_synthetic / bk_synthetic_value (splitmix64 of the linear global cell index)."""
    ax = [np.mod(np.arange(lo[d], hi[d], dtype=np.int64), int(global_cells[d])).astype(np.uint64) for d in range(3)]
    g0, g1 = np.uint64(global_cells[0]), np.uint64(global_cells[1])
    lin = (ax[2][:, None, None] * g1 + ax[1][None, :, None]) * g0 + ax[0][None, None, :]
    with np.errstate(over="ignore"):
        z = np.uint64(int(seed) & (2 ** 64 - 1)) + (lin + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def compare_storage(grid, lo, hi, brick_a, brick_b, tol=1e-12, stream=None):
    """compareBrick between two brick storages over the brick box [lo,hi): (ok, mismatching cells, max relative diff)"""
    bad = C.c_ulonglong()
    rel = C.c_double()
    check(load().bk_compare_storage(grid.dev.ptr, _u3(grid.dims), _u3(lo), _u3(hi), brick_a.dat, brick_a.step,
                                    brick_b.dat, brick_b.step, tol, C.byref(bad), C.byref(rel), stream))
    return bad.value == 0, bad.value, rel.value


def _field(b_in, b_out):
    assert b_in.info is b_out.info or b_in.info.adj.ptr == b_out.info.adj.ptr
    return Field(b_in.info.adj.ptr, b_in.dat, b_in.step, b_out.dat, b_out.step)


def _coeff(coeff):
    if coeff is None:
        return None
    c = np.ascontiguousarray(coeff, dtype=np.float64)
    return c.ctypes.data_as(_lib.dp)


def stencil(stencil_id, grid, b_in, b_out, lo=None, hi=None, coeff=None, kernel=_lib.KERNEL_AUTO, stream=None):
    """out = stencil(in) over the brick box [lo,hi) of `grid` -- the brick_kernel<<<...>>> launch."""
    lo = (0, 0, 0) if lo is None else lo
    hi = grid.dims if hi is None else hi
    f = _field(b_in, b_out)
    check(load().bk_stencil_apply(stencil_id, C.byref(f), grid.dev.ptr, _u3(grid.dims), _u3(lo), _u3(hi),
                                  _coeff(coeff), kernel, stream))


def stencil_part(stencil_id, grid, b_in, b_out, lo, hi, ready_lo, ready_hi, part, coeff=None, stream=None):
    """one half of a split sweep (bk_stencil_apply_part): part = PART_READY (CTAs that read only bricks of the ready
    box) or PART_REST (all the others)"""
    f = _field(b_in, b_out)
    check(load().bk_stencil_apply_part(stencil_id, C.byref(f), grid.dev.ptr, _u3(grid.dims), _u3(lo), _u3(hi),
                                       _coeff(coeff), _u3(ready_lo), _u3(ready_hi), part, stream))


def stencil_advance(stencil_id, steps, grid, b_in, b_out, lo=None, hi=None, ready=None, part=_lib.PART_ALL, coeff=None,
                    stream=None, remote=None):
    """`steps` (1 or 2) time steps in one pass (bk_stencil_advance); ready = (ready_lo, ready_hi) for split launches.
    remote = (table, ghost_lo, ghost_n): the exchange inside the sweep (bk_stencil_advance_remote) -- ghost brick
    ghost_lo + g of the input is read from the address table[g] (DeviceBuffer of device-visible pointers) instead of
    the own storage.  Raises Unsupported when there is no such kernel for this stencil/layout."""
    lo = (0, 0, 0) if lo is None else lo
    hi = grid.dims if hi is None else hi
    f = _field(b_in, b_out)
    rl, rh = (_u3(ready[0]), _u3(ready[1])) if ready else (None, None)
    if remote is not None:
        table, g_lo, g_n = remote
        rc = load().bk_stencil_advance_remote(stencil_id, steps, C.byref(f), grid.dev.ptr, _u3(grid.dims), _u3(lo), _u3(hi),
                                              _coeff(coeff), rl, rh, part, table.ptr, int(g_lo), int(g_n), stream)
    else:
        rc = load().bk_stencil_advance(stencil_id, steps, C.byref(f), grid.dev.ptr, _u3(grid.dims), _u3(lo), _u3(hi),
                                       _coeff(coeff), rl, rh, part, stream)
    if rc == _lib.BK_EUNSUPPORTED:
        raise Unsupported(f"no {steps}-step kernel for stencil {stencil_id}")
    check(rc)


class Unsupported(_lib.BrickError):
    pass


def stencil_list(stencil_id, ids_dev, n, b_in, b_out, coeff=None, stream=None):
    f = _field(b_in, b_out)
    check(load().bk_stencil_apply_list(stencil_id, C.byref(f), ids_dev.ptr, n, _coeff(coeff), stream))


class BrickDecomp:
    """BrickDecomp<3,8,8,8>(dims, depth) + initialize(skin3d_good) (brick-mpi.h:178-513)."""

    def __init__(self, dims, depth=8):
        h = C.c_void_p()
        check(load().bk_decomp_create(C.byref(h), _u3(dims), depth))
        self._h = h
        L = load()
        self.dims, self.depth = tuple(int(x) for x in dims), depth
        self.nbricks = L.bk_decomp_nbricks(h)
        sep, t = (C.c_uint * 3)(), (C.c_uint * 3)()
        check(L.bk_decomp_sep_pos(h, sep))
        check(L.bk_decomp_tdims(h, t))
        self.sep_pos, self.tdims = tuple(sep), tuple(t)
        n = self.tdims[0] * self.tdims[1] * self.tdims[2]
        self.grid = np.ctypeslib.as_array(L.bk_decomp_grid(h), shape=(n,)).reshape(self.tdims[::-1]).copy()
        self.adj = np.ctypeslib.as_array(L.bk_decomp_adj(h), shape=(self.nbricks * 27,)).reshape(-1, 27).copy()
        self.ghost, self.skin = [], []
        for which, dst in ((0, self.ghost), (1, self.skin)):
            for i in range(L.bk_decomp_nregions(h)):
                r = Region()
                check(L.bk_decomp_region(h, which, i, C.byref(r)))
                dst.append(r)
        ss = (C.c_long * 26)()
        check(L.bk_decomp_skin_size(h, ss))
        self.skin_size = list(ss)
        self.rank_map = {}
        self._info = None

    def populate(self, cart, coo):
        """populate(comm, bDecomp, 0, 1, coo) for a periodic Cartesian grid of processes."""
        sets, ranks = (C.c_uint64 * 27)(), (C.c_int * 27)()
        check(load().bk_rank_map((C.c_int * 3)(*cart), (C.c_int * 3)(*coo), sets, ranks))
        self.rank_map = {int(s): int(r) for s, r in zip(sets, ranks)}
        return self.rank_map

    def getBrickInfo(self):
        if self._info is None:
            self._info = BrickInfo(self.adj)
        return self._info

    def id_list(self, which):
        n = load().bk_decomp_list(self._h, which, None)
        ids = np.zeros(n, dtype=np.uint32)
        load().bk_decomp_list(self._h, which, ids.ctypes.data_as(_lib.up))
        return ids

    def exchange_bytes(self, step=BRICK):
        return sum(g.len for g in self.ghost) * step * 8

    def __del__(self):
        try:
            load().bk_decomp_destroy(self._h)
        except Exception:
            pass


def section_range(rank, allsubs, size):
    """Z-Morton id range [l, r) of `rank` when `allsubs` subdomains are dealt to `size` ranks (strong/args.cpp:104-113):
    the first allsubs % size ranks get one more."""
    shift, length = allsubs % size, allsubs // size
    mylen = length + (1 if shift > rank else 0)
    lo = rank * mylen + (0 if shift > rank else shift)
    return lo, lo + mylen


def zmort_decode(idx):
    c = (C.c_ulong * 3)()
    check(load().bk_zmort_decode(idx, c))
    return tuple(c)


def zmort_encode(coord):
    return int(load().bk_zmort_encode((C.c_ulong * 3)(*coord)))


class StitchedGrid:
    """A rank's Z-Morton section of subdomains presented as ONE dense brick grid (bk_stitch_*): entries are global brick
    ids q*nbricks + local id; same-GPU ghosts alias their owners' bricks, only surface regions need the exchange.
    GPU analogue of the mmap ghost aliasing of strong/main.cpp:205-262."""

    def __init__(self, decomp, first, count, subdim):
        self.decomp, self.first, self.count, self.subdim = decomp, int(first), int(count), int(subdim)
        self.box = _lib.StitchBox()
        check(load().bk_stitch_box(self.first, self.count, self.subdim, C.byref(self.box)))
        self.is_box = bool(self.box.is_box)
        self.wrap = tuple(bool(w) for w in self.box.wrap)
        self.n = tuple(int(x) for x in self.box.n)
        self.lo = tuple(int(x) for x in self.box.lo)
        d = (C.c_uint * 3)()
        check(load().bk_stitch_dims(decomp._h, C.byref(self.box), d))
        self.dims = tuple(d)
        self.grid = None
        if self.is_box:
            g = np.zeros(self.dims[0] * self.dims[1] * self.dims[2], dtype=np.uint32)
            check(load().bk_stitch_grid(decomp._h, C.byref(self.box), g.ctypes.data_as(_lib.up)))
            self.grid = g.reshape(self.dims[::-1])

    def region_needed(self, sub_id, region):
        rc = load().bk_stitch_region_needed(self.decomp._h, C.byref(self.box), int(sub_id), int(region))
        if rc < 0:
            raise ValueError("bad subdomain id or region index")
        return bool(rc)

    def sweep_box(self, last=False):
        """brick box a sweep covers: the shell is swept only where it is real ghost storage (communication avoiding),
        never where it aliases my own interior; the last sweep of a period covers the interior only"""
        lo = tuple(1 if (w or last) else 0 for w in self.wrap)
        hi = tuple(d - 1 if (w or last) else d for w, d in zip(self.wrap, self.dims))
        return lo, hi


def section_owner(sub_id, allsubs, size):
    """(rank, index inside the rank) of a Z-Morton id -- inverse of section_range (strong/args.cpp:47-55)"""
    shift, length = allsubs % size, allsubs // size
    split = shift * ((allsubs + size - 1) // size)
    if sub_id < split:
        return sub_id // (length + 1), sub_id % (length + 1)
    return (sub_id - shift) // length, (sub_id - shift) % length


def strong_pull_plan(decomp, rank, size, subdim, stitched=None):
    """The strong-scaling exchange of one rank as pure host data (what drivers/strong.cpp feeds bk_xplan_create):
    one (owner rank, owner subdomain index, source brick, my subdomain index, destination brick, bricks) per ghost region
    -- ghost[i] of my subdomain <- skin[i] of the subdomain in direction ghost[i].neighbor, periodic in the Z-Morton
    arrangement (strong/args.cpp:36-56, strong/main.cu:188-247).  With a StitchedGrid only the regions on the surface of
    my box are kept (bk_stitch_region_needed)."""
    allsubs = subdim ** 3
    lo, hi = section_range(rank, allsubs, size)
    plan = []
    for q in range(hi - lo):
        c = zmort_decode(lo + q)
        for i, (g, sk) in enumerate(zip(decomp.ghost, decomp.skin)):
            if stitched is not None and not stitched.region_needed(lo + q, i):
                continue
            s = int(g.neighbor)
            nb = tuple((c[a] + subdim + (1 if (s >> (a + 1)) & 1 else -1 if (s >> (31 + a + 1)) & 1 else 0)) % subdim
                       for a in range(3))
            owner, sub = section_owner(zmort_encode(nb), allsubs, size)
            plan.append((owner, sub, sk.pos, q, g.pos, g.len))
    return plan


class ExchangeView:
    """One fused pull of every ghost region of a storage: ghost[i] <- peer_base[rank_map[ghost[i].neighbor]] skin[i].

    Stands in for ExchangeView::exchange / BrickDecomp::exchange (brick-mpi.h:96-123, :466-495).  `peer_ptrs[r]` is the
    device address of rank r's storage as seen from THIS process (own pointer, or a CUDA-IPC mapping)."""

    @staticmethod
    def plan(decomp, step=BRICK):
        """the pull plan as pure host data: one (peer rank, src byte offset in the PEER's storage, dst byte offset in
        MINE, bytes) per ghost region -- ghost[i] <- skin[i] of rank_map[ghost[i].neighbor] (brick-mpi.h:476-485)"""
        return [(decomp.rank_map[g.neighbor], s.pos * step * 8, g.pos * step * 8, g.len * step * 8)
                for g, s in zip(decomp.ghost, decomp.skin)]

    def __init__(self, decomp, storage, peer_ptrs, my_rank=0):
        plan = self.plan(decomp, storage.step)
        segs = (Seg * len(plan))()
        self.remote_bytes = 0
        for i, (peer, src_off, dst_off, nbytes) in enumerate(plan):
            segs[i].src = peer_ptrs[peer] + src_off
            segs[i].dst = storage.dat.ptr + dst_off
            segs[i].bytes = nbytes
            if peer != my_rank:
                self.remote_bytes += nbytes
        h = C.c_void_p()
        check(load().bk_xplan_create(C.byref(h), segs, len(decomp.ghost)))
        self._h = h
        self.bytes = load().bk_xplan_bytes(h)

    def set_shape(self, ctas=0, threads=0):
        """launch shape of the pull kernel: (0, 0) = wide default, e.g. (32, 1024) = narrow (see bk_xplan_set_shape)"""
        check(load().bk_xplan_set_shape(self._h, ctas, threads))

    def exchange(self, stream=None):
        check(load().bk_xplan_run(self._h, stream))

    def exchange_sync(self, wait_flags, signal_flags, epoch, stream=None):
        w = (C.c_void_p * max(1, len(wait_flags)))(*wait_flags)
        s = (C.c_void_p * max(1, len(signal_flags)))(*signal_flags)
        check(load().bk_xplan_run_sync(self._h, w, len(wait_flags), s, len(signal_flags), epoch, stream))

    def exchange_gate(self, wait_flags, signal_flags, gate_ptr, epoch, stream=None):
        """pull, then the kernel's last CTA stores `epoch` to the local gate and to the peers' done flags"""
        w = (C.c_void_p * max(1, len(wait_flags)))(*wait_flags)
        s = (C.c_void_p * max(1, len(signal_flags)))(*signal_flags)
        check(load().bk_xplan_run_gate(self._h, w, len(wait_flags), s, len(signal_flags), gate_ptr, epoch, stream))

    def exchange_ce(self, wait_flags=(), signal_flags=(), epoch=0, stream=None):
        """the same plan on the copy engines (no SM taken from the sweeps); flags as in exchange_sync"""
        w = (C.c_void_p * max(1, len(wait_flags)))(*wait_flags)
        s = (C.c_void_p * max(1, len(signal_flags)))(*signal_flags)
        check(load().bk_xplan_run_ce(self._h, w, len(wait_flags), s, len(signal_flags), epoch, stream))

    def __del__(self):
        try:
            load().bk_xplan_destroy(self._h)
        except Exception:
            pass


def array_stencil(stencil_id, arr_in, arr_out, extent, lo, hi, coeff=None, stream=None):
    """the array-layout baseline kernel (arr_kernel, weak/main.cu:27-33): out = stencil(in) for cells lo <= (i,j,k) < hi of
    a plain array with `extent` cells per axis (i first); arr_in / arr_out are DeviceBuffers"""
    check(load().bk_array_stencil_apply(stencil_id, arr_in.ptr, arr_out.ptr, _l3(extent), _l3(lo), _l3(hi), _coeff(coeff),
                                        stream))


def bitset_of(direction):
    """BitSet.set of a direction vector (di, dj, dk) in {-1,0,1}^3: +axis a is bit a, -axis a bit 31+a, axis 1 = i"""
    s = 0
    for a, v in enumerate(direction, start=1):
        if v > 0:
            s |= 1 << a
        elif v < 0:
            s |= 1 << (31 + a)
    return s


class ArrayExchangeView:
    """exchangeArr<3> (include/array-mpi.h:146-213) for DEVICE arrays: the 26 ghost regions of a padded array are pulled
    straight from the neighbours' interior faces / edges / corners as strided boxes by one kernel (bk_xplan_create_boxes)
    -- the reference packs them into 26 buffers, sends, receives and unpacks.  `peer_ptrs[r]` = device address of rank
    r's array as seen from this process; rank_map as filled by populate()."""

    @staticmethod
    def boxes(dom, pad, ghost):
        """pure host data: (direction, src offset, dst offset, cells (i,j,k)) per neighbour, offsets in elements"""
        ext = [dom[a] + 2 * (pad[a] + ghost[a]) for a in range(3)]
        stride = (1, ext[0], ext[0] * ext[1])
        out = []
        for dk in (-1, 0, 1):
            for dj in (-1, 0, 1):
                for di in (-1, 0, 1):
                    v = (di, dj, dk)
                    if v == (0, 0, 0):
                        continue
                    n, so, do = [], 0, 0
                    for a in range(3):
                        p, g, d = pad[a], ghost[a], dom[a]
                        if v[a] > 0:      # my upper ghost <- the neighbour's lowest interior cells
                            n.append(g); src, dst = p + g, p + g + d
                        elif v[a] < 0:    # my lower ghost <- the neighbour's highest interior cells
                            n.append(g); src, dst = p + d, p
                        else:
                            n.append(d); src = dst = p + g
                        so += src * stride[a]
                        do += dst * stride[a]
                    out.append((v, so, do, tuple(n)))
        return out, tuple(ext)

    def __init__(self, dom, pad, ghost, rank_map, arr, peer_ptrs, my_rank=0):
        plan, ext = self.boxes(dom, pad, ghost)
        bx = (_lib.Box * len(plan))()
        self.bytes = 0
        for i, (v, so, do, n) in enumerate(plan):
            peer = rank_map[bitset_of(v)]
            bx[i].src = peer_ptrs[peer] + so * 8
            bx[i].dst = arr.ptr + do * 8
            bx[i].n = (C.c_long * 3)(*n)
            bx[i].src_stride = (C.c_long * 2)(ext[0], ext[0] * ext[1])
            bx[i].dst_stride = (C.c_long * 2)(ext[0], ext[0] * ext[1])
            self.bytes += n[0] * n[1] * n[2] * 8
        h = C.c_void_p()
        check(load().bk_xplan_create_boxes(C.byref(h), bx, len(plan)))
        self._h = h

    def exchange(self, stream=None):
        check(load().bk_xplan_run(self._h, stream))

    def exchange_sync(self, wait_flags, signal_flags, epoch, stream=None):
        w = (C.c_void_p * max(1, len(wait_flags)))(*wait_flags)
        s = (C.c_void_p * max(1, len(signal_flags)))(*signal_flags)
        check(load().bk_xplan_run_sync(self._h, w, len(wait_flags), s, len(signal_flags), epoch, stream))

    def __del__(self):
        try:
            load().bk_xplan_destroy(self._h)
        except Exception:
            pass


def device_sync():
    check(load().bk_device_sync())


class Event:
    def __init__(self):
        e = C.c_void_p()
        check(load().bk_event_create(C.byref(e)))
        self.h = e

    def record(self, stream=None):
        check(load().bk_event_record(self.h, stream))

    def sync(self):
        check(load().bk_event_sync(self.h))

    def elapsed_ms(self, later):
        ms = C.c_float()
        check(load().bk_event_elapsed_ms(self.h, later.h, C.byref(ms)))
        return ms.value
