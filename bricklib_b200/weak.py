"""The weak-scaling time loop (weak/main.cu:246-287) on one GPU per process.

Per exchange period:  ghost <- neighbours' skin  (one fused pull kernel over the 42 contiguous brick ranges, peers
reached through CUDA-IPC mappings over NVLink; self-neighbours are ordinary device copies in the same kernel), then
ST_ITER sweeps ping-ponging in -> out -> in.  The last sweep of a period skips the ghost shell (weak/main.cpp:209).

Overlap: the first sweep of a period is split in two launches over the SAME tile decomposition of the whole grid
(bk_stencil_apply_part): the CTAs that read only the subdomain's own bricks start at once on the compute stream, the
CTAs that touch the ghost shell run on the high-priority exchange stream right behind the pull kernel.  Both halves
run concurrently; later sweeps wait for both.  Stream-ordered only: nothing spins on the device for the sweep.
"""
import ctypes as C
import os

import numpy as np

from . import _lib, core
from ._lib import check, load

PADDING = 8
GZ = 8


class Handshake:
    """Cross-process flags in IPC-shared device memory.

    flags[r*2+0] on rank p = "rank r's skin is final for epoch e" (written by r, polled by p before pulling),
    flags[r*2+1] on rank p = "rank r has finished pulling p's skin for epoch e" (p polls before overwriting skin)."""

    def __init__(self, world):
        self.world = world
        self.buf = core.DeviceBuffer(max(1, world) * 2 * 8)
        self.buf.zero()
        core.device_sync()
        self.peer = {}  # rank -> base pointer of that rank's flag buffer as seen here

    def ready_flag_on(self, peer, writer):
        return self.peer[peer] + (writer * 2 + 0) * 8

    def done_flag_on(self, peer, writer):
        return self.peer[peer] + (writer * 2 + 1) * 8


class WeakDomain:
    """One subdomain: decomposition, two storages, exchange plan, sweep schedule."""

    def __init__(self, dom, stencil_id, cart=(1, 1, 1), coo=(0, 0, 0), rank=0, kernel=_lib.KERNEL_AUTO):
        self.dom = tuple(dom)
        self.stencil = stencil_id
        self.kernel = kernel
        self.rank = rank
        self.cart = tuple(cart)
        self.coo = tuple(coo)
        self.world = cart[0] * cart[1] * cart[2]
        L = load()
        self.st_iter = L.bk_stencil_st_iter(stencil_id)
        self.decomp = core.BrickDecomp(dom, GZ)
        self.decomp.populate(cart, coo)
        self.info = self.decomp.getBrickInfo()
        self.grid = core.DeviceGrid(self.decomp.grid)
        self.storage = [self.info.allocate(core.BRICK), self.info.allocate(core.BRICK)]
        self.bricks = [core.Brick(self.info, s, 0) for s in self.storage]
        self.peers = sorted({self.decomp.rank_map[g.neighbor] for g in self.decomp.ghost} - {rank})
        self.view = None
        self.hs = None
        self.epoch = 0
        self.comm_stream = None
        self.ev_comm = self.ev_comp = None
        # "ce": ghost ranges move on the copy engines (no SM taken from the sweeps -- the marching kernels fill an SM's
        # register file, so a pull kernel would time-share SMs with them); "kernel": one pull kernel (k_xplan)
        # ghost transport: "kernel" = one pull kernel over NVLink peer mappings (k_xplan; the default -- measured faster at
        # every N), "ce" = faces on the copy engines + narrow kernel for edges/corners (bk_xplan_run_ce; takes no SM, but
        # 104 MB from 7 peers took 0.55 ms at N=8 against ~0.15 ms for the kernel: profiles/r01c_multi_gpu.md)
        self.transport = "kernel"
        # split sweeps with thin k segments for the ghost-dependent layers (BK_PART_THIN): None = when ghost ranges cross
        # NVLink and the stencil radius is <= 2 (measured at N=8, profiles/r01c_multi_gpu.md: 13-pt +8 %, 125-pt +3 %,
        # 7-pt +1 %, 25-pt -2 %); True / False force it
        self.thin = None
        self.trace = None    # developer aid: a list collects (label, Event) marks of one period (tools/period_trace.py)
        self.fuse = 2        # time steps per pass where a fused kernel exists (7/13-point); 1 = one sweep per pass
        # "pull": ghost bricks are filled by the pull kernel, then swept; "direct": THE EXCHANGE INSIDE THE SWEEP -- the
        # launches of pass 0 that touch ghosts read them in place from the neighbours' storages (no pull, no ghost copy)
        self.exchange_mode = os.environ.get("BK_EXCHANGE_MODE", "pull")
        self.remap = None    # direct mode: DeviceBuffer of one address per ghost brick (built by connect())
        # submit the READY half of pass 0 ahead of the pull (see period()); BK_READY_FIRST=1 makes it the default
        self.ready_first = os.environ.get("BK_READY_FIRST", "0") not in ("", "0")

    # ---- wiring -------------------------------------------------------------------------------------------------
    def connect(self, peer_storage_ptrs=None, handshake=None):
        """peer_storage_ptrs[r]: device pointer of rank r's storage[0] visible here (self entry may be omitted)."""
        ptrs = dict(peer_storage_ptrs or {})
        ptrs[self.rank] = self.storage[0].dat.ptr
        self.view = core.ExchangeView(self.decomp, self.storage[0], ptrs, self.rank)
        self.hs = handshake
        # the same plan as one address per ghost BRICK: where the skin brick it mirrors lies (a peer's storage or my own)
        lo, hi = self.decomp.sep_pos[1], self.decomp.sep_pos[2]
        table = np.zeros(hi - lo, dtype=np.uint64)
        brick = self.storage[0].step * 8
        for peer, src_off, dst_off, nbytes in core.ExchangeView.plan(self.decomp, self.storage[0].step):
            first, n = dst_off // brick - lo, nbytes // brick
            table[first:first + n] = ptrs[peer] + src_off + np.arange(n, dtype=np.uint64) * np.uint64(brick)
        assert np.all(table != 0), "every ghost brick mirrors a skin brick"
        self.remap = (core.DeviceBuffer.from_numpy(table), lo, hi - lo)

    def set_pull_shape(self, ctas=0, threads=0):
        self.view.set_shape(ctas, threads)

    def enable_overlap(self):
        s = C.c_void_p()
        check(load().bk_stream_create_priority(C.byref(s), 1))
        self.comm_stream = s
        self.ev_comm, self.ev_comp = core.Event(), core.Event()

    # ---- data ---------------------------------------------------------------------------------------------------
    def load_interior(self, field):
        """field[k][j][i] = interior cells; ghost cells are left zero (filled by the first exchange)."""
        ext = tuple(n + 2 * (PADDING + GZ) for n in self.dom[::-1])
        arr = np.zeros(ext)
        o = PADDING + GZ
        arr[o:-o, o:-o, o:-o] = field
        dev = core.DeviceBuffer.from_numpy(arr)
        strideg = tuple(n + 2 * GZ for n in self.dom)
        core.copyToBrick(strideg, (PADDING,) * 3, (0,) * 3, dev, self.grid, self.bricks[0])
        core.device_sync()
        dev.free()

    def global_origin(self):
        """cell origin (i,j,k) of this subdomain inside the periodic global array.  populate() pairs set element +d with
        Cartesian coordinate c-1 (brick-mpi.h:740-751), i.e. coordinates run AGAINST the axes; axis d (1=i) <-> coo[3-d]"""
        return tuple((self.cart[2 - a] - 1 - self.coo[2 - a]) * self.dom[a] for a in range(3))

    def global_cells(self):
        return tuple(self.cart[2 - a] * self.dom[a] for a in range(3))

    def fill_synthetic(self, seed, which=0, stream=None):
        """storage[which] <- the position-addressable synthetic field (bk_fill_synthetic), ghost shell included (it holds
        the periodic neighbours' values, exactly what the first exchange delivers)"""
        o = self.global_origin()
        core.fill_synthetic(self.grid, self.bricks[which], tuple(x - GZ for x in o), self.global_cells(), seed, stream)

    def read_bricks(self, lo, hi, which=0):
        """cells of the INTERIOR brick box [lo,hi) (brick coordinates without the ghost shell) as a [k][j][i] array"""
        g = GZ // 8
        n = tuple(h - l for l, h in zip(lo, hi))
        out = np.empty((n[2] * 8, n[1] * 8, n[0] * 8))
        one = np.empty(512)
        L = load()
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    b = int(self.decomp.grid[lo[2] + g + k, lo[1] + g + j, lo[0] + g + i])
                    check(L.bk_memcpy_d2h(one.ctypes.data, self.storage[which].brick_ptr(b), 4096, None))
                    check(L.bk_stream_sync(None))
                    out[k * 8:k * 8 + 8, j * 8:j * 8 + 8, i * 8:i * 8 + 8] = one.reshape(8, 8, 8)
        return out

    def read_interior(self, which=0):
        ext = tuple(n + 2 * (PADDING + GZ) for n in self.dom[::-1])
        dev = core.DeviceBuffer(int(np.prod(ext)) * 8)
        dev.zero()
        core.copyFromBrick(self.dom, (PADDING,) * 3, (GZ,) * 3, dev, self.grid, self.bricks[which])
        arr = dev.download(np.float64).reshape(ext)
        dev.free()
        o = PADDING + GZ
        return np.ascontiguousarray(arr[o:-o, o:-o, o:-o])

    # ---- the time loop ------------------------------------------------------------------------------------------
    def _sweep(self, src, dst, lo, hi, stream):
        core.stencil(self.stencil, self.grid, self.bricks[src], self.bricks[dst], lo, hi, None, self.kernel, stream)

    def _announce(self, stream):
        """tell every peer that my skin is final for this epoch (no-op without peers)"""
        if self.hs is None or not self.peers:
            return
        sig = (C.c_void_p * len(self.peers))(*[self.hs.ready_flag_on(p, self.rank) for p in self.peers])
        check(load().bk_flags_signal(sig, len(self.peers), self.epoch, stream))

    def _pull(self, stream, fused_signal=False):
        """pull the neighbours' skins once they have announced them, then tell them I am done reading"""
        hs, e = self.hs, self.epoch
        ce = self._remote() and fused_signal      # only next to overlapped sweeps; else the kernel is faster
        if hs is None or not self.peers:
            if ce:
                self.view.exchange_ce(stream=stream)
            else:
                self.view.exchange(stream)
            return
        waits = [hs.ready_flag_on(self.rank, p) for p in self.peers]
        dones = [hs.done_flag_on(p, self.rank) for p in self.peers]
        if ce:
            self.view.exchange_ce(waits, dones, e, stream)
        elif fused_signal:   # the pull kernel's last CTA raises the done flags itself
            self.view.exchange_gate(waits, dones, None, e, stream)
        else:
            self.view.exchange_sync(waits, dones, e, stream)

    def _exchange(self, stream, fused_signal=False):
        self._announce(stream)
        self._pull(stream, fused_signal)

    def _remote(self):
        """does the exchange run on the copy engines (and the split sweep with thin ghost-dependent segments)?"""
        return self.transport == "ce"

    def _thin(self):
        if self.thin is None:
            return bool(self.peers) and load().bk_stencil_radius(self.stencil) <= 2
        return bool(self.thin)

    def _wait_peers_done(self, stream):
        if self.hs is None or not self.peers:
            return
        w = (C.c_void_p * len(self.peers))(*[self.hs.done_flag_on(self.rank, p) for p in self.peers])
        check(load().bk_flags_wait(w, len(self.peers), self.epoch, stream))

    def _mark(self, label, stream):
        if self.trace is not None:
            ev = core.Event()
            ev.record(stream)
            self.trace.append((label, ev))

    def period(self, stream=None):
        """one exchange + ST_ITER time steps; returns the number of kernel launches issued.

        The steps are issued as passes of `self.fuse` (1 or 2) time steps each (bk_stencil_advance); the first pass is
        split into a READY half on the compute stream and a REST half behind the pull when overlap is enabled."""
        n0 = load().bk_launch_count()
        self.epoch += 1
        t = self.grid.dims
        g = GZ // 8
        full = ((0, 0, 0), t)
        own = ((g,) * 3, tuple(x - g for x in t))        # bricks that are final without the exchange
        fuse = self.steps_per_pass()
        cs = self.comm_stream
        overlap = cs is not None and self.kernel != _lib.KERNEL_BRICK
        self._mark("start", stream)
        if cs is not None:
            self.ev_comp.record(stream)
            check(load().bk_stream_wait_event(cs, self.ev_comp.h))  # previous period's sweeps wrote the skin
        xs = cs if cs is not None else stream
        self._announce(xs)
        # ready_first: the READY half of pass 0 is SUBMITTED before the pull (same streams, same dependencies -- only the
        # order in which the two independent streams are fed changes): its CTAs are on the SMs when the pull's arrive, and
        # the high-priority pull trickles in as they retire instead of taking the machine first.  Which order is faster
        # is a measurement (bench.py times both); correctness does not depend on it.
        if self.direct_active(fuse):
            return self._period_direct(stream, n0, fuse, full, own, cs)
        early = False
        if self.ready_first and overlap and fuse < self.st_iter:
            try:
                thin = _lib.PART_THIN if self._thin() else 0
                self._advance(fuse, 0, 1, full[0], full[1], own, _lib.PART_READY | thin, stream)
                self._mark("pass 0 READY done", stream)
                early = True
            except core.Unsupported:
                pass
        self._pull(xs, fused_signal=cs is not None)
        self._mark("exchange done", xs)
        done, p = 0, 0
        while done < self.st_iter:
            src, dst = p % 2, 1 - p % 2
            last = done + fuse >= self.st_iter
            lo, hi = own if last else full
            if p == 1:
                self._wait_peers_done(stream)  # pass 1 rewrites storage[0], whose skin the peers were reading
            if p == 0 and overlap and not last:
                try:
                    thin = _lib.PART_THIN if self._thin() else 0
                    if not early:
                        self._advance(fuse, src, dst, lo, hi, own, _lib.PART_READY | thin, stream)
                        self._mark("pass 0 READY done", stream)
                    self._advance(fuse, src, dst, lo, hi, own, _lib.PART_REST | thin, cs)
                    self._mark("pass 0 REST done", cs)
                    self.ev_comm.record(cs)
                    check(load().bk_stream_wait_event(stream, self.ev_comm.h))
                    done, p = done + fuse, p + 1
                    continue
                except core.Unsupported:
                    if early:
                        raise
                    if fuse == 2:
                        fuse = 1
                        continue
                    overlap = False
            if p == 0 and cs is not None:
                self.ev_comm.record(cs)
                check(load().bk_stream_wait_event(stream, self.ev_comm.h))
            try:
                self._advance(fuse, src, dst, lo, hi, None, _lib.PART_ALL, stream)
            except core.Unsupported:
                if fuse == 1:
                    raise
                fuse = 1
                continue
            self._mark(f"pass {p} done", stream)
            done, p = done + fuse, p + 1
        assert p % 2 == 0 or self.st_iter % 2 == 1, "result must end in storage[0]"
        return load().bk_launch_count() - n0

    def direct_active(self, fuse=None):
        """does period() run with the exchange inside the sweep?  Asked for (exchange_mode == "direct"), wired (connect()),
        and a kernel for it exists: every marching kernel but the STAGED two-step one (7-point needs
        BK_FUSED_VARIANT=composed|wide, or one sweep per pass); else period() uses the pull"""
        if self.exchange_mode != "direct" or self.remap is None or self.kernel == _lib.KERNEL_BRICK:
            return False
        fuse = self.steps_per_pass() if fuse is None else fuse
        return not (fuse == 2 and load().bk_stencil_fused_variant_get() == _lib.FUSED_STAGED)

    def _period_direct(self, stream, n0, fuse, full, own, cs):
        """the rest of a period with THE EXCHANGE INSIDE THE SWEEP (after the announcement): pass 0 reads the ghost bricks
        in place from the neighbours' storages (bk_stencil_advance_remote), so there is no pull, no ghost write and no
        re-read.  With an exchange stream: READY half on the compute stream at once, the REST half (the only launches that
        touch ghosts) on the exchange stream behind a wait for the neighbours' announcements; then I tell them I am done
        reading their skins.  Pass 1 waits for their same message as in pull mode."""
        have_peers = self.hs is not None and bool(self.peers)
        xs = cs if cs is not None else stream
        waits = [self.hs.ready_flag_on(self.rank, p) for p in self.peers] if have_peers else []
        dones = [self.hs.done_flag_on(p, self.rank) for p in self.peers] if have_peers else []

        def wait_for_neighbours(s):
            if have_peers:
                w = (C.c_void_p * len(waits))(*waits)
                check(load().bk_flags_wait(w, len(waits), self.epoch, s))

        def done_reading(s):
            if have_peers:
                d = (C.c_void_p * len(dones))(*dones)
                check(load().bk_flags_signal(d, len(dones), self.epoch, s))

        last0 = fuse >= self.st_iter
        lo, hi = own if last0 else full
        thin = _lib.PART_THIN if self._thin() else 0
        if cs is not None and not last0:
            self._advance(fuse, 0, 1, lo, hi, own, _lib.PART_READY | thin, stream)          # reads no ghost: plain kernel
            self._mark("pass 0 READY done", stream)
            wait_for_neighbours(cs)
            self._advance(fuse, 0, 1, lo, hi, own, _lib.PART_REST | thin, cs, remote=self.remap)
            done_reading(cs)
            self._mark("pass 0 REST done", cs)
            self.ev_comm.record(cs)
            check(load().bk_stream_wait_event(stream, self.ev_comm.h))
        else:
            wait_for_neighbours(xs)
            core.stencil_advance(self.stencil, fuse, self.grid, self.bricks[0], self.bricks[1], lo, hi, None, _lib.PART_ALL, None,
                                 xs, remote=self.remap)
            done_reading(xs)
            if cs is not None:
                self.ev_comm.record(cs)
                check(load().bk_stream_wait_event(stream, self.ev_comm.h))
            self._mark("pass 0 done", xs)
        done, p = fuse, 1
        while done < self.st_iter:
            src, dst = p % 2, 1 - p % 2
            last = done + fuse >= self.st_iter
            lo, hi = own if last else full
            if p == 1:
                self._wait_peers_done(stream)  # pass 1 rewrites storage[0], whose skin the peers were reading
            self._advance(fuse, src, dst, lo, hi, None, _lib.PART_ALL, stream)
            self._mark(f"pass {p} done", stream)
            done, p = done + fuse, p + 1
        assert p % 2 == 0 or self.st_iter % 2 == 1, "result must end in storage[0]"
        return load().bk_launch_count() - n0

    def steps_per_pass(self):
        """time steps one launch advances: 2 where the fused kernel pays off (7-point), else 1"""
        fuse = min(max(1, self.fuse), load().bk_stencil_fused_steps(self.stencil))
        if self.kernel == _lib.KERNEL_BRICK or self.st_iter % fuse:
            fuse = 1
        return fuse

    def _advance(self, steps, src, dst, lo, hi, ready, part, stream, remote=None):
        if steps == 1 and part == _lib.PART_ALL and remote is None:
            self._sweep(src, dst, lo, hi, stream)
        else:
            core.stencil_advance(self.stencil, steps, self.grid, self.bricks[src], self.bricks[dst], lo, hi, ready, part,
                                 None, stream, remote=remote)


class ArrayDomain:
    """The reference's array-layout baseline loop ("Arr:" block, weak/main.cu:161-213) on the device: two padded arrays,
    per period one exchangeArr (ArrayExchangeView: 26 strided boxes pulled by one kernel) and ST_ITER arr_kernel sweeps
    over the WHOLE ghost-inclusive region, ping-ponging in -> out -> in (the rim the sweeps cannot compute correctly
    grows by one radius per sweep and is exactly consumed by the ghost depth, as with bricks)."""

    def __init__(self, dom, stencil_id, cart=(1, 1, 1), coo=(0, 0, 0), rank=0):
        self.dom, self.stencil, self.rank = tuple(dom), stencil_id, rank
        self.cart, self.coo = tuple(cart), tuple(coo)
        self.st_iter = load().bk_stencil_st_iter(stencil_id)
        self.ext = tuple(n + 2 * (PADDING + GZ) for n in self.dom)
        nbytes = int(np.prod(self.ext)) * 8
        self.arr = [core.DeviceBuffer(nbytes), core.DeviceBuffer(nbytes)]
        for a in self.arr:
            a.zero()
        sets, ranks = (C.c_uint64 * 27)(), (C.c_int * 27)()
        check(load().bk_rank_map((C.c_int * 3)(*cart), (C.c_int * 3)(*coo), sets, ranks))
        self.rank_map = {int(s): int(r) for s, r in zip(sets, ranks)}
        self.peers = sorted(set(self.rank_map.values()) - {rank})
        self.view = None

    def connect(self, peer_ptrs=None):
        ptrs = dict(peer_ptrs or {})
        ptrs[self.rank] = self.arr[0].ptr
        self.view = core.ArrayExchangeView(self.dom, (PADDING,) * 3, (GZ,) * 3, self.rank_map, self.arr[0], ptrs, self.rank)

    def load_interior(self, field):
        host = np.zeros(self.ext[::-1])
        o = PADDING + GZ
        host[o:-o, o:-o, o:-o] = field
        self.arr[0].upload(host)

    def read_interior(self, which=0):
        o = PADDING + GZ
        return np.ascontiguousarray(self.arr[which].download(np.float64).reshape(self.ext[::-1])[o:-o, o:-o, o:-o])

    def sweeps(self, stream=None):
        lo = (PADDING,) * 3
        hi = tuple(e - PADDING for e in self.ext)
        for s in range(self.st_iter):
            core.array_stencil(self.stencil, self.arr[s % 2], self.arr[1 - s % 2], self.ext, lo, hi, None, stream)

    def period(self, stream=None):
        """exchange + ST_ITER sweeps; single process (peers on the same host thread order themselves by the stream)"""
        self.view.exchange(stream)
        self.sweeps(stream)


def shell_boxes(lo, hi, in_lo, in_hi):
    """the six slabs of box [lo,hi) minus inner box [in_lo,in_hi): two k-slabs, two j-slabs, two i-slabs"""
    out = []
    if in_lo[2] > lo[2]:
        out.append((lo, (hi[0], hi[1], in_lo[2])))
    if hi[2] > in_hi[2]:
        out.append(((lo[0], lo[1], in_hi[2]), hi))
    if in_lo[1] > lo[1]:
        out.append(((lo[0], lo[1], in_lo[2]), (hi[0], in_lo[1], in_hi[2])))
    if hi[1] > in_hi[1]:
        out.append(((lo[0], in_hi[1], in_lo[2]), (hi[0], hi[1], in_hi[2])))
    if in_lo[0] > lo[0]:
        out.append(((lo[0], in_lo[1], in_lo[2]), (in_lo[0], in_hi[1], in_hi[2])))
    if hi[0] > in_hi[0]:
        out.append(((in_hi[0], in_lo[1], in_lo[2]), (hi[0], in_hi[1], in_hi[2])))
    return out


class FieldPipeline:
    """Fields that live in HOST memory streamed through the GPU: every step uploads the field's interior bricks from pinned
    host memory (H2D), runs one period (exchange + ST_ITER sweeps) and downloads the result bricks (D2H).  Steps are
    independent fields, so len(doms) of them are kept in flight (one uploading, one computing, one downloading), as a
    user streaming fields through the GPU would do.  Each slot has its own upload, compute and download stream, so both
    PCIe directions are busy all the time: the step time is bounded by one direction's copy."""

    def __init__(self, doms, one_period_in_flight=True):
        self.L, self.ck = load(), check
        L, ck = self.L, self.ck
        d0 = doms[0]
        lo, hi = 1, d0.decomp.sep_pos[1]        # inner + skin bricks = the interior; ghosts come from the exchange
        self.off, self.nbytes = lo * 512 * 8, (hi - lo) * 512 * 8
        self.serial, self.last_run, self.slots = one_period_in_flight, None, []
        for d in doms:
            hin, hout = C.c_void_p(), C.c_void_p()
            ck(L.bk_host_alloc(C.byref(hin), self.nbytes))
            ck(L.bk_host_alloc(C.byref(hout), self.nbytes))
            st = [C.c_void_p() for _ in range(3)]   # upload, compute, download
            for x in st:
                ck(L.bk_stream_create(C.byref(x)))
            ev = [core.Event() for _ in range(3)]     # uploaded, computed, downloaded
            ck(L.bk_memcpy_d2h(hin, d.storage[0].dat.ptr + self.off, self.nbytes, None))
            self.slots.append((d, hin, hout, st, ev))
        core.device_sync()
        for _, _, _, _, ev in self.slots:
            for e in ev:
                e.record(None)
        core.device_sync()

    def step(self, i):
        L, ck = self.L, self.ck
        d, hin, hout, (s_up, s_run, s_down), (e_up, e_run, e_down) = self.slots[i % len(self.slots)]
        ck(L.bk_stream_wait_event(s_up, e_down.h))         # this slot's previous result has left the device
        ck(L.bk_memcpy_h2d(d.storage[0].dat.ptr + self.off, hin, self.nbytes, s_up))
        e_up.record(s_up)
        ck(L.bk_stream_wait_event(s_run, e_up.h))
        # ONE period in flight per GPU: a period (2 ms next to 25 ms of copies) starts when the previous step's period has
        # finished, so every rank runs the periods of all slots in one global order -- the ordering the single-domain loop
        # is proven with.  Periods of different slots in flight at once would put kernels that wait for a peer's flag on
        # several streams, and two ranks could then wait for each other (cross-slot, through shared hardware queues):
        # tests/test_hostdev.py reproduces that deadlock on the CPU stand-in for the device.
        if self.serial and self.last_run is not None:
            ck(L.bk_stream_wait_event(s_run, self.last_run.h))
        d.period(s_run)
        e_run.record(s_run)
        self.last_run = e_run
        ck(L.bk_stream_wait_event(s_down, e_run.h))
        ck(L.bk_memcpy_d2h(hout, d.storage[0].dat.ptr + self.off, self.nbytes, s_down))
        e_down.record(s_down)

    def host_in(self, slot):
        """slot's pinned input buffer as a float64 array: the interior bricks (ids [1, sep_pos[1])), 512 cells each"""
        return np.ctypeslib.as_array((C.c_double * (self.nbytes // 8)).from_address(self.slots[slot][1].value))

    def host_out(self, slot):
        """slot's pinned result buffer (valid after the step that used the slot has been synchronised)"""
        return np.ctypeslib.as_array((C.c_double * (self.nbytes // 8)).from_address(self.slots[slot][2].value))

    def sync(self):
        for _, _, _, st, _ in self.slots:
            for x in st:
                self.ck(self.L.bk_stream_sync(x))

    def close(self):
        for _, hin, hout, st, _ in self.slots:
            self.L.bk_host_free(hin)
            self.L.bk_host_free(hout)
            for x in st:
                self.L.bk_stream_destroy(x)
        self.slots = []
