"""hostdev -- an in-process stand-in for the DEVICE side of libbrick_b200's C ABI (test infrastructure, CPU only).

The orchestration above the C ABI -- WeakDomain.period (exchange, flag handshake, split first pass, fused passes), the
end-to-end leg's three fields in flight, tests/handshake_case.py, tools/composed_trial.py -- is Python that only ever runs
on a GPU box.  This module lets the SAME Python run here: `install()` swaps `load()` for a proxy whose device entry points
(bk_dev_alloc, bk_memcpy_*, streams, events, bk_stencil_*, bk_xplan_*, bk_flags_*, layout calls) are implemented on host
memory, while the host-side entry points (decomposition, rank maps, metadata, stitching) stay the real library's.

What it models
  * streams as FIFO queues of operations, executed only when the host blocks (stream / event / device synchronise) --
    so the host runs ahead exactly as it does over CUDA;
  * events with record tickets (a wait captures the latest record at the time of the call, as cudaStreamWaitEvent does);
  * kernels that wait for a flag (k_wait, bk_flags_wait) as operations that are not READY until the flag holds the value;
  * hardware queues: with `hw_queues = k` the streams of a PROCESS share k queues and an operation can only run when
    every earlier operation of its queue has run (the "false dependency" of streams that share a channel); None = a queue
    per stream.  `dev.process = r` before a rank creates its streams says whose they are (ranks emulated in one Python
    process still own their queues, as separate processes do);
  * a schedule policy: "fifo" (oldest runnable operation first) or a seeded random choice among the runnable ones;
  * a deadlock is detected, not suffered: if the host blocks and nothing can run, `Deadlock` names the stuck operations.
  * one SM-resource effect, on request: with `wide_pull_spin=True` the gated pull behaves like the kernel this project
    shipped until round 2 -- the wait for the peers' flags is a spin inside EVERY CTA of the wide pull, so once the pull is
    at the head of its stream it occupies the GPU (no other kernel of that process starts) until the flags arrive.
  * several ranks as THREADS of this process (`RankThreads`): each thread is one rank's host program, blocking calls
    (synchronise, collectives of the stand-in for torch.distributed) hand the turn to the other threads, and a deadlock is
    declared when every thread is blocked and nothing can run.  This is how bench.py's whole main() runs here for N ranks.
What it does not model: SM resources beyond that, timing.  CUDA IPC handles are the addresses themselves.

The arithmetic is the oracle's (oracle/oracle.c through oracle.port(): brick sweeps through the adjacency list, layout
copies) -- this module is a checker's tool like the oracle itself and is imported by tests only.
"""
import collections
import ctypes as C
import functools
import random
import threading

import numpy as np

import bricklib_b200 as bk
from bricklib_b200 import _lib, core, weak

BK_EUNSUPPORTED = -4


class Deadlock(RuntimeError):
    pass


def _raw_sid(stream):
    if stream is None:
        return 0
    return int(stream.value or 0) if hasattr(stream, "value") else int(stream)


def _ival(x):
    return int(x.value or 0) if hasattr(x, "value") else int(x or 0)


def _u64(addr):
    return C.c_uint64.from_address(addr)


class Op:
    __slots__ = ("seq", "stream", "label", "ready", "run", "kernel", "exclusive")

    def __init__(self, seq, stream, label, run, ready=None, kernel=False, exclusive=False):
        self.seq, self.stream, self.label, self.run, self.ready, self.kernel = seq, stream, label, run, ready, kernel
        self.exclusive = exclusive      # while it waits at the head of its stream it holds every CTA slot of its GPU


class HostDev:
    def __init__(self, real, hw_queues=None, policy="fifo", seed=0, wide_pull_spin=False):
        import oracle
        self.real = real                    # the real library: host-side logic only
        self.P = oracle.port()
        self.hw_queues, self.policy, self.rng = hw_queues, policy, random.Random(seed)
        self.bufs = {}                      # address -> numpy buffer (device and pinned allocations)
        self.streams = {0: collections.deque()}
        self.stream_index = {0: ("null", 0)}    # stream -> (owning process, index among that process's streams)
        self._tls = threading.local()
        self.mutex = threading.RLock()
        self.cond = threading.Condition(self.mutex)
        self.live_threads, self.waiters, self.dead = 1, {}, None     # rank threads alive / blocked (with what would let them go on)
        self.wide_pull_spin = wide_pull_spin
        self.resident = {}                  # process -> the exclusive operation that currently occupies its GPU
        self.events = {}                    # handle -> [last recorded ticket, last fired ticket, virtual time of the last firing]
        self.plans = {}
        self.adj_checked = set()
        self.next_handle = 0x1000
        self.seq = 0
        self.clock = 0.0
        self.launches = 0
        self.executed = []                  # labels in execution order (tests look at it)

    # ---- plumbing --------------------------------------------------------------------------------------------------
    def __getattr__(self, name):            # everything not emulated: the real library (host logic only)
        return getattr(self.real, name)

    @property
    def process(self):                      # whose streams are being created: per host thread (= rank)
        return getattr(self._tls, "process", 0)

    @process.setter
    def process(self, value):
        self._tls.process = value

    def block(self, what, can_go_on):
        """the calling rank thread cannot go on until another thread acts: hand over the turn, or declare the deadlock
        when every live thread is blocked, none of them could go on by now, and nothing can run (call with the mutex held;
        returns after a wake-up)"""
        if self.dead is not None:
            raise Deadlock(self.dead)
        me = threading.get_ident()
        self.waiters[me] = can_go_on
        try:
            if len(self.waiters) >= self.live_threads and not any(f() for f in self.waiters.values()) and not self._runnable():
                heads = sorted((q[0].seq, f"stream {sid:#x}: {q[0].label}") for sid, q in self.streams.items() if q)
                self.dead = f"{what}: every rank is blocked and nothing can run; queue heads: " + "; ".join(x for _, x in heads)
                self.dead += self._where_the_ranks_are()
                self.cond.notify_all()
                raise Deadlock(self.dead)
            self.cond.wait(timeout=5.0)
            if self.dead is not None:
                raise Deadlock(self.dead)
        finally:
            self.waiters.pop(me, None)

    def _where_the_ranks_are(self):
        """the innermost frames of every blocked rank thread outside this module (what the deadlock report needs most)"""
        import sys
        import traceback
        out = []
        frames = sys._current_frames()
        for ident in self.waiters:
            fr = frames.get(ident)
            stack = [f for f in traceback.extract_stack(fr) if "hostdev.py" not in f.filename and "threading.py" not in f.filename]
            out.append(" <- ".join(f"{f.name}:{f.lineno}" for f in reversed(stack[-4:])))
        return " || blocked in: " + " || ".join(out)

    def _handle(self):
        self.next_handle += 0x10
        return self.next_handle

    def _sid(self, stream):
        """stream handle -> queue key; the NULL stream is per process (= per rank thread), as every process has its own"""
        sid = _raw_sid(stream)
        if sid != 0 or self.process == 0:
            return sid
        key = -1 - int(self.process)
        if key not in self.streams:
            self.streams[key] = collections.deque()
            self.stream_index[key] = (self.process, 0)
        return key

    def _enqueue(self, stream, label, run, ready=None, kernel=False, exclusive=False):
        sid = self._sid(stream)
        if sid not in self.streams:
            raise AssertionError(f"operation on an unknown stream {sid:#x}")
        self.seq += 1
        self.streams[sid].append(Op(self.seq, sid, label, run, ready, kernel, exclusive))
        if kernel:
            self.launches += 1
        return 0

    def _runnable(self):
        heads = [q[0] for q in self.streams.values() if q]
        if self.hw_queues:
            first = {}
            for sid, q in self.streams.items():
                if q:
                    h = self._queue_of(sid)
                    first[h] = min(first.get(h, q[0].seq), q[0].seq)
            heads = [op for op in heads if op.seq == first[self._queue_of(op.stream)]]
        # a wide kernel that spins in every CTA: once at the head of its stream it becomes resident on its GPU and stays
        # until its flags arrive; meanwhile no other kernel of that process can start
        for proc in {self.stream_index[op.stream][0] for op in heads}:
            if self.resident.get(proc) is None:
                waiting = [op for op in heads if op.exclusive and self.stream_index[op.stream][0] == proc and not op.ready()]
                if waiting:
                    self.resident[proc] = min(waiting, key=lambda o: o.seq) if self.policy == "fifo" else self.rng.choice(waiting)
        out = []
        for op in heads:
            holder = self.resident.get(self.stream_index[op.stream][0])
            if holder is not None and holder is not op and op.kernel:
                continue
            if op.ready is None or op.ready():
                out.append(op)
        return out

    def _queue_of(self, sid):
        proc, idx = self.stream_index[sid]
        return (proc, idx % self.hw_queues)

    def pump(self, until=None, what="device synchronise"):
        """run queued operations until `until()` holds (or everything has run)"""
        while True:
            if until is not None and until():
                return
            ready = self._runnable()
            if not ready:
                if until is None and not any(self.streams.values()):
                    return
                if self.live_threads > 1:       # another rank's host program may still enqueue what this one waits for
                    self.block(what, lambda: (until is not None and until()) or bool(self._runnable()))
                    continue
                if any(self.streams.values()):
                    stuck = sorted((q[0].seq, f"stream {sid:#x}: {q[0].label}") for sid, q in self.streams.items() if q)
                    raise Deadlock(f"{what}: nothing can run; queue heads: " + "; ".join(s for _, s in stuck))
                raise Deadlock(f"{what}: all queues are empty and the condition does not hold")
            op = min(ready, key=lambda o: o.seq) if self.policy == "fifo" else self.rng.choice(ready)
            self.streams[op.stream].popleft()
            if self.resident.get(self.stream_index[op.stream][0]) is op:
                self.resident[self.stream_index[op.stream][0]] = None
            self.clock += 1.0 if op.kernel else 0.01
            op.run()
            self.executed.append(op.label)
            if self.live_threads > 1:
                self.cond.notify_all()

    def _span(self, addr, nbytes=None):
        """(numpy uint8 view from addr to the end of its allocation)"""
        for base, buf in self.bufs.items():
            if base <= addr < base + buf.nbytes:
                v = buf[addr - base:]
                assert nbytes is None or nbytes <= v.nbytes, "access beyond the allocation"
                return v
        raise AssertionError(f"address {addr:#x} is not inside any allocation of this device")

    # ---- runtime ---------------------------------------------------------------------------------------------------
    def bk_device_count(self, n):
        n._obj.value = 1
        return 0

    def bk_set_device(self, dev):
        return 0

    def bk_bind_host_to_device(self):
        return 0

    def _alloc(self, out, nbytes):
        buf = np.zeros(max(int(nbytes), 1) + 64, dtype=np.uint8)
        addr = (buf.ctypes.data + 63) & ~63
        self.bufs[addr] = buf[addr - buf.ctypes.data:addr - buf.ctypes.data + max(int(nbytes), 1)]
        out._obj.value = addr
        return 0

    def bk_dev_alloc(self, out, nbytes):
        return self._alloc(out, nbytes)

    def bk_host_alloc(self, out, nbytes):
        return self._alloc(out, nbytes)

    def bk_dev_free(self, ptr):
        self.bk_device_sync()               # cudaFree synchronises the device
        self.bufs.pop(_ival(ptr), None)
        return 0

    def bk_host_free(self, ptr):
        self.bufs.pop(_ival(ptr), None)
        return 0

    def bk_dev_memset(self, ptr, byte, nbytes, stream):
        a, n = _ival(ptr), int(nbytes)
        return self._enqueue(stream, "memset", lambda: self._span(a, n)[:n].fill(byte))

    def _copy(self, stream, label, dst, src, nbytes):
        d, s, n = _ival(dst), _ival(src), int(nbytes)
        return self._enqueue(stream, label, lambda: C.memmove(d, s, n))

    def bk_memcpy_h2d(self, dev, host, nbytes, stream):
        return self._copy(stream, "h2d", dev, host, nbytes)

    def bk_memcpy_d2h(self, host, dev, nbytes, stream):
        return self._copy(stream, "d2h", host, dev, nbytes)

    def bk_memcpy_d2d(self, dst, src, nbytes, stream):
        return self._copy(stream, "d2d", dst, src, nbytes)

    def bk_stream_create(self, out):
        h = self._handle()
        self.streams[h] = collections.deque()
        self.stream_index[h] = (self.process, sum(1 for p, _ in self.stream_index.values() if p == self.process))
        out._obj.value = h
        return 0

    def bk_stream_create_priority(self, out, high):
        return self.bk_stream_create(out)

    def bk_stream_destroy(self, stream):
        sid = self._sid(stream)
        self.pump(lambda: not self.streams[sid], "bk_stream_destroy")
        del self.streams[sid]
        return 0

    def bk_stream_sync(self, stream):
        sid = self._sid(stream)
        self.pump(lambda: not self.streams[sid], f"bk_stream_sync({sid:#x})")
        return 0

    def _mine(self, me):
        return [q for sid, q in self.streams.items() if self.stream_index[sid][0] == me or (me == 0 and sid == 0)]

    def bk_device_sync(self):
        if self.live_threads > 1:           # a rank's device: its own streams (and the null stream)
            me = self.process               # captured: other rank threads evaluate this condition too (deadlock check)
            self.pump(lambda: not any(self._mine(me)), f"bk_device_sync (rank {me})")
        else:
            self.pump(what="bk_device_sync")
        return 0

    # CUDA IPC between rank threads: the handle is the address
    def bk_ipc_export(self, ptr, handle):
        C.memmove(handle, C.byref(C.c_uint64(_ival(ptr))), 8)
        return 0

    def bk_ipc_open(self, handle, out):
        raw = handle if isinstance(handle, (bytes, bytearray)) else bytes(handle)
        out._obj.value = int.from_bytes(raw[:8], "little")
        return 0

    def bk_ipc_close(self, ptr):
        return 0

    def bk_event_create(self, out):
        h = self._handle()
        self.events[h] = [0, 0, 0.0]
        out._obj.value = h
        return 0

    def bk_event_destroy(self, ev):
        return 0

    def bk_event_record(self, ev, stream):
        e = self.events[_ival(ev)]
        e[0] += 1
        ticket = e[0]

        def fire():
            e[1], e[2] = max(e[1], ticket), self.clock
        return self._enqueue(stream, f"record event {_ival(ev):#x}#{ticket}", fire)

    def bk_event_sync(self, ev):
        e = self.events[_ival(ev)]
        ticket = e[0]
        self.pump(lambda: e[1] >= ticket, f"bk_event_sync({_ival(ev):#x})")
        return 0

    def bk_event_elapsed_ms(self, a, b, out):
        out._obj.value = max(self.events[_ival(b)][2] - self.events[_ival(a)][2], 1e-3)
        return 0

    def bk_stream_wait_event(self, stream, ev):
        e = self.events[_ival(ev)]
        ticket = e[0]                       # the latest record at the time of THIS call
        if ticket == 0:
            return 0
        return self._enqueue(stream, f"wait event {_ival(ev):#x}#{ticket}", lambda: None, lambda: e[1] >= ticket)

    def bk_launch_count(self):
        return self.launches

    # ---- layout ----------------------------------------------------------------------------------------------------
    def _longs(self, v):
        return (C.c_long * 3)(*[int(x) for x in v])

    def bk_copy_to_brick(self, dl, pad, gz, arr, grid, dat, step, stream, direction=0):
        dl, pad, gz = self._longs(dl), self._longs(pad), self._longs(gz)
        a, g, d, st = _ival(arr), _ival(grid), _ival(dat), int(step)
        run = lambda: self.P.L.orc_copy_brick(direction, dl, pad, gz, C.c_void_p(a), C.c_void_p(g), C.c_void_p(d), C.c_size_t(st),  # noqa: E731
                                              C.c_size_t(0))
        return self._enqueue(stream, "copyToBrick" if direction == 0 else "copyFromBrick", run, kernel=True)

    def bk_copy_from_brick(self, dl, pad, gz, arr, grid, dat, step, stream):
        return self.bk_copy_to_brick(dl, pad, gz, arr, grid, dat, step, stream, direction=1)

    def _grid(self, grid, gdims):
        gd = [int(x) for x in gdims]
        n = gd[0] * gd[1] * gd[2]
        return np.frombuffer(self._span(_ival(grid), 4 * n)[:4 * n], dtype=np.uint32).reshape(gd[2], gd[1], gd[0]), gd

    def bk_fill_synthetic(self, grid, gdims, org, glob, seed, dat, step, stream):
        org, glob, d, st = [int(x) for x in org], [int(x) for x in glob], _ival(dat), int(step)
        ga, gd = self._grid(grid, gdims)
        ga = ga.copy()

        def run():
            field = core.synthetic_field(seed, glob, org, [org[a] + 8 * gd[a] for a in range(3)])
            cells = field.reshape(gd[2], 8, gd[1], 8, gd[0], 8).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 512)
            store = np.frombuffer(self._span(d), dtype=np.float64)
            for pos, b in enumerate(ga.ravel()):
                if b:
                    store[int(b) * st:int(b) * st + 512] = cells[pos]
        return self._enqueue(stream, "fill_synthetic", run, kernel=True)

    def bk_compare_storage(self, grid, gdims, lo, hi, a, a_step, b, b_step, tol, mism, maxrel, stream):
        self.bk_device_sync()               # the real call synchronises its stream and frees its scratch (device-wide)
        ga, _ = self._grid(grid, gdims)
        lo, hi = [int(x) for x in lo], [int(x) for x in hi]
        ids = ga[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]].ravel().astype(np.int64)
        A = np.frombuffer(self._span(_ival(a)), dtype=np.float64)
        B = np.frombuffer(self._span(_ival(b)), dtype=np.float64)
        idx = ids[:, None] * int(a_step) + np.arange(512)[None, :]
        jdx = ids[:, None] * int(b_step) + np.arange(512)[None, :]
        x, y = A[idx], B[jdx]
        diff, mag = np.abs(x - y), np.abs(x) + np.abs(y)
        bad = ~((diff < tol) | (diff < mag * tol))
        mism._obj.value = int(bad.sum())
        if maxrel is not None:
            with np.errstate(invalid="ignore", divide="ignore"):
                maxrel._obj.value = float(np.where(mag > 0, diff / np.where(mag > 0, mag, 1.0), 0.0).max())
        self.launches += 1
        return 0

    # ---- stencils --------------------------------------------------------------------------------------------------
    def _sweep(self, stencil, adj, src, src_step, dst, dst_step, grid, gd, lo, hi, coeff):
        cf = None if coeff is None else (C.c_double * 7)(*coeff)
        rc = self.P.L.orc_sweep_brick(stencil, C.c_void_p(grid), self._longs(gd), self._longs(lo), self._longs(hi), C.c_void_p(adj),
                                      C.c_void_p(src), C.c_size_t(src_step), C.c_size_t(0), C.c_void_p(dst), C.c_size_t(dst_step),
                                      C.c_size_t(0), cf)
        assert rc == 0, rc

    def _adjacency_check(self, f, grid, lo, hi, steps, stream):
        """bk_stencil.cu: marching_matches_adjacency -- the first marching launch over a new (adj, grid, box) triple runs a
        check kernel, SYNCHRONISES its stream and frees its result word -- a device-wide wait (cached afterwards; BK_SKIP_ADJ_CHECK skips it).  A host that blocks here
        cannot enqueue anything else meanwhile, which matters when one host thread feeds several emulated ranks."""
        import os
        if os.environ.get("BK_SKIP_ADJ_CHECK"):
            return
        key = (_ival(f._obj.adj), _ival(grid), tuple(int(x) for x in lo), tuple(int(x) for x in hi), steps - 1)
        if key in self.adj_checked:
            return
        self.adj_checked.add(key)
        self._enqueue(stream, "adjacency check", lambda: None, kernel=True)
        self.bk_device_sync()               # cudaStreamSynchronize, then cudaFree of the result word: the whole device

    def _advance(self, stencil, steps, f, grid, gdims, boxes, coeff, stream, label):
        fld = f._obj
        adj, src, dst = _ival(fld.adj), _ival(fld.inp), _ival(fld.out)
        s_in, s_out = int(fld.in_step), int(fld.out_step)
        g, gd = _ival(grid), [int(x) for x in gdims]
        cf = None if coeff is None else [coeff[i] for i in range(7)]
        boxes = [([int(x) for x in lo], [int(x) for x in hi]) for lo, hi in boxes]

        def run():
            if steps == 1:
                for lo, hi in boxes:
                    self._sweep(stencil, adj, src, s_in, dst, s_out, g, gd, lo, hi, cf)
                return
            # two steps = one step over the whole grid (intermediate zero outside it: the null brick), one over the box
            n = self._span(src).nbytes // 8
            tmp = np.zeros(n, dtype=np.float64)
            self._sweep(stencil, adj, src, s_in, tmp.ctypes.data, s_in, g, gd, [0, 0, 0], gd, cf)
            tmp[:512] = 0.0
            for lo, hi in boxes:
                self._sweep(stencil, adj, tmp.ctypes.data, s_in, dst, s_out, g, gd, lo, hi, cf)
        return self._enqueue(stream, label, run, kernel=True)

    def bk_stencil_apply(self, stencil, f, grid, gdims, lo, hi, coeff, flags, stream):
        if flags != _lib.KERNEL_BRICK:
            self._adjacency_check(f, grid, lo, hi, 1, stream)
        return self._advance(stencil, 1, f, grid, gdims, [(lo, hi)], coeff, stream, f"sweep st{stencil}")

    def bk_stencil_advance(self, stencil, steps, f, grid, gdims, lo, hi, coeff, ready_lo, ready_hi, part, stream):
        if steps == 2 and self.real.bk_stencil_radius(stencil) > 2:
            return BK_EUNSUPPORTED
        lo, hi = [int(x) for x in lo], [int(x) for x in hi]
        if not part & _lib.PART_GRID_TOPOLOGY:
            self._adjacency_check(f, grid, lo, hi, steps, stream)
        part &= ~(_lib.PART_THIN | _lib.PART_GRID_TOPOLOGY)
        boxes, name = [(lo, hi)], "ALL"
        if part != _lib.PART_ALL:
            # READY = the bricks whose one-brick neighbourhood lies in the ready box; REST = all the others
            rl, rh = [int(x) for x in ready_lo], [int(x) for x in ready_hi]
            in_lo = [min(max(lo[a], rl[a] + 1), hi[a]) for a in range(3)]
            in_hi = [max(min(hi[a], rh[a] - 1), in_lo[a]) for a in range(3)]
            empty = any(in_hi[a] <= in_lo[a] for a in range(3))
            if part == _lib.PART_READY:
                boxes, name = ([] if empty else [(in_lo, in_hi)]), "READY"
            else:
                boxes, name = ([(lo, hi)] if empty else weak.shell_boxes(tuple(lo), tuple(hi), tuple(in_lo), tuple(in_hi))), "REST"
        return self._advance(stencil, steps, f, grid, gdims, boxes, coeff, stream, f"advance st{stencil} x{steps} {name}")

    def bk_stencil_advance_remote(self, stencil, steps, f, grid, gdims, lo, hi, coeff, ready_lo, ready_hi, part, remap, ghost_lo,
                                  ghost_n, stream):
        """the exchange inside the sweep: when the launch EXECUTES, every ghost brick of the input is read through the
        address table (here: copied through it into the ghost brick first, which no other launch reads in this mode)"""
        if steps == 2 and self.real.bk_stencil_fused_variant_get() == _lib.FUSED_STAGED:
            return BK_EUNSUPPORTED
        fld = f._obj
        src, s_in, table, g0, gn = _ival(fld.inp), int(fld.in_step), _ival(remap), int(ghost_lo), int(ghost_n)

        def through_the_table():
            addr = np.frombuffer(self._span(table, 8 * gn)[:8 * gn], dtype=np.uint64)
            for g in range(gn):
                self._span(int(addr[g]), 4096)          # must be a live brick of this (stand-in) node
                C.memmove(src + (g0 + g) * s_in * 8, int(addr[g]), 4096)
        self._enqueue(stream, "ghost bricks through the address table", through_the_table)
        return self.bk_stencil_advance(stencil, steps, f, grid, gdims, lo, hi, coeff, ready_lo, ready_hi, part, stream)

    def bk_stencil_apply_part(self, stencil, f, grid, gdims, lo, hi, coeff, ready_lo, ready_hi, part, stream):
        return self.bk_stencil_advance(stencil, 1, f, grid, gdims, lo, hi, coeff, ready_lo, ready_hi, part, stream)

    # ---- exchange --------------------------------------------------------------------------------------------------
    def bk_xplan_create(self, out, segs, nseg):
        h = self._handle()
        self.plans[h] = [(_ival(segs[i].src), _ival(segs[i].dst), int(segs[i].bytes)) for i in range(nseg)]
        out._obj.value = h
        return 0

    def bk_xplan_destroy(self, h):
        self.plans.pop(_ival(h), None)
        return 0

    def bk_xplan_bytes(self, h):
        return sum(b for _, _, b in self.plans[_ival(h)])

    def bk_xplan_set_shape(self, h, ctas, threads):
        return 0

    def _pull(self, h):
        for s, d, n in self.plans[_ival(h)]:
            self._span(s, n), self._span(d, n)      # both ends must be live allocations of this device
            C.memmove(d, s, n)

    def _flags(self, arr, n):
        return [_ival(arr[i]) for i in range(n)]

    def bk_xplan_run(self, h, stream):
        return self._enqueue(stream, "pull", lambda: self._pull(h), kernel=True)

    def bk_flags_wait(self, flags, n, value, stream):
        fl, v = self._flags(flags, n), int(value)
        return self._enqueue(stream, f"k_wait(>= {v}) on {len(fl)} flag(s)", lambda: None, lambda: all(_u64(a).value >= v for a in fl),
                             kernel=True)

    def bk_flags_signal(self, flags, n, value, stream):
        fl, v = self._flags(flags, n), int(value)

        def run():
            for a in fl:
                self._span(a, 8)
                _u64(a).value = v
        return self._enqueue(stream, f"k_signal({v}) to {len(fl)} flag(s)", run, kernel=True)

    def bk_xplan_run_sync(self, h, wait, nwait, signal, nsignal, epoch, stream):
        if nwait and self.wide_pull_spin:
            fl, v = self._flags(wait, nwait), int(epoch)
            self._enqueue(stream, "pull that spins in every CTA", lambda: self._pull(h), lambda: all(_u64(a).value >= v for a in fl),
                          kernel=True, exclusive=True)
        else:
            if nwait:
                self.bk_flags_wait(wait, nwait, epoch, stream)
            self.bk_xplan_run(h, stream)
        if nsignal:
            self.bk_flags_signal(signal, nsignal, epoch, stream)
        return 0

    def bk_xplan_run_ce(self, h, wait, nwait, signal, nsignal, epoch, stream):
        """the copy-engine transport: narrow wait kernel, the ranges as copies, a signal kernel -- the same order of events"""
        return self.bk_xplan_run_sync(h, wait, nwait, signal, nsignal, epoch, stream)

    def bk_xplan_run_gate(self, h, wait, nwait, signal, nsignal, gate, epoch, stream):
        spin = bool(nwait) and self.wide_pull_spin
        if nwait and not spin:
            self.bk_flags_wait(wait, nwait, epoch, stream)
        sg, v, g = self._flags(signal, nsignal), int(epoch), _ival(gate) if gate is not None else 0
        fl = self._flags(wait, nwait) if spin else []

        def run():
            self._pull(h)
            for a in sg + ([g] if g else []):
                _u64(a).value = v
        return self._enqueue(stream, "pull that spins in every CTA + done flags" if spin else "pull + done flags", run,
                             (lambda: all(_u64(a).value >= v for a in fl)) if spin else None, kernel=True, exclusive=spin)


def _locked(fn):
    @functools.wraps(fn)
    def call(self, *a, **k):
        with self.mutex:
            if self.dead is not None:
                raise Deadlock(self.dead)
            return fn(self, *a, **k)
    return call


for _name, _fn in list(vars(HostDev).items()):
    if _name.startswith("bk_") and callable(_fn):
        setattr(HostDev, _name, _locked(_fn))


class RankDist:
    """torch.distributed for rank threads: collectives that really wait for every rank (and count as blocked)"""

    def __init__(self, dev, world):
        self.dev, self.world, self.gen, self.slots = dev, world, 0, {}
        self.log = []

    def _collect(self, rank, value, what):
        dev = self.dev
        with dev.mutex:
            gen = self.gen
            box = self.slots.setdefault(gen, {})
            box[rank] = value
            if len(box) == self.world:
                self.gen += 1
                dev.cond.notify_all()
            while len(box) < self.world:
                dev.block(f"collective {what} (rank {rank})", lambda: len(box) >= self.world)
            return [box[r] for r in range(self.world)]

    def for_rank(self, rank):
        return _RankView(self, rank)


class _RankView:
    def __init__(self, dist, rank):
        self.dist, self.rank = dist, rank

    def barrier(self, group=None):
        self.dist._collect(self.rank, None, "host barrier" if group is not None else "barrier")

    def all_gather_object(self, out, obj):
        for i, v in enumerate(self.dist._collect(self.rank, obj, "all_gather_object")):
            out[i] = v

    def allreduce(self, value, op):
        vals = self.dist._collect(self.rank, value, "all_reduce")
        return max(vals) if op == "max" else sum(vals)

    def destroy_process_group(self):
        self.dist._collect(self.rank, None, "destroy")


def run_ranks(dev, world, target):
    """target(rank) on `world` threads, one per rank; returns the list of results; the first exception is re-raised"""
    results, errors = [None] * world, []

    def body(r):
        try:
            dev.process = r
            results[r] = target(r)
        except BaseException as exc:  # noqa: BLE001
            errors.append((r, exc))
            with dev.mutex:
                if dev.dead is None and not isinstance(exc, Deadlock):
                    dev.dead = f"rank {r} failed: {exc!r}"
        finally:
            with dev.mutex:
                dev.live_threads -= 1
                dev.cond.notify_all()

    with dev.mutex:
        dev.live_threads = world
    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    with dev.mutex:
        dev.live_threads = 1
    if errors:
        errors.sort(key=lambda e: isinstance(e[1], Deadlock))      # a real failure first, the deadlocks it caused after
        raise errors[0][1]
    assert not any(t.is_alive() for t in threads), "a rank thread is still running"
    return results


class installed:
    """with hostdev.installed(hw_queues=..., policy=..., seed=...) as dev: ... -- bricklib_b200's load() returns the stand-in"""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.saved = [(m, m.load) for m in (_lib, core, weak, bk)]
        dev = HostDev(_lib.load(), **self.kw)
        for m, _ in self.saved:
            m.load = lambda dev=dev: dev
        self.dev = dev
        return dev

    def __exit__(self, *exc):
        # objects that outlive the block must not hand stand-in addresses / handles to the real library in their __del__
        import gc
        gc.collect()
        for obj in gc.get_objects():
            try:
                if isinstance(obj, core.DeviceBuffer) and getattr(obj, "ptr", None) is not None:
                    obj._owned, obj.ptr = False, None
                elif isinstance(obj, (core.ExchangeView, core.ArrayExchangeView)) and getattr(obj, "_h", None) is not None:
                    obj._h = None
            except ReferenceError:
                pass
        for m, f in self.saved:
            m.load = f
        return False
