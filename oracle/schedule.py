"""oracle/schedule.py -- TEST INFRASTRUCTURE ONLY.

The weak-scaling time loop of the reference (weak/main.cu:246-287, weak/main.cpp:172-213) restated over an abstract
backend so the SAME driver runs (a) the compiled reference (RefBackend: reference BrickDecomp, reference exchange()
over the in-process MPI stand-in, reference generated brick code), (b) the C port (PortBackend) and (c) the CUDA
product (tests/ supply that backend through the C-ABI).  Per period:  exchange ghost<-skin on the `in` storage, then
ST_ITER sweeps over the whole ghost-inclusive brick grid, ping-ponging in->out->in; the result ends in `in`.
"""
import numpy as np

import oracle

PAD = 8
GZ = 8


def cart_coords(cart):
    """rank -> (c0,c1,c2), c0 slowest: MPI row-major Cartesian order (weak/args.cpp:105-111)."""
    out = []
    for c0 in range(cart[0]):
        for c1 in range(cart[1]):
            for c2 in range(cart[2]):
                out.append((c0, c1, c2))
    return out


def global_origin(cart, coo, dom):
    """Cell origin (i,j,k) of a rank's subdomain inside the periodic global array.

    populate() pairs set element +d with coordinate c-1 (brick-mpi.h:740-751): the rank that supplies my upper-i
    ghost is the one at coo[2]-1, i.e. Cartesian coordinates run AGAINST the axes.  Axis d (1=i) <-> coo[3-d]."""
    return tuple((cart[2 - a] - 1 - coo[2 - a]) * dom[a] for a in range(3))


def split_global(glob, cart, dom):
    fields = []
    for coo in cart_coords(cart):
        o = global_origin(cart, coo, dom)
        fields.append(np.ascontiguousarray(glob[o[2]:o[2] + dom[2], o[1]:o[1] + dom[1], o[0]:o[0] + dom[0]]))
    return fields


def join_global(fields, cart, dom):
    glob = np.zeros((cart[0] * dom[2], cart[1] * dom[1], cart[2] * dom[0]))
    for coo, f in zip(cart_coords(cart), fields):
        o = global_origin(cart, coo, dom)
        glob[o[2]:o[2] + dom[2], o[1]:o[1] + dom[1], o[0]:o[0] + dom[0]] = f
    return glob


def periodic_steps(stencil, glob, steps, coeff=None):
    """independent check: the periodic global array advanced `steps` times with the C port's array form."""
    P = oracle.port()
    r = oracle.RADIUS[stencil]
    cur = glob
    for _ in range(steps):
        padded = np.pad(cur, r, mode="wrap")
        lo = (r, r, r)
        hi = tuple(r + n for n in cur.shape[::-1])
        out = P.sweep_array(stencil, padded, lo, hi, coeff)
        cur = np.ascontiguousarray(out[r:-r, r:-r, r:-r])
    return cur


class _CpuBackend:
    """shared host-side plumbing of RefBackend / PortBackend (storage = aligned numpy, step 512)."""
    step = 512

    def setup(self, dom, cart):
        self.dom, self.cart = dom, cart
        self.coos = cart_coords(cart)
        self.layouts = [self.make_layout(dom, cart, coo) for coo in self.coos]
        nb = self.layouts[0]["nbricks"]
        self.store = [[oracle.aligned_zeros(nb * self.step), oracle.aligned_zeros(nb * self.step)] for _ in self.coos]

    def load(self, fields):
        ext = tuple(n + 2 * GZ for n in self.dom)
        for r, f in enumerate(fields):
            arr = np.zeros(tuple(n + 2 * (PAD + GZ) for n in self.dom[::-1]))
            arr[PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ] = f
            self.to_brick(r, ext, arr)

    def unload(self, which=0):
        out = []
        for r in range(len(self.coos)):
            arr = np.zeros(tuple(n + 2 * (PAD + GZ) for n in self.dom[::-1]))
            self.from_brick(r, arr, which)
            out.append(np.ascontiguousarray(arr[PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ, PAD + GZ:-PAD - GZ]))
        return out


class RefBackend(_CpuBackend):
    def __init__(self):
        self.R = oracle.ref()

    def make_layout(self, dom, cart, coo):
        return self.R.decomp(dom, GZ, cart, coo, keep=True)

    def to_brick(self, r, ext, arr):
        L = self.layouts[r]
        self.R.copy_to_brick(ext, (PAD,) * 3, (0,) * 3, arr, L["grid"], L["adj"], self.store[r][0], self.step)

    def from_brick(self, r, arr, which):
        L = self.layouts[r]
        self.R.copy_from_brick(self.dom, (PAD,) * 3, (GZ,) * 3, arr, L["grid"], L["adj"], self.store[r][which],
                               self.step)

    def exchange(self):
        # the reference's own BrickDecomp::exchange (brick-mpi.h:466-495), one call per rank, then delivery
        for r, L in enumerate(self.layouts):
            self.R.exchange_post(L["handle"], r, self.store[r][0], L["nbricks"], self.step)
        self.R.deliver()

    def sweep(self, stencil, src, dst, skip=0):
        for r, L in enumerate(self.layouts):
            t = L["tdims"]
            self.R.sweep_brick(stencil, L["grid"], (skip,) * 3, tuple(x - skip for x in t), L["adj"],
                               self.store[r][src], self.step, 0, self.store[r][dst], self.step, 0)


class PortBackend(_CpuBackend):
    def __init__(self):
        self.P = oracle.port()

    def make_layout(self, dom, cart, coo):
        L = self.P.decomp(dom, GZ)
        L["rank_map"] = self.P.rank_map(cart, coo)
        return L

    def to_brick(self, r, ext, arr):
        self.P.copy_brick(0, ext, (PAD,) * 3, (0,) * 3, arr, self.layouts[r]["grid"], self.store[r][0], self.step)

    def from_brick(self, r, arr, which):
        self.P.copy_brick(1, self.dom, (PAD,) * 3, (GZ,) * 3, arr, self.layouts[r]["grid"], self.store[r][which],
                          self.step)

    def exchange(self):
        # ghost[i] of rank r <- skin[i] of rank_map[ghost[i].neighbor] (brick-mpi.h:476-485)
        for r, L in enumerate(self.layouts):
            for (nb, _, _, gpos, glen), (_, _, _, spos, _) in zip(L["ghost"], L["skin"]):
                src = self.store[L["rank_map"][nb]][0]
                self.store[r][0][gpos * self.step:(gpos + glen) * self.step] = \
                    src[spos * self.step:(spos + glen) * self.step]

    def sweep(self, stencil, src, dst, skip=0):
        for r, L in enumerate(self.layouts):
            t = L["tdims"]
            self.P.sweep_brick(stencil, L["grid"], (skip,) * 3, tuple(x - skip for x in t), L["adj"],
                               self.store[r][src], self.step, 0, self.store[r][dst], self.step, 0)


def weak_run(backend, stencil, dom, cart, periods, fields, skip_last=False):
    """advance `periods` exchange periods (= periods*ST_ITER time steps); returns per-rank interior arrays."""
    backend.setup(dom, cart)
    backend.load(fields)
    it = oracle.ST_ITER[stencil]
    for _ in range(periods):
        backend.exchange()
        for s in range(it):
            last = s == it - 1
            backend.sweep(stencil, s % 2, 1 - s % 2, skip=1 if (last and skip_last) else 0)
    return backend.unload(0)


def _pointwise(x, pw):
    if not pw:
        return x
    op, c = pw
    return {"max": lambda: np.maximum(x, c), "min": lambda: np.minimum(x, c), "abs": lambda: np.abs(x)}[op]()


def taps_sweep(arr, taps, lo, hi, pre=None, post=None):
    """The meaning of a lowered stencil script, in numpy:
        out[k,j,i] = post( sum_t c_t * pre( arr[k+dk_t, j+dj_t, i+di_t] ) )   for lo <= (i,j,k) < hi, zero elsewhere
    (what codegen/vecscatter's generated loop nest computes for a stencils/*.py expression; pre / post = the pointwise
    clamps of stencils/cond.py, ("max", c) / ("min", c) / ("abs", 0), pinned against the reference's generated cond.py
    code in tests/test_oracle_pin.py).  taps = [((di, dj, dk), c)], summed in the given order."""
    out = np.zeros_like(arr)
    (i0, j0, k0), (i1, j1, k1) = lo, hi
    src = _pointwise(arr, pre)
    acc = np.zeros((k1 - k0, j1 - j0, i1 - i0))
    for (di, dj, dk), c in taps:
        acc += c * src[k0 + dk:k1 + dk, j0 + dj:j1 + dj, i0 + di:i1 + di]
    out[k0:k1, j0:j1, i0:i1] = _pointwise(acc, post)
    return out
