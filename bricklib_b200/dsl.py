"""Lowering of stencil scripts (the reference's stencils/*.py expressions) to tap lists and compiled GPU stencils.

Reference flow: `brick("7pt.py", "CUDA", (8,8,8), (4,8), b)` is expanded at BUILD time by codegen/vecscatter, which
executes the script with the `st` package and prints vector code (vecscatter:82-109).  Here the same script is executed
at RUN time with bricklib_b200.st, which evaluates it into a linear form; `lower()` binds the ConstRef symbols and
returns the taps; `compile_stencil()` hands them to bk_stencil_compile, which picks the kernel family (marching star /
marching cube / per-brick tap table).  Scripts shipped in bricklib_b200/stencils/ restate the reference's five specs.
"""
import ctypes as C
import os
import re
import sys

from . import _lib, st as _st
from ._lib import check, load
from .st import expr as _expr, func as _func, grid as _grid

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT_DIR = os.path.join(HERE, "stencils")

# stencils/fake.h:11-33 -- the MPI_* macros the reference passes with -D to its generated kernels
FAKE_H_CONSTANTS = {
    "MPI_ALPHA": 0.4, "MPI_BETA": 0.1,
    "MPI_A0": 0.1, "MPI_A1": 0.06, "MPI_A2": 0.045, "MPI_A3": 0.03, "MPI_A4": 0.015,
    "MPI_B0": 0.4, "MPI_B1": 0.07, "MPI_B2": 0.03,
    "MPI_C0": 0.1, "MPI_C1": 0.04, "MPI_C2": 0.03, "MPI_C3": 0.01, "MPI_C4": 0.006, "MPI_C5": 0.004,
    "MPI_C6": 0.005, "MPI_C7": 0.002, "MPI_C8": 0.003, "MPI_C9": 0.001,
}


class LoweringError(ValueError):
    pass


class StencilScript:
    """the evaluated script: output grid name, and the linear form of its right-hand side"""

    def __init__(self, path, out_grid, form):
        self.path, self.out_grid, self.form = path, out_grid, form
        self.in_grids = sorted({k[0] for k in form.taps})
        self.dims = len(next(iter(form.taps))[1]) if form.taps else 0
        pres = {k[2] for k in form.taps}
        # pointwise clamps: of every value read (one for the whole stencil) and of the sum
        self.pre = next(iter(pres)) if len(pres) == 1 else (None if not pres else "mixed")
        self.post = form.post
        syms = set(form.free.symbols())
        for p in form.taps.values():
            syms.update(p.symbols())
        self.symbols = sorted(syms)


def script_path(name_or_path):
    if os.path.exists(name_or_path):
        return name_or_path
    cand = os.path.join(SCRIPT_DIR, name_or_path if name_or_path.endswith(".py") else name_or_path + ".py")
    if os.path.exists(cand):
        return cand
    raise FileNotFoundError(name_or_path)


def load_script(name_or_path):
    """execute a stencil script with bricklib_b200.st standing in for the reference's `st` package"""
    path = script_path(name_or_path)
    shim = {"st": _st, "st.expr": _expr, "st.grid": _grid, "st.func": _func}
    saved = {k: sys.modules.get(k) for k in shim}
    sys.modules.update(shim)
    sys.path.insert(0, os.path.dirname(os.path.abspath(path)))   # scripts may share helpers (stencils/_star.py)
    try:
        ns = {"__name__": "__stencil__", "__file__": path}
        exec(compile(open(path).read(), path, "exec"), ns)
    finally:
        sys.path.pop(0)
        sys.modules.pop("_star", None)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    outs = ns.get("STENCIL")
    if not outs:
        raise LoweringError(f"{path}: the script must end with STENCIL = [output grid, ...]")
    if len(outs) != 1:
        raise LoweringError(f"{path}: {len(outs)} output grids; one launch computes one")
    g = outs[0]
    if g.out is None:
        raise LoweringError(f"{path}: grid {g.name} is listed in STENCIL but never assigned")
    return StencilScript(path, g.name, g.out[1])


def _resolver(consts):
    table = dict(FAKE_H_CONSTANTS)
    table.update(consts or {})

    def lookup(name):
        if name in table:
            return float(table[name])
        m = re.fullmatch(r"(\w+)\[(\d+)\]", name)
        if m and m.group(1) in table:
            return float(table[m.group(1)][int(m.group(2))])
        raise LoweringError(f"no value for constant {name!r}: pass consts={{...}}")

    return lookup


def lower(name_or_path, consts=None):
    """-> (taps, script): taps = [((d0, d1, ...), coefficient)] in offset order, d0 = offset along Index(0) = i;
    script.pre / script.post = pointwise clamps (op, constant) of the values read / of the sum, or None:
        out = post( sum_t c_t * pre( in(. + d_t) ) ).
    Raises LoweringError for other non-linear scripts, several input grids or a non-zero grid-free term."""
    sc = name_or_path if isinstance(name_or_path, StencilScript) else load_script(name_or_path)
    f = sc.form
    if f.opaque:
        raise LoweringError(f"{sc.path}: not a linear stencil ({f.opaque}); no tap list exists")
    if len(sc.in_grids) != 1:
        raise LoweringError(f"{sc.path}: reads {len(sc.in_grids)} grids; a sweep has one input field")
    if sc.pre == "mixed":
        raise LoweringError(f"{sc.path}: some reads are clamped and some are not; a kernel applies ONE clamp to the input")
    look = _resolver(consts)
    if f.free.evaluate(look) != 0.0:
        raise LoweringError(f"{sc.path}: constant term {f.free.evaluate(look)}; kernels compute a pure tap sum")
    taps = sorted(((offs, p.evaluate(look)) for (_, offs, _pre), p in f.taps.items()), key=lambda t: t[0][::-1])
    return taps, sc


class CompiledStencil:
    """a stencil lowered from a script and compiled by bk_stencil_compile; apply() is the brick_kernel launch"""
    KINDS = {0: "star", 1: "cube", 2: "taps", 3: "generated"}

    def __init__(self, name_or_path, consts=None):
        taps, sc = lower(name_or_path, consts)
        if sc.dims != 3:
            raise LoweringError(f"{sc.path}: {sc.dims}-D stencil; the brick kernels are 3-D")
        self.script, self.taps = sc, taps
        arr = (_lib.Tap * len(taps))()
        for t, ((di, dj, dk), c) in zip(arr, taps):
            t.di, t.dj, t.dk, t.c = di, dj, dk, c
        h = C.c_void_p()
        self.pre, self.post = sc.pre, sc.post
        if sc.pre is None and sc.post is None:
            rc = load().bk_stencil_compile(C.byref(h), arr, len(taps))
        else:
            pw = [_lib.Pointwise(_lib.POINTWISE_OPS[x[0]], x[1]) if x else _lib.Pointwise(0, 0.0) for x in (sc.pre, sc.post)]
            rc = load().bk_stencil_compile_pointwise(C.byref(h), arr, len(taps), C.byref(pw[0]), C.byref(pw[1]))
        if rc == _lib.BK_EUNSUPPORTED:
            raise LoweringError(f"{sc.path}: {load().bk_last_error().decode()}")
        check(rc)
        self._h = h
        k, r, n, it, fs = (C.c_int() for _ in range(5))
        check(load().bk_stencil_def_info(h, C.byref(k), C.byref(r), C.byref(n), C.byref(it), C.byref(fs)))
        self.kind, self.radius, self.ntaps, self.st_iter, self.fused_steps = (self.KINDS[k.value], r.value, n.value,
                                                                             it.value, fs.value)

    def advance(self, steps, grid, b_in, b_out, lo=None, hi=None, ready=None, part=_lib.PART_ALL,
                kernel=_lib.KERNEL_AUTO, stream=None):
        from . import core
        lo = (0, 0, 0) if lo is None else lo
        hi = grid.dims if hi is None else hi
        f = core._field(b_in, b_out)
        rl, rh = (core._u3(ready[0]), core._u3(ready[1])) if ready else (None, None)
        rc = load().bk_stencil_def_advance(self._h, steps, C.byref(f), grid.dev.ptr, core._u3(grid.dims), core._u3(lo),
                                           core._u3(hi), rl, rh, part, kernel, stream)
        if rc == _lib.BK_EUNSUPPORTED:
            raise core.Unsupported(load().bk_last_error().decode())
        check(rc)

    def apply(self, grid, b_in, b_out, lo=None, hi=None, kernel=_lib.KERNEL_AUTO, stream=None):
        self.advance(1, grid, b_in, b_out, lo, hi, None, _lib.PART_ALL, kernel, stream)

    def source(self):
        """the CUDA source the library generated for this stencil (kind "generated": a marching kernel specialised to the
        tap pattern, compiled with NVRTC -- the counterpart of the text codegen/vecscatter prints); None otherwise"""
        n = C.c_size_t()
        if load().bk_stencil_def_source(self._h, None, 0, C.byref(n)) != _lib.BK_OK:
            return None
        buf = C.create_string_buffer(n.value + 1)
        check(load().bk_stencil_def_source(self._h, buf, n.value + 1, None))
        return buf.value.decode()

    def __del__(self):
        try:
            load().bk_stencil_def_destroy(self._h)
        except Exception:
            pass


def compile_stencil(name_or_path, consts=None):
    return CompiledStencil(name_or_path, consts)


def emit_c(name_or_path, symbol, consts=None):
    """C source of the tap table of a script, for C/C++ callers of bk_stencil_compile -- the build-time role the
    reference gives codegen/vecscatter (a script becomes code), reduced to data:
        static const bk_tap_t <symbol>[] = {{di, dj, dk, c}, ...};  enum { <symbol>_count = N };"""
    taps, sc = lower(name_or_path, consts)
    if sc.dims != 3:
        raise LoweringError(f"{sc.path}: {sc.dims}-D stencil; bk_tap_t is 3-D")
    rows = ",\n".join("  {%d, %d, %d, %r}" % (o[0], o[1], o[2], c) for o, c in taps)
    return (f"/* generated from {os.path.basename(sc.path)} by bricklib_b200.dsl: {sc.out_grid}(i,j,k) = sum of "
            f"{len(taps)} taps of {sc.in_grids[0]} */\n"
            f"static const bk_tap_t {symbol}[] = {{\n{rows}\n}};\nenum {{ {symbol}_count = {len(taps)} }};\n")


def _main(argv):
    import argparse
    ap = argparse.ArgumentParser(prog="python -m bricklib_b200.dsl",
                                 description="lower a stencil script to its tap list (JSON, or a C table with --emit-c)")
    ap.add_argument("script")
    ap.add_argument("--const", action="append", default=[], metavar="NAME=VALUE[,VALUE...]")
    ap.add_argument("--emit-c", metavar="SYMBOL")
    ap.add_argument("--emit-cuda", action="store_true", help="print the CUDA kernel the library generates for the script")
    a = ap.parse_args(argv)
    consts = {}
    for kv in a.const:
        k, v = kv.split("=", 1)
        vals = [float(x) for x in v.split(",")]
        consts[k] = vals if len(vals) > 1 else vals[0]
    if a.emit_cuda:
        src = CompiledStencil(a.script, consts).source()
        if src is None:
            raise SystemExit("this script lowers to a built-in kernel (star / symmetric cube): nothing is generated")
        sys.stdout.write(src)
    elif a.emit_c:
        sys.stdout.write(emit_c(a.script, a.emit_c, consts))
    else:
        import json
        taps, sc = lower(a.script, consts)
        json.dump({"script": sc.path, "in": sc.in_grids, "out": sc.out_grid, "dims": sc.dims,
                   "taps": [list(o) + [c] for o, c in taps]}, sys.stdout)
        sys.stdout.write("\n")


if __name__ == "__main__":
    _main(sys.argv[1:])
