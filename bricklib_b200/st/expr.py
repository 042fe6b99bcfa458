"""Index, ConstRef, If and the linear-form arithmetic behind stencil expressions (reference API: codegen/st/expr.py)."""
import numbers


class NonLinear(Exception):
    """the expression is not a linear combination of grid references with grid-free coefficients"""


class Poly:
    """A polynomial in ConstRef symbols with float coefficients: {sorted tuple of symbol names: factor}."""
    __slots__ = ("terms",)

    def __init__(self, terms=None):
        self.terms = {k: v for k, v in (terms or {}).items() if v != 0}

    @classmethod
    def const(cls, v):
        return cls({(): float(v)})

    @classmethod
    def symbol(cls, name):
        return cls({(name,): 1.0})

    def __add__(self, o):
        t = dict(self.terms)
        for k, v in o.terms.items():
            t[k] = t.get(k, 0.0) + v
        return Poly(t)

    def __neg__(self):
        return Poly({k: -v for k, v in self.terms.items()})

    def __mul__(self, o):
        t = {}
        for ka, va in self.terms.items():
            for kb, vb in o.terms.items():
                k = tuple(sorted(ka + kb))
                t[k] = t.get(k, 0.0) + va * vb
        return Poly(t)

    def is_number(self):
        return all(k == () for k in self.terms)

    def number(self):
        return self.terms.get((), 0.0)

    def evaluate(self, lookup):
        """lookup(name) -> float"""
        total = 0.0
        for syms, fac in self.terms.items():
            for s in syms:
                fac *= lookup(s)
            total += fac
        return total

    def symbols(self):
        return sorted({s for k in self.terms for s in k})


class Expr:
    """A stencil expression as a linear form: self.taps[(grid name, offsets)] = Poly, self.free = Poly (grid-free part).
    `opaque` marks a non-linear sub-expression (why lowering will refuse)."""

    def __init__(self, taps=None, free=None, opaque=None):
        self.taps = taps or {}
        self.free = free if free is not None else Poly()
        self.opaque = opaque

    # -- construction helpers
    @staticmethod
    def lift(x):
        if isinstance(x, Expr):
            return x
        if isinstance(x, numbers.Real):
            return Expr(free=Poly.const(x))
        raise TypeError(f"cannot use {type(x).__name__} in a stencil expression")

    def _grid_free(self):
        return not self.taps and self.opaque is None

    def _scaled(self, p):
        return Expr({k: v * p for k, v in self.taps.items()}, self.free * p, self.opaque)

    # -- arithmetic
    def __add__(self, o):
        o = Expr.lift(o)
        taps = dict(self.taps)
        for k, v in o.taps.items():
            taps[k] = taps[k] + v if k in taps else v
        return Expr(taps, self.free + o.free, self.opaque or o.opaque)

    __radd__ = __add__

    def __neg__(self):
        return self._scaled(Poly.const(-1.0))

    def __sub__(self, o):
        return self + (-Expr.lift(o))

    def __rsub__(self, o):
        return Expr.lift(o) + (-self)

    def __mul__(self, o):
        o = Expr.lift(o)
        if self.opaque or o.opaque:
            return Expr(opaque=self.opaque or o.opaque)
        if o._grid_free():
            return self._scaled(o.free)
        if self._grid_free():
            return o._scaled(self.free)
        return Expr(opaque="product of two grid references")

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Expr.lift(o)
        if o._grid_free() and o.free.is_number() and o.free.number() != 0:
            return self._scaled(Poly.const(1.0 / o.free.number()))
        return Expr(opaque="division by a symbol or a grid reference")

    def __rtruediv__(self, o):
        return Expr.lift(o) / self

    # comparisons / logic only exist to be fed to If(): they make the result non-linear
    def _cmp(self, o):
        return Expr(opaque="comparison")

    __lt__ = __le__ = __gt__ = __ge__ = _cmp

    def __hash__(self):
        return id(self)


class ConstRef(Expr):
    """A named run-time constant -- a macro (`MPI_B0`), an array element (`coeff[3]`) or a numeric literal (`0.2`)."""

    def __init__(self, v):
        self.name = str(v)
        try:
            poly = Poly.const(float(self.name))
        except ValueError:
            poly = Poly.symbol(self.name)
        super().__init__(free=poly)


class Index:
    """Index(n): the n-th loop index (0 = i, the unit-stride axis).  `i + 2` / `i - 1` give an offset index."""

    def __init__(self, n, offset=0):
        self.n, self.offset = int(n), int(offset)

    def __add__(self, o):
        if not isinstance(o, numbers.Integral):
            raise ValueError("an index may only be shifted by an integer")
        return Index(self.n, self.offset + int(o))

    __radd__ = __add__

    def __sub__(self, o):
        if not isinstance(o, numbers.Integral):
            raise ValueError("an index may only be shifted by an integer")
        return Index(self.n, self.offset - int(o))


def If(cond, then, otherwise):
    """conditional expression (stencils/cond.py) -- representable, never linear"""
    return Expr(opaque="If")
