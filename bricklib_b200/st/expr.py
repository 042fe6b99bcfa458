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
    """A stencil expression as a linear form: self.taps[(grid name, offsets, pre)] = Poly, self.free = Poly (grid-free
    part).  Two pointwise non-linearities are representable because a kernel can apply them for free:
      pre  -- a clamp of the VALUE READ, e.g. max(in(i+1,j,k), 0): part of the tap key, ("max", 0.0) / ("min", c) / ("abs", 0.0)
      post -- a clamp of the WHOLE sum, e.g. If(e > 0, e, -e) = abs(e): self.post, same encoding; nothing may be added
              to or multiplied with such an expression afterwards.
    `opaque` marks any other non-linear sub-expression (why lowering will refuse)."""

    def __init__(self, taps=None, free=None, opaque=None, post=None):
        self.taps = taps or {}
        self.free = free if free is not None else Poly()
        self.opaque = opaque
        self.post = post

    # -- construction helpers
    @staticmethod
    def lift(x):
        if isinstance(x, Expr):
            return x
        if isinstance(x, numbers.Real):
            return Expr(free=Poly.const(x))
        raise TypeError(f"cannot use {type(x).__name__} in a stencil expression")

    def _grid_free(self):
        return not self.taps and self.opaque is None and self.post is None

    def _sealed(self):
        """an expression with a post-clamp is final: arithmetic on it is no longer of the form clamp(sum of taps)"""
        return "arithmetic on a clamped sum" if self.post is not None else None

    def _scaled(self, p):
        why = self.opaque or self._sealed()
        return Expr({k: v * p for k, v in self.taps.items()}, self.free * p, why)

    def same_form(self, o):
        """structurally the same linear form (used to recognise If(e > 0, e, -e) as abs(e))"""
        return (self.opaque is None and o.opaque is None and self.post == o.post and
                {k: v.terms for k, v in self.taps.items() if v.terms} == {k: v.terms for k, v in o.taps.items() if v.terms}
                and self.free.terms == o.free.terms)

    # -- arithmetic
    def __add__(self, o):
        o = Expr.lift(o)
        taps = dict(self.taps)
        for k, v in o.taps.items():
            taps[k] = taps[k] + v if k in taps else v
        return Expr(taps, self.free + o.free, self.opaque or o.opaque or self._sealed() or o._sealed())

    __radd__ = __add__

    def __neg__(self):
        return self._scaled(Poly.const(-1.0))

    def __sub__(self, o):
        return self + (-Expr.lift(o))

    def __rsub__(self, o):
        return Expr.lift(o) + (-self)

    def __mul__(self, o):
        o = Expr.lift(o)
        if self.opaque or o.opaque or self._sealed() or o._sealed():
            return Expr(opaque=self.opaque or o.opaque or self._sealed() or o._sealed())
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

    # comparisons only exist to be fed to If()
    def __gt__(self, o):
        return Comparison(self, ">", Expr.lift(o))

    def __ge__(self, o):
        return Comparison(self, ">=", Expr.lift(o))

    def __lt__(self, o):
        return Comparison(self, "<", Expr.lift(o))

    def __le__(self, o):
        return Comparison(self, "<=", Expr.lift(o))

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


class Comparison:
    def __init__(self, lhs, op, rhs):
        self.lhs, self.op, self.rhs = lhs, op, rhs


def clamp_whole(e, op, c):
    """post-clamp of a whole expression: op in max / min / abs"""
    e = Expr.lift(e)
    if e.opaque or e.post is not None:
        return Expr(opaque=e.opaque or "nested clamps")
    return Expr(dict(e.taps), e.free, None, (op, float(c)))


def If(cond, then, otherwise):
    """conditional expression (stencils/cond.py).  The one shape a kernel can take is the absolute value written as a
    select -- If(e > 0, e, -e) and its mirror images; anything else is non-linear."""
    then, otherwise = Expr.lift(then), Expr.lift(otherwise)
    if isinstance(cond, Comparison) and cond.rhs._grid_free() and cond.rhs.free.is_number() and cond.rhs.free.number() == 0:
        e = cond.lhs
        pos, neg = (then, otherwise) if cond.op in (">", ">=") else (otherwise, then)
        if pos.same_form(e) and neg.same_form(-e):
            return clamp_whole(e, "abs", 0.0)
    return Expr(opaque="If")
