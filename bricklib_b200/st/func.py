"""Func: an external function applied to expressions, e.g. Func("max", 2) (reference API: codegen/st/func.py).

max / min against a number and fabs / abs are pointwise clamps a kernel applies for free: on a single grid reference
they clamp the value read (stencils/cond.py: max(in(i+1,j,k), 0.0)), on a whole sum they clamp the result.  Any other call
makes the expression non-linear, and it will refuse to lower."""
from .expr import Expr, clamp_whole
from .grid import GridRef


def _number(x):
    x = Expr.lift(x)
    if x._grid_free() and x.free.is_number():
        return x.free.number()
    return None


class Func:
    def __init__(self, name, arity):
        self.name, self.arity = name, arity

    def __call__(self, *args):
        if len(args) != self.arity:
            raise ValueError("Func {} passed wrong number of arguments".format(self.name))
        op = {"max": "max", "fmax": "max", "min": "min", "fmin": "min", "abs": "abs", "fabs": "abs"}.get(self.name)
        target, c = None, 0.0
        if op in ("max", "min") and self.arity == 2:
            for a, b in (args, args[::-1]):
                if _number(b) is not None and _number(a) is None:
                    target, c = a, _number(b)
        elif op == "abs" and self.arity == 1:
            target = args[0]
        if target is None:
            return Expr(opaque=f"call of {self.name}")
        if isinstance(target, GridRef):
            ref = target.clamped(op, c)
            return ref if ref is not None else Expr(opaque="nested clamps")
        return clamp_whole(target, op, c)
