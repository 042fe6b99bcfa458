"""Func: an external function applied to expressions, e.g. Func("max", 2) (reference API: codegen/st/func.py).
A call is never linear, so a stencil using one cannot be lowered to a tap list."""
from .expr import Expr


class Func:
    def __init__(self, name, arity):
        self.name, self.arity = name, arity

    def __call__(self, *args):
        if len(args) != self.arity:
            raise ValueError("Func {} passed wrong number of arguments".format(self.name))
        return Expr(opaque=f"call of {self.name}")
