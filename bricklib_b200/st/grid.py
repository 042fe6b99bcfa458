"""Grid: a named field; grid(i + 1, j, k) is a reference at an offset, grid(i, j, k).assign(expr) defines the output
(reference API: codegen/st/grid.py)."""
from .expr import Expr, Index, Poly


class GridRef(Expr):
    def __init__(self, grid, offsets, pre=None):
        super().__init__({(grid.name, offsets, pre): Poly.const(1.0)})
        self.grid, self.offsets, self.pre = grid, offsets, pre

    def clamped(self, op, c):
        """the value read, clamped: max(ref, c) / min(ref, c) / abs(ref)"""
        return GridRef(self.grid, self.offsets, (op, float(c))) if self.pre is None else None

    def assign(self, rhs):
        if any(self.offsets):
            raise ValueError("the output is written at the centre point")
        self.grid.out = (self, Expr.lift(rhs))


class Grid:
    def __init__(self, src_name, dims):
        self.name, self.dims, self.out = str(src_name), int(dims), None

    def __call__(self, *indices):
        if len(indices) != self.dims:
            raise ValueError("Index list not consistent with dimensions")
        offs = [0] * self.dims
        for pos, ix in enumerate(indices):
            if not isinstance(ix, Index):
                raise ValueError("grid arguments are indices, optionally shifted by an integer")
            if ix.n != pos:
                raise ValueError("transposed index order is not supported: argument %d uses Index(%d)" % (pos, ix.n))
            offs[pos] = ix.offset
        if self.out is not None:      # a grid that has been assigned reads as its defining expression ...
            e = self.out[1]
            if not any(offs):
                return e
            # ... moved to the point it is read at: tmp(i+1,j,k) shifts every tap of tmp's definition by (+1,0,0)
            if e.post is not None:
                return Expr(opaque="shifted read of a clamped sum")
            taps = {}
            for (name, o, pre), poly in e.taps.items():
                if len(o) != len(offs):
                    return Expr(opaque="shifted read across grids of different rank")
                taps[(name, tuple(a + b for a, b in zip(o, offs)), pre)] = poly
            return Expr(taps, e.free, e.opaque)
        return GridRef(self, tuple(offs))
