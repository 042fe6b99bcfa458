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
        if self.out is not None:      # a grid that has been assigned reads as its defining expression
            return self.out[1]
        if len(indices) != self.dims:
            raise ValueError("Index list not consistent with dimensions")
        offs = [0] * self.dims
        for pos, ix in enumerate(indices):
            if not isinstance(ix, Index):
                raise ValueError("grid arguments are indices, optionally shifted by an integer")
            if ix.n != pos:
                raise ValueError("transposed index order is not supported: argument %d uses Index(%d)" % (pos, ix.n))
            offs[pos] = ix.offset
        return GridRef(self, tuple(offs))
