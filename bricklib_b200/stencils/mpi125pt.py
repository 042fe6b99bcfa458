# 125-point cube of radius 2.  The weight of a point depends only on its sorted absolute offsets, ten classes
# MPI_C0..MPI_C9 for (0,0,0) (0,0,1) (0,0,2) (0,1,1) (0,1,2) (0,2,2) (1,1,1) (1,1,2) (1,2,2) (2,2,2)
# (spec: reference stencils/mpi125pt.py:13-32).
from itertools import product

from st.expr import ConstRef, Index
from st.grid import Grid

idx = [Index(a) for a in range(3)]
src, dst = Grid("in", 3), Grid("out", 3)
classes = [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2), (1, 2, 2), (2, 2, 2)]
weight = {c: ConstRef(f"MPI_C{n}") for n, c in enumerate(classes)}

total = None
for off in product(range(-2, 3), repeat=3):
    term = weight[tuple(sorted(abs(o) for o in off))] * src(*[ix + o for ix, o in zip(idx, off)])
    total = term if total is None else total + term
dst(*idx).assign(total)
STENCIL = [dst]
