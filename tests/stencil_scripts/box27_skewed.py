# 27-point box whose weight depends on the SIGNED offsets: not symmetric, lowers to the general tap-table kernel
from itertools import product

from st.expr import ConstRef, Index
from st.grid import Grid

ix = [Index(a) for a in range(3)]
a, b = Grid("a", 3), Grid("b", 3)
acc = 0
for n, off in enumerate(product((-1, 0, 1), repeat=3)):
    acc = acc + ConstRef(str((n + 1) / 100.0)) * a(*[x + o for x, o in zip(ix, off)])
b(*ix).assign(acc)
STENCIL = [b]
