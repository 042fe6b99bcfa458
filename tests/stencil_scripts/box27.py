# 27-point box: sign- and permutation-symmetric weights (lowers to the cube kernel), plus a skewed variant below
from itertools import product

from st.expr import ConstRef, Index
from st.grid import Grid

ix = [Index(a) for a in range(3)]
a, b = Grid("a", 3), Grid("b", 3)
w = [ConstRef(f"w{n}") for n in range(4)]          # by number of non-zero offsets
acc = 0
for off in product((-1, 0, 1), repeat=3):
    acc = acc + w[sum(o != 0 for o in off)] * a(*[x + o for x, o in zip(ix, off)])
b(*ix).assign(acc)
STENCIL = [b]
