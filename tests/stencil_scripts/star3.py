# radius-3 star with direction-dependent weights: runs on the radius-4 marching kernel
from st.expr import ConstRef, Index
from st.grid import Grid

ix = [Index(a) for a in range(3)]
f, g = Grid("f", 3), Grid("g", 3)
acc = ConstRef("c[0]") * f(*ix)
n = 1
for axis in range(3):
    for d in (1, 2, 3):
        for s in (+1, -1):
            acc = acc + ConstRef(f"c[{n}]") * f(*[x + (s * d if a == axis else 0) for a, x in enumerate(ix)])
            n += 1
g(*ix).assign(acc)
STENCIL = [g]
