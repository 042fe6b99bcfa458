# 7-point star with seven independent run-time coefficients coeff[0..6] (spec: reference stencils/7pt.py;
# coefficient order centre, i+1, i-1, j+1, j-1, k+1, k-1 as single/cpu.cpp:11-17 fills them).
from st.expr import ConstRef, Index
from st.grid import Grid

idx = [Index(a) for a in range(3)]
src, dst = Grid("bIn", 3), Grid("bOut", 3)


def at(axis=None, shift=0):
    return src(*[ix + (shift if a == axis else 0) for a, ix in enumerate(idx)])


total = ConstRef("coeff[0]") * at()
slot = 1
for axis in range(3):
    for shift in (+1, -1):
        total = total + ConstRef(f"coeff[{slot}]") * at(axis, shift)
        slot += 1
dst(*idx).assign(total)
STENCIL = [dst]
