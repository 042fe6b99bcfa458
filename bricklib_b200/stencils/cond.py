# rectified 7-point star: every value read is clamped at zero, and the weighted sum is returned as its absolute value,
# written -- as in the reference spec -- with a two-argument max and a select (spec: reference stencils/cond.py)
from st.expr import ConstRef, If, Index
from st.func import Func
from st.grid import Grid

idx = [Index(a) for a in range(3)]
src, dst = Grid("bIn", 3), Grid("bOut", 3)
relu = Func("max", 2)
shifts = [(None, 0)] + [(axis, s) for axis in range(3) for s in (+1, -1)]   # coeff[0..6]: centre, i+1, i-1, j+1, ...

total = 0
for slot, (axis, s) in enumerate(shifts):
    total = total + ConstRef(f"coeff[{slot}]") * relu(src(*[ix + (s if a == axis else 0) for a, ix in enumerate(idx)]), 0.0)
dst(*idx).assign(If(total > 0, total, -total))
STENCIL = [dst]
