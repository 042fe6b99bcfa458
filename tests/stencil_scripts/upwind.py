# asymmetric, off-axis taps with literal and named weights: lowers to the general tap-table kernel
from st.expr import ConstRef, Index
from st.grid import Grid

i, j, k = Index(0), Index(1), Index(2)
u, v = Grid("u", 3), Grid("v", 3)
w = ConstRef("W")
rhs = 0.5 * u(i, j, k) - w * u(i - 1, j, k) + 2 * w * u(i - 2, j + 1, k) + u(i + 1, j - 1, k + 3) / 4 \
    - ConstRef("0.125") * u(i, j, k - 3) + 0.25 * (u(i, j + 2, k) - u(i, j - 2, k))
v(i, j, k).assign(rhs)
STENCIL = [v]
