# shared by the mpi*pt star scripts: centre weight `weights[0]`, then `weights[d]` for the six points at distance d
from st.expr import ConstRef, Index
from st.grid import Grid


def star(weight_names):
    idx = [Index(a) for a in range(3)]
    src, dst = Grid("in", 3), Grid("out", 3)
    w = [ConstRef(n) for n in weight_names]
    total = w[0] * src(*idx)
    for d in range(1, len(w)):
        for axis in range(3):
            for shift in (+d, -d):
                total = total + w[d] * src(*[ix + (shift if a == axis else 0) for a, ix in enumerate(idx)])
    dst(*idx).assign(total)
    return dst
