#!/usr/bin/env python3
"""oracle/gen_golden_taps.py -- TEST INFRASTRUCTURE.  Golden tap lists from the REFERENCE's own DSL.

Imports the reference's `st` package (/root/reference/codegen/st) -- unmodified -- executes each of its stencil scripts
(/root/reference/stencils/*.py) and walks the AST the reference's code generator would consume: a sum of
coefficient * grid-reference products.  Output: tests/golden/stencil_taps.json =
  {script: {"dims": n, "in": grid, "out": grid, "taps": [[offsets..., "symbolic coefficient"], ...]}}  (cond.py: linear=false)
in the order the script writes its terms.  tests/test_dsl.py checks bricklib_b200.st / dsl.py against it (on the
shipped scripts, and -- where /root/reference exists -- on the reference's scripts themselves).
The reference is not present on the GPU box: run here, commit the JSON.
"""
import glob
import json
import os
import sys

REF = os.environ.get("BRICK_REF", "/root/reference")
sys.path.insert(0, os.path.join(REF, "codegen"))

import st.expr as E  # noqa: E402
import st.grid as G  # noqa: E402
from st.alop import BinaryOperators as B  # noqa: E402


class NotLinear(Exception):
    pass


def terms(node, sign=1):
    """flatten a +/- chain into signed product terms"""
    if isinstance(node, E.BinOp) and node.operator in (B.Add, B.Sub):
        yield from terms(node.lhs, sign)
        yield from terms(node.rhs, sign if node.operator is B.Add else -sign)
    else:
        yield sign, node


def product(node):
    """(coefficient text, GridRef) of a product term"""
    if isinstance(node, G.GridRef):
        return "1", node
    if isinstance(node, E.BinOp) and node.operator is B.Mul:
        sides = [node.lhs, node.rhs]
        ref = [s for s in sides if isinstance(s, G.GridRef)]
        cof = [s for s in sides if isinstance(s, (E.ConstRef, E.IntLiteral, E.FloatLiteral))]
        if len(ref) == 1 and len(cof) == 1:
            return str(cof[0].val), ref[0]
    raise NotLinear(type(node).__name__)


def main():
    out = {}
    for path in sorted(glob.glob(os.path.join(REF, "stencils", "*.py"))):
        ns = {}
        exec(compile(open(path).read(), path, "exec"), ns)
        grid = ns["STENCIL"][0]
        lhs, rhs = grid.out
        name = os.path.basename(path)
        try:
            taps = []
            for sign, t in terms(rhs):
                c, ref = product(t)
                taps.append(list(ref.offsets) + [c if sign > 0 else "-" + c])
            out[name] = {"linear": True, "dims": grid.dims, "out": grid.name, "in": ref.grid.name, "taps": taps}
        except NotLinear as e:
            out[name] = {"linear": False, "why": str(e)}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "stencil_taps.json")
    json.dump(out, open(dst, "w"), indent=0, sort_keys=True)
    print({k: (len(v["taps"]) if v["linear"] else v["why"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
