"""`st` -- the stencil expression language of the reference's stencil scripts (stencils/*.py import `st.expr`,
`st.grid`, `st.func` from codegen/st/), re-implemented for lowering instead of code generation.

The reference builds an AST and prints vector code from it at build time (codegen/vecscatter).  Here an expression is
evaluated straight into a LINEAR FORM -- {(grid, offsets): coefficient polynomial in the ConstRef symbols} -- because
that is all a B200 kernel launch needs: a tap list (bricklib_b200/dsl.py -> bk_stencil_compile).  Non-linear
constructs (Func calls such as max, If, comparisons: stencils/cond.py) are representable but refuse to lower.
"""
