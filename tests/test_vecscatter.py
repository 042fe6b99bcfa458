"""brick(...) statements (include/vecscatter.h) and the build step that replaces them (python -m bricklib_b200.vecscatter,
the counterpart of the reference's codegen/vecscatter)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "vecscatter.h"
typedef Brick<Dim<8, 8, 8>, Dim<4, 8>> Brick3D;
#define ST_SCRTPT "%(scripts)s/mpi13pt.py"
#define VSVEC "CUDA"
#define BDIM 8, 8, 8
#define VFOLD 4, 8
#define MPI_B0 0.4
#define MPI_B1 0.07
#define MPI_B2 0.03
void brick_kernel(const BrickLaunch &b, Brick3D &in, Brick3D &out) {
  brick(ST_SCRTPT, VSVEC, (BDIM), (VFOLD), b);
}
void upwind(const BrickLaunch &where, Brick3D &u, Brick3D &v, double W) { brick("%(tests)s/upwind.py", VSVEC, (BDIM), (VFOLD), where); }
'''


def run_tool(tmp_path, text):
    src, dst = tmp_path / "k.cpp", tmp_path / "k-out.cpp"
    src.write_text(text)
    r = subprocess.run([sys.executable, "-m", "bricklib_b200.vecscatter", str(src), str(dst), "--", "-I", os.path.join(ROOT, "include")],
                       capture_output=True, text=True, cwd=ROOT)
    return r, dst


def test_brick_statements_become_tap_tables_and_launches(tmp_path):
    text = SRC % {"scripts": os.path.join(ROOT, "bricklib_b200", "stencils"), "tests": os.path.join(ROOT, "tests", "stencil_scripts")}
    r, dst = run_tool(tmp_path, text)
    assert r.returncode == 0, r.stderr
    assert "2 brick statement(s)" in r.stderr
    out = dst.read_text()
    assert "brick(ST_SCRTPT" not in out and out.count("bk_vs::launch(") == 2
    # macros expanded by the preprocessor pass select the script; its constants are pasted verbatim (free variables)
    assert "mpi13pt.py, 13 taps" in out and "{0, 0, 0, (double) ((MPI_B0))}" in out and "{-2, 0, 0, (double) ((MPI_B2))}" in out
    assert ", in, out, b); } while (false)" in out
    # literal factors and named constants combine into C expressions; the launch descriptor keeps its name
    assert "{-2, 1, 0, (double) (2.0 * (W))}" in out and "{1, -1, 3, (double) (0.25)}" in out and ", u, v, where); }" in out
    assert out.count("\n") == text.count("\n") + 1            # one statement per line: line numbers survive (#line header)
    cc = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(dst)],
                        capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]


def test_nonlinear_scripts_and_tile_are_refused(tmp_path):
    bad = tmp_path / "sq.py"
    bad.write_text("from st.expr import Index\nfrom st.grid import Grid\ni, j, k = Index(0), Index(1), Index(2)\n"
                   "a, b = Grid('a', 3), Grid('b', 3)\nb(i, j, k).assign(a(i, j, k) * a(i + 1, j, k))\nSTENCIL = [b]\n")
    r, _ = run_tool(tmp_path, '#include "vecscatter.h"\nvoid f(const BrickLaunch &L, int &a, int &b) { brick("%s", "CUDA", (8,8,8), (4,8), L); }\n' % bad)
    assert r.returncode != 0 and "not a linear stencil" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_scripts_driver_runs_every_statement_and_validates():
    exe = os.path.join(ROOT, "drivers", "scripts")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is not built: run __graft_entry__.build()")
    r = subprocess.run([exe, "-n", "64", "-r", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("mpi7pt.py", "mpi125pt.py", "7pt.py", "cond.py", "box27_skewed.py", "upwind.py"):
        line = [ln for ln in r.stdout.splitlines() if ln.startswith(name + ":")]
        assert line and "result match" in line[0], r.stdout
