"""The composed two-step kernel (bricklib_b200/csrc/bk_diamond.h) replayed on the host: tests/cpp/diamond_emulation.cpp
restates the marching kernel's shared-memory layout, copy list, thread mapping and plane loop around the SAME
diamond_plane() the kernel calls and compares with bk_stencil_advance(steps=2) by definition (one step over the grid,
intermediate zero outside it, a second step over the box).  Runs without a GPU."""
import os
import subprocess

import bricklib_b200 as bk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_replay_of_the_composed_kernel_matches_two_plain_steps(tmp_path):
    exe = str(tmp_path / "diamond_emulation")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I", os.path.join(ROOT, "bricklib_b200", "csrc"),
                        os.path.join(ROOT, "tests", "cpp", "diamond_emulation.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "diamond emulation ok" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count(" 0 cells off") == 24   # 8 cases x the three shipped geometries


def test_replay_notices_a_missing_face_correction(tmp_path):
    """the check has teeth: without the grid-face correction every case that touches a grid face must fail"""
    hdr = open(os.path.join(ROOT, "bricklib_b200", "csrc", "bk_diamond.h")).read()
    assert "if (edge != 0u) {" in hdr
    (tmp_path / "bk_diamond.h").write_text(hdr.replace("if (edge != 0u) {", "if (false && edge != 0u) {"))
    exe = str(tmp_path / "mutant")
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I", str(tmp_path), os.path.join(ROOT, "tests", "cpp", "diamond_emulation.cpp"),
                        "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "FAILED (21)" in r.stdout   # all but the interior-box case of each geometry


def test_fused_variant_switch_without_a_gpu():
    before = bk.fused_variant()
    assert before in (bk.FUSED_STAGED, bk.FUSED_COMPOSED, bk.FUSED_COMPOSED_WIDE)
    assert bk.fused_variant(bk.FUSED_COMPOSED) == before and bk.fused_variant() == bk.FUSED_COMPOSED
    assert bk.fused_variant(bk.FUSED_COMPOSED_WIDE) == bk.FUSED_COMPOSED and bk.fused_variant() == bk.FUSED_COMPOSED_WIDE
    assert bk.fused_variant(bk.FUSED_STAGED) == bk.FUSED_COMPOSED_WIDE and bk.fused_variant() == bk.FUSED_STAGED
    try:
        bk.fused_variant(5)
        raise AssertionError("an unknown variant must be refused")
    except bk.BrickError:
        pass
    bk.fused_variant(before)
    env = dict(os.environ, BK_FUSED_VARIANT="composed")
    r = subprocess.run(["python", "-c", "import bricklib_b200 as bk; print(bk.fused_variant())"], capture_output=True, text=True, env=env,
                       cwd=ROOT)
    assert r.stdout.strip() == "1", r.stdout + r.stderr
    r = subprocess.run(["python", "-c", "import bricklib_b200 as bk; print(bk.fused_variant())"], capture_output=True, text=True,
                       env=dict(os.environ, BK_FUSED_VARIANT="wide"), cwd=ROOT)
    assert r.stdout.strip() == "2", r.stdout + r.stderr
