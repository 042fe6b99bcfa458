"""The reference-shaped C++ surface that needs no GPU, compiled with g++ and run: populate(MPI_Comm&, ...) with the
reference's five arguments, the collective mpi_statistics(double, MPI_Comm) over rank threads, BrickDecomp(dims, depth,
numfield), Dim<2,2> (AVX2-fold) bricks through the accessor and refoldBrick (tests/cpp/comm_surface.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_surface_and_folds(tmp_path):
    exe = tmp_path / "comm_surface"
    lib = os.path.join(ROOT, "bricklib_b200")
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-fopenmp", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "cpp", "comm_surface.cpp"), "-o", str(exe), "-L", lib, "-lbrick_b200",
                         f"-Wl,-rpath,{lib}", "-lpthread"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "comm surface ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
