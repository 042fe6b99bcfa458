# 25-point star of radius 4: MPI_A0 centre, MPI_A1..MPI_A4 at distance 1..4 (spec: reference stencils/mpi25pt.py)
from _star import star

STENCIL = [star(["MPI_A0", "MPI_A1", "MPI_A2", "MPI_A3", "MPI_A4"])]
