# 13-point star of radius 2: MPI_B0 centre, MPI_B1 / MPI_B2 at distance 1 / 2 (spec: reference stencils/mpi13pt.py)
from _star import star

STENCIL = [star(["MPI_B0", "MPI_B1", "MPI_B2"])]
