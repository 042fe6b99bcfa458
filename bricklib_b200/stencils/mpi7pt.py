# 7-point star, MPI_ALPHA at the centre and MPI_BETA on the six neighbours (spec: reference stencils/mpi7pt.py)
from _star import star

STENCIL = [star(["MPI_ALPHA", "MPI_BETA"])]
