/*
 * oracle/ref_array.cpp -- TEST INFRASTRUCTURE ONLY.
 * Exposes the reference's scalar array form ST_CPU (stencils/fake.h:44-352; selected at compile time by
 * -DMPI_13PT / -DMPI_25PT / -DMPI_125PT, default 7pt) as one extern "C" sweep over the cell box [lo,hi) of a
 * padded row-major array, the way weak/main.cpp's array_stencil loops use it.  Compiled once per stencil by
 * oracle/Makefile with -DREF_ARRAY_FN=<name>.
 */
#include "stencils/fake.h"

extern "C" void REF_ARRAY_FN(const long *ext, const long *lo, const long *hi, const double *in, double *out) {
  typedef const double (*cin_t)[ext[1]][ext[0]];
  typedef double (*cout_t)[ext[1]][ext[0]];
  cin_t arrIn = (cin_t) in;
  cout_t arrOut = (cout_t) out;
#pragma omp parallel for collapse(2)
  for (long k = lo[2]; k < hi[2]; ++k)
    for (long j = lo[1]; j < hi[1]; ++j)
#pragma omp simd
      for (long i = lo[0]; i < hi[0]; ++i)
        ST_CPU;
}
