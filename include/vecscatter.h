/*
 * vecscatter.h -- the brick(...) statement of the reference (include/vecscatter.h:62-90), for the B200 build.
 *
 * Reference flow: inside a __global__ function one writes
 *     brick("../stencils/mpi7pt.py", "CUDA", (8, 8, 8), (4, 8), b);
 * the macro leaves a `#pragma vecscatter Scatter Brick(file, line, script, ...)` behind, and the build runs
 * codegen/vecscatter over the translation unit: it executes the script with the `st` package and REPLACES the statement
 * with generated vector code for brick `b`; the grids the script names (bIn/bOut, in/out) and its constants (MPI_ALPHA,
 * coeff[3]) are free variables of that code, resolved in the caller's scope.
 *
 * Here the statement, the pragma it leaves and the build step are the same -- `python -m bricklib_b200.vecscatter in.cpp
 * out.cpp -- <compiler flags>` replaces it -- but the replacement is HOST code: it lowers the script to its tap table
 * (coefficient expressions pasted verbatim, so they still resolve in the caller's scope), hands it to
 * bk_stencil_compile (star / cube marching kernels, or a kernel GENERATED for the tap pattern and compiled with NVRTC) and
 * launches it over a whole brick box.  One documented difference follows from that: the last argument is not the index
 * of ONE brick inside a device kernel but a BrickLaunch -- the box of bricks the reference's <<<grid>>> would have
 * enumerated (grid, extents, [lo,hi), stream).  A reference kernel + launch
 *     __global__ void brick_kernel(unsigned *grid, Brick3D bIn, Brick3D bOut, unsigned *stride) {
 *       unsigned b = grid[blockIdx.x + (blockIdx.y + blockIdx.z * stride[1]) * stride[0]];
 *       brick(ST_SCRTPT, VSVEC, (BDIM), (VFOLD), b); }
 *     brick_kernel<<<dim3(strideb...), 32>>>(grid_dev, bIn_dev, bOut_dev, stride_dev);
 * becomes
 *     void brick_kernel(const BrickLaunch &b, Brick3D &bIn, Brick3D &bOut) { brick(ST_SCRTPT, VSVEC, (BDIM), (VFOLD), b); }
 *     brick_kernel(BrickLaunch(grid_dev, strideb), bIn_dev, bOut_dev);
 * tile(...) (the array-layout form) is not generated: the array baseline covers the five named stencils only
 * (arrayStencil, array-mpi.h).
 */
#ifndef BRICK_VECSCATTER_H
#define BRICK_VECSCATTER_H

#include <memory>
#include <vector>
#include "brick-b200.h"

#ifndef bElem
#define bElem double
#endif

#define VS_STRING(...) #__VA_ARGS__
#define VS_TOSTR(...) VS_STRING(__VA_ARGS__)
#define _SELECTMACRO(_v0, _v1, _v2, _v3, _v4, _v5, NAME, ...) NAME

/// the brick box a brick(...) statement sweeps: what the reference's launch grid enumerates
struct BrickLaunch {
  const unsigned *grid_dev;
  std::vector<long> gdims, lo, hi;
  void *stream;
  unsigned kernel;
  BrickLaunch(const unsigned *grid_dev, const std::vector<long> &gdims, void *stream = nullptr)
      : grid_dev(grid_dev), gdims(gdims), lo{0, 0, 0}, hi(gdims), stream(stream), kernel(BK_KERNEL_AUTO) {}
  BrickLaunch(const unsigned *grid_dev, const std::vector<long> &gdims, const std::vector<long> &lo, const std::vector<long> &hi,
              void *stream = nullptr)
      : grid_dev(grid_dev), gdims(gdims), lo(lo), hi(hi), stream(stream), kernel(BK_KERNEL_AUTO) {}
};

namespace bk_vs {

/// the compiled stencil behind one brick(...) statement; recompiled only when a coefficient VALUE changes
struct Compiled {
  std::unique_ptr<BrickStencilDef> def;
  std::vector<bk_tap_t> taps;
};

template <typename T>
void launch(Compiled &c, const bk_tap_t *taps, int n, bk_pointwise_t pre, bk_pointwise_t post, T &in, T &out, const BrickLaunch &b) {
  bool same = c.def && (int) c.taps.size() == n;
  for (int i = 0; same && i < n; ++i) same = c.taps[i].c == taps[i].c;
  if (!same) {
    c.taps.assign(taps, taps + n);
    c.def.reset(new BrickStencilDef(c.taps, pre, post));
  }
  c.def->launch(b.grid_dev, b.gdims, in, out, b.lo, b.hi, b.stream, b.kernel);
}

}  // namespace bk_vs

/* the statement: identical pragma text to the reference's _brick5 / _brick6 (vecscatter.h:76,90) */
#define brick(...) _SELECTMACRO(__VA_ARGS__, _brick6, _brick5)(__VA_ARGS__)
#define _brick5(file, vec, vsdim, vsfold, brickIdx) do { _Pragma(VS_TOSTR(vecscatter Scatter Brick(__FILE__, __LINE__, file, VS_TOSTR(bElem), vec, bidx=VS_TOSTR(brickIdx), dim=vsdim, fold=vsfold))) } while (false)
#define _brick6(file, vec, vsdim, vsfold, brickIdx, stri) do { _Pragma(VS_TOSTR(vecscatter Scatter Brick(__FILE__, __LINE__, file, VS_TOSTR(bElem), vec, bidx=VS_TOSTR(brickIdx), dim=vsdim, fold=vsfold, stride=stri))) } while (false)

#endif  // BRICK_VECSCATTER_H
