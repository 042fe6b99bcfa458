/*
 * bricksetup.h -- host-side grid set-up and array <-> brick conversion.
 *
 * Same entry points as the reference's include/bricksetup.h: init_grid<dims> (:73-90), copyToBrick<dims> (:172-200),
 * copyFromBrick<dims> (:213-221).  Vectors are contiguous-axis FIRST ({i,j,k}), as in the reference.  These run on
 * the host over host-resident bricks (set-up / validation); the device-side twins are copyToBrickDevice /
 * copyFromBrickDevice in brick-b200.h (C ABI: bk_copy_to_brick / bk_copy_from_brick).
 */
#ifndef BRICKSETUP_H
#define BRICKSETUP_H

#include <vector>
#include "brick.h"

/// Regular non-periodic grid: grid[p] = p; neighbours at +-stride, anything outside the LINEAR range maps to brick 0.
template <unsigned dims>
BrickInfo<dims> init_grid(unsigned *&grid_ptr, const std::vector<long> &dimlist) {
  std::vector<long> stride(dims);
  long size = 1;
  for (unsigned d = 0; d < dims; ++d) stride[d] = size, size *= dimlist[d];
  grid_ptr = (unsigned *) malloc(size * sizeof(unsigned));
  BrickInfo<dims> info((unsigned) size);
  constexpr unsigned nn = static_power<3, dims>::value;
#pragma omp parallel for
  for (long p = 0; p < size; ++p) {
    grid_ptr[p] = (unsigned) p;
    for (unsigned s = 0; s < nn; ++s) {
      long q = p;
      unsigned t = s;
      for (unsigned d = 0; d < dims; ++d, t /= 3) q += ((long) (t % 3) - 1) * stride[d];
      info.adj[p][s] = (q < 0 || q >= size) ? 0u : (unsigned) q;
    }
  }
  return info;
}

namespace brick_detail {
template <unsigned dims, class T, class F>
void for_each_cell(const std::vector<long> &dimlist, const std::vector<long> &padding, const std::vector<long> &ghost,
                   bElem *arr, unsigned *grid_ptr, T &brick, F f) {
  static_assert(dims == T::DIMS && dims <= 4, "dimension mismatch");
  std::vector<long> tile(dims), strideA(dims), strideB(dims), nb(dims);
  long ext[dims];  // brick extents, contiguous axis first
  T::extents_fast_first(ext);
  long sizeA = 1, sizeB = 1, cells = 1, nbricks = 1;
  for (unsigned d = 0; d < dims; ++d) {
    tile[d] = ext[d];
    strideA[d] = sizeA, strideB[d] = sizeB;
    sizeA *= dimlist[d] + 2 * (padding[d] + ghost[d]);
    sizeB *= (dimlist[d] + 2 * ghost[d]) / tile[d];
    nb[d] = dimlist[d] / tile[d];
    cells *= tile[d], nbricks *= nb[d];
  }
#pragma omp parallel for
  for (long n = 0; n < nbricks; ++n) {
    long rem = n, aoff = 0, goff = 0;
    for (unsigned d = 0; d < dims; ++d) {
      const long s = ghost[d] / tile[d] + rem % nb[d];
      rem /= nb[d];
      aoff += (padding[d] + s * tile[d]) * strideA[d];
      goff += s * strideB[d];
    }
    const unsigned b = grid_ptr[goff];
    for (long c = 0; c < cells; ++c) {
      long r2 = c, o = 0;
      int idx[dims];
      for (unsigned d = 0; d < dims; ++d) {
        const long e = r2 % tile[d];
        r2 /= tile[d];
        o += e * strideA[d];
        idx[dims - 1 - d] = (int) e;  // Brick::elem wants the slowest axis first
      }
      f(brick.elem(b, idx), arr + aoff + o);
    }
  }
}
}  // namespace brick_detail

template <unsigned dims, typename T>
inline void copyToBrick(const std::vector<long> &dimlist, const std::vector<long> &padding, const std::vector<long> &ghost,
                        bElem *arr, unsigned *grid_ptr, T &brick) {
  brick_detail::for_each_cell<dims>(dimlist, padding, ghost, arr, grid_ptr, brick, [](bElem &b, bElem *a) { b = *a; });
}
template <unsigned dims, typename T>
inline void copyToBrick(const std::vector<long> &dimlist, bElem *arr, unsigned *grid_ptr, T &brick) {
  const std::vector<long> zero(dimlist.size(), 0);
  copyToBrick<dims>(dimlist, zero, zero, arr, grid_ptr, brick);
}
template <unsigned dims, typename T>
inline void copyFromBrick(const std::vector<long> &dimlist, const std::vector<long> &padding,
                          const std::vector<long> &ghost, bElem *arr, unsigned *grid_ptr, T &brick) {
  brick_detail::for_each_cell<dims>(dimlist, padding, ghost, arr, grid_ptr, brick, [](bElem &b, bElem *a) { *a = b; });
}

#endif  // BRICKSETUP_H
