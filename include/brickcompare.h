/*
 * brickcompare.h -- element-wise comparison of a brick field with a plain array (host side).
 *
 * compareBrick<dims>(dimlist, padding, ghost, arr, grid, brick) as in the reference's include/brickcompare.h:30-78,
 * with the tolerance rule of :36-37 (|a-b| < tol or |a-b| < (|a|+|b|)*tol).  The reference fixes tol = 1e-6
 * (include/cmpconst.h:9); this build's contract is 1e-12, so the tolerance is a variable (default 1e-12).
 * Device-side twin: compareBrickDevice in brick-b200.h (C ABI: bk_compare_brick).
 */
#ifndef BRICKCOMPARE_H
#define BRICKCOMPARE_H

#include <cmath>
#include <iostream>
#include "bricksetup.h"

#ifndef BRICK_TOLERANCE
#define BRICK_TOLERANCE 1e-12
#endif

namespace brick_detail {
inline bool close_enough(bElem a, bElem b, double tol) {
  const double diff = std::abs(a - b);
  return diff < tol || diff < (std::abs(a) + std::abs(b)) * tol;
}
}  // namespace brick_detail

template <unsigned dims, typename T>
inline bool compareBrick(const std::vector<long> &dimlist, const std::vector<long> &padding,
                         const std::vector<long> &ghost, bElem *arr, unsigned *grid_ptr, T &brick,
                         double tol = BRICK_TOLERANCE) {
  long bad = 0;
  brick_detail::for_each_cell<dims>(dimlist, padding, ghost, arr, grid_ptr, brick, [&](bElem &b, bElem *a) {
    if (!brick_detail::close_enough(b, *a, tol)) {
#pragma omp atomic
      ++bad;
    }
  });
  if (bad) std::cout << "brick compare: " << bad << " cells differ beyond " << tol << std::endl;
  return bad == 0;
}
template <unsigned dims, typename T>
inline bool compareBrick(const std::vector<long> &dimlist, bElem *arr, unsigned *grid_ptr, T &brick) {
  const std::vector<long> zero(dimlist.size(), 0);
  return compareBrick<dims>(dimlist, zero, zero, arr, grid_ptr, brick);
}

/// plain array vs plain array over the interior (reference: compareArray, src/multiarray.cpp:51-63)
inline bool compareArray(const std::vector<long> &extent, const bElem *a, const bElem *b, double tol = BRICK_TOLERANCE) {
  long n = 1;
  for (long e : extent) n *= e;
  long bad = 0;
#pragma omp parallel for reduction(+ : bad)
  for (long i = 0; i < n; ++i) bad += !brick_detail::close_enough(a[i], b[i], tol);
  return bad == 0;
}

#endif  // BRICKCOMPARE_H
