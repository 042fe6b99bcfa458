// drivers/common.h -- what the three drivers share: problem constants, stencil table, seeded inputs, a plain CPU
// array sweep used ONLY to validate GPU results (the reference drivers do the same: arr_func vs brick_func), timing.
// Constants and coefficient values: stencils/stencils.h:11-28 and stencils/fake.h:11-33 of the reference.
#pragma once
#include <omp.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <random>
#include <string>
#include <vector>
#include "brick-mpi.h"
#include "brickcompare.h"

#define TILE 8
#define GZ TILE
#define PADDING 8
#define BDIM TILE, TILE, TILE
#define VFOLD 4, 8 /* stencils/cudavfold.h:9 */

typedef Brick<Dim<BDIM>, Dim<VFOLD>> Brick3D;

struct StencilDef {
  const char *name;   // CLI name
  const char *script; // the reference's stencil expression
  int id;             // BK_ST_*
  int radius, st_iter, points;
};
static const StencilDef kStencils[] = {
    {"7pt", "stencils/7pt.py", BK_ST_7PT, 1, 8, 7},
    {"mpi7pt", "stencils/mpi7pt.py", BK_ST_MPI7PT, 1, 8, 7},
    {"mpi13pt", "stencils/mpi13pt.py", BK_ST_MPI13PT, 2, 4, 13},
    {"mpi25pt", "stencils/mpi25pt.py", BK_ST_MPI25PT, 4, 2, 25},
    {"mpi125pt", "stencils/mpi125pt.py", BK_ST_MPI125PT, 2, 4, 125},
};
inline const StencilDef *find_stencil(const std::string &n) {
  for (auto &s : kStencils)
    if (n == s.name) return &s;
  for (auto &s : kStencils)  // accept "13pt" for "mpi13pt"
    if ("mpi" + n == s.name) return &s;
  return nullptr;
}

/// tap list (dk,dj,di,coefficient) of a stencil, straight from the .py expressions
struct Tap {
  int dk, dj, di;
  double c;
};
inline std::vector<Tap> stencil_taps(int id, const double *coeff) {
  std::vector<Tap> t;
  static const double A[5] = {0.1, 0.06, 0.045, 0.03, 0.015}, B[3] = {0.4, 0.07, 0.03};
  static const double C[10] = {0.1, 0.04, 0.03, 0.01, 0.006, 0.004, 0.005, 0.002, 0.003, 0.001};
  auto star = [&](int r, const double *w) {
    t.push_back({0, 0, 0, w[0]});
    for (int d = 1; d <= r; ++d)
      for (int s = -1; s <= 1; s += 2) {
        t.push_back({0, 0, s * d, w[d]});
        t.push_back({0, s * d, 0, w[d]});
        t.push_back({s * d, 0, 0, w[d]});
      }
  };
  switch (id) {
    case BK_ST_7PT:  // coeff[0] centre, [1] i+1, [2] i-1, [3] j+1, [4] j-1, [5] k+1, [6] k-1 (stencils/7pt.py)
      t = {{0, 0, 0, coeff[0]}, {0, 0, 1, coeff[1]}, {0, 0, -1, coeff[2]}, {0, 1, 0, coeff[3]},
           {0, -1, 0, coeff[4]}, {1, 0, 0, coeff[5]}, {-1, 0, 0, coeff[6]}};
      break;
    case BK_ST_MPI7PT: {
      const double w[2] = {0.4, 0.1};
      star(1, w);
      break;
    }
    case BK_ST_MPI13PT: star(2, B); break;
    case BK_ST_MPI25PT: star(4, A); break;
    case BK_ST_MPI125PT:  // coefficient class = sorted (|di|,|dj|,|dk|) (stencils/mpi125pt.py:13-32)
      for (int dk = -2; dk <= 2; ++dk)
        for (int dj = -2; dj <= 2; ++dj)
          for (int di = -2; di <= 2; ++di) {
            int v[3] = {std::abs(di), std::abs(dj), std::abs(dk)};
            std::sort(v, v + 3);
            static const int cls[3][3][3] = {{{0, 1, 2}, {-1, 3, 4}, {-1, -1, 5}},
                                             {{-1, -1, -1}, {-1, 6, 7}, {-1, -1, 8}},
                                             {{-1, -1, -1}, {-1, -1, -1}, {-1, -1, 9}}};
            t.push_back({dk, dj, di, C[cls[v[0]][v[1]][v[2]]]});
          }
      break;
  }
  return t;
}

/// VALIDATION ONLY: out = stencil(in) on a plain array with strides {1, sx, sx*sy}, over [lo,hi) per axis (i first)
inline void cpu_array_sweep(const std::vector<Tap> &taps, const bElem *in, bElem *out, const std::vector<long> &stride,
                            const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long k = lo[2]; k < hi[2]; ++k)
    for (long j = lo[1]; j < hi[1]; ++j)
      for (long i = lo[0]; i < hi[0]; ++i) {
        const long p = i + j * stride[1] + k * stride[2];
        double acc = 0;
        for (const Tap &t : taps) acc += t.c * in[p + t.di + t.dj * stride[1] + t.dk * stride[2]];
        out[p] = acc;
      }
}

/// VALIDATION ONLY: stencils/cond.py on a plain array -- every value read clamped at zero, the sum returned as |sum|
/// (3axis.cu:229-240 d3cond_arr)
inline void cpu_array_sweep_cond(const std::vector<Tap> &taps, const bElem *in, bElem *out, const std::vector<long> &stride,
                                 const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long k = lo[2]; k < hi[2]; ++k)
    for (long j = lo[1]; j < hi[1]; ++j)
      for (long i = lo[0]; i < hi[0]; ++i) {
        const long p = i + j * stride[1] + k * stride[2];
        double acc = 0;
        for (const Tap &t : taps) acc += t.c * std::max(in[p + t.di + t.dj * stride[1] + t.dk * stride[2]], 0.0);
        out[p] = acc > 0 ? acc : -acc;
      }
}

/// U[0,1) doubles, reproducible (the reference seeds from std::random_device, src/multiarray.cpp:13-23)
inline bElem *randomArray(const std::vector<long> &extent, uint64_t seed = 0x5EED) {
  size_t n = 1;
  for (long e : extent) n *= (size_t) e;
  bElem *a = (bElem *) aligned_alloc(ALIGN, (n * sizeof(bElem) + ALIGN - 1) / ALIGN * ALIGN);
#pragma omp parallel
  {
    std::mt19937_64 rng(seed + 7919ull * omp_get_thread_num());
    std::uniform_real_distribution<bElem> d(0, 1);
#pragma omp for schedule(static)
    for (size_t i = 0; i < n; ++i) a[i] = d(rng);
  }
  return a;
}
inline bElem *zeroArray(const std::vector<long> &extent) {
  size_t n = 1;
  for (long e : extent) n *= (size_t) e;
  bElem *a = (bElem *) aligned_alloc(ALIGN, (n * sizeof(bElem) + ALIGN - 1) / ALIGN * ALIGN);
  std::memset(a, 0, n * sizeof(bElem));
  return a;
}

/// time_func of stencils/stencils.h:40-53 with a configurable budget (seconds)
template <typename T>
double time_func(T func, double budget = 1.0) {
  int it = 1;
  func();
  double st = omp_get_wtime(), ed = st;
  while (ed < st + budget) {
    for (int i = 0; i < it; ++i) func();
    it <<= 1;
    ed = omp_get_wtime();
  }
  return (ed - st) / (it - 1);
}

/// VALIDATION ONLY: advance a periodic global array (extents G, i first) `steps` time steps on the host, in place
inline void cpu_periodic_steps(bElem *field, const long *G, const std::vector<Tap> &taps, int R, int steps) {
  const long P[3] = {G[0] + 2 * R, G[1] + 2 * R, G[2] + 2 * R};
  const std::vector<long> ps = {1, P[0], P[0] * P[1]};
  bElem *a = zeroArray({P[0], P[1], P[2]}), *b = zeroArray({P[0], P[1], P[2]});
  for (int s = 0; s < steps; ++s) {
#pragma omp parallel for collapse(2)
    for (long k = 0; k < P[2]; ++k)  // periodic padding
      for (long j = 0; j < P[1]; ++j)
        for (long i = 0; i < P[0]; ++i)
          a[i + j * ps[1] + k * ps[2]] =
              field[(i - R + G[0]) % G[0] + ((j - R + G[1]) % G[1]) * G[0] + ((k - R + G[2]) % G[2]) * G[0] * G[1]];
    const long lo[3] = {R, R, R}, hi[3] = {R + G[0], R + G[1], R + G[2]};
    cpu_array_sweep(taps, a, b, ps, lo, hi);
#pragma omp parallel for collapse(2)
    for (long k = 0; k < G[2]; ++k)
      for (long j = 0; j < G[1]; ++j)
        for (long i = 0; i < G[0]; ++i) field[i + j * G[0] + k * G[0] * G[1]] = b[i + R + (j + R) * ps[1] + (k + R) * ps[2]];
  }
  free(a);
  free(b);
}

/// a host barrier for the rank threads (C++17 has no std::barrier)
#include <condition_variable>
#include <mutex>
struct Barrier {
  std::mutex m;
  std::condition_variable cv;
  int n, waiting = 0, gen = 0;
  explicit Barrier(int n) : n(n) {}
  void wait() {
    std::unique_lock<std::mutex> l(m);
    const int g = gen;
    if (++waiting == n) {
      waiting = 0, ++gen;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return g != gen; });
    }
  }
};
