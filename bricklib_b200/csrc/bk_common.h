// bk_common.h -- internals shared by the translation units of libbrick_b200.so
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "bricklib_b200.h"

namespace bk {

void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  set_error("%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
  return BK_ECUDA;
}

#define BK_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return bk::cuda_fail(_e, #call, __FILE__, __LINE__);    \
  } while (0)

#define BK_LAUNCHED()                                                              \
  do {                                                                             \
    bk::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) return bk::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define BK_REQUIRE(cond, msg)                                                      \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      bk::set_error("%s: %s", __func__, msg);                                      \
      return BK_EINVAL;                                                            \
    }                                                                              \
  } while (0)

constexpr int BRICK_EDGE = 8;
constexpr int BRICK_ELEMS = 512;

// Coefficients in the form the kernels consume.
//   star:  c0 * in(0) + sum_{axis a, d=1..R} ( cp[a][d-1] * in(+d along a) + cm[a][d-1] * in(-d along a) )
//   cube:  sum_{dz,dy,dx in -2..2} cc[|dz|][|dy|][|dx|] * in(dx,dy,dz)
struct StarCoef {
  double c0;
  double cp[3][4];
  double cm[3][4];
};
struct CubeCoef {
  double cc[3][3][3];
};

int star_coef_for(int stencil, const double *coeff_host, StarCoef *out);   // BK_ST_7PT..BK_ST_MPI25PT
int cube_coef_for(int stencil, CubeCoef *out);                            // BK_ST_MPI125PT

// What a kernel launch needs to know about a stencil: built from a BK_ST_* id (coef_spec_for) or lowered from a tap list
// (bk_stencil_compile).  radius is the KERNEL radius (1, 2 or 4 for stars -- a radius-3 star runs on the radius-4
// kernel with zero outer coefficients; 2 for the cube).
struct CoefSpec {
  int kind = 0;    // 0 star, 1 sign- and permutation-symmetric cube
  int radius = 0;
  int fused_ok = 0;  // a two-steps-per-pass kernel exists (radius 1 and 2 stars)
  StarCoef sc;
  CubeCoef cc;
};
int coef_spec_for(int stencil, const double *coeff_host, CoefSpec *out);

// launch geometry of a generated marching kernel (bk_codegen.cu -> bk_stencil_tiled.cu: launch_generated)
struct GenGeom {
  int TI, TJ, ovh, threads;
  size_t smem;
};

}  // namespace bk
