// bk_array.cu -- the array-layout baseline: the same stencils over a plain padded array.
//
// The reference times every stencil twice, on a conventional array ("Arr:" lines; arr_kernel, weak/main.cu:27-33 with
// ST_GPU = the scalar forms of stencils/fake.h:44-352; d3pt7_arr, stencils/3axis.cu) and on bricks ("Bri:" lines), so
// that a report can put the two layouts side by side.  The reference kernel is one thread per point in 8x8x8 blocks.
// Here a CTA owns a 64 x 4 (i x j) column and marches KT planes along k: the k taps come from a register window that
// slides one plane per step (one global load per point for the whole k direction), the in-plane taps are read through
// L1 (neighbouring threads of the same CTA fetch the same lines).  It is a baseline, not the product: no shared-memory
// staging, no bulk copies -- the marching brick kernels in bk_stencil_tiled.cu are what the array layout is compared to.
//
// Summation order per point: centre, then by distance d = 1..R: +i, -i, +j, -j, +k, -k (star); dz, dy, dx ascending
// (cube).  Parity with the oracle is to 1e-12 relative, like every other kernel.
#include "bk_common.h"

namespace {

using bk::CubeCoef;
using bk::StarCoef;

struct ArrArgs {
  const double *in;
  double *out;
  long sy, sz;  // elements between consecutive j rows / k planes
  long lo[3];
  int n[3];
};

constexpr int TX = 64, TY = 4, KT = 16;

template <int R>
__global__ void __launch_bounds__(TX *TY) k_array_star(ArrArgs a, StarCoef cf) {
  const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
  const int k0 = blockIdx.z * KT;
  if (i >= a.n[0] || j >= a.n[1]) return;
  const long col = (a.lo[0] + i) + (a.lo[1] + j) * a.sy;
  const double *p = a.in + col + (a.lo[2] + k0) * a.sz;
  double *q = a.out + col + (a.lo[2] + k0) * a.sz;
  double w[2 * R + 1];  // the column's values at planes k-R .. k+R
#pragma unroll
  for (int d = 0; d < 2 * R; ++d) w[d] = __ldg(p + (d - R) * a.sz);
  const int kn = min(KT, a.n[2] - k0);
  for (int k = 0; k < kn; ++k) {
    w[2 * R] = __ldg(p + R * a.sz);
    double acc = cf.c0 * w[R];
#pragma unroll
    for (int d = 1; d <= R; ++d) {
      acc = fma(cf.cp[0][d - 1], __ldg(p + d), acc);
      acc = fma(cf.cm[0][d - 1], __ldg(p - d), acc);
      acc = fma(cf.cp[1][d - 1], __ldg(p + d * a.sy), acc);
      acc = fma(cf.cm[1][d - 1], __ldg(p - d * a.sy), acc);
      acc = fma(cf.cp[2][d - 1], w[R + d], acc);
      acc = fma(cf.cm[2][d - 1], w[R - d], acc);
    }
    *q = acc;
#pragma unroll
    for (int d = 0; d < 2 * R; ++d) w[d] = w[d + 1];
    p += a.sz, q += a.sz;
  }
}

__global__ void __launch_bounds__(TX *TY) k_array_cube(ArrArgs a, CubeCoef cf) {
  constexpr int R = 2;
  const int i = blockIdx.x * TX + threadIdx.x, j = blockIdx.y * TY + threadIdx.y;
  const int k0 = blockIdx.z * KT;
  if (i >= a.n[0] || j >= a.n[1]) return;
  const long col = (a.lo[0] + i) + (a.lo[1] + j) * a.sy;
  const int kn = min(KT, a.n[2] - k0);
  for (int k = 0; k < kn; ++k) {
    const double *p = a.in + col + (a.lo[2] + k0 + k) * a.sz;
    double acc = 0.0;
#pragma unroll
    for (int dz = -R; dz <= R; ++dz)
#pragma unroll
      for (int dy = -R; dy <= R; ++dy)
#pragma unroll
        for (int dx = -R; dx <= R; ++dx)
          acc = fma(cf.cc[dz < 0 ? -dz : dz][dy < 0 ? -dy : dy][dx < 0 ? -dx : dx], __ldg(p + dx + dy * a.sy + dz * a.sz), acc);
    a.out[col + (a.lo[2] + k0 + k) * a.sz] = acc;
  }
}

}  // namespace

extern "C" int bk_array_stencil_apply(int stencil, const double *in, double *out, const long *extent, const long *lo,
                                      const long *hi, const double *coeff, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(in && out && extent && lo && hi, "null argument");
  BK_REQUIRE(in != out, "in-place sweep is not defined");
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  const int R = bk_stencil_radius(stencil);  // of the stencil (the kernel radius of CoefSpec rounds 3 up to 4)
  ArrArgs a;
  a.in = in, a.out = out, a.sy = extent[0], a.sz = extent[0] * extent[1];
  for (int d = 0; d < 3; ++d) {
    BK_REQUIRE(lo[d] >= R && hi[d] >= lo[d] && hi[d] + R <= extent[d], "the box plus the stencil radius must lie inside the array");
    BK_REQUIRE(hi[d] - lo[d] < (1l << 31), "box too large");
    a.lo[d] = lo[d], a.n[d] = (int) (hi[d] - lo[d]);
  }
  if (a.n[0] == 0 || a.n[1] == 0 || a.n[2] == 0) return BK_OK;
  const dim3 grid((a.n[0] + TX - 1) / TX, (a.n[1] + TY - 1) / TY, (a.n[2] + KT - 1) / KT), block(TX, TY, 1);
  BK_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "box too large");
  cudaStream_t s = (cudaStream_t) stream;
  if (spec.kind == 1) k_array_cube<<<grid, block, 0, s>>>(a, spec.cc);
  else if (R == 1) k_array_star<1><<<grid, block, 0, s>>>(a, spec.sc);
  else if (R == 2) k_array_star<2><<<grid, block, 0, s>>>(a, spec.sc);
  else k_array_star<4><<<grid, block, 0, s>>>(a, spec.sc);
  BK_LAUNCHED();
  return BK_OK;
}
