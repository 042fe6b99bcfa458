// bk_layout.cu -- array <-> brick conversion and on-device comparison.
// Device versions of copyToBrick / copyFromBrick (include/bricksetup.h:139-221: iter_grid walks bricks
// [ghost/8, (dim+ghost)/8) per axis, array origin = padding + brick*8) and compareBrick (include/brickcompare.h:30-57).
// One CTA per brick, 256 threads, two cells per thread; brick side is fully coalesced (4 KiB contiguous), array side
// moves 64-byte rows.
#include "bk_common.h"

namespace {

struct LayoutArgs {
  long sA1, sA2;        // array strides (elements) of j and k
  long pad[3];          // array origin offset per axis
  int b0[3], nb[3];     // first brick and brick count per axis
  long sB1, sB2;        // brick-grid strides
};

template <int MODE>  // 0 array->brick, 1 brick->array, 2 compare
__global__ void __launch_bounds__(256) k_layout(LayoutArgs a, const double *__restrict__ arr_in, double *arr_out,
                                               const unsigned *__restrict__ grid, const double *__restrict__ dat_in,
                                               double *dat_out, size_t step, double tol,
                                               unsigned long long *mismatch, unsigned long long *maxrel_bits) {
  const int bi = a.b0[0] + blockIdx.x, bj = a.b0[1] + blockIdx.y, bk_ = a.b0[2] + blockIdx.z;
  const unsigned id = grid[bi + bj * a.sB1 + bk_ * a.sB2];
  const size_t bbase = (size_t) id * step;
  const long abase = (a.pad[0] + bi * 8L) + (a.pad[1] + bj * 8L) * a.sA1 + (a.pad[2] + bk_ * 8L) * a.sA2;
  unsigned long long bad = 0;
  double worst = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = threadIdx.x + h * 256;
    const int i = e & 7, j = (e >> 3) & 7, k = e >> 6;
    const long ap = abase + i + j * a.sA1 + k * a.sA2;
    if (MODE == 0) {
      dat_out[bbase + e] = arr_in[ap];
    } else if (MODE == 1) {
      arr_out[ap] = dat_in[bbase + e];
    } else {
      const double x = dat_in[bbase + e], y = arr_in[ap];
      const double diff = fabs(x - y), mag = fabs(x) + fabs(y);
      if (!(diff < tol || diff < mag * tol)) ++bad;   // brickcompare.h:36-37 (NaN counts as a mismatch)
      const double rel = mag > 0.0 ? diff / mag : 0.0;
      worst = fmax(worst, rel);
    }
  }
  if (MODE == 2) {
    if (bad) atomicAdd(mismatch, bad);
    if (worst > 0.0) atomicMax(maxrel_bits, (unsigned long long) __double_as_longlong(worst));  // positive doubles order as ints
  }
}

int make_args(const long *dl, const long *pad, const long *gz, LayoutArgs *a, dim3 *g) {
  for (int d = 0; d < 3; ++d) {
    if (dl[d] <= 0 || dl[d] % 8 || gz[d] % 8 || pad[d] < 0 || gz[d] < 0) return BK_EINVAL;
    a->pad[d] = pad[d];
    a->b0[d] = (int) (gz[d] / 8);
    a->nb[d] = (int) (dl[d] / 8);
  }
  const long e0 = dl[0] + 2 * (pad[0] + gz[0]), e1 = dl[1] + 2 * (pad[1] + gz[1]);
  a->sA1 = e0;
  a->sA2 = e0 * e1;
  a->sB1 = (dl[0] + 2 * gz[0]) / 8;
  a->sB2 = a->sB1 * ((dl[1] + 2 * gz[1]) / 8);
  if (a->nb[1] > 65535 || a->nb[2] > 65535) return BK_EINVAL;
  *g = dim3(a->nb[0], a->nb[1], a->nb[2]);
  return BK_OK;
}

// ---- synthetic fields and storage comparison ------------------------------------------------------------------------
// The reference fills its arrays on the host with randomArray (src/multiarray.cpp:33-45: per-thread mt19937_64, U[0,1))
// and copies them into bricks.  Here the field is a counter-based hash of the GLOBAL periodic cell coordinate, written
// straight into the bricks on the device: any rank (and the CPU checker) can evaluate any cell of the global field
// without communication, which is what lets bench.py check a multi-GPU run against the oracle on sampled boxes.
__host__ __device__ __forceinline__ double synthetic_value(unsigned long long seed, unsigned long long lin) {
  unsigned long long z = seed + (lin + 1ull) * 0x9E3779B97F4A7C15ull;  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (double) (z >> 11) * (1.0 / 9007199254740992.0);  // 53 bits -> [0,1)
}

struct FillArgs {
  unsigned gd[3];
  long org[3], glob[3];
  unsigned long long seed;
};

__global__ void __launch_bounds__(256) k_fill_synthetic(FillArgs a, const unsigned *__restrict__ grid, double *dat, size_t step) {
  const unsigned bi = blockIdx.x, bj = blockIdx.y, bk_ = blockIdx.z;
  const unsigned id = grid[bi + (bj + (size_t) bk_ * a.gd[1]) * a.gd[0]];
  if (id == 0) return;  // the null brick stays zero
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = threadIdx.x + h * 256;
    long c[3] = {a.org[0] + bi * 8L + (e & 7), a.org[1] + bj * 8L + ((e >> 3) & 7), a.org[2] + bk_ * 8L + (e >> 6)};
#pragma unroll
    for (int d = 0; d < 3; ++d) c[d] = ((c[d] % a.glob[d]) + a.glob[d]) % a.glob[d];
    const unsigned long long lin = ((unsigned long long) c[2] * a.glob[1] + c[1]) * a.glob[0] + c[0];
    dat[(size_t) id * step + e] = synthetic_value(a.seed, lin);
  }
}

__global__ void __launch_bounds__(256) k_compare_storage(const unsigned *__restrict__ grid, unsigned gx, unsigned gy, unsigned lx,
                                                         unsigned ly, unsigned lz, const double *__restrict__ A, size_t sa,
                                                         const double *__restrict__ B, size_t sb, double tol,
                                                         unsigned long long *mismatch, unsigned long long *maxrel_bits) {
  const unsigned id = grid[(lx + blockIdx.x) + ((ly + blockIdx.y) + (size_t) (lz + blockIdx.z) * gy) * gx];
  unsigned long long bad = 0;
  double worst = 0.0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = threadIdx.x + h * 256;
    const double x = A[(size_t) id * sa + e], y = B[(size_t) id * sb + e];
    const double diff = fabs(x - y), mag = fabs(x) + fabs(y);
    if (!(diff < tol || diff < mag * tol)) ++bad;  // brickcompare.h:36-37 (NaN counts as a mismatch)
    worst = fmax(worst, mag > 0.0 ? diff / mag : 0.0);
  }
  if (bad) atomicAdd(mismatch, bad);
  if (worst > 0.0) atomicMax(maxrel_bits, (unsigned long long) __double_as_longlong(worst));
}

}  // namespace

extern "C" {

int bk_copy_to_brick(const long *dl, const long *pad, const long *gz, const double *arr, const unsigned *grid,
                     double *dat, size_t step, void *stream) {
  BK_REQUIRE(dl && pad && gz && arr && grid && dat, "null argument");
  LayoutArgs a;
  dim3 g;
  BK_REQUIRE(make_args(dl, pad, gz, &a, &g) == BK_OK, "extents must be positive multiples of 8");
  k_layout<0><<<g, 256, 0, (cudaStream_t) stream>>>(a, arr, nullptr, grid, nullptr, dat, step, 0.0, nullptr, nullptr);
  BK_LAUNCHED();
  return BK_OK;
}

int bk_copy_from_brick(const long *dl, const long *pad, const long *gz, double *arr, const unsigned *grid,
                       const double *dat, size_t step, void *stream) {
  BK_REQUIRE(dl && pad && gz && arr && grid && dat, "null argument");
  LayoutArgs a;
  dim3 g;
  BK_REQUIRE(make_args(dl, pad, gz, &a, &g) == BK_OK, "extents must be positive multiples of 8");
  k_layout<1><<<g, 256, 0, (cudaStream_t) stream>>>(a, nullptr, arr, grid, dat, nullptr, step, 0.0, nullptr, nullptr);
  BK_LAUNCHED();
  return BK_OK;
}

int bk_compare_brick(const long *dl, const long *pad, const long *gz, const double *arr, const unsigned *grid,
                     const double *dat, size_t step, double tol, unsigned long long *mismatches, double *max_rel,
                     void *stream) {
  BK_REQUIRE(dl && pad && gz && arr && grid && dat && mismatches, "null argument");
  LayoutArgs a;
  dim3 g;
  BK_REQUIRE(make_args(dl, pad, gz, &a, &g) == BK_OK, "extents must be positive multiples of 8");
  unsigned long long *acc = nullptr;
  BK_CUDA(cudaMalloc(&acc, 2 * sizeof(unsigned long long)));
  BK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), (cudaStream_t) stream));
  k_layout<2><<<g, 256, 0, (cudaStream_t) stream>>>(a, arr, nullptr, grid, dat, nullptr, step, tol, acc, acc + 1);
  BK_LAUNCHED();
  unsigned long long host[2];
  BK_CUDA(cudaMemcpyAsync(host, acc, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  BK_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
  BK_CUDA(cudaFree(acc));
  *mismatches = host[0];
  if (max_rel) {
    double r;
    memcpy(&r, &host[1], sizeof(r));
    *max_rel = r;
  }
  return BK_OK;
}

int bk_fill_synthetic(const unsigned *grid, const unsigned *gdims, const long *origin, const long *global_cells,
                      uint64_t seed, double *dat, size_t step, void *stream) {
  BK_REQUIRE(grid && gdims && origin && global_cells && dat, "null argument");
  BK_REQUIRE(step >= 512 && gdims[0] && gdims[1] && gdims[2] && gdims[1] <= 65535 && gdims[2] <= 65535, "bad grid extents");
  FillArgs a;
  for (int d = 0; d < 3; ++d) {
    BK_REQUIRE(global_cells[d] > 0, "global extents must be positive");
    a.gd[d] = gdims[d], a.org[d] = origin[d], a.glob[d] = global_cells[d];
  }
  a.seed = seed;
  k_fill_synthetic<<<dim3(gdims[0], gdims[1], gdims[2]), 256, 0, (cudaStream_t) stream>>>(a, grid, dat, step);
  BK_LAUNCHED();
  return BK_OK;
}

double bk_synthetic_value(uint64_t seed, uint64_t linear_cell) { return synthetic_value(seed, linear_cell); }

int bk_compare_storage(const unsigned *grid, const unsigned *gdims, const unsigned *lo, const unsigned *hi, const double *a,
                       size_t a_step, const double *b, size_t b_step, double tol, unsigned long long *mismatches,
                       double *max_rel, void *stream) {
  BK_REQUIRE(grid && gdims && lo && hi && a && b && mismatches, "null argument");
  for (int d = 0; d < 3; ++d) BK_REQUIRE(lo[d] < hi[d] && hi[d] <= gdims[d], "brick box outside the grid");
  BK_REQUIRE(hi[1] - lo[1] <= 65535 && hi[2] - lo[2] <= 65535, "box too large");
  unsigned long long *acc = nullptr;
  BK_CUDA(cudaMalloc(&acc, 2 * sizeof(unsigned long long)));
  BK_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), (cudaStream_t) stream));
  k_compare_storage<<<dim3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]), 256, 0, (cudaStream_t) stream>>>(
      grid, gdims[0], gdims[1], lo[0], lo[1], lo[2], a, a_step, b, b_step, tol, acc, acc + 1);
  BK_LAUNCHED();
  unsigned long long host[2];
  BK_CUDA(cudaMemcpyAsync(host, acc, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  BK_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
  BK_CUDA(cudaFree(acc));
  *mismatches = host[0];
  if (max_rel) {
    double r;
    memcpy(&r, &host[1], sizeof(r));
    *max_rel = r;
  }
  return BK_OK;
}

}  // extern "C"
