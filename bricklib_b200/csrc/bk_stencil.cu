// bk_stencil.cu -- stencil entry points of the C ABI + the per-brick kernel family.
//
// What the reference generates per stencil with codegen/vecscatter (CUDA backend: one warp per brick, 8-byte loads,
// dev_shl shuffles) is replaced by two hand-written kernel families:
//   * k_brick  (this file)       one CTA per brick, neighbours resolved through the adjacency list exactly like the
//                                accessor does (include/brick.h:234-246).  Works for ANY brick set (id lists, multi-
//                                subdomain launches, irregular adjacency) -- the general path.
//   * k_tiled  (bk_stencil_tiled.cu)  a CTA marches a column of bricks of a dense grid box with an async smem pipeline;
//                                the fast path for box-shaped launches (what every reference driver issues).
// Stencil definitions: stencils/{7pt,mpi7pt,mpi13pt,mpi25pt,mpi125pt}.py; constants stencils/fake.h:11-33.
// Summation order (fixed, documented in DESIGN.md): centre first, then by distance d = 1..R: +i,-i,+j,-j,+k,-k (star);
// dz,dy,dx ascending (cube).  The reference's order is codegen-dependent, parity is to 1e-12 relative, not bitwise.
#include "bk_common.h"
#include "bk_codegen.h"
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace bk {

// stencils/fake.h:11-33
static const double kAlpha = 0.4, kBeta = 0.1;
static const double kA[5] = {0.1, 0.06, 0.045, 0.03, 0.015};
static const double kB[3] = {0.4, 0.07, 0.03};
static const double kC[10] = {0.1, 0.04, 0.03, 0.01, 0.006, 0.004, 0.005, 0.002, 0.003, 0.001};

int star_coef_for(int stencil, const double *coeff, StarCoef *o) {
  *o = StarCoef();
  switch (stencil) {
    case BK_ST_7PT:  // stencils/7pt.py: coeff[0] centre, [1] i+1, [2] i-1, [3] j+1, [4] j-1, [5] k+1, [6] k-1
      if (!coeff) {
        set_error("BK_ST_7PT needs coeff[0..6]");
        return BK_EINVAL;
      }
      o->c0 = coeff[0];
      for (int a = 0; a < 3; ++a) {
        o->cp[a][0] = coeff[1 + 2 * a];
        o->cm[a][0] = coeff[2 + 2 * a];
      }
      return 1;
    case BK_ST_MPI7PT:
      o->c0 = kAlpha;
      for (int a = 0; a < 3; ++a) o->cp[a][0] = o->cm[a][0] = kBeta;
      return 1;
    case BK_ST_MPI13PT:
      o->c0 = kB[0];
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 2; ++d) o->cp[a][d] = o->cm[a][d] = kB[d + 1];
      return 2;
    case BK_ST_MPI25PT:
      o->c0 = kA[0];
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 4; ++d) o->cp[a][d] = o->cm[a][d] = kA[d + 1];
      return 4;
  }
  set_error("not a star stencil: %d", stencil);
  return BK_EINVAL;
}

int cube_coef_for(int stencil, CubeCoef *o) {
  if (stencil != BK_ST_MPI125PT) {
    set_error("not a cube stencil: %d", stencil);
    return BK_EINVAL;
  }
  // stencils/mpi125pt.py:13-32: the coefficient depends on the sorted (|dx|,|dy|,|dz|) triple, 10 classes
  for (int z = 0; z < 3; ++z)
    for (int y = 0; y < 3; ++y)
      for (int x = 0; x < 3; ++x) {
        int v[3] = {x, y, z};
        for (int p = 0; p < 2; ++p)
          for (int q = 0; q < 2 - p; ++q)
            if (v[q] > v[q + 1]) {
              int t = v[q];
              v[q] = v[q + 1];
              v[q + 1] = t;
            }
        static const int cls[3][3][3] = {  // [lo][mid][hi] -> class index
            {{0, 1, 2}, {-1, 3, 4}, {-1, -1, 5}}, {{-1, -1, -1}, {-1, 6, 7}, {-1, -1, 8}}, {{-1, -1, -1}, {-1, -1, -1}, {-1, -1, 9}}};
        o->cc[z][y][x] = kC[cls[v[0]][v[1]][v[2]]];
      }
  return 2;
}

int coef_spec_for(int stencil, const double *coeff, CoefSpec *o) {
  *o = CoefSpec();
  if (stencil == BK_ST_MPI125PT) {
    o->kind = 1;
    o->radius = cube_coef_for(stencil, &o->cc);
    return o->radius < 0 ? BK_EINVAL : BK_OK;
  }
  o->radius = star_coef_for(stencil, coeff, &o->sc);
  if (o->radius < 0) return BK_EINVAL;
  o->fused_ok = o->radius <= 2;
  return BK_OK;
}

// implemented in bk_stencil_tiled.cu
int launch_tiled(const CoefSpec &spec, const bk_field_t &f, const bk_field_t *multi_dev, unsigned nsub, const unsigned *grid,
                 const unsigned *gdims, const unsigned *lo, const unsigned *hi, cudaStream_t s, int part = BK_PART_ALL,
                 const unsigned *ready_lo = nullptr, const unsigned *ready_hi = nullptr, int steps = 1);
// implemented in bk_stencil_remote.cu
int launch_tiled_remote(const CoefSpec &spec, const bk_field_t &f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
                        const unsigned *hi, cudaStream_t s, int part, const unsigned *ready_lo, const unsigned *ready_hi, int steps,
                        const double *const *remap, unsigned ghost_lo, unsigned ghost_n);

}  // namespace bk

namespace {

using bk::CubeCoef;
using bk::StarCoef;

struct Select {
  const unsigned *grid;     // dense id array or nullptr
  const unsigned *ids;      // explicit list or nullptr
  const bk_field_t *multi;  // per-subdomain fields (strong driver) or nullptr
  unsigned gsx, gsy;        // grid strides in bricks
  unsigned lo[3];
  unsigned nx;              // bricks per subdomain along i (multi only)
};

template <int R, bool CUBE>
struct CoefOf {
  using type = StarCoef;
};
template <int R>
struct CoefOf<R, true> {
  using type = CubeCoef;
};

template <int R, bool CUBE>
__global__ void __launch_bounds__(256) k_brick(Select sel, bk_field_t f, typename CoefOf<R, CUBE>::type cf) {
  constexpr int W = 8 + 2 * R;
  __shared__ double box[W * W * W];
  __shared__ unsigned nb[27];

  unsigned b;
  if (sel.ids) {
    b = sel.ids[blockIdx.x];
  } else {
    unsigned bx = blockIdx.x;
    if (sel.multi) {  // strong/main.cu:85-99: blockIdx.x = subdomain * strideb + bi
      f = sel.multi[bx / sel.nx];
      bx %= sel.nx;
    }
    b = sel.grid[(sel.lo[0] + bx) + ((sel.lo[1] + blockIdx.y) + (size_t) (sel.lo[2] + blockIdx.z) * sel.gsy) * sel.gsx];
  }
  if (threadIdx.x < 27) nb[threadIdx.x] = f.adj[(size_t) b * 27 + threadIdx.x];
  __syncthreads();

  // gather the (8+2R)^3 neighbourhood; a star stencil never touches edge/corner neighbours
  for (int idx = threadIdx.x; idx < W * W * W; idx += 256) {
    const int x = idx % W, y = (idx / W) % W, z = idx / (W * W);
    const int gx = x + 8 - R, gy = y + 8 - R, gz = z + 8 - R;  // 0..23 across the 3 bricks of an axis
    const int ox = gx >> 3, oy = gy >> 3, oz = gz >> 3;
    if (!CUBE && ((ox != 1) + (oy != 1) + (oz != 1) > 1)) continue;
    box[idx] = f.in[(size_t) nb[oz * 9 + oy * 3 + ox] * f.in_step + ((gz & 7) << 6) + ((gy & 7) << 3) + (gx & 7)];
  }
  __syncthreads();

#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = threadIdx.x + h * 256;
    const int i = e & 7, j = (e >> 3) & 7, k = e >> 6;
    const double *c = &box[((k + R) * W + (j + R)) * W + (i + R)];
    double acc;
    if constexpr (!CUBE) {
      acc = cf.c0 * c[0];
#pragma unroll
      for (int d = 1; d <= R; ++d) {
        acc = fma(cf.cp[0][d - 1], c[d], acc);
        acc = fma(cf.cm[0][d - 1], c[-d], acc);
        acc = fma(cf.cp[1][d - 1], c[d * W], acc);
        acc = fma(cf.cm[1][d - 1], c[-d * W], acc);
        acc = fma(cf.cp[2][d - 1], c[d * W * W], acc);
        acc = fma(cf.cm[2][d - 1], c[-d * W * W], acc);
      }
    } else {
      acc = 0.0;
#pragma unroll
      for (int dz = -R; dz <= R; ++dz)
#pragma unroll
        for (int dy = -R; dy <= R; ++dy)
#pragma unroll
          for (int dx = -R; dx <= R; ++dx)
            acc = fma(cf.cc[dz < 0 ? -dz : dz][dy < 0 ? -dy : dy][dx < 0 ? -dx : dx], c[(dz * W + dy) * W + dx], acc);
    }
    f.out[(size_t) b * f.out_step + e] = acc;
  }
}

int launch_brick(const bk::CoefSpec &spec, const Select &sel, const bk_field_t &f, dim3 grid, cudaStream_t s) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return BK_OK;
  if (spec.kind == 1) {
    k_brick<2, true><<<grid, 256, 0, s>>>(sel, f, spec.cc);
  } else {
    if (spec.radius == 1) k_brick<1, false><<<grid, 256, 0, s>>>(sel, f, spec.sc);
    if (spec.radius == 2) k_brick<2, false><<<grid, 256, 0, s>>>(sel, f, spec.sc);
    if (spec.radius == 4) k_brick<4, false><<<grid, 256, 0, s>>>(sel, f, spec.sc);
  }
  BK_LAUNCHED();
  return BK_OK;
}

// ---- arbitrary linear stencils: one CTA per brick, taps from a device table -------------------------------------------
// The general lowering target of bk_stencil_compile (what codegen/vecscatter does for any stencils/*.py expression that
// is neither a star nor the symmetric cube): the (8+2R)^3 neighbourhood is gathered through the adjacency list like
// k_brick does, then every thread walks the tap table (offset into the box, coefficient) for its two cells.
struct TapDev {
  int off;  // ((dk * W) + dj) * W + di in the shared box
  double c;
};

__device__ __forceinline__ double pointwise(double x, int op, double c) {
  return op == BK_OP_MAX ? fmax(x, c) : op == BK_OP_MIN ? fmin(x, c) : op == BK_OP_ABS ? fabs(x) : x;
}

template <int R>
__global__ void __launch_bounds__(256) k_taps(Select sel, bk_field_t f, const TapDev *__restrict__ taps, int ntaps,
                                              bk_pointwise_t pre, bk_pointwise_t post) {
  constexpr int W = 8 + 2 * R;
  __shared__ double box[W * W * W];
  __shared__ unsigned nb[27];
  const unsigned b = sel.ids ? sel.ids[blockIdx.x]
                             : sel.grid[(sel.lo[0] + blockIdx.x) +
                                        ((sel.lo[1] + blockIdx.y) + (size_t) (sel.lo[2] + blockIdx.z) * sel.gsy) * sel.gsx];
  if (threadIdx.x < 27) nb[threadIdx.x] = f.adj[(size_t) b * 27 + threadIdx.x];
  __syncthreads();
  for (int idx = threadIdx.x; idx < W * W * W; idx += 256) {
    const int x = idx % W, y = (idx / W) % W, z = idx / (W * W);
    const int gx = x + 8 - R, gy = y + 8 - R, gz = z + 8 - R;
    box[idx] = pointwise(
        f.in[(size_t) nb[(gz >> 3) * 9 + (gy >> 3) * 3 + (gx >> 3)] * f.in_step + ((gz & 7) << 6) + ((gy & 7) << 3) + (gx & 7)],
        pre.op, pre.c);
  }
  __syncthreads();
  const int e0 = threadIdx.x, e1 = threadIdx.x + 256;
  const double *c0 = &box[(((e0 >> 6) + R) * W + (((e0 >> 3) & 7) + R)) * W + ((e0 & 7) + R)];
  const double *c1 = &box[(((e1 >> 6) + R) * W + (((e1 >> 3) & 7) + R)) * W + ((e1 & 7) + R)];
  double a0 = 0.0, a1 = 0.0;
  for (int t = 0; t < ntaps; ++t) {
    const TapDev tp = taps[t];
    a0 = fma(tp.c, c0[tp.off], a0);
    a1 = fma(tp.c, c1[tp.off], a1);
  }
  f.out[(size_t) b * f.out_step + e0] = pointwise(a0, post.op, post.c);
  f.out[(size_t) b * f.out_step + e1] = pointwise(a1, post.op, post.c);
}

// ---- is the dense grid a faithful picture of the adjacency list? -------------------------------------------------------
// The marching kernels (built-in and generated) take neighbour ids from the dense `grid` array and read brick 0 outside
// it; the reference's accessor -- and k_brick / k_taps here -- follow BrickInfo::adj (include/brick.h:234-246).  The two
// agree for every grid the drivers build (BrickDecomp, the interior of init_grid), but a caller may hand in ANY adjacency
// (periodic wrap, holes).  So before a (adj, grid, box) triple is first swept by a marching kernel, one small kernel
// checks  adj[grid[p]][s] == grid[p + delta_s]  (0 outside the grid) for every position p the launch's reads start from
// and every neighbour slot s the stencil uses; the verdict is cached.  On a mismatch BK_KERNEL_AUTO falls back to the
// adjacency-following family and BK_KERNEL_TILED reports BK_EUNSUPPORTED -- never silently different numbers.
__global__ void __launch_bounds__(256) k_check_adj(const unsigned *__restrict__ adj, const unsigned *__restrict__ grid, int gx, int gy,
                                                   int gz, int lx, int ly, int lz, int nx, int ny, int nz, unsigned slots,
                                                   unsigned *bad) {
  const long n = (long) nx * ny * nz * 27;
  for (long t = blockIdx.x * 256L + threadIdx.x; t < n; t += (long) gridDim.x * 256) {
    const int s = (int) (t % 27);
    if (!((slots >> s) & 1)) continue;
    long r = t / 27;
    const int i = lx + (int) (r % nx), j = ly + (int) ((r / nx) % ny), k = lz + (int) (r / ((long) nx * ny));
    const unsigned id = grid[((size_t) k * gy + j) * gx + i];
    if (id == 0) continue;  // the null brick is never an output; what is read THROUGH it is zero either way
    const int qi = i + s % 3 - 1, qj = j + (s / 3) % 3 - 1, qk = k + s / 9 - 1;
    const bool in = qi >= 0 && qi < gx && qj >= 0 && qj < gy && qk >= 0 && qk < gz;
    const unsigned want = in ? grid[((size_t) qk * gy + qj) * gx + qi] : 0u;
    if (adj[(size_t) id * 27 + s] != want) atomicAdd(bad, 1u);
  }
}

struct AdjKey {
  const void *a, *g;
  unsigned d[3], l[3], h[3], slots;
  bool operator<(const AdjKey &o) const { return memcmp(this, &o, sizeof(AdjKey)) < 0; }
};
std::mutex adj_mu;
std::map<AdjKey, int> adj_cache;

constexpr unsigned kSlotsStar = (1u << 13) | (1u << 12) | (1u << 14) | (1u << 10) | (1u << 16) | (1u << 4) | (1u << 22);
constexpr unsigned kSlotsAll = (1u << 27) - 1;

// 1 = grid and adjacency agree on everything a marching launch over [lo,hi) reads (expanded by `grow` bricks for the
// two-step kernel, whose first step runs on the neighbours too), 0 = they differ, negative = CUDA error
int marching_matches_adjacency(const unsigned *adj, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
                               const unsigned *hi, int grow, unsigned slots, cudaStream_t s) {
  if (getenv("BK_SKIP_ADJ_CHECK")) return 1;  // developer knob
  using Key = AdjKey;
  std::mutex &mu = adj_mu;
  std::map<AdjKey, int> &cache = adj_cache;
  Key key;
  memset(&key, 0, sizeof(key));
  key.a = adj, key.g = grid, key.slots = slots;
  for (int d = 0; d < 3; ++d) {
    key.d[d] = gdims[d];
    key.l[d] = lo[d] > (unsigned) grow ? lo[d] - grow : 0;
    key.h[d] = std::min(gdims[d], hi[d] + grow);
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  const int nx = (int) (key.h[0] - key.l[0]), ny = (int) (key.h[1] - key.l[1]), nz = (int) (key.h[2] - key.l[2]);
  int verdict = 1;
  if (nx > 0 && ny > 0 && nz > 0) {
    unsigned *bad = nullptr, host = 0;
    BK_CUDA(cudaMalloc(&bad, sizeof(unsigned)));
    BK_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned), s));
    const long n = (long) nx * ny * nz * 27;
    k_check_adj<<<(unsigned) std::min<long>((n + 255) / 256, 148 * 16), 256, 0, s>>>(adj, grid, (int) gdims[0], (int) gdims[1],
                                                                                  (int) gdims[2], (int) key.l[0], (int) key.l[1],
                                                                                  (int) key.l[2], nx, ny, nz, slots, bad);
    BK_LAUNCHED();
    BK_CUDA(cudaMemcpyAsync(&host, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    BK_CUDA(cudaStreamSynchronize(s));
    BK_CUDA(cudaFree(bad));
    verdict = host == 0;
  }
  std::lock_guard<std::mutex> lk(mu);
  cache[key] = verdict;
  return verdict;
}

int adjacency_mismatch() {
  bk::set_error("the dense grid does not match the adjacency list: use bk_stencil_apply (it follows the adjacency) for this brick set");
  return BK_EUNSUPPORTED;
}

int check_box(const unsigned *gdims, const unsigned *lo, const unsigned *hi) {
  for (int a = 0; a < 3; ++a)
    if (!(lo[a] <= hi[a] && hi[a] <= gdims[a])) return BK_EINVAL;
  if (hi[1] - lo[1] > 65535 || hi[2] - lo[2] > 65535) return BK_EINVAL;
  return BK_OK;
}

}  // namespace

extern "C" {

int bk_adjacency_forget(const void *adj_or_grid_dev) {
  std::lock_guard<std::mutex> lk(adj_mu);
  for (auto it = adj_cache.begin(); it != adj_cache.end();)
    it = (!adj_or_grid_dev || it->first.a == adj_or_grid_dev || it->first.g == adj_or_grid_dev) ? adj_cache.erase(it) : std::next(it);
  return BK_OK;
}

int bk_stencil_radius(int s) {
  static const int r[BK_ST_COUNT] = {1, 1, 2, 4, 2};
  return (s < 0 || s >= BK_ST_COUNT) ? BK_EINVAL : r[s];
}
int bk_stencil_st_iter(int s) {  // stencils/fake.h:39-344: ghost depth 8 cells / radius
  static const int it[BK_ST_COUNT] = {8, 8, 4, 2, 4};
  return (s < 0 || s >= BK_ST_COUNT) ? BK_EINVAL : it[s];
}
int bk_stencil_fused_steps(int s) {  // time steps per pass that pay off (the radius-2 fused kernel exists but is slower)
  static const int f[BK_ST_COUNT] = {2, 2, 1, 1, 1};
  return (s < 0 || s >= BK_ST_COUNT) ? BK_EINVAL : f[s];
}
int bk_stencil_points(int s) {
  static const int p[BK_ST_COUNT] = {7, 7, 13, 25, 125};
  return (s < 0 || s >= BK_ST_COUNT) ? BK_EINVAL : p[s];
}

static int apply_spec(const bk::CoefSpec &spec, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                      const unsigned *lo, const unsigned *hi, unsigned flags, cudaStream_t s) {
  if (flags != BK_KERNEL_BRICK) {
    const int ok = marching_matches_adjacency(f->adj, grid, gdims, lo, hi, 0, spec.kind == 1 ? kSlotsAll : kSlotsStar, s);
    if (ok < 0) return ok;
    if (ok) {
      int rc = bk::launch_tiled(spec, *f, nullptr, 1, grid, gdims, lo, hi, s);
      if (rc != BK_EUNSUPPORTED || flags == BK_KERNEL_TILED) return rc;
    } else if (flags == BK_KERNEL_TILED) {
      bk::set_error("the dense grid does not match the adjacency list: the marching kernels would read other neighbours");
      return BK_EUNSUPPORTED;
    }
  }
  Select sel = {grid, nullptr, nullptr, gdims[0], gdims[1], {lo[0], lo[1], lo[2]}, 0};
  return launch_brick(spec, sel, *f, dim3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]), s);
}

int bk_stencil_apply(int stencil, const bk_field_t *f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
                     const unsigned *hi, const double *coeff, unsigned flags, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi, "null argument");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  cudaStream_t s = (cudaStream_t) stream;
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  return apply_spec(spec, f, grid, gdims, lo, hi, flags, s);
}

int bk_stencil_apply_part(int stencil, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                          const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                          const unsigned *ready_hi, int part, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi && ready_lo && ready_hi, "null argument");
  const bool trust_grid = (part & BK_PART_GRID_TOPOLOGY) != 0;
  part &= ~BK_PART_GRID_TOPOLOGY;
  BK_REQUIRE((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST,
             "part must be BK_PART_READY or BK_PART_REST (optionally | BK_PART_THIN)");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  // only the marching kernel has the split enumeration; the caller falls back to whole-box launches on EUNSUPPORTED
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  const int ok = trust_grid ? 1 : marching_matches_adjacency(f->adj, grid, gdims, lo, hi, 0, spec.kind == 1 ? kSlotsAll : kSlotsStar, (cudaStream_t) stream);
  if (ok <= 0) return ok < 0 ? ok : adjacency_mismatch();
  return bk::launch_tiled(spec, *f, nullptr, 1, grid, gdims, lo, hi, (cudaStream_t) stream, part, ready_lo, ready_hi);
}

int bk_stencil_advance(int stencil, int steps, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                       const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                       const unsigned *ready_hi, int part, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(steps == 1 || steps == 2, "steps must be 1 or 2");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi, "null argument");
  const bool trust_grid = (part & BK_PART_GRID_TOPOLOGY) != 0;
  part &= ~BK_PART_GRID_TOPOLOGY;
  BK_REQUIRE(part == BK_PART_ALL ||
                 (ready_lo && ready_hi && ((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST)),
             "bad part");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  const int ok = trust_grid ? 1 : marching_matches_adjacency(f->adj, grid, gdims, lo, hi, steps - 1, spec.kind == 1 ? kSlotsAll : kSlotsStar,
                                            (cudaStream_t) stream);
  if (ok <= 0) return ok < 0 ? ok : adjacency_mismatch();
  return bk::launch_tiled(spec, *f, nullptr, 1, grid, gdims, lo, hi, (cudaStream_t) stream, part, ready_lo, ready_hi,
                          steps);
}

// bk_stencil_advance with the ghost bricks of the INPUT read in place from the neighbours' storages: brick id
// ghost_lo + g comes from remap_dev[g] (device array of ghost_n device-visible addresses, each one brick of 512 doubles)
// instead of f->in + id * in_step.  The exchange then happens inside the sweep: no pull, no ghost write and re-read.
int bk_stencil_advance_remote(int stencil, int steps, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                              const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                              const unsigned *ready_hi, int part, const double *const *remap_dev, unsigned ghost_lo,
                              unsigned ghost_n, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(steps == 1 || steps == 2, "steps must be 1 or 2");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi && remap_dev && ghost_n > 0, "null argument");
  const bool trust_grid = (part & BK_PART_GRID_TOPOLOGY) != 0;
  part &= ~BK_PART_GRID_TOPOLOGY;
  BK_REQUIRE(part == BK_PART_ALL ||
                 (ready_lo && ready_hi && ((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST)),
             "bad part");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  const int ok = trust_grid ? 1 : marching_matches_adjacency(f->adj, grid, gdims, lo, hi, steps - 1, spec.kind == 1 ? kSlotsAll : kSlotsStar,
                                                            (cudaStream_t) stream);
  if (ok <= 0) return ok < 0 ? ok : adjacency_mismatch();
  return bk::launch_tiled_remote(spec, *f, grid, gdims, lo, hi, (cudaStream_t) stream, part, ready_lo, ready_hi, steps, remap_dev,
                                 ghost_lo, ghost_n);
}

int bk_stencil_apply_list(int stencil, const bk_field_t *f, const unsigned *ids, size_t n, const double *coeff,
                          void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && (ids || n == 0), "null argument");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(n < (1ull << 31), "list too long");
  Select sel = {nullptr, ids, nullptr, 0, 0, {0, 0, 0}, 0};
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  return launch_brick(spec, sel, *f, dim3((unsigned) n, 1, 1), (cudaStream_t) stream);
}

int bk_stencil_apply_multi(int stencil, const bk_field_t *fields_dev, unsigned nsub, const unsigned *grid,
                           const unsigned *gdims, const unsigned *lo, const unsigned *hi, const double *coeff,
                           void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(fields_dev && grid && gdims && lo && hi && nsub > 0, "null argument");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  const unsigned nx = hi[0] - lo[0];
  BK_REQUIRE((unsigned long long) nx * nsub < (1ull << 31) && nsub <= 65535, "launch too wide");
  bk::CoefSpec spec;
  if (bk::coef_spec_for(stencil, coeff, &spec) != BK_OK) return BK_EINVAL;
  if (!getenv("BK_MULTI_BRICK")) {  // developer knob: force the per-brick family
    bk_field_t none = {nullptr, nullptr, 512, nullptr, 512};
    int rc = bk::launch_tiled(spec, none, fields_dev, nsub, grid, gdims, lo, hi, (cudaStream_t) stream);
    if (rc != BK_EUNSUPPORTED) return rc;
  }
  Select sel = {grid, nullptr, fields_dev, gdims[0], gdims[1], {lo[0], lo[1], lo[2]}, nx};
  bk_field_t dummy = {nullptr, nullptr, 512, nullptr, 512};
  return launch_brick(spec, sel, dummy, dim3(nx * nsub, hi[1] - lo[1], hi[2] - lo[2]), (cudaStream_t) stream);
}

// ---- stencils lowered from tap lists (the stencils/*.py expressions) -------------------------
}  // extern "C"

struct bk_stencil_def {
  int kind = BK_KIND_TAPS;  // BK_KIND_*
  int radius = 0;           // of the expression
  int ntaps = 0;
  bk::CoefSpec spec;        // kind STAR / CUBE
  int krad = 0;             // kernel radius (1, 2 or 4)
  std::vector<TapDev> taps_host;  // the tap table of k_taps, uploaded at its first launch (compiling needs no device)
  TapDev *taps_dev = nullptr;
  std::mutex mu;
  bk_pointwise_t pre = {BK_OP_NONE, 0.0}, post = {BK_OP_NONE, 0.0};
  bk::GenStencil *gen = nullptr;  // BK_KIND_GENERATED: the marching kernel generated for these taps (bk_codegen.cu)
  std::string gen_why;            // why there is none
};

extern "C" {

int bk_stencil_compile(bk_stencil_def_t **out, const bk_tap_t *taps, int ntaps) {
  return bk_stencil_compile_pointwise(out, taps, ntaps, nullptr, nullptr);
}

int bk_stencil_compile_pointwise(bk_stencil_def_t **out, const bk_tap_t *taps, int ntaps, const bk_pointwise_t *pre,
                                 const bk_pointwise_t *post) {
  BK_REQUIRE(out && taps && ntaps > 0 && ntaps <= 4096, "bad arguments");
  BK_REQUIRE((!pre || (pre->op >= BK_OP_NONE && pre->op <= BK_OP_ABS)) && (!post || (post->op >= BK_OP_NONE && post->op <= BK_OP_ABS)),
             "unknown pointwise op");
  const bool nonlinear = (pre && pre->op != BK_OP_NONE) || (post && post->op != BK_OP_NONE);
  // merge repeated offsets, drop zero coefficients
  std::map<std::array<int, 3>, double> m;
  for (int t = 0; t < ntaps; ++t) m[{taps[t].dk, taps[t].dj, taps[t].di}] += taps[t].c;
  int radius = 0;
  bool star = true;
  for (auto it = m.begin(); it != m.end();) {
    if (it->second == 0.0) {
      it = m.erase(it);
      continue;
    }
    int nz = 0;
    for (int a = 0; a < 3; ++a) radius = std::max(radius, std::abs(it->first[a])), nz += it->first[a] != 0;
    star = star && nz <= 1;
    ++it;
  }
  BK_REQUIRE(!m.empty(), "every coefficient is zero");
  if (radius > 4) {
    bk::set_error("bk_stencil_compile: radius %d (kernels cover radius <= 4 = half a brick)", radius);
    return BK_EUNSUPPORTED;
  }
  bk_stencil_def *d = new bk_stencil_def();
  d->radius = radius, d->ntaps = (int) m.size();
  d->krad = radius <= 1 ? 1 : radius <= 2 ? 2 : 4;
  auto coef = [&](int dk, int dj, int di) {
    auto it = m.find({dk, dj, di});
    return it == m.end() ? 0.0 : it->second;
  };
  if (pre) d->pre = *pre;
  if (post) d->post = *post;
  if (star && !nonlinear) {  // any coefficients: the marching kernels take one per tap
    d->kind = BK_KIND_STAR;
    d->spec.kind = 0, d->spec.radius = d->krad, d->spec.fused_ok = d->krad <= 2;
    d->spec.sc = bk::StarCoef();
    d->spec.sc.c0 = coef(0, 0, 0);
    for (int r = 1; r <= radius; ++r) {
      d->spec.sc.cp[0][r - 1] = coef(0, 0, r), d->spec.sc.cm[0][r - 1] = coef(0, 0, -r);
      d->spec.sc.cp[1][r - 1] = coef(0, r, 0), d->spec.sc.cm[1][r - 1] = coef(0, -r, 0);
      d->spec.sc.cp[2][r - 1] = coef(r, 0, 0), d->spec.sc.cm[2][r - 1] = coef(-r, 0, 0);
    }
    *out = d;
    return BK_OK;
  }
  bool sym = radius <= 2 && !nonlinear;  // c(dx,dy,dz) a function of the sorted (|dx|,|dy|,|dz|) only?
  if (sym) {
    for (int z = -2; z <= 2 && sym; ++z)
      for (int y = -2; y <= 2 && sym; ++y)
        for (int x = -2; x <= 2 && sym; ++x) {
          int v[3] = {std::abs(x), std::abs(y), std::abs(z)};
          std::sort(v, v + 3);
          sym = coef(z, y, x) == coef(v[2], v[1], v[0]);
        }
  }
  if (sym) {
    d->kind = BK_KIND_CUBE;
    d->spec.kind = 1, d->spec.radius = 2, d->spec.fused_ok = 0;
    for (int z = 0; z < 3; ++z)
      for (int y = 0; y < 3; ++y)
        for (int x = 0; x < 3; ++x) d->spec.cc.cc[z][y][x] = coef(z, y, x);
    *out = d;
    return BK_OK;
  }
  d->kind = BK_KIND_TAPS;
  if (!getenv("BK_NO_CODEGEN")) {  // developer knob: keep the tap-table kernel
    std::vector<bk::GenTap> gt;
    for (auto &kv : m) gt.push_back({kv.first[2], kv.first[1], kv.first[0], kv.second});
    d->gen = bk::gen_create(gt, d->pre, d->post, &d->gen_why);
    if (d->gen) d->kind = BK_KIND_GENERATED;
  }
  const int W = 8 + 2 * d->krad;
  for (auto &kv : m) d->taps_host.push_back({(kv.first[0] * W + kv.first[1]) * W + kv.first[2], kv.second});
  *out = d;
  return BK_OK;
}

int bk_stencil_def_destroy(bk_stencil_def_t *d) {
  if (!d) return BK_OK;
  if (d->taps_dev) cudaFree(d->taps_dev);
  bk::gen_destroy(d->gen);
  delete d;
  return BK_OK;
}

int bk_stencil_def_info(const bk_stencil_def_t *d, int *kind, int *radius, int *ntaps, int *st_iter, int *fused_steps) {
  BK_REQUIRE(d, "null stencil");
  if (kind) *kind = d->kind;
  if (radius) *radius = d->radius;
  if (ntaps) *ntaps = d->ntaps;
  if (st_iter) *st_iter = d->radius ? 8 / d->radius : 8;  // sweeps a ghost depth of 8 cells allows between two exchanges
  if (fused_steps) *fused_steps = (d->kind == BK_KIND_STAR && d->krad == 1) ? 2 : 1;
  return BK_OK;
}

int bk_stencil_def_advance(const bk_stencil_def_t *d, int steps, const bk_field_t *f, const unsigned *grid,
                           const unsigned *gdims, const unsigned *lo, const unsigned *hi, const unsigned *ready_lo,
                           const unsigned *ready_hi, int part, unsigned flags, void *stream) {
  BK_REQUIRE(d && f && f->adj && f->in && f->out && grid && gdims && lo && hi, "null argument");
  BK_REQUIRE(steps == 1 || steps == 2, "steps must be 1 or 2");
  const bool trust_grid = (part & BK_PART_GRID_TOPOLOGY) != 0;
  part &= ~BK_PART_GRID_TOPOLOGY;
  BK_REQUIRE(part == BK_PART_ALL ||
                 (ready_lo && ready_hi && ((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST)),
             "bad part");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  cudaStream_t s = (cudaStream_t) stream;
  if (d->kind == BK_KIND_STAR || d->kind == BK_KIND_CUBE) {
    if (steps == 1 && part == BK_PART_ALL && !trust_grid) return apply_spec(d->spec, f, grid, gdims, lo, hi, flags, s);
    const int ok = trust_grid ? 1 : marching_matches_adjacency(f->adj, grid, gdims, lo, hi, steps - 1, d->kind == BK_KIND_CUBE ? kSlotsAll : kSlotsStar, s);
    if (ok <= 0) return ok < 0 ? ok : adjacency_mismatch();
    return bk::launch_tiled(d->spec, *f, nullptr, 1, grid, gdims, lo, hi, s, part, ready_lo, ready_hi, steps);
  }
  if (d->kind == BK_KIND_GENERATED && steps == 1 && flags != BK_KERNEL_BRICK) {
    const dim3 box(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]);
    if (box.x == 0 || box.y == 0 || box.z == 0) return BK_OK;
    const int ok = trust_grid ? 1 : marching_matches_adjacency(f->adj, grid, gdims, lo, hi, 0, kSlotsAll, s);
    if (ok < 0) return ok;
    if (ok) {
      const int rc = bk::gen_launch(d->gen, *f, grid, gdims, lo, hi, s, part, ready_lo, ready_hi);
      if (rc != BK_EUNSUPPORTED || part != BK_PART_ALL || flags == BK_KERNEL_TILED) return rc;
    } else if (part != BK_PART_ALL || flags == BK_KERNEL_TILED) {
      return adjacency_mismatch();
    }
  }
  if (steps != 1 || part != BK_PART_ALL) return BK_EUNSUPPORTED;  // general taps: whole-box single sweeps only
  const dim3 g(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]);
  if (g.x == 0 || g.y == 0 || g.z == 0) return BK_OK;
  {
    bk_stencil_def *md = const_cast<bk_stencil_def *>(d);
    std::lock_guard<std::mutex> lk(md->mu);
    if (!md->taps_dev) {
      BK_CUDA(cudaMalloc(&md->taps_dev, sizeof(TapDev) * md->taps_host.size()));
      BK_CUDA(cudaMemcpy(md->taps_dev, md->taps_host.data(), sizeof(TapDev) * md->taps_host.size(), cudaMemcpyHostToDevice));
    }
  }
  Select sel = {grid, nullptr, nullptr, gdims[0], gdims[1], {lo[0], lo[1], lo[2]}, 0};
  if (d->krad == 1) k_taps<1><<<g, 256, 0, s>>>(sel, *f, d->taps_dev, d->ntaps, d->pre, d->post);
  if (d->krad == 2) k_taps<2><<<g, 256, 0, s>>>(sel, *f, d->taps_dev, d->ntaps, d->pre, d->post);
  if (d->krad == 4) k_taps<4><<<g, 256, 0, s>>>(sel, *f, d->taps_dev, d->ntaps, d->pre, d->post);
  BK_LAUNCHED();
  return BK_OK;
}

int bk_stencil_def_source(const bk_stencil_def_t *d, char *buf, size_t cap, size_t *len) {
  BK_REQUIRE(d, "null stencil");
  const std::string &src = d->gen ? bk::gen_source(d->gen) : d->gen_why;
  if (len) *len = src.size();
  if (buf && cap) {
    const size_t n = std::min(cap - 1, src.size());
    memcpy(buf, src.data(), n);
    buf[n] = 0;
  }
  return d->gen ? BK_OK : BK_EUNSUPPORTED;
}

int bk_stencil_def_apply(const bk_stencil_def_t *d, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                         const unsigned *lo, const unsigned *hi, unsigned flags, void *stream) {
  return bk_stencil_def_advance(d, 1, f, grid, gdims, lo, hi, nullptr, nullptr, BK_PART_ALL, flags, stream);
}

}  // extern "C"
