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

// implemented in bk_stencil_tiled.cu
int launch_tiled(int stencil, const bk_field_t &f, const bk_field_t *multi_dev, unsigned nsub, const unsigned *grid,
                 const unsigned *gdims, const unsigned *lo, const unsigned *hi, const double *coeff, cudaStream_t s,
                 int part = BK_PART_ALL, const unsigned *ready_lo = nullptr, const unsigned *ready_hi = nullptr,
                 int steps = 1);

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

int launch_brick(int stencil, const Select &sel, const bk_field_t &f, dim3 grid, const double *coeff, cudaStream_t s) {
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return BK_OK;
  if (stencil == BK_ST_MPI125PT) {
    CubeCoef cc;
    if (bk::cube_coef_for(stencil, &cc) < 0) return BK_EINVAL;
    k_brick<2, true><<<grid, 256, 0, s>>>(sel, f, cc);
  } else {
    StarCoef sc;
    const int r = bk::star_coef_for(stencil, coeff, &sc);
    if (r < 0) return BK_EINVAL;
    if (r == 1) k_brick<1, false><<<grid, 256, 0, s>>>(sel, f, sc);
    if (r == 2) k_brick<2, false><<<grid, 256, 0, s>>>(sel, f, sc);
    if (r == 4) k_brick<4, false><<<grid, 256, 0, s>>>(sel, f, sc);
  }
  BK_LAUNCHED();
  return BK_OK;
}

int check_box(const unsigned *gdims, const unsigned *lo, const unsigned *hi) {
  for (int a = 0; a < 3; ++a)
    if (!(lo[a] <= hi[a] && hi[a] <= gdims[a])) return BK_EINVAL;
  if (hi[1] - lo[1] > 65535 || hi[2] - lo[2] > 65535) return BK_EINVAL;
  return BK_OK;
}

}  // namespace

extern "C" {

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

int bk_stencil_apply(int stencil, const bk_field_t *f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
                     const unsigned *hi, const double *coeff, unsigned flags, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi, "null argument");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  cudaStream_t s = (cudaStream_t) stream;
  if (flags != BK_KERNEL_BRICK) {
    int rc = bk::launch_tiled(stencil, *f, nullptr, 1, grid, gdims, lo, hi, coeff, s);
    if (rc != BK_EUNSUPPORTED || flags == BK_KERNEL_TILED) return rc;
  }
  Select sel = {grid, nullptr, nullptr, gdims[0], gdims[1], {lo[0], lo[1], lo[2]}, 0};
  return launch_brick(stencil, sel, *f, dim3(hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]), coeff, s);
}

int bk_stencil_apply_part(int stencil, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                          const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                          const unsigned *ready_hi, int part, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi && ready_lo && ready_hi, "null argument");
  BK_REQUIRE((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST,
             "part must be BK_PART_READY or BK_PART_REST (optionally | BK_PART_THIN)");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  // only the marching kernel has the split enumeration; the caller falls back to whole-box launches on EUNSUPPORTED
  return bk::launch_tiled(stencil, *f, nullptr, 1, grid, gdims, lo, hi, coeff, (cudaStream_t) stream, part, ready_lo,
                          ready_hi);
}

int bk_stencil_advance(int stencil, int steps, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                       const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                       const unsigned *ready_hi, int part, void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(steps == 1 || steps == 2, "steps must be 1 or 2");
  BK_REQUIRE(f && f->adj && f->in && f->out && grid && gdims && lo && hi, "null argument");
  BK_REQUIRE(part == BK_PART_ALL ||
                 (ready_lo && ready_hi && ((part & ~BK_PART_THIN) == BK_PART_READY || (part & ~BK_PART_THIN) == BK_PART_REST)),
             "bad part");
  BK_REQUIRE(f->in_step >= 512 && f->out_step >= 512, "brick step smaller than a brick");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  return bk::launch_tiled(stencil, *f, nullptr, 1, grid, gdims, lo, hi, coeff, (cudaStream_t) stream, part, ready_lo,
                          ready_hi, steps);
}

int bk_stencil_apply_list(int stencil, const bk_field_t *f, const unsigned *ids, size_t n, const double *coeff,
                          void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(f && f->adj && f->in && f->out && (ids || n == 0), "null argument");
  BK_REQUIRE(f->in != f->out, "in-place sweep is not defined");
  BK_REQUIRE(n < (1ull << 31), "list too long");
  Select sel = {nullptr, ids, nullptr, 0, 0, {0, 0, 0}, 0};
  return launch_brick(stencil, sel, *f, dim3((unsigned) n, 1, 1), coeff, (cudaStream_t) stream);
}

int bk_stencil_apply_multi(int stencil, const bk_field_t *fields_dev, unsigned nsub, const unsigned *grid,
                           const unsigned *gdims, const unsigned *lo, const unsigned *hi, const double *coeff,
                           void *stream) {
  BK_REQUIRE(stencil >= 0 && stencil < BK_ST_COUNT, "unknown stencil id");
  BK_REQUIRE(fields_dev && grid && gdims && lo && hi && nsub > 0, "null argument");
  BK_REQUIRE(check_box(gdims, lo, hi) == BK_OK, "brick box outside the grid");
  const unsigned nx = hi[0] - lo[0];
  BK_REQUIRE((unsigned long long) nx * nsub < (1ull << 31) && nsub <= 65535, "launch too wide");
  if (!getenv("BK_MULTI_BRICK")) {  // developer knob: force the per-brick family
    bk_field_t none = {nullptr, nullptr, 512, nullptr, 512};
    int rc = bk::launch_tiled(stencil, none, fields_dev, nsub, grid, gdims, lo, hi, coeff, (cudaStream_t) stream);
    if (rc != BK_EUNSUPPORTED) return rc;
  }
  Select sel = {grid, nullptr, fields_dev, gdims[0], gdims[1], {lo[0], lo[1], lo[2]}, nx};
  bk_field_t dummy = {nullptr, nullptr, 512, nullptr, 512};
  return launch_brick(stencil, sel, dummy, dim3(nx * nsub, hi[1] - lo[1], hi[2] - lo[2]), coeff, (cudaStream_t) stream);
}

}  // extern "C"
