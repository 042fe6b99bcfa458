// bk_stencil_tiled.cu -- the fast path for box-shaped launches of the STAR stencils (7/13/25-point).
//
// Replaces the body codegen/vecscatter generates for brick("mpi*pt.py","CUDA",(8,8,8),(4,8),b) (one warp per brick,
// 8-byte loads, dev_shl shuffles; launched by brick_kernel, weak/main.cu:35-43) with a TMA-staged marching kernel:
//
//   * a CTA owns a tile of TI x TJ bricks in (i,j) and marches along k through KL brick layers (a "segment");
//   * one producer warp streams the column into a ring of D shared-memory stages, G planes per stage, with 1-D bulk
//     async copies (cp.async.bulk ... mbarrier::complete_tx): a brick's G planes are G*512 contiguous bytes, so one
//     copy per own brick and per i-neighbour brick, and G copies of R rows (R*64 contiguous bytes) per j-neighbour;
//     full/empty mbarriers, no __syncthreads in the steady state;
//   * consumer threads own an (x-pair) x (YT rows) patch of one brick and keep 2R+1 partial outputs per point in
//     registers: plane t is read ONCE from shared memory and scattered along k into the outputs t-R..t+R, the i/j
//     neighbours of plane t are read from the same stage (128-bit LDS, bank-conflict free thanks to a 64-byte skew of
//     odd brick columns), finished planes go straight to global memory with 128-bit stores.
// HBM sees every input byte once (halo re-reads hit L2 because neighbouring tiles run concurrently) and every output
// byte once: 16 B per point, the roofline SURVEY.md section 8(d) counts.
//
// Brick ids come from the dense `grid` array (BrickDecomp::operator[] / init_grid order); positions outside the grid
// read brick 0, the null brick every out-of-domain adjacency entry points to (include/brick-mpi.h:275-277).  The
// per-brick family in bk_stencil.cu remains the general path for arbitrary adjacency.
//
// Summation order per output point (fixed): k-minus taps d=R..1, centre, i taps d=1..R (+d then -d), j taps d=1..R
// (+d then -d), k-plus taps d=1..R.  Parity with the reference is to 1e-12 relative, not bitwise.
#include "bk_common.h"
#include "bk_diamond.h"
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <tuple>
#include <vector>

namespace {

using bk::StarCoef;

#include "bk_march_common.h"
static_assert(sizeof(bk_field_dev) == sizeof(bk_field_t) && offsetof(bk_field_dev, out_step) == offsetof(bk_field_t, out_step),
              "bk_field_dev restates bk_field_t");

// ---- compile-time geometry --------------------------------------------------------------------------------------
template <int R_, int YT_, int TI_, int TJ_, int G_, int D_, int MAXREG_ = 255, int NPW_ = 1, bool CUBE_ = false,
          int CREG_ = 0, int PREG_ = 40, int FLAGS_ = 0>
struct Cfg {
  // CREG > 0: register re-balancing between the roles (setmaxnreg, warpgroup granular): the kernel starts with
  // 65536/NT registers per thread, the producer warpgroup shrinks to PREG and the consumer warpgroups grow to CREG
  [[maybe_unused]] static constexpr int CREG = CREG_, PREG = PREG_, FLAGS = FLAGS_;
  static constexpr bool CUBE = CUBE_;                 // (2R+1)^3 cube stencil instead of a star
  // FLAGS & 2: the composed two-step update of a radius-1 star (bk_diamond.h) on the radius-2 geometry
  static constexpr bool DIAM = (FLAGS_ & 2) != 0;
  using Coef = typename std::conditional<CUBE_, bk::CubeCoef,
                                         typename std::conditional<DIAM, bk::DiamondCoef, bk::StarCoef>::type>::type;
  static constexpr int R = R_, YT = YT_, TI = TI_, TJ = TJ_, G = G_, D = D_;
  static constexpr int W = 2 * R + 1;                 // partial outputs in flight per point
  static constexpr int RUP = ((R + G - 1) / G) * G;   // halo planes streamed before/after a segment (multiple of G)
  static constexpr int SW = TI + 2;                   // slot columns: i-halo, TI own, i-halo
  static constexpr int SH = TJ + 1;                   // slot rows: row 0 = shared j-halo slots, then TJ own rows
  static constexpr int SLOTP = G * 512 + 64;          // slot pitch = 64 mod 128: i-adjacent bricks land in opposite bank halves
  static constexpr int STAGE = ((SH * SW * SLOTP + 127) / 128) * 128;
  static constexpr int NCONS = TI * TJ * 32 / YT;     // consumer threads: 4 x-pairs * 8/YT row groups per brick
  static constexpr int NCW = NCONS / 32;
  static constexpr int NPW = NPW_;                    // producer warps (bulk-copy issue is instruction bound)
  static constexpr int NT = NCONS + 32 * NPW;
  static constexpr int NJH = (CUBE || DIAM) ? TI + 2 : TI;  // j-halo columns: cube and diamond also need the corner bricks
  static constexpr int NCOPY = TI * TJ + 2 * TJ + 2 * NJH;  // copy jobs per stage (j-halo jobs issue G copies)
  static constexpr int JOBS = (NCOPY + 32 * NPW - 1) / (32 * NPW);
  [[maybe_unused]] static constexpr int MAXREG = MAXREG_;              // register cap (chosen so that the intended CTAs/SM fit)
  static constexpr size_t SMEM = (size_t) D * STAGE + 2 * D * 8 + 128;
  static constexpr int OVH = 2 * RUP + 2;             // cost-model overhead planes per segment (halo planes + fill)
  static constexpr bool FUSED = false;
  static_assert(8 % G == 0 && 8 % YT == 0 && TI % 2 == 0 && 2 * R <= 8, "geometry");
  static_assert(!DIAM || (R == 2 && !CUBE), "the composed two-step update runs on the radius-2 star geometry");
  static_assert(CREG == 0 || (NCW % 4 == 0 && NPW == 4 && NCONS * CREG + 128 * PREG <= 65536), "setmaxnreg needs warpgroups");
  // SW is even, so the bank half of a slot depends on its column only: a quarter warp (4 x-pairs of 2 i-adjacent
  // bricks) reads 8 distinct 16-B bank groups
  __host__ __device__ static constexpr int slotoff(int bi, int bj) { return (bj * SW + bi) * SLOTP; }
};

// bk_stencil_remote.cu compiles this file a second time with BK_REMOTE_TU defined: the same marching body with ONE more
// line in the producer (ghost bricks are read in place from the neighbours' storages, RemoteArgs) and its own entry
// point.  Keeping that variant in a translation unit of its own leaves the code of the kernels below untouched.
#ifdef BK_REMOTE_TU
#define BK_REM_PARAM , const RemoteArgs &rem
#else
#define BK_REM_PARAM
#endif
template <class C>
__device__ __forceinline__ void march_body(const TiledArgs &a, const typename C::Coef &cf BK_REM_PARAM) {
  constexpr int R = C::R, YT = C::YT, TI = C::TI, TJ = C::TJ, G = C::G, D = C::D, W = C::W, RUP = C::RUP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // dynamic shared memory is only guaranteed 16-B aligned: align the ring to 128 B by hand
  unsigned char *ring = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar_full = ring_u32 + D * C::STAGE;
  const uint32_t bar_empty = bar_full + D * 8;

  const double *fin = a.in;
  double *fout = a.out;
  size_t in_step = a.in_step, out_step = a.out_step;
  if (a.multi) {  // strong/main.cu:85-99: many subdomains share grid and adjacency, each has its own storages
    const bk_field_dev f = a.multi[blockIdx.z];
    fin = f.in, fout = f.out, in_step = f.in_step, out_step = f.out_step;
  }
  const int tid = threadIdx.x;
  int bq = 0, brel = (int) blockIdx.x;
  while (bq + 1 < a.nbox && brel >= a.box[bq + 1].first) ++bq;
  brel -= a.box[bq].first;
  const int tx = a.box[bq].lo[0] + brel % a.box[bq].dim[0];
  brel /= a.box[bq].dim[0];
  const int ty = a.box[bq].lo[1] + brel % a.box[bq].dim[1];
  const int tseg = a.box[bq].lo[2] + brel / a.box[bq].dim[1];
  const int i0 = a.lo[0] + tx * TI, j0 = a.lo[1] + ty * TJ;
  int kb0, nl;  // first brick layer and number of layers of this segment
  seg_range(a, tseg, kb0, nl);
  const int P = nl * 8 + 2 * RUP;            // planes streamed; plane t is absolute plane kb0*8 - RUP + t
  const int NS = P / G;

  if (tid == 0) {
    for (int s = 0; s < D; ++s) {
      mbar_init(bar_full + 8 * s, 32 * C::NPW);  // every producer lane arrives once per fill
      mbar_init(bar_empty + 8 * s, C::NCW);   // one arrival per consumer warp per drain
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if constexpr (C::CREG > 0) {
    if (tid >= C::NCONS)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::PREG));
    else
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::CREG));
  }
  if (tid >= C::NCONS) {
    // ================================================ producer warp ============================================
    // jobs are dealt round-robin over the producer warps, then over lanes
    const int pw = (tid - C::NCONS) >> 5, lane = tid & 31;
    int sbi[C::JOBS], sbj[C::JOBS], kind[C::JOBS];  // slot coordinates; kind 0 none, 1 whole slot, 2 low rows, 3 high rows
    uint32_t dsto[C::JOBS];
    unsigned idn[C::JOBS];
#pragma unroll
    for (int q = 0; q < C::JOBS; ++q) {
      int job = pw + C::NPW * (lane + 32 * q);
      kind[q] = 0;
      sbi[q] = sbj[q] = 0;
      if (job < TI * TJ) {
        kind[q] = 1, sbi[q] = 1 + job % TI, sbj[q] = 1 + job / TI;
      } else if ((job -= TI * TJ) < 2 * TJ) {
        kind[q] = 1, sbi[q] = (job & 1) ? TI + 1 : 0, sbj[q] = 1 + (job >> 1);
      } else if ((job -= 2 * TJ) < 2 * C::NJH) {
        kind[q] = 2 + (job & 1), sbi[q] = ((C::CUBE || C::DIAM) ? 0 : 1) + (job >> 1), sbj[q] = (job & 1) ? TJ + 1 : 0;
      }
      // both j-halo kinds land in slot row 0: rows [8-R,8) from below (kind 2), rows [0,R) from above (kind 3)
      dsto[q] = C::slotoff(sbi[q], kind[q] >= 2 ? 0 : sbj[q]) + (kind[q] == 2 ? (8 - R) * 64 : 0);
    }
    auto brick_id = [&](int q, int kb) -> unsigned {
      const int gi = i0 + sbi[q] - 1, gj = j0 + sbj[q] - 1;
      if (kind[q] == 0 || gi < 0 || gi >= a.gx || gj < 0 || gj >= a.gy || kb < 0 || kb >= a.gz) return 0u;
      return __ldg(a.grid + ((size_t) kb * a.gy + gj) * a.gx + gi);
    };
    const int z_first = kb0 * 8 - RUP;  // may be negative (only when kb0 == 0)
    int kb = (z_first >= 0) ? z_first / 8 : -((7 - z_first) / 8);
    int pz = z_first - kb * 8;
#pragma unroll
    for (int q = 0; q < C::JOBS; ++q) idn[q] = brick_id(q, kb);
    unsigned idc[C::JOBS];
    int st = 0;
    uint32_t ph = 0;
    bool fresh = true;  // idn holds the ids of layer kb
    for (int n = 0; n < NS; ++n) {
      if (fresh) {
#pragma unroll
        for (int q = 0; q < C::JOBS; ++q) idc[q] = idn[q], idn[q] = brick_id(q, kb + 1);  // prefetch next layer
        fresh = false;
      }
      if (n >= D) mbar_wait(bar_empty + 8 * st, ph ^ 1);
      uint32_t bytes = 0;
#pragma unroll
      for (int q = 0; q < C::JOBS; ++q) bytes += kind[q] == 0 ? 0 : kind[q] == 1 ? G * 512 : G * R * 64;
      const uint32_t fb = bar_full + 8 * st;
      mbar_expect_tx(fb, bytes);
      const uint32_t sb = ring_u32 + st * C::STAGE;
#pragma unroll
      for (int q = 0; q < C::JOBS; ++q) {
        const double *src = fin + (size_t) idc[q] * in_step + pz * 64;
#ifdef BK_REMOTE_TU
        {  // a ghost brick: the neighbour's skin brick it mirrors, where it lies (over NVLink for a peer's)
          const unsigned g = idc[q] - rem.ghost_lo;
          if (g < rem.ghost_n) src = rem.remap[g] + pz * 64;
        }
#endif
        if (kind[q] == 1) {
          bulk_g2s(sb + dsto[q], src, G * 512, fb);
        } else if (kind[q] >= 2) {
          src += (kind[q] == 2 ? (8 - R) * 8 : 0);
#pragma unroll
          for (int g = 0; g < G; ++g) bulk_g2s(sb + dsto[q] + g * 512, src + g * 64, R * 64, fb);
        }
      }
      pz += G;
      if (pz == 8) pz = 0, ++kb, fresh = true;
      if (++st == D) st = 0, ph ^= 1;
    }
    return;
  }

  // ================================================== consumers ================================================
  const int c = tid & 3;                 // x-pair inside the brick row: cells x0 = 2c, 2c+1
  const int e = (tid >> 2) & 1;          // brick parity inside a pair of i-adjacent bricks
  int rest = tid >> 3;
  const bool swp = (R % 2 == 1) && (rest & 1);  // odd row group of the half warp (see the i-halo loads)
  const int y0 = (rest % (8 / YT)) * YT;  // first of my YT rows
  rest /= (8 / YT);
  const int bi = (rest % (TI / 2)) * 2 + e, bj = rest / (TI / 2);
  const int own_slot = C::slotoff(bi + 1, bj + 1);
  const int own_off = own_slot + y0 * 64 + c * 16;
  // byte offsets (from the stage base, plane 0) of the 2R j-halo rows: below rows y0-R..y0-1, above rows y0+YT..
  int joff[2 * R];
#pragma unroll
  for (int h = 0; h < 2 * R; ++h) {
    const int ya = (h < R) ? y0 - R + h : y0 + YT + (h - R);
    int base;
    if (ya < 0)
      base = C::slotoff(bi + 1, bj) + (8 + ya) * 64;  // bj == 0 -> slot row 0 = the shared j-halo slot
    else if (ya >= 8)
      base = C::slotoff(bi + 1, (bj == TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
    else
      base = own_slot + ya * 64;
    joff[h] = base + c * 16;
  }
  // i-halo offsets relative to the address of my x-pair in a row
  const int dl = C::slotoff(bi, bj + 1) - own_slot, dr = C::slotoff(bi + 2, bj + 1) - own_slot;
  constexpr int NI = (R % 2 == 0) ? R / 2 : R;  // loads per side: 16-B chunks for even R, single cells for odd R
  int ioffL[NI], ioffR[NI];
#pragma unroll
  for (int m = 1; m <= NI; ++m) {
    if (R % 2 == 0) {
      ioffL[m - 1] = -16 * m + ((c - m < 0) ? dl + 64 : 0);
      ioffR[m - 1] = 16 * m + ((c + m > 3) ? dr - 64 : 0);
    } else {  // cell x0-m / x0+1+m
      ioffL[m - 1] = -8 * m + ((2 * c - m < 0) ? dl + 64 : 0);
      ioffR[m - 1] = 8 * (1 + m) + ((2 * c + 1 + m > 7) ? dr - 64 : 0);
      if (swp) {
        const int t = ioffL[m - 1];
        ioffL[m - 1] = ioffR[m - 1], ioffR[m - 1] = t;
      }
    }
  }

  const bool mine = (i0 + bi < a.hi[0]) && (j0 + bj < a.hi[1]);  // bricks of a partial tile are staged, not stored
  const unsigned *gcol = a.grid + ((size_t) kb0 * a.gy + (j0 + bj)) * a.gx + (i0 + bi);
  const size_t glayer = (size_t) a.gy * a.gx;
  unsigned id_next = mine ? __ldg(gcol) : 0u;
  double *outp = fout;
  [[maybe_unused]] unsigned edge_ij = 0;  // diamond: which of my cells lie on a grid face along i / j (bk_diamond.h)
  if constexpr (C::DIAM) {
    const int gi = i0 + bi, gj = j0 + bj;
    if (gi == 0 && c == 0) edge_ij |= 1u;
    if (gi == a.gx - 1 && c == 3) edge_ij |= 2u;
#pragma unroll
    for (int r = 0; r < YT; ++r)
      if ((gj == 0 && y0 + r == 0) || (gj == a.gy - 1 && y0 + r == 7)) edge_ij |= 4u << r;
  }

  double2 acc[W][YT];
#pragma unroll
  for (int w = 0; w < W; ++w)
#pragma unroll
    for (int r = 0; r < YT; ++r) acc[w][r] = make_double2(0.0, 0.0);

  int st = 0, pl = 0;
  uint32_t ph = 0;
  int orel = -R - RUP;  // output plane finished by the current iteration, relative to the segment start
  const int nout = nl * 8;

  // plane `orel` of the segment is complete: 128-bit stores straight into the output brick
  auto store_plane = [&](const double2 (&fin)[YT]) {
    if (orel >= 0 && orel < nout) {
      const int oz = orel & 7;
      if (oz == 0) {
        outp = fout + (size_t) id_next * out_step + y0 * 8 + c * 2;
        if (mine && (orel >> 3) + 1 < nl) id_next = __ldg(gcol + ((orel >> 3) + 1) * glayer);
      }
      if (mine) {
#pragma unroll
        for (int r = 0; r < YT; ++r) *reinterpret_cast<double2 *>(outp + oz * 64 + r * 8) = fin[r];
      }
    }
  };

#pragma unroll 1
  for (int tb = 0; tb < P; tb += W) {
#pragma unroll
    for (int u = 0; u < W; ++u) {
      if (tb + u < P) {
        if (pl == 0) mbar_wait(bar_full + 8 * st, ph);
        const unsigned char *pb = ring + st * C::STAGE + pl * 512;
        const int sF = ((u - R) % W + W) % W;  // slot of output t-R: finished by this plane
        const int s0 = u % W;
        const int sN = (u + R) % W;            // slot of output t+R: first contribution, (re)initialises the slot
        if constexpr (C::CUBE) {
          // Sign symmetry of the cube coefficients (c depends on |dx|,|dy|,|dz| only, stencils/mpi125pt.py:13-32):
          // fold +-dx, then +-dy, then weight the 9 folded values once per |dz| and scatter along k.
          // 2*(YT+4)/YT + 6 + 3 adds, 18 FMA and 3 adds per point instead of 125 FMA.
          static_assert(!C::CUBE || R == 2, "cube path is written for radius 2");
          double2 X0[YT + 4], X1[YT + 4], X2[YT + 4];
#pragma unroll
          for (int rr = 0; rr < YT + 4; ++rr) {
            const unsigned char *pr = (rr < 2) ? pb + joff[rr] : (rr >= YT + 2) ? pb + joff[rr - YT] : pb + own_off + (rr - 2) * 64;
            const double2 ctr = *reinterpret_cast<const double2 *>(pr);
            const double2 lft = *reinterpret_cast<const double2 *>(pr + ioffL[0]);
            const double2 rgt = *reinterpret_cast<const double2 *>(pr + ioffR[0]);
            X0[rr] = ctr;
            X1[rr] = make_double2(lft.y + ctr.y, ctr.x + rgt.x);
            X2[rr] = make_double2(lft.x + rgt.x, lft.y + rgt.y);
          }
#pragma unroll
          for (int r = 0; r < YT; ++r) {
            const int rr = r + 2;
            double2 f[3][3];  // [|dy|][|dx|]
            f[0][0] = X0[rr], f[0][1] = X1[rr], f[0][2] = X2[rr];
            f[1][0] = make_double2(X0[rr - 1].x + X0[rr + 1].x, X0[rr - 1].y + X0[rr + 1].y);
            f[1][1] = make_double2(X1[rr - 1].x + X1[rr + 1].x, X1[rr - 1].y + X1[rr + 1].y);
            f[1][2] = make_double2(X2[rr - 1].x + X2[rr + 1].x, X2[rr - 1].y + X2[rr + 1].y);
            f[2][0] = make_double2(X0[rr - 2].x + X0[rr + 2].x, X0[rr - 2].y + X0[rr + 2].y);
            f[2][1] = make_double2(X1[rr - 2].x + X1[rr + 2].x, X1[rr - 2].y + X1[rr + 2].y);
            f[2][2] = make_double2(X2[rr - 2].x + X2[rr + 2].x, X2[rr - 2].y + X2[rr + 2].y);
            // the coefficient is also symmetric under permutation of (|dx|,|dy|,|dz|): cc[az][ay][ax] == cc[az][ax][ay],
            // so the mirrored folds share one FMA per |dz| (6 instead of 9)
            const double2 g[6] = {f[0][0],
                                  make_double2(f[0][1].x + f[1][0].x, f[0][1].y + f[1][0].y),
                                  make_double2(f[0][2].x + f[2][0].x, f[0][2].y + f[2][0].y),
                                  f[1][1],
                                  make_double2(f[1][2].x + f[2][1].x, f[1][2].y + f[2][1].y),
                                  f[2][2]};
            constexpr int GY[6] = {0, 0, 0, 1, 1, 2}, GX[6] = {0, 1, 2, 1, 2, 2};
            double2 p0 = acc[s0][r], p1, p2;
            p1 = make_double2(cf.cc[1][0][0] * g[0].x, cf.cc[1][0][0] * g[0].y);
            p2 = make_double2(cf.cc[2][0][0] * g[0].x, cf.cc[2][0][0] * g[0].y);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              p0.x = fma(cf.cc[0][GY[q]][GX[q]], g[q].x, p0.x), p0.y = fma(cf.cc[0][GY[q]][GX[q]], g[q].y, p0.y);
              if (q > 0) {
                p1.x = fma(cf.cc[1][GY[q]][GX[q]], g[q].x, p1.x), p1.y = fma(cf.cc[1][GY[q]][GX[q]], g[q].y, p1.y);
                p2.x = fma(cf.cc[2][GY[q]][GX[q]], g[q].x, p2.x), p2.y = fma(cf.cc[2][GY[q]][GX[q]], g[q].y, p2.y);
              }
            }
            const int sm1 = ((u - 1) % W + W) % W, sp1 = (u + 1) % W;
            acc[sF][r].x += p2.x, acc[sF][r].y += p2.y;
            acc[sm1][r].x += p1.x, acc[sm1][r].y += p1.y;
            acc[s0][r] = p0;
            acc[sp1][r].x += p1.x, acc[sp1][r].y += p1.y;
            acc[sN][r] = p2;
          }
          store_plane(acc[sF]);
        } else if constexpr (C::DIAM) {
          // two time steps of a radius-1 star as one composed update (bk_diamond.h); plane tb+u is absolute plane zt
          static_assert(!C::DIAM || (W == 5 && R == 2), "diamond_plane keeps five partial outputs");
          double2 v[YT];
#pragma unroll
          for (int r = 0; r < YT; ++r) v[r] = *reinterpret_cast<const double2 *>(pb + own_off + r * 64);
          const int zt = kb0 * 8 - RUP + tb + u;
          const unsigned edge = edge_ij | ((zt == 0 || zt == a.gz * 8 - 1) ? bk::kDiamondEdgeK : 0u);
          bk::diamond_plane<YT>(pb, own_off, joff[0], joff[1], joff[2 * R - 2], joff[2 * R - 1], ioffL[0], ioffR[0],
                                ((tid >> 3) & 1) != 0, cf, acc, u, edge, v);
          store_plane(acc[sF]);
        } else {
          double2 v[YT];
#pragma unroll
          for (int r = 0; r < YT; ++r) v[r] = *reinterpret_cast<const double2 *>(pb + own_off + r * 64);

          // ---- k taps, first half: this plane is the +d neighbour of outputs t-d ------------------------------------
#pragma unroll
          for (int r = 0; r < YT; ++r) {
            acc[sF][r].x = fma(cf.cp[2][R - 1], v[r].x, acc[sF][r].x);
            acc[sF][r].y = fma(cf.cp[2][R - 1], v[r].y, acc[sF][r].y);
          }
          store_plane(acc[sF]);
#pragma unroll
          for (int d = R - 1; d >= 1; --d) {
            const int s = ((u - d) % W + W) % W;
#pragma unroll
            for (int r = 0; r < YT; ++r) {
              acc[s][r].x = fma(cf.cp[2][d - 1], v[r].x, acc[s][r].x);
              acc[s][r].y = fma(cf.cp[2][d - 1], v[r].y, acc[s][r].y);
            }
          }
          // ---- centre + in-plane taps of output t ------------------------------------------------------------------
          double2 rows[YT + 2 * R];
#pragma unroll
          for (int h = 0; h < R; ++h) rows[h] = *reinterpret_cast<const double2 *>(pb + joff[h]);
#pragma unroll
          for (int r = 0; r < YT; ++r) rows[R + r] = v[r];
#pragma unroll
          for (int h = 0; h < R; ++h) rows[R + YT + h] = *reinterpret_cast<const double2 *>(pb + joff[R + h]);
#pragma unroll
          for (int r = 0; r < YT; ++r) {
            double line[2 * R + 2];  // cells x0-R .. x0+1+R of row r
            line[R] = v[r].x, line[R + 1] = v[r].y;
            const unsigned char *pr = pb + own_off + r * 64;
            if constexpr (R % 2 == 0) {
#pragma unroll
              for (int m = 1; m <= R / 2; ++m) {
                const double2 lft = *reinterpret_cast<const double2 *>(pr + ioffL[m - 1]);
                const double2 rgt = *reinterpret_cast<const double2 *>(pr + ioffR[m - 1]);
                line[R - 2 * m] = lft.x, line[R - 2 * m + 1] = lft.y;
                line[R + 2 * m] = rgt.x, line[R + 2 * m + 1] = rgt.y;
              }
            } else {
#pragma unroll
              for (int m = 1; m <= R; ++m) {  // odd row groups load right first: a half warp then covers all 32 banks
                const double q0 = *reinterpret_cast<const double *>(pr + ioffL[m - 1]);
                const double q1 = *reinterpret_cast<const double *>(pr + ioffR[m - 1]);
                line[R - m] = swp ? q1 : q0;
                line[R + 1 + m] = swp ? q0 : q1;
              }
            }
            double ax = fma(cf.c0, v[r].x, acc[s0][r].x), ay = fma(cf.c0, v[r].y, acc[s0][r].y);
            if constexpr (C::FLAGS & 1) {
              // shorter dependency chains: the i taps are summed on their own and added at the end
              double ix = cf.cp[0][0] * line[R + 1], iy = cf.cp[0][0] * line[R + 2];
              ix = fma(cf.cm[0][0], line[R - 1], ix), iy = fma(cf.cm[0][0], line[R], iy);
#pragma unroll
              for (int d = 2; d <= R; ++d) {
                ix = fma(cf.cp[0][d - 1], line[R + d], ix);
                iy = fma(cf.cp[0][d - 1], line[R + 1 + d], iy);
                ix = fma(cf.cm[0][d - 1], line[R - d], ix);
                iy = fma(cf.cm[0][d - 1], line[R + 1 - d], iy);
              }
#pragma unroll
              for (int d = 1; d <= R; ++d) {
                ax = fma(cf.cp[1][d - 1], rows[R + r + d].x, ax);
                ay = fma(cf.cp[1][d - 1], rows[R + r + d].y, ay);
                ax = fma(cf.cm[1][d - 1], rows[R + r - d].x, ax);
                ay = fma(cf.cm[1][d - 1], rows[R + r - d].y, ay);
              }
              acc[s0][r].x = ax + ix, acc[s0][r].y = ay + iy;
              continue;
            }
#pragma unroll
            for (int d = 1; d <= R; ++d) {
              ax = fma(cf.cp[0][d - 1], line[R + d], ax);
              ay = fma(cf.cp[0][d - 1], line[R + 1 + d], ay);
              ax = fma(cf.cm[0][d - 1], line[R - d], ax);
              ay = fma(cf.cm[0][d - 1], line[R + 1 - d], ay);
            }
#pragma unroll
            for (int d = 1; d <= R; ++d) {
              ax = fma(cf.cp[1][d - 1], rows[R + r + d].x, ax);
              ay = fma(cf.cp[1][d - 1], rows[R + r + d].y, ay);
              ax = fma(cf.cm[1][d - 1], rows[R + r - d].x, ax);
              ay = fma(cf.cm[1][d - 1], rows[R + r - d].y, ay);
            }
            acc[s0][r].x = ax, acc[s0][r].y = ay;
          }
          // ---- k taps, second half: this plane is the -d neighbour of outputs t+d -----------------------------------
#pragma unroll
          for (int d = 1; d <= R - 1; ++d) {
            const int s = (u + d) % W;
#pragma unroll
            for (int r = 0; r < YT; ++r) {
              acc[s][r].x = fma(cf.cm[2][d - 1], v[r].x, acc[s][r].x);
              acc[s][r].y = fma(cf.cm[2][d - 1], v[r].y, acc[s][r].y);
            }
          }
#pragma unroll
          for (int r = 0; r < YT; ++r) {
            acc[sN][r].x = cf.cm[2][R - 1] * v[r].x;
            acc[sN][r].y = cf.cm[2][R - 1] * v[r].y;
          }

        }
        ++orel;
        if (++pl == G) {
          pl = 0;
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(bar_empty + 8 * st);
          if (++st == D) st = 0, ph ^= 1;
        }
      }
    }
  }
}

#ifndef BK_REMOTE_TU
template <class C>
__global__ void __launch_bounds__(C::NT) k_star(const __grid_constant__ TiledArgs a,
                                                const __grid_constant__ typename C::Coef cf) {
  march_body<C>(a, cf);
}
// same body under an explicit register cap (so that two CTAs fit one SM)
template <class C>
__global__ void __launch_bounds__(C::NT) __maxnreg__(C::MAXREG) k_star_capped(const __grid_constant__ TiledArgs a,
                                                                               const __grid_constant__ typename C::Coef cf) {
  march_body<C>(a, cf);
}
#else
// the same three flavours (plain, register cap, register re-balancing) with the RemoteArgs table as a third argument
template <class C>
__global__ void __launch_bounds__(C::NT) k_star_remote(const __grid_constant__ TiledArgs a, const __grid_constant__ typename C::Coef cf,
                                                       const __grid_constant__ RemoteArgs r) {
  march_body<C>(a, cf, r);
}
template <class C>
__global__ void __launch_bounds__(C::NT) __maxnreg__(C::MAXREG)
    k_star_capped_remote(const __grid_constant__ TiledArgs a, const __grid_constant__ typename C::Coef cf, const __grid_constant__ RemoteArgs r) {
  march_body<C>(a, cf, r);
}
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
    k_star_rebal_remote(const __grid_constant__ TiledArgs a, const __grid_constant__ typename C::Coef cf, const __grid_constant__ RemoteArgs r) {
  march_body<C>(a, cf, r);
}
#endif

#ifndef BK_REMOTE_TU
// ==================================================================================================================
// Two time steps per pass (temporal blocking) for the star stencils.
//
// The reference exchanges ghost zones once per ST_ITER sweeps and recomputes the ghost shell in between
// (weak/main.cu:246-287): between two exchanges consecutive sweeps have no outside dependency, so two of them can be
// applied in ONE pass over HBM.  The CTA streams input planes exactly as above (ring of D one-plane stages, halo depth
// 2R), stage A turns input plane t into intermediate plane t-R (own tile plus an R-wide strip around it) and writes it
// to a 2-plane ring in shared memory, stage B turns that intermediate plane into output plane t-2R with the same
// register scatter along k.  The intermediate field never touches HBM: 16 B of traffic per point per TWO steps.
// Semantics = bk_stencil_apply over the whole grid followed by bk_stencil_apply over [lo,hi): the intermediate is
// computed at every in-grid cell the second step reads and is zero outside the grid (the null brick).
template <int R_, int YT_, int TI_, int TJ_, int D_, int NPW_, int MINB_ = 1, int CREG_ = 0, int PREG_ = 40, bool LAG_ = false,
          bool EW_ = false, int ABL_ = 0>
struct FCfg {
  static constexpr int ABL = ABL_;                     // developer ablations (WRONG results): 1 no strip, 2 no i-halo loads, 4 no masks, 8 no mid store
  static constexpr bool EW = EW_;                      // wait for the next input plane BEFORE stage B (no spin loop between B(n) and A(n+1))
  static constexpr bool LAG = LAG_;                    // stage B runs one plane behind stage A (independent work between barriers)
  static constexpr int MINB = MINB_;                   // CTAs per SM the register allocation must allow
  static constexpr int CREG = CREG_, PREG = PREG_;     // setmaxnreg re-balancing as in Cfg (0 = off)
  static constexpr bool CUBE = false, FUSED = true;
  using Coef = bk::StarCoef;
  static constexpr int R = R_, YT = YT_, TI = TI_, TJ = TJ_, D = D_;
  static constexpr int W = 2 * R + 1;
  static constexpr int H = 2 * R;                      // halo depth of the loads: two steps
  static constexpr int SW = TI + 2, SH = TJ + 1;
  static constexpr int SLOTP = 512 + 64;               // one plane per stage; same bank skew as Cfg
  static constexpr int STAGE = ((SH * SW * SLOTP + 127) / 128) * 128;
  static constexpr int NCONS = TI * TJ * 32 / YT, NCW = NCONS / 32, NPW = NPW_, NT = NCONS + 32 * NPW;
  static constexpr int NCOPY = TI * TJ + 2 * TJ + 2 * (TI + 2);   // own, i-halo bricks, j-halo rows incl. corner columns
  static constexpr int JOBS = (NCOPY + 32 * NPW - 1) / (32 * NPW);
  static constexpr int NSTRIP = 16 * R * (TI + TJ);    // intermediate points outside the tile, one per thread
  static constexpr int M = LAG ? 3 : 2;                // intermediate planes in flight
  static constexpr size_t SMEM = (size_t) (D + M) * STAGE + 2 * D * 8 + 128;
  static constexpr int OVH = 2 * H + 3;                // cost-model overhead planes per segment
  [[maybe_unused]] static constexpr int MAXREG = 255;
  static_assert(NSTRIP <= NCONS && 8 % YT == 0 && TI % 2 == 0 && 2 * H <= 8 && !(EW && LAG), "geometry");
  static_assert(CREG == 0 || (NCW % 4 == 0 && NPW % 4 == 0 && NCONS * CREG + 32 * NPW * PREG <= 65536), "setmaxnreg");
  __host__ __device__ static constexpr int slotoff(int bi, int bj) { return (bj * SW + bi) * SLOTP; }
  // byte offset of the cell at tile-relative (X,Y), X in [-8, 8TI+8), Y in [-H, 8TJ+H): rows below/above the tile live in
  // slot row 0 (rows [8-H,8) and [0,H))
  __device__ static int cell(int X, int Y) {
    const int sx = (X + 8) >> 3, cx = (X + 8) & 7;
    int sy, ry;
    if (Y < 0) sy = 0, ry = 8 + Y;
    else if (Y >= 8 * TJ) sy = 0, ry = Y - 8 * TJ;
    else sy = 1 + (Y >> 3), ry = Y & 7;
    return slotoff(sx, sy) + ry * 64 + cx * 8;
  }
};

template <class C>
__global__ void __launch_bounds__(C::NT, C::MINB) k_star2(const __grid_constant__ TiledArgs a, const __grid_constant__ bk::StarCoef cf) {
  constexpr int R = C::R, YT = C::YT, TI = C::TI, TJ = C::TJ, D = C::D, W = C::W, H = C::H;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *ring = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127);
  unsigned char *mid = ring + D * C::STAGE;  // M intermediate planes, same slot layout as an input stage
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar_full = ring_u32 + (D + C::M) * C::STAGE;
  const uint32_t bar_empty = bar_full + D * 8;

  const double *fin = a.in;
  double *fout = a.out;
  size_t in_step = a.in_step, out_step = a.out_step;
  if (a.multi) {
    const bk_field_dev f = a.multi[blockIdx.z];
    fin = f.in, fout = f.out, in_step = f.in_step, out_step = f.out_step;
  }
  const int tid = threadIdx.x;
  int bq = 0, brel = (int) blockIdx.x;
  while (bq + 1 < a.nbox && brel >= a.box[bq + 1].first) ++bq;
  brel -= a.box[bq].first;
  const int tx = a.box[bq].lo[0] + brel % a.box[bq].dim[0];
  brel /= a.box[bq].dim[0];
  const int ty = a.box[bq].lo[1] + brel % a.box[bq].dim[1];
  const int tseg = a.box[bq].lo[2] + brel / a.box[bq].dim[1];
  const int i0 = a.lo[0] + tx * TI, j0 = a.lo[1] + ty * TJ;
  int kb0, nl;
  seg_range(a, tseg, kb0, nl);
  const int nout = nl * 8;
  const int P = nout + 2 * H;  // input planes streamed: relative planes -H .. nout+H-1

  if (tid == 0) {
    for (int s = 0; s < D; ++s) {
      mbar_init(bar_full + 8 * s, 32 * C::NPW);
      mbar_init(bar_empty + 8 * s, 1);  // one consumer arrives after the consumers' plane barrier
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if constexpr (C::CREG > 0) {
    if (tid >= C::NCONS)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::PREG));
    else
      asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::CREG));
  }
  if (tid >= C::NCONS) {
    // ================================================ producer warps ===========================================
    // (a bulk copy takes its operands from uniform registers: the per-lane copies of a warp are issued one after the
    // other, ~35 cycles each, so the copy rate scales with the number of producer WARPS)
    const int pw = (tid - C::NCONS) >> 5, lane = tid & 31;
    int sbi[C::JOBS], sbj[C::JOBS], kind[C::JOBS];
    uint32_t dsto[C::JOBS];
    unsigned idn[C::JOBS], idc[C::JOBS];
#pragma unroll
    for (int q = 0; q < C::JOBS; ++q) {
      int job = pw + C::NPW * (lane + 32 * q);
      kind[q] = 0;
      sbi[q] = sbj[q] = 0;
      if (job < TI * TJ) {
        kind[q] = 1, sbi[q] = 1 + job % TI, sbj[q] = 1 + job / TI;
      } else if ((job -= TI * TJ) < 2 * TJ) {
        kind[q] = 1, sbi[q] = (job & 1) ? TI + 1 : 0, sbj[q] = 1 + (job >> 1);
      } else if ((job -= 2 * TJ) < 2 * (TI + 2)) {
        kind[q] = 2 + (job & 1), sbi[q] = job >> 1, sbj[q] = (job & 1) ? TJ + 1 : 0;
      }
      dsto[q] = C::slotoff(sbi[q], kind[q] >= 2 ? 0 : sbj[q]) + (kind[q] == 2 ? (8 - H) * 64 : 0);
    }
    auto brick_id = [&](int q, int kb) -> unsigned {
      const int gi = i0 + sbi[q] - 1, gj = j0 + sbj[q] - 1;
      if (kind[q] == 0 || gi < 0 || gi >= a.gx || gj < 0 || gj >= a.gy || kb < 0 || kb >= a.gz) return 0u;
      return __ldg(a.grid + ((size_t) kb * a.gy + gj) * a.gx + gi);
    };
    const int z_first = kb0 * 8 - H;
    int kb = (z_first >= 0) ? z_first / 8 : -((7 - z_first) / 8);
    int pz = z_first - kb * 8;
#pragma unroll
    for (int q = 0; q < C::JOBS; ++q) idn[q] = brick_id(q, kb);
    int st = 0;
    uint32_t ph = 0;
    bool fresh = true;
    uint32_t bytes = 0;
#pragma unroll
    for (int q = 0; q < C::JOBS; ++q) bytes += kind[q] == 0 ? 0 : kind[q] == 1 ? 512 : H * 64;
    for (int n = 0; n < P; ++n) {
      if (fresh) {
#pragma unroll
        for (int q = 0; q < C::JOBS; ++q) idc[q] = idn[q], idn[q] = brick_id(q, kb + 1);
        fresh = false;
      }
      if (n >= D) mbar_wait(bar_empty + 8 * st, ph ^ 1);
      const uint32_t fb = bar_full + 8 * st;
      mbar_expect_tx(fb, bytes);
      const uint32_t sb = ring_u32 + st * C::STAGE;
#pragma unroll
      for (int q = 0; q < C::JOBS; ++q) {
        const double *src = fin + (size_t) idc[q] * in_step + pz * 64;
        if (kind[q] == 1) bulk_g2s(sb + dsto[q], src, 512, fb);
        else if (kind[q] >= 2) bulk_g2s(sb + dsto[q], src + (kind[q] == 2 ? (8 - H) * 8 : 0), H * 64, fb);
      }
      if (++pz == 8) pz = 0, ++kb, fresh = true;
      if (++st == D) st = 0, ph ^= 1;
    }
    return;
  }

  // ================================================== consumers ================================================
  const int c = tid & 3, e = (tid >> 2) & 1;
  int rest = tid >> 3;
  const bool swp = (R % 2 == 1) && (rest & 1);  // odd row group of the half warp (see the i-halo loads)
  const int y0 = (rest % (8 / YT)) * YT;
  rest /= (8 / YT);
  const int bi = (rest % (TI / 2)) * 2 + e, bj = rest / (TI / 2);
  const int own_slot = C::slotoff(bi + 1, bj + 1);
  const int own_off = own_slot + y0 * 64 + c * 16;
  int joff[2 * R];
#pragma unroll
  for (int h = 0; h < 2 * R; ++h) {
    const int ya = (h < R) ? y0 - R + h : y0 + YT + (h - R);
    int base;
    if (ya < 0) base = C::slotoff(bi + 1, bj) + (8 + ya) * 64;
    else if (ya >= 8) base = C::slotoff(bi + 1, (bj == TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
    else base = own_slot + ya * 64;
    joff[h] = base + c * 16;
  }
  const int dl = C::slotoff(bi, bj + 1) - own_slot, dr = C::slotoff(bi + 2, bj + 1) - own_slot;
  constexpr int NI = (R % 2 == 0) ? R / 2 : R;
  int ioffL[NI], ioffR[NI];
#pragma unroll
  for (int m = 1; m <= NI; ++m) {
    if (R % 2 == 0) {
      ioffL[m - 1] = -16 * m + ((c - m < 0) ? dl + 64 : 0);
      ioffR[m - 1] = 16 * m + ((c + m > 3) ? dr - 64 : 0);
    } else {
      ioffL[m - 1] = -8 * m + ((2 * c - m < 0) ? dl + 64 : 0);
      ioffR[m - 1] = 8 * (1 + m) + ((2 * c + 1 + m > 7) ? dr - 64 : 0);
      if (swp) {
        const int t = ioffL[m - 1];
        ioffL[m - 1] = ioffR[m - 1], ioffR[m - 1] = t;
      }
    }
  }
  // my strip point (intermediate cell outside the tile that the second step reads), if any
  // every warp takes an equal share of the strip: NI lanes on the two i-strip columns (a column maps to two bank groups
  // only, so few lanes per instruction), NJ lanes on the j-strip rows
  constexpr int SNI = 2 * R * 8 * TJ / C::NCW, SNJ = 2 * R * 8 * TI / C::NCW;
  static_assert(SNI * C::NCW == 2 * R * 8 * TJ && SNJ * C::NCW == 2 * R * 8 * TI && SNI % 2 == 0 && SNI + SNJ <= 32, "strip split");
  const int wrp = tid >> 5, ln = tid & 31;
  const bool has_strip = ln < SNI + SNJ;
  int sX = 0, sY = 0;
  if (ln < SNI) {
    const int side = ln >= SNI / 2, idx = wrp * (SNI / 2) + (ln - side * (SNI / 2));  // idx in [0, R*8TJ)
    sY = idx % (8 * TJ);
    sX = side ? 8 * TI + idx / (8 * TJ) : -1 - idx / (8 * TJ);
  } else if (has_strip) {
    const int q = wrp * SNJ + (ln - SNI), side = q / (R * 8 * TI), rem = q % (R * 8 * TI);
    sX = rem % (8 * TI);
    sY = side ? 8 * TJ + rem / (8 * TI) : -1 - rem / (8 * TI);
  }
  const int soff = C::cell(sX, sY);
  int sxm[R], sxp[R], sym[R], syp[R];
#pragma unroll
  for (int d = 1; d <= R; ++d) {
    sxm[d - 1] = C::cell(sX - d, sY), sxp[d - 1] = C::cell(sX + d, sY);
    sym[d - 1] = C::cell(sX, sY - d), syp[d - 1] = C::cell(sX, sY + d);
  }
  bool strip_in;  // is my strip point's brick column inside the grid?
  {
    const int gi = i0 + ((sX + 8) >> 3) - 1, gj = j0 + (sY < 0 ? -1 : sY >= 8 * TJ ? TJ : (sY >> 3));
    strip_in = has_strip && gi >= 0 && gi < a.gx && gj >= 0 && gj < a.gy;
  }
  const bool own_in = (i0 + bi < a.gx) && (j0 + bj < a.gy);
  const bool mine = (i0 + bi < a.hi[0]) && (j0 + bj < a.hi[1]);
  const unsigned *gcol = a.grid + ((size_t) kb0 * a.gy + (j0 + bj)) * a.gx + (i0 + bi);
  const size_t glayer = (size_t) a.gy * a.gx;
  unsigned id_next = mine ? __ldg(gcol) : 0u;
  double *outp = fout;

  double2 accA[W][YT], accB[W][YT];
  double accS[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    accS[w] = 0.0;
#pragma unroll
    for (int r = 0; r < YT; ++r) accA[w][r] = accB[w][r] = make_double2(0.0, 0.0);
  }

  // one plane of the star update on my patch: `pb` = plane base, slot u%W of `acc` belongs to this plane's own output;
  // `v` = my own cells of the plane (stage A loads them, stage B takes them from stage A's registers); on return slot
  // (u-R)%W holds the plane this input completes (output index t-R)
  auto star_plane = [&](const unsigned char *pb, double2 (&acc)[W][YT], const int u, const double2 (&v)[YT]) {
    const int sF = ((u - R) % W + W) % W, s0 = u % W, sN = (u + R) % W;
#pragma unroll
    for (int r = 0; r < YT; ++r) {
      acc[sF][r].x = fma(cf.cp[2][R - 1], v[r].x, acc[sF][r].x);
      acc[sF][r].y = fma(cf.cp[2][R - 1], v[r].y, acc[sF][r].y);
    }
#pragma unroll
    for (int d = R - 1; d >= 1; --d) {
      const int s = ((u - d) % W + W) % W;
#pragma unroll
      for (int r = 0; r < YT; ++r) {
        acc[s][r].x = fma(cf.cp[2][d - 1], v[r].x, acc[s][r].x);
        acc[s][r].y = fma(cf.cp[2][d - 1], v[r].y, acc[s][r].y);
      }
    }
    double2 rows[YT + 2 * R];
#pragma unroll
    for (int h = 0; h < R; ++h) rows[h] = *reinterpret_cast<const double2 *>(pb + joff[h]);
#pragma unroll
    for (int r = 0; r < YT; ++r) rows[R + r] = v[r];
#pragma unroll
    for (int h = 0; h < R; ++h) rows[R + YT + h] = *reinterpret_cast<const double2 *>(pb + joff[R + h]);
#pragma unroll
    for (int r = 0; r < YT; ++r) {
      double line[2 * R + 2];
      line[R] = v[r].x, line[R + 1] = v[r].y;
      const unsigned char *pr = pb + own_off + r * 64;
      if constexpr (R % 2 == 0) {
#pragma unroll
        for (int m = 1; m <= R / 2; ++m) {
          const double2 lft = *reinterpret_cast<const double2 *>(pr + ioffL[m - 1]);
          const double2 rgt = *reinterpret_cast<const double2 *>(pr + ioffR[m - 1]);
          line[R - 2 * m] = lft.x, line[R - 2 * m + 1] = lft.y;
          line[R + 2 * m] = rgt.x, line[R + 2 * m + 1] = rgt.y;
        }
      } else {
#pragma unroll
        for (int m = 1; m <= R; ++m) {  // odd row groups load right first: a half warp then covers all 32 banks
          if constexpr (C::ABL & 2) {
            line[R - m] = v[r].y, line[R + 1 + m] = v[r].x;
            continue;
          }
          const double q0 = *reinterpret_cast<const double *>(pr + ioffL[m - 1]);
          const double q1 = *reinterpret_cast<const double *>(pr + ioffR[m - 1]);
          line[R - m] = swp ? q1 : q0;
          line[R + 1 + m] = swp ? q0 : q1;
        }
      }
      double ax = fma(cf.c0, v[r].x, acc[s0][r].x), ay = fma(cf.c0, v[r].y, acc[s0][r].y);
#pragma unroll
      for (int d = 1; d <= R; ++d) {
        ax = fma(cf.cp[0][d - 1], line[R + d], ax);
        ay = fma(cf.cp[0][d - 1], line[R + 1 + d], ay);
        ax = fma(cf.cm[0][d - 1], line[R - d], ax);
        ay = fma(cf.cm[0][d - 1], line[R + 1 - d], ay);
      }
#pragma unroll
      for (int d = 1; d <= R; ++d) {
        ax = fma(cf.cp[1][d - 1], rows[R + r + d].x, ax);
        ay = fma(cf.cp[1][d - 1], rows[R + r + d].y, ay);
        ax = fma(cf.cm[1][d - 1], rows[R + r - d].x, ax);
        ay = fma(cf.cm[1][d - 1], rows[R + r - d].y, ay);
      }
      acc[s0][r].x = ax, acc[s0][r].y = ay;
    }
#pragma unroll
    for (int d = 1; d <= R - 1; ++d) {
      const int s = (u + d) % W;
#pragma unroll
      for (int r = 0; r < YT; ++r) {
        acc[s][r].x = fma(cf.cm[2][d - 1], v[r].x, acc[s][r].x);
        acc[s][r].y = fma(cf.cm[2][d - 1], v[r].y, acc[s][r].y);
      }
    }
#pragma unroll
    for (int r = 0; r < YT; ++r) {
      acc[sN][r].x = cf.cm[2][R - 1] * v[r].x;
      acc[sN][r].y = cf.cm[2][R - 1] * v[r].y;
    }
  };

  int st = 0, msl = 0, mprev = C::M - 1;
  uint32_t ph = 0;
  const int zabs0 = kb0 * 8, zmax = a.gz * 8;
  // iteration n: stage B consumes intermediate plane n-1 (written in the previous iteration, published by the plane
  // barrier) while stage A produces intermediate plane n from input plane n -- the two are independent, so their
  // shared-memory latencies overlap; one CTA barrier per plane
#pragma unroll 1
  for (int tb = 0; tb <= P; tb += W) {
#pragma unroll
    for (int u = 0; u < W; ++u) {
      const int n = tb + u;
      if (n <= P) {
        double2 vmid[YT];  // my own cells of the intermediate plane produced in this iteration
        auto stage_b = [&](const int nb, const int slot, const int uB) {  // intermediate plane nb-H-R -> output nb-2H
          const unsigned char *pmB = mid + slot * C::STAGE;
          const int orel = nb - 2 * H;
          if constexpr (C::LAG) {
#pragma unroll
            for (int r = 0; r < YT; ++r) vmid[r] = *reinterpret_cast<const double2 *>(pmB + own_off + r * 64);
          }
          star_plane(pmB, accB, uB, vmid);
          if (orel >= 0) {
            const int sF = ((uB - R) % W + W) % W;
            const int oz = orel & 7;
            if (oz == 0) {
              outp = fout + (size_t) id_next * out_step + y0 * 8 + c * 2;
              if (mine && (orel >> 3) + 1 < nl) id_next = __ldg(gcol + ((orel >> 3) + 1) * glayer);
            }
            if (mine) {
#pragma unroll
              for (int r = 0; r < YT; ++r) *reinterpret_cast<double2 *>(outp + oz * 64 + r * 8) = accB[sF][r];
            }
          }
        };
        if (C::LAG && n - 1 >= H) stage_b(n - 1, mprev, u % W);  // (n - 1 - H) mod W with H = W - 1
        // ---- stage A: input plane n-H completes intermediate plane n-H-R ------------------------------------------
        if (n < P) {
          if (!C::EW || n == 0) mbar_wait(bar_full + 8 * st, ph);
          const unsigned char *pbA = ring + st * C::STAGE;
          unsigned char *pm = mid + msl * C::STAGE;
          const int zm = zabs0 + n - H - R;
          const bool zin = zm >= 0 && zm < zmax;
          {
            double2 vin[YT];
#pragma unroll
            for (int r = 0; r < YT; ++r) vin[r] = *reinterpret_cast<const double2 *>(pbA + own_off + r * 64);
            star_plane(pbA, accA, u, vin);
            const int sF = ((u - R) % W + W) % W;
            // cells outside the grid hold zero (null-brick semantics): only boundary tiles / planes ever mask, so the
            // common case takes a warp-uniform branch around the selects
            const bool ok = (C::ABL & 4) ? true : (zin && own_in);
            if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
              for (int r = 0; r < YT; ++r) vmid[r] = accA[sF][r];
            } else {
#pragma unroll
              for (int r = 0; r < YT; ++r) vmid[r] = ok ? accA[sF][r] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int r = 0; r < YT; ++r)
              if (!(C::ABL & 8)) *reinterpret_cast<double2 *>(pm + own_off + r * 64) = vmid[r];
          }
          if (has_strip && !(C::ABL & 1)) {
            const int sF = ((u - R) % W + W) % W, s0 = u % W, sN = (u + R) % W;
            const double sv = *reinterpret_cast<const double *>(pbA + soff);
            accS[sF] = fma(cf.cp[2][R - 1], sv, accS[sF]);
            *reinterpret_cast<double *>(pm + soff) = (zin && strip_in) ? accS[sF] : 0.0;
#pragma unroll
            for (int d = R - 1; d >= 1; --d)
              accS[((u - d) % W + W) % W] = fma(cf.cp[2][d - 1], sv, accS[((u - d) % W + W) % W]);
            double t = fma(cf.c0, sv, accS[s0]);
#pragma unroll
            for (int d = 1; d <= R; ++d) {
              t = fma(cf.cp[0][d - 1], *reinterpret_cast<const double *>(pbA + sxp[d - 1]), t);
              t = fma(cf.cm[0][d - 1], *reinterpret_cast<const double *>(pbA + sxm[d - 1]), t);
            }
#pragma unroll
            for (int d = 1; d <= R; ++d) {
              t = fma(cf.cp[1][d - 1], *reinterpret_cast<const double *>(pbA + syp[d - 1]), t);
              t = fma(cf.cm[1][d - 1], *reinterpret_cast<const double *>(pbA + sym[d - 1]), t);
            }
            accS[s0] = t;
#pragma unroll
            for (int d = 1; d <= R - 1; ++d) accS[(u + d) % W] = fma(cf.cm[2][d - 1], sv, accS[(u + d) % W]);
            accS[sN] = cf.cm[2][R - 1] * sv;
          }
        }
        // every consumer has read the input stage / the previous intermediate plane and written its part of the new one
        asm volatile("bar.sync 1, %0;" ::"n"(C::NCONS) : "memory");
        if (n < P) {
          if (tid == 0) mbar_arrive(bar_empty + 8 * st);
          if (++st == D) st = 0, ph ^= 1;
          // early wait: the next plane's data is awaited here, so that stage B of this plane and stage A of the next one
          // form one straight-line block whose shared-memory latencies overlap
          if (C::EW && n + 1 < P) mbar_wait(bar_full + 8 * st, ph);
          if (!C::LAG && n >= H) stage_b(n, msl, (u + 1) % W);  // (n - H) mod W
        }
        mprev = msl;
        msl = (msl == C::M - 1) ? 0 : msl + 1;
      }
    }
  }
}

// register re-balancing (Cfg::CREG > 0): ptxas needs the register count at entry, i.e. min-blocks in the launch bounds
template <class C>
__global__ void __launch_bounds__(C::NT, 1) k_star_rebal(const __grid_constant__ TiledArgs a,
                                                         const __grid_constant__ typename C::Coef cf) {
  march_body<C>(a, cf);
}

#endif  // !BK_REMOTE_TU

// Brick layers per k segment.  Every CTA streams 8*layers + ovh planes (ovh = halo planes + pipeline fill) and CTAs are
// dealt to the `slots` resident CTA slots in blockIdx order (all tiles of segment 0, then segment 1, ...; the last
// segment may be shorter).  Long segments amortise ovh, short ones fill the last wave: the launch is list-scheduled
// here for every candidate length and the shortest makespan wins (model checked against profiles/r01_segments.md and
// the fused-kernel sweeps).  Results are cached per (tiles, layers, slots, ovh).
int pick_segment_layers(long tiles, int nz, int slots, int ovh) {
  static std::mutex mu;
  static std::map<std::tuple<long, int, int, int>, int> cache;
  const auto key = std::make_tuple(tiles, nz, slots, ovh);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
  }
  int best_kl = nz;
  double best = -1.0;
  int last_segs = -1;
  for (int kl = nz; kl >= 1; --kl) {
    const int segs = (nz + kl - 1) / kl;
    if (segs == last_segs) continue;  // same segment count with a longer segment: never better
    last_segs = segs;
    if ((double) tiles * segs > 64.0 * slots && best >= 0) break;  // far too many tiny CTAs
    const int kl_min = (nz + segs - 1) / segs;                     // most even split for this segment count
    std::priority_queue<double, std::vector<double>, std::greater<double>> free_at;
    for (int i = 0; i < slots; ++i) free_at.push(0.0);
    double makespan = 0.0;
    for (int q = 0; q < segs; ++q) {
      const int len = std::min(kl_min, nz - q * kl_min);
      const double d = 8.0 * len + ovh;
      for (long t = 0; t < tiles; ++t) {
        const double start = free_at.top();
        free_at.pop();
        free_at.push(start + d);
        makespan = std::max(makespan, start + d);
      }
    }
    // CTAs of one wave do not finish together on real hardware: charge a fraction of a CTA time for the ragged tail
    const double cost = makespan + 0.15 * (8.0 * kl_min + ovh);
    if (best < 0 || cost < best) best = cost, best_kl = kl_min;
  }
  std::lock_guard<std::mutex> lk(mu);
  cache[key] = best_kl;
  return best_kl;
}

// The launch of a marching kernel of tile TI x TJ bricks: k segments (cost model above), the CTA boxes of a whole or
// split launch, then `launch(grid, args)`.  Shared by the compiled-in kernels (launch_cfg) and the generated ones.
struct Geom {
  int TI, TJ, OVH;
};
template <class LaunchFn>
int launch_geom(const Geom &g, int slots, const TiledArgs &a0, unsigned nsub, int part, const int *rdy_lo, const int *rdy_hi,
                LaunchFn &&launch) {
  TiledArgs a = a0;
  const int nx = a.hi[0] - a.lo[0], ny = a.hi[1] - a.lo[1], nz = a.hi[2] - a.lo[2];
  if (nx <= 0 || ny <= 0 || nz <= 0) return BK_OK;
  a.ntx = (nx + g.TI - 1) / g.TI;
  const int nty = (ny + g.TJ - 1) / g.TJ;
  // split launches: the layers whose CTAs read not-yet-ready bricks get thin segments of their own, so that the READY
  // part (which overlaps the exchange) is as large as possible
  a.kh = a.kt = 0;
  const bool thin = (part & BK_PART_THIN) != 0;
  part &= ~BK_PART_THIN;
  if (part != BK_PART_ALL && thin) {
    a.kh = std::max(0, std::min(nz, rdy_lo[2] + 1 - a.lo[2]));
    a.kt = std::max(0, std::min(nz - a.kh, a.hi[2] - (rdy_hi[2] - 1)));
    if (nz - a.kh - a.kt <= 0) a.kh = a.kt = 0;
  }
  const int nmid = nz - a.kh - a.kt;
  // k segments: every CTA streams kl*8 + 2*RUP planes and the launch takes ceil(CTAs / resident slots) rounds, so pick
  // the segment count that minimises rounds * planes (long segments amortise the halo planes, short ones fill the
  // last round)
  a.kl = pick_segment_layers((long) a.ntx * nty * nsub, nmid, slots, g.OVH);
  if (const char *e = getenv("BK_STAR_KL")) a.kl = atoi(e) > 0 ? atoi(e) : a.kl;  // developer knob
  const int segs = (nmid + a.kl - 1) / a.kl + (a.kh > 0) + (a.kt > 0);
  // CTA boxes.  "inner" = CTAs whose whole read footprint (tile + 1 brick all round) lies in the ready box
  int in_lo[3] = {0, 0, 0}, in_hi[3] = {0, 0, 0};
  const int ext[3] = {a.ntx, nty, segs};
  if (part != BK_PART_ALL) {
    const int T[2] = {g.TI, g.TJ};
    for (int d = 0; d < 2; ++d) {
      int l = 0, h = ext[d];
      while (l < h && a.lo[d] + l * T[d] - 1 < rdy_lo[d]) ++l;
      while (h > l && a.lo[d] + (h - 1) * T[d] + T[d] + 1 > rdy_hi[d]) --h;
      in_lo[d] = l, in_hi[d] = h;
    }
    int l = 0, h = segs;
    auto seg_lo = [&](int q) { int b, n; seg_range(a, q, b, n); return b; };
    auto seg_end = [&](int q) { int b, n; seg_range(a, q, b, n); return b + n; };
    while (l < h && seg_lo(l) - 1 < rdy_lo[2]) ++l;
    while (h > l && seg_end(h - 1) + 1 > rdy_hi[2]) --h;
    in_lo[2] = l, in_hi[2] = h;
  }
  a.nbox = 0;
  int first = 0;
  auto add_box = [&](int x0, int x1, int y0, int y1, int z0, int z1) {
    if (x1 <= x0 || y1 <= y0 || z1 <= z0) return;
    TiledArgs::Box &b = a.box[a.nbox++];
    b.lo[0] = x0, b.lo[1] = y0, b.lo[2] = z0, b.dim[0] = x1 - x0, b.dim[1] = y1 - y0, b.dim[2] = z1 - z0;
    b.first = first;
    first += b.dim[0] * b.dim[1] * b.dim[2];
  };
  const bool has_inner = in_hi[0] > in_lo[0] && in_hi[1] > in_lo[1] && in_hi[2] > in_lo[2];
  if (part == BK_PART_ALL) {
    add_box(0, ext[0], 0, ext[1], 0, ext[2]);
  } else if (part == BK_PART_READY) {
    if (has_inner) add_box(in_lo[0], in_hi[0], in_lo[1], in_hi[1], in_lo[2], in_hi[2]);
  } else if (!has_inner) {
    add_box(0, ext[0], 0, ext[1], 0, ext[2]);
  } else {
    // long CTAs first (blockIdx order = dispatch order): the thin head / tail segments then fill the ragged end
    add_box(0, ext[0], 0, in_lo[1], in_lo[2], in_hi[2]);                 // tile rows before / after
    add_box(0, ext[0], in_hi[1], ext[1], in_lo[2], in_hi[2]);
    add_box(0, in_lo[0], in_lo[1], in_hi[1], in_lo[2], in_hi[2]);        // tile columns left / right
    add_box(in_hi[0], ext[0], in_lo[1], in_hi[1], in_lo[2], in_hi[2]);
    add_box(0, ext[0], 0, ext[1], 0, in_lo[2]);                          // segments below / above
    add_box(0, ext[0], 0, ext[1], in_hi[2], ext[2]);
  }
  if (first == 0) return BK_OK;
  return launch(dim3((unsigned) first, 1, nsub), a);
}

int device_slots(const void *kern, int threads, size_t smem) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  return sms * (per_sm > 0 ? per_sm : 1);
}

#ifndef BK_REMOTE_TU
template <class C>
int launch_cfg(const TiledArgs &a0, const typename C::Coef &cf, cudaStream_t s, unsigned nsub, int part,
               const int *rdy_lo, const int *rdy_hi) {
  void (*kern)(const TiledArgs, const typename C::Coef);
  if constexpr (C::FUSED) kern = k_star2<C>;
  else if constexpr (C::CREG > 0) kern = k_star_rebal<C>;
  else if constexpr (C::MAXREG < 255) kern = k_star_capped<C>;
  else kern = k_star<C>;
  BK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM));
  if (getenv("BK_DEBUG")) {
    cudaFuncAttributes fa;
    int nb = -1;
    cudaFuncGetAttributes(&fa, kern);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, C::NT, C::SMEM);
    fprintf(stderr, "[bk] %s R=%d YT=%d tile=%dx%d D=%d: %d thr, %d regs, %zu B smem, %zu B local, %d CTA/SM\n",
            C::FUSED ? "k_star2" : "k_star", C::R, C::YT, C::TI, C::TJ, C::D, C::NT, fa.numRegs, C::SMEM, fa.localSizeBytes, nb);
  }
  static std::atomic<int> slots_cached{0};  // several rank threads (drivers/*) may get here at once: benign double init
  int slots = slots_cached.load(std::memory_order_relaxed);
  if (!slots) {
    slots = device_slots((const void *) kern, C::NT, C::SMEM);
    slots_cached.store(slots, std::memory_order_relaxed);
  }
  const Geom g = {C::TI, C::TJ, C::OVH};
  return launch_geom(g, slots, a0, nsub, part, rdy_lo, rdy_hi, [&](dim3 grid, const TiledArgs &a) -> int {
    kern<<<grid, C::NT, C::SMEM, s>>>(a, cf);
    BK_LAUNCHED();
    return BK_OK;
  });
}

std::atomic<int> g_fused_variant{-1};  // -1: not initialised (the environment decides at first use)
int fused_variant() {
  int v = g_fused_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = getenv("BK_FUSED_VARIANT");
    v = BK_FUSED_STAGED;  // "staged" | "composed" | "wide" (or 0 | 1 | 2)
    if (e && (e[0] == 'c' || e[0] == 'C' || e[0] == '1')) v = BK_FUSED_COMPOSED;
    if (e && (e[0] == 'w' || e[0] == 'W' || e[0] == '2')) v = BK_FUSED_COMPOSED_WIDE;
    int expect = -1;
    if (!g_fused_variant.compare_exchange_strong(expect, v)) v = expect;
  }
  return v;
}

}  // namespace

extern "C" {
int bk_stencil_fused_variant_set(int variant) {
  BK_REQUIRE(variant == BK_FUSED_STAGED || variant == BK_FUSED_COMPOSED || variant == BK_FUSED_COMPOSED_WIDE,
             "BK_FUSED_STAGED, BK_FUSED_COMPOSED or BK_FUSED_COMPOSED_WIDE");
  const int before = fused_variant();
  g_fused_variant.store(variant, std::memory_order_relaxed);
  return before;
}
int bk_stencil_fused_variant_get(void) { return fused_variant(); }
}

namespace bk {
int fused_variant_now() { return fused_variant(); }  // for bk_stencil_remote.cu
}

namespace bk {

int launch_tiled(const CoefSpec &spec, const bk_field_t &f, const bk_field_t *multi_dev, unsigned nsub, const unsigned *grid,
                 const unsigned *gdims, const unsigned *lo, const unsigned *hi, cudaStream_t s, int part,
                 const unsigned *ready_lo, const unsigned *ready_hi, int steps) {
  if (!multi_dev && (((size_t) f.in | (size_t) f.out) & 15 || (f.in_step & 1) || (f.out_step & 1))) return BK_EUNSUPPORTED;
  if (steps == 2 && (spec.kind != 0 || spec.radius > 2)) return BK_EUNSUPPORTED;
  TiledArgs a;
  a.in = f.in, a.out = f.out, a.in_step = f.in_step, a.out_step = f.out_step, a.grid = grid;
  a.gx = (int) gdims[0], a.gy = (int) gdims[1], a.gz = (int) gdims[2];
  for (int d = 0; d < 3; ++d) a.lo[d] = (int) lo[d], a.hi[d] = (int) hi[d];
  a.ntx = a.kl = a.kh = a.kt = 0;
  a.multi = reinterpret_cast<const bk_field_dev *>(multi_dev);
  a.nbox = 0;
  int rdy_lo[3] = {0, 0, 0}, rdy_hi[3] = {0, 0, 0};
  if (part != BK_PART_ALL)
    for (int d = 0; d < 3; ++d) rdy_lo[d] = (int) ready_lo[d], rdy_hi[d] = (int) ready_hi[d];
  int v = 0;
  if (const char *e = getenv("BK_STAR_VARIANT")) v = atoi(e);  // developer knob: one alternative geometry per stencil
  // Geometries (R, YT, TI, TJ, G, D, register cap, producer warps), chosen on B200 -- see DESIGN.md section 4:
  //   radius 1/2: 4x4-brick tiles, 4 consumer warps + 2 producer warps, 2 CTAs per SM (<= 128 registers)
  //   radius 4  : 6x4-brick tiles, 2 rows per thread (12 consumer warps at 152 registers) + 4 producer warps (40
  //               registers, setmaxnreg), 3-stage ring (130 KB: leaves L1 for the id/adjacency reads), 1 CTA per SM
  //   cube      : same shape: 6x4-brick tiles, 12 consumer warps (152 registers) + 4 producer warps, 1 CTA per SM
  if (spec.kind == 1) {
    const CubeCoef &cc = spec.cc;
    if (v == 1) return launch_cfg<Cfg<2, 2, 6, 4, 2, 3, 128, 4, true>>(a, cc, s, nsub, part, rdy_lo, rdy_hi);
    return launch_cfg<Cfg<2, 2, 6, 4, 2, 3, 255, 4, true, 152, 40>>(a, cc, s, nsub, part, rdy_lo, rdy_hi);
  }
  const StarCoef &sc = spec.sc;
  const int r = spec.radius;
  if (steps == 2) {  // two time steps per pass: (R, YT, TI, TJ, D, producer warps)
    if (r == 1 && fused_variant() != BK_FUSED_STAGED) {
      // the composed operator on the radius-2 star geometry.  BK_FUSED_COMPOSED: 4x4-brick tiles, 4 consumer + 2 producer
      // warps, 2 CTAs per SM (the 13-point kernel's shape); BK_FUSED_COMPOSED_WIDE: 8x4-brick tiles, 8 consumer + 4
      // producer warps with register re-balancing, 1 CTA per SM (the 25-point kernel's shape: fewer halo bricks per point)
      const bk::DiamondCoef dc = bk::diamond_coef(sc.c0, sc.cp[0][0], sc.cm[0][0], sc.cp[1][0], sc.cm[1][0], sc.cp[2][0], sc.cm[2][0]);
      if (v == 1) return launch_cfg<Cfg<2, 2, 4, 4, 2, 3, 128, 2, false, 0, 40, 2>>(a, dc, s, nsub, part, rdy_lo, rdy_hi);
      if (fused_variant() == BK_FUSED_COMPOSED_WIDE)
        return launch_cfg<Cfg<2, 4, 8, 4, 2, 3, 255, 4, false, 232, 40, 2>>(a, dc, s, nsub, part, rdy_lo, rdy_hi);
      return launch_cfg<Cfg<2, 4, 4, 4, 2, 3, 168, 2, false, 0, 40, 2>>(a, dc, s, nsub, part, rdy_lo, rdy_hi);
    }
    if (r == 1) {
      if (v == 1) return launch_cfg<FCfg<1, 4, 8, 4, 4, 4>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 2) return launch_cfg<FCfg<1, 2, 4, 4, 3, 2, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 3) return launch_cfg<FCfg<1, 4, 8, 4, 4, 4, 1, 0, 40, true>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 4) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, true>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 5) return launch_cfg<FCfg<1, 4, 4, 4, 4, 2, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);                    // deeper ring
      if (v == 6) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, true>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);  // early wait
      if (v == 7) return launch_cfg<FCfg<1, 4, 4, 4, 4, 2, 2, 0, 40, false, true>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);  // both
      if (v == 11) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 1>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);   // ablations
      if (v == 12) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 13) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 4>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 14) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 8>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 15) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 15>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      if (v == 16) return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2, 0, 40, false, false, 3>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
      return launch_cfg<FCfg<1, 4, 4, 4, 3, 2, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
    }
    if (r == 2) return launch_cfg<FCfg<2, 2, 4, 4, 4, 4>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
    return BK_EUNSUPPORTED;
  }
  if (r == 1) {
    if (v == 1) return launch_cfg<Cfg<1, 4, 6, 4, 1, 4, 128, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
    return launch_cfg<Cfg<1, 4, 4, 4, 2, 3, 255, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  }
  if (r == 2) {
    if (v == 1) return launch_cfg<Cfg<2, 4, 6, 4, 1, 4, 128, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
    return launch_cfg<Cfg<2, 4, 4, 4, 2, 3, 128, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  }
  if (v == 1) return launch_cfg<Cfg<4, 2, 4, 4, 2, 5, 255, 4>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  if (v == 5) return launch_cfg<Cfg<4, 2, 6, 4, 2, 3, 255, 4, false, 152, 40>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  if (v == 2) return launch_cfg<Cfg<4, 4, 8, 4, 2, 4, 255, 4, false, 232, 40>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  if (v == 3) return launch_cfg<Cfg<4, 4, 6, 4, 2, 4, 232, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  if (v == 4) return launch_cfg<Cfg<4, 4, 6, 4, 2, 3, 232, 2>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  // radius 4 has two tuned shapes with the same cost per (padded) brick column (profiles/r02_kernels.md): 6x4-brick
  // tiles with 2 rows per thread, and 8x4-brick tiles with 4 rows per thread (56 instead of 72 B of LDS per point).  The
  // box decides: whichever pads it less wins (64 bricks: 8x4 exact vs 66 for 6x4; 66 bricks: 6x4 exact vs 72).
  if (v == 0) {
    const long nx = a.hi[0] - a.lo[0], ny = a.hi[1] - a.lo[1];
    const long pad64 = ((nx + 5) / 6) * 6 * ((ny + 3) / 4) * 4, pad84 = ((nx + 7) / 8) * 8 * ((ny + 3) / 4) * 4;
    if (pad84 <= pad64) return launch_cfg<Cfg<4, 4, 8, 4, 2, 4, 255, 4, false, 232, 40>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
  }
  return launch_cfg<Cfg<4, 2, 6, 4, 2, 3, 255, 4, false, 152, 40>>(a, sc, s, nsub, part, rdy_lo, rdy_hi);
}


// A kernel generated by bk_codegen.cu (cudaKernel_t from cudaLibraryGetKernel): same launch descriptor, same CTA
// enumeration, coefficients passed by value as the kernel's second parameter (`coef`, `coef_bytes`).
int launch_generated(const void *kernel, const GenGeom &gg, const void *coef, const bk_field_t &f, const unsigned *grid,
                     const unsigned *gdims, const unsigned *lo, const unsigned *hi, cudaStream_t s, int part,
                     const unsigned *ready_lo, const unsigned *ready_hi) {
  if (((size_t) f.in | (size_t) f.out) & 15 || (f.in_step & 1) || (f.out_step & 1)) return BK_EUNSUPPORTED;
  TiledArgs a;
  a.in = f.in, a.out = f.out, a.in_step = f.in_step, a.out_step = f.out_step, a.grid = grid;
  a.gx = (int) gdims[0], a.gy = (int) gdims[1], a.gz = (int) gdims[2];
  for (int d = 0; d < 3; ++d) a.lo[d] = (int) lo[d], a.hi[d] = (int) hi[d];
  a.ntx = a.kl = a.kh = a.kt = 0;
  a.multi = nullptr;
  a.nbox = 0;
  int rdy_lo[3] = {0, 0, 0}, rdy_hi[3] = {0, 0, 0};
  if (part != BK_PART_ALL)
    for (int d = 0; d < 3; ++d) rdy_lo[d] = (int) ready_lo[d], rdy_hi[d] = (int) ready_hi[d];
  BK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gg.smem));
  const int slots = device_slots(kernel, gg.threads, gg.smem);
  const Geom g = {gg.TI, gg.TJ, gg.ovh};
  return launch_geom(g, slots, a, 1, part, rdy_lo, rdy_hi, [&](dim3 cta_grid, const TiledArgs &args) -> int {
    void *params[2] = {const_cast<TiledArgs *>(&args), const_cast<void *>(coef)};
    BK_CUDA(cudaLaunchKernel(kernel, cta_grid, dim3((unsigned) gg.threads), params, gg.smem, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return BK_OK;
  });
}

size_t tiled_args_bytes() { return sizeof(TiledArgs); }

}  // namespace bk
#else  // BK_REMOTE_TU ================================================================================================

template <class C>
int launch_cfg_remote(const TiledArgs &a0, const typename C::Coef &cf, const RemoteArgs &rem, cudaStream_t s, int part,
                      const int *rdy_lo, const int *rdy_hi) {
  void (*kern)(const TiledArgs, const typename C::Coef, const RemoteArgs);
  if constexpr (C::CREG > 0) kern = k_star_rebal_remote<C>;
  else if constexpr (C::MAXREG < 255) kern = k_star_capped_remote<C>;
  else kern = k_star_remote<C>;
  BK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) C::SMEM));
  static std::atomic<int> slots_cached{0};
  int slots = slots_cached.load(std::memory_order_relaxed);
  if (!slots) {
    slots = device_slots((const void *) kern, C::NT, C::SMEM);
    slots_cached.store(slots, std::memory_order_relaxed);
  }
  const Geom g = {C::TI, C::TJ, C::OVH};
  return launch_geom(g, slots, a0, 1, part, rdy_lo, rdy_hi, [&](dim3 grid, const TiledArgs &a) -> int {
    kern<<<grid, C::NT, C::SMEM, s>>>(a, cf, rem);
    BK_LAUNCHED();
    return BK_OK;
  });
}

}  // namespace

namespace bk {

int fused_variant_now();  // bk_stencil_tiled.cu

// The default geometry of every stencil (and of the composed two-step update) with the ghost bricks of the input read in
// place through `remap` (bk_stencil_advance_remote).  The staged two-step kernel and the developer geometries have no such
// variant: BK_EUNSUPPORTED.
int launch_tiled_remote(const CoefSpec &spec, const bk_field_t &f, const unsigned *grid, const unsigned *gdims, const unsigned *lo,
                        const unsigned *hi, cudaStream_t s, int part, const unsigned *ready_lo, const unsigned *ready_hi, int steps,
                        const double *const *remap, unsigned ghost_lo, unsigned ghost_n) {
  if (((size_t) f.in | (size_t) f.out) & 15 || (f.in_step & 1) || (f.out_step & 1)) return BK_EUNSUPPORTED;
  TiledArgs a;
  a.in = f.in, a.out = f.out, a.in_step = f.in_step, a.out_step = f.out_step, a.grid = grid;
  a.gx = (int) gdims[0], a.gy = (int) gdims[1], a.gz = (int) gdims[2];
  for (int d = 0; d < 3; ++d) a.lo[d] = (int) lo[d], a.hi[d] = (int) hi[d];
  a.ntx = a.kl = a.kh = a.kt = 0;
  a.multi = nullptr;
  a.nbox = 0;
  int rdy_lo[3] = {0, 0, 0}, rdy_hi[3] = {0, 0, 0};
  if (part != BK_PART_ALL)
    for (int d = 0; d < 3; ++d) rdy_lo[d] = (int) ready_lo[d], rdy_hi[d] = (int) ready_hi[d];
  const RemoteArgs rem = {remap, ghost_lo, ghost_n};
  if (getenv("BK_STAR_VARIANT") && atoi(getenv("BK_STAR_VARIANT")) != 0) return BK_EUNSUPPORTED;
  if (spec.kind == 1) {
    if (steps != 1) return BK_EUNSUPPORTED;
    return launch_cfg_remote<Cfg<2, 2, 6, 4, 2, 3, 255, 4, true, 152, 40>>(a, spec.cc, rem, s, part, rdy_lo, rdy_hi);
  }
  const StarCoef &sc = spec.sc;
  const int r = spec.radius;
  if (steps == 2) {
    if (r != 1 || fused_variant_now() == BK_FUSED_STAGED) {
      set_error("ghost bricks read in place: the staged two-step kernel has no such variant (BK_FUSED_VARIANT=composed|wide has)");
      return BK_EUNSUPPORTED;
    }
    const bk::DiamondCoef dc = bk::diamond_coef(sc.c0, sc.cp[0][0], sc.cm[0][0], sc.cp[1][0], sc.cm[1][0], sc.cp[2][0], sc.cm[2][0]);
    if (fused_variant_now() == BK_FUSED_COMPOSED_WIDE)
      return launch_cfg_remote<Cfg<2, 4, 8, 4, 2, 3, 255, 4, false, 232, 40, 2>>(a, dc, rem, s, part, rdy_lo, rdy_hi);
    return launch_cfg_remote<Cfg<2, 4, 4, 4, 2, 3, 168, 2, false, 0, 40, 2>>(a, dc, rem, s, part, rdy_lo, rdy_hi);
  }
  if (r == 1) return launch_cfg_remote<Cfg<1, 4, 4, 4, 2, 3, 255, 2>>(a, sc, rem, s, part, rdy_lo, rdy_hi);
  if (r == 2) return launch_cfg_remote<Cfg<2, 4, 4, 4, 2, 3, 128, 2>>(a, sc, rem, s, part, rdy_lo, rdy_hi);
  const long nx = a.hi[0] - a.lo[0], ny = a.hi[1] - a.lo[1];
  const long pad64 = ((nx + 5) / 6) * 6 * ((ny + 3) / 4) * 4, pad84 = ((nx + 7) / 8) * 8 * ((ny + 3) / 4) * 4;
  if (pad84 <= pad64) return launch_cfg_remote<Cfg<4, 4, 8, 4, 2, 4, 255, 4, false, 232, 40>>(a, sc, rem, s, part, rdy_lo, rdy_hi);
  return launch_cfg_remote<Cfg<4, 2, 6, 4, 2, 3, 255, 4, false, 152, 40>>(a, sc, rem, s, part, rdy_lo, rdy_hi);
}

}  // namespace bk
#endif  // BK_REMOTE_TU
