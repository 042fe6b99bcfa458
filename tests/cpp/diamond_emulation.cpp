// diamond_emulation.cpp -- host replay of the composed two-step kernel (bricklib_b200/csrc/bk_diamond.h as used by
// march_body<Cfg<2,4,4,4,2,3,168,2,false,0,40,2>> in bk_stencil_tiled.cu) against a plain two-step reference.
//
// There is no GPU where the CPU tests run, so this program restates, line by line, what surrounds diamond_plane() in the
// kernel -- the producer's copy list into the shared-memory stage (slot layout, j-halo rows, corner columns), the
// consumer's thread mapping (x-pair, row group, brick of the tile, halo offsets), the plane loop with its five rotating
// partial outputs and store_plane() -- and calls the SAME diamond_plane() the kernel calls.  Unloaded shared memory is
// poisoned with NaN, every offset is checked against the stage size.  The reference is bk_stencil_advance(steps = 2) by
// definition: one step over the whole grid, intermediate zero outside the grid, a second step over the box.
//   g++ -O1 -std=c++17 -I bricklib_b200/csrc tests/cpp/diamond_emulation.cpp -o diamond_emulation && ./diamond_emulation
#include "bk_diamond.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace {

// ---- Cfg<2, YT, TI, TJ, G, ...> restated (bk_stencil_tiled.cu: struct Cfg), DIAM: corner columns too ----------------
constexpr int R = 2, W = 2 * R + 1;
template <int YT_, int TI_, int TJ_, int G_>
struct Geo {
  static constexpr int YT = YT_, TI = TI_, TJ = TJ_, G = G_;
  static constexpr int RUP = ((R + G - 1) / G) * G;
  static constexpr int SW = TI + 2, SH = TJ + 1;
  static constexpr int SLOTP = G * 512 + 64;
  static constexpr int STAGE = ((SH * SW * SLOTP + 127) / 128) * 128;
  static constexpr int NCONS = TI * TJ * 32 / YT;
  static constexpr int NJH = TI + 2;
  static constexpr int NCOPY = TI * TJ + 2 * TJ + 2 * NJH;
  static constexpr int slotoff(int bi, int bj) { return (bj * SW + bi) * SLOTP; }
};

struct Grid {
  int gx, gy, gz;
  std::vector<unsigned> id;  // dense ids, brick 0 = null
  unsigned at(int i, int j, int k) const {
    if (i < 0 || i >= gx || j < 0 || j >= gy || k < 0 || k >= gz) return 0u;
    return id[((size_t) k * gy + j) * gx + i];
  }
};

double frand() { return (double) rand() / RAND_MAX; }

int failures = 0;
#define CHECK(cond, ...)                    \
  do {                                      \
    if (!(cond)) {                          \
      if (failures++ < 20) {                \
        printf("CHECK failed: " __VA_ARGS__); \
        printf("\n");                       \
      }                                     \
    }                                       \
  } while (0)

// the producer's fill of one stage (planes n*G .. n*G+G-1 of the CTA's plane sequence), bk_stencil_tiled.cu: producer warp
template <class C>
void fill_stage(std::vector<unsigned char> &stage, const Grid &g, const std::vector<double> &in, int i0, int j0, int kb0, int n) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (size_t b = 0; b + 8 <= stage.size(); b += 8) memcpy(&stage[b], &nan, 8);
  const int z_first = kb0 * 8 - C::RUP;
  const int zabs = z_first + n * C::G;
  const int kb = (zabs >= 0) ? zabs / 8 : -((7 - zabs) / 8);
  const int pz = zabs - kb * 8;
  CHECK(pz >= 0 && pz + C::G <= 8, "stage crosses a brick layer: pz %d", pz);
  for (int job0 = 0; job0 < C::NCOPY; ++job0) {
    int job = job0, kind = 0, sbi = 0, sbj = 0;
    if (job < C::TI * C::TJ) {
      kind = 1, sbi = 1 + job % C::TI, sbj = 1 + job / C::TI;
    } else if ((job -= C::TI * C::TJ) < 2 * C::TJ) {
      kind = 1, sbi = (job & 1) ? C::TI + 1 : 0, sbj = 1 + (job >> 1);
    } else if ((job -= 2 * C::TJ) < 2 * C::NJH) {
      kind = 2 + (job & 1), sbi = 0 + (job >> 1), sbj = (job & 1) ? C::TJ + 1 : 0;
    }
    const int dsto = C::slotoff(sbi, kind >= 2 ? 0 : sbj) + (kind == 2 ? (8 - R) * 64 : 0);
    const unsigned id = g.at(i0 + sbi - 1, j0 + sbj - 1, kb);
    const double *src = in.data() + (size_t) id * 512 + pz * 64;
    if (kind == 1) {
      CHECK(dsto >= 0 && dsto + C::G * 512 <= C::STAGE, "slot copy outside the stage");
      memcpy(&stage[dsto], src, C::G * 512);
    } else {
      src += (kind == 2 ? (8 - R) * 8 : 0);
      for (int q = 0; q < C::G; ++q) {
        CHECK(dsto + q * 512 >= 0 && dsto + q * 512 + R * 64 <= C::STAGE, "row copy outside the stage");
        memcpy(&stage[dsto + q * 512], src + q * 64, R * 64);
      }
    }
  }
}

// one CTA: tile (tx, ty), k segment [kb0, kb0 + nl), box [lo, hi) -- bk_stencil_tiled.cu: march_body, consumers
template <class C>
void run_cta(const Grid &g, const std::vector<double> &in, std::vector<double> &out, const bk::DiamondCoef &cf, const int lo[3],
             const int hi[3], int tx, int ty, int kb0, int nl) {
  const int i0 = lo[0] + tx * C::TI, j0 = lo[1] + ty * C::TJ;
  const int P = nl * 8 + 2 * C::RUP, NS = P / C::G;
  std::vector<std::vector<unsigned char>> stages(NS, std::vector<unsigned char>(C::STAGE));
  for (int n = 0; n < NS; ++n) fill_stage<C>(stages[n], g, in, i0, j0, kb0, n);
  for (int tid = 0; tid < C::NCONS; ++tid) {
    const int c = tid & 3;
    const int e = (tid >> 2) & 1;
    int rest = tid >> 3;
    const int y0 = (rest % (8 / C::YT)) * C::YT;
    rest /= (8 / C::YT);
    const int bi = (rest % (C::TI / 2)) * 2 + e, bj = rest / (C::TI / 2);
    const int own_slot = C::slotoff(bi + 1, bj + 1);
    const int own_off = own_slot + y0 * 64 + c * 16;
    int joff[2 * R];
    for (int h = 0; h < 2 * R; ++h) {
      const int ya = (h < R) ? y0 - R + h : y0 + C::YT + (h - R);
      int base;
      if (ya < 0)
        base = C::slotoff(bi + 1, bj) + (8 + ya) * 64;
      else if (ya >= 8)
        base = C::slotoff(bi + 1, (bj == C::TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
      else
        base = own_slot + ya * 64;
      joff[h] = base + c * 16;
    }
    const int dl = C::slotoff(bi, bj + 1) - own_slot, dr = C::slotoff(bi + 2, bj + 1) - own_slot;
    const int ioffL = -16 * 1 + ((c - 1 < 0) ? dl + 64 : 0);
    const int ioffR = 16 * 1 + ((c + 1 > 3) ? dr - 64 : 0);
    // every address diamond_plane forms must stay inside the stage (plane 0 and plane C::G-1)
    for (int pl = 0; pl < C::G; ++pl) {
      const int offs[] = {own_off, own_off + (C::YT - 1) * 64, joff[0], joff[1], joff[2], joff[3]};
      for (int o : offs)
        for (int d : {0, ioffL, ioffR}) CHECK(o + d + pl * 512 >= 0 && o + d + pl * 512 + 16 <= C::STAGE && (o + d) % 16 == 0, "offset");
    }
    const bool mine = (i0 + bi < hi[0]) && (j0 + bj < hi[1]);
    const size_t glayer = (size_t) g.gy * g.gx;
    const size_t gcol = ((size_t) kb0 * g.gy + (j0 + bj)) * g.gx + (i0 + bi);
    unsigned id_next = mine ? g.id[gcol] : 0u;
    size_t outp = 0;
    unsigned edge_ij = 0;
    {
      const int gi = i0 + bi, gj = j0 + bj;
      if (gi == 0 && c == 0) edge_ij |= 1u;
      if (gi == g.gx - 1 && c == 3) edge_ij |= 2u;
      for (int r = 0; r < C::YT; ++r)
        if ((gj == 0 && y0 + r == 0) || (gj == g.gy - 1 && y0 + r == 7)) edge_ij |= 4u << r;
    }
    double2 acc[W][C::YT];
    for (int w = 0; w < W; ++w)
      for (int r = 0; r < C::YT; ++r) acc[w][r] = make_double2(0.0, 0.0);
    int orel = -R - C::RUP;
    const int nout = nl * 8;
    for (int tb = 0; tb < P; tb += W)
      for (int u = 0; u < W; ++u) {
        if (tb + u >= P) continue;
        const int t = tb + u;
        const unsigned char *pb = stages[t / C::G].data() + (t % C::G) * 512;
        const int sF = ((u - R) % W + W) % W;
        double2 v[C::YT];
        for (int r = 0; r < C::YT; ++r) v[r] = bk::diamond_ld2(pb + own_off + r * 64);
        const int zt = kb0 * 8 - C::RUP + tb + u;
        const unsigned edge = edge_ij | ((zt == 0 || zt == g.gz * 8 - 1) ? bk::kDiamondEdgeK : 0u);
        bk::diamond_plane<C::YT>(pb, own_off, joff[0], joff[1], joff[2 * R - 2], joff[2 * R - 1], ioffL, ioffR, ((tid >> 3) & 1) != 0, cf,
                                 acc, u, edge, v);
        if (orel >= 0 && orel < nout) {  // store_plane
          const int oz = orel & 7;
          if (oz == 0) {
            outp = (size_t) id_next * 512 + y0 * 8 + c * 2;
            if (mine && (orel >> 3) + 1 < nl) id_next = g.id[gcol + ((orel >> 3) + 1) * glayer];
          }
          if (mine)
            for (int r = 0; r < C::YT; ++r) {
              out[outp + oz * 64 + r * 8] = acc[sF][r].x;
              out[outp + oz * 64 + r * 8 + 1] = acc[sF][r].y;
            }
        }
        ++orel;
      }
  }
}

// ---- shared-memory bank model of the consumer loads --------------------------------------------------------------
// 32 banks x 4 B.  A 128-bit load is served a quarter warp at a time (8 lanes x 16 B = 128 B: one wavefront if the eight
// chunks fall into eight different 16-byte bank groups), a 64-bit load half a warp at a time (16 lanes x 8 B).  Lanes that
// read the SAME address share a wavefront.  Returns wavefronts per warp-instruction, summed over the instruction list.
template <class C>
void bank_report(const char *geo) {
  const int i0 = 0, j0 = 0;
  (void) i0, (void) j0;
  long total = 0, ideal = 0;
  int worst = 0;
  for (int warp = 0; warp < C::NCONS / 32; ++warp) {
    struct Ld {
      int off[32];
      int bytes;
    };
    std::vector<Ld> loads;
    for (int which = 0; which < 24; ++which) {
      Ld ld;
      ld.bytes = 16;
      bool used = true;
      for (int lane = 0; lane < 32; ++lane) {
        const int tid = warp * 32 + lane;
        const int c = tid & 3, e = (tid >> 2) & 1;
        int rest = tid >> 3;
        const int y0 = (rest % (8 / C::YT)) * C::YT;
        rest /= (8 / C::YT);
        const int bi = (rest % (C::TI / 2)) * 2 + e, bj = rest / (C::TI / 2);
        const int own_slot = C::slotoff(bi + 1, bj + 1), own_off = own_slot + y0 * 64 + c * 16;
        int joff[4];
        for (int h = 0; h < 4; ++h) {
          const int ya = (h < R) ? y0 - R + h : y0 + C::YT + (h - R);
          int base;
          if (ya < 0) base = C::slotoff(bi + 1, bj) + (8 + ya) * 64;
          else if (ya >= 8) base = C::slotoff(bi + 1, (bj == C::TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
          else base = own_slot + ya * 64;
          joff[h] = base + c * 16;
        }
        const int dl = C::slotoff(bi, bj + 1) - own_slot, dr = C::slotoff(bi + 2, bj + 1) - own_slot;
        const int iL = -16 + ((c - 1 < 0) ? dl + 64 : 0), iR = 16 + ((c + 1 > 3) ? dr - 64 : 0);
        // instruction list of diamond_plane: YT own rows, 4 halo rows, then left / right chunks of rows y0-1 .. y0+YT
        // (the two halo rows of that range only feed one cell each: 64-bit loads, see bk_diamond.h)
        int o = 0;
        if (which < C::YT) o = own_off + which * 64;
        else if (which < C::YT + 4) o = joff[which - C::YT];
        else {
          const int q = (which - C::YT - 4) / 2, side = (which - C::YT - 4) % 2;
          if (q >= C::YT + 2) {
            used = false;
            break;
          }
          const int row = (q == 0) ? joff[1] : (q == C::YT + 1) ? joff[2] : own_off + (q - 1) * 64;
          o = row + (side ? iR : iL);
          if (q == 0 || q == C::YT + 1) {   // 64-bit loads of lf.y / rg.x; odd row groups load the right cell first
            const bool swp = ((tid >> 3) & 1) != 0;
            const int first = swp ? iR : iL + 8, second = swp ? iL + 8 : iR;
            ld.bytes = 8, o = row + (side ? second : first);
          }
        }
        ld.off[lane] = o;
      }
      if (used) loads.push_back(ld);
    }
    for (const Ld &ld : loads) {
      const int lanes = ld.bytes == 16 ? 8 : 16, groups = 128 / ld.bytes;
      int wf = 0;
      for (int ph = 0; ph < 32 / lanes; ++ph) {
        int mult[16] = {0};
        std::vector<int> seen;
        for (int l = 0; l < lanes; ++l) {
          const int a = ld.off[ph * lanes + l];
          bool dup = false;
          for (int x : seen) dup = dup || x == a;
          if (dup) continue;
          seen.push_back(a);
          ++mult[(a / ld.bytes) % groups];
        }
        int m = 0;
        for (int gidx = 0; gidx < groups; ++gidx) m = std::max(m, mult[gidx]);
        wf += m;
      }
      total += wf, ideal += 32 * ld.bytes / 128;
      worst = std::max(worst, wf);
    }
  }
  const int warps = C::NCONS / 32;
  printf("bank model %-24s %ld wavefronts per plane-tile for the consumer loads (%.1f per warp; conflict-free would be %ld; worst "
         "instruction %d), %.3f per point\n", geo, total, (double) total / warps, ideal, worst, (double) total / (C::TI * C::TJ * 64));
}

struct Star {
  double c0, cp[3], cm[3];
};

// bk_stencil_advance(steps = 2) by definition, on cells
void reference(const Grid &g, const std::vector<double> &in, std::vector<double> &out, const Star &s, const int lo[3], const int hi[3]) {
  const int nx = g.gx * 8, ny = g.gy * 8, nz = g.gz * 8;
  auto cell = [&](const std::vector<double> &f, int x, int y, int z) -> double {
    if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz) return 0.0;
    return f[(size_t) g.at(x >> 3, y >> 3, z >> 3) * 512 + (z & 7) * 64 + (y & 7) * 8 + (x & 7)];
  };
  auto step = [&](const std::vector<double> &f, int x, int y, int z) {
    double a = s.c0 * cell(f, x, y, z);
    a += s.cp[0] * cell(f, x + 1, y, z) + s.cm[0] * cell(f, x - 1, y, z);
    a += s.cp[1] * cell(f, x, y + 1, z) + s.cm[1] * cell(f, x, y - 1, z);
    a += s.cp[2] * cell(f, x, y, z + 1) + s.cm[2] * cell(f, x, y, z - 1);
    return a;
  };
  std::vector<double> tmp(in.size(), 0.0);
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x)
        tmp[(size_t) g.at(x >> 3, y >> 3, z >> 3) * 512 + (z & 7) * 64 + (y & 7) * 8 + (x & 7)] = step(in, x, y, z);
  for (int z = lo[2] * 8; z < hi[2] * 8; ++z)
    for (int y = lo[1] * 8; y < hi[1] * 8; ++y)
      for (int x = lo[0] * 8; x < hi[0] * 8; ++x)
        out[(size_t) g.at(x >> 3, y >> 3, z >> 3) * 512 + (z & 7) * 64 + (y & 7) * 8 + (x & 7)] = step(tmp, x, y, z);
}

template <class C>
double run_case(int gx, int gy, int gz, const int lo[3], const int hi[3], int kl, const Star &s, const char *what) {
  Grid g{gx, gy, gz, {}};
  g.id.resize((size_t) gx * gy * gz);
  // ids in a scrambled order: the kernel must go through the grid, never assume id = position
  const unsigned nb = (unsigned) g.id.size();
  for (unsigned p = 0; p < nb; ++p) g.id[p] = 1 + (p * 7 + 3) % nb;
  if (nb % 7 == 0)
    for (unsigned p = 0; p < nb; ++p) g.id[p] = 1 + p;
  std::vector<double> in((size_t) (nb + 1) * 512, 0.0), got(in.size(), -7.0), want(in.size(), -7.0);
  for (size_t q = 512; q < in.size(); ++q) in[q] = frand();
  const bk::DiamondCoef cf = bk::diamond_coef(s.c0, s.cp[0], s.cm[0], s.cp[1], s.cm[1], s.cp[2], s.cm[2]);
  const int ntx = (hi[0] - lo[0] + C::TI - 1) / C::TI, nty = (hi[1] - lo[1] + C::TJ - 1) / C::TJ, nz = hi[2] - lo[2];
  for (int q = 0; q * kl < nz; ++q)
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx) run_cta<C>(g, in, got, cf, lo, hi, tx, ty, lo[2] + q * kl, std::min(kl, nz - q * kl));
  reference(g, in, want, s, lo, hi);
  double worst = 0.0;
  size_t bad = 0;
  for (size_t q = 0; q < got.size(); ++q) {
    const double d = std::fabs(got[q] - want[q]);
    if (!(d <= 1e-14)) ++bad;  // also catches NaN
    if (d > worst || d != d) worst = (d != d) ? 1e300 : d;
  }
  printf("%-46s grid %dx%dx%d box [%d,%d)x[%d,%d)x[%d,%d) kl %d: max |diff| %.3g, %zu cells off\n", what, gx, gy, gz, lo[0], hi[0], lo[1],
         hi[1], lo[2], hi[2], kl, worst, bad);
  if (bad) ++failures;
  return worst;
}

}  // namespace

template <class C>
void all_cases(const char *geo) {
  printf("---- %s: tile %dx%d bricks, %d rows per thread, %d planes per stage, %d consumer threads, stage %d B\n", geo, C::TI, C::TJ,
         C::YT, C::G, C::NCONS, C::STAGE);
  const Star mpi7 = {0.4, {0.1, 0.1, 0.1}, {0.1, 0.1, 0.1}};              // stencils/mpi7pt.py, fake.h:11-12
  const Star coeff = {0.31, {0.11, 0.23, 0.07}, {0.19, 0.05, 0.29}};      // stencils/7pt.py with seven different coeff[]
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {5, 6, 3};
    run_case<C>(5, 6, 3, lo, hi, 3, coeff, "whole grid, partial tiles, one segment");
    run_case<C>(5, 6, 3, lo, hi, 1, coeff, "whole grid, one layer per segment");
    run_case<C>(5, 6, 3, lo, hi, 2, mpi7, "whole grid, mpi7pt, ragged last segment");
  }
  {
    const int lo[3] = {1, 1, 1}, hi[3] = {5, 5, 3};
    run_case<C>(6, 6, 4, lo, hi, 2, coeff, "interior box (ghost shell skipped)");
  }
  {
    const int lo[3] = {0, 2, 1}, hi[3] = {3, 7, 2};
    run_case<C>(9, 7, 3, lo, hi, 1, coeff, "box touching two faces, off the tile grid");
  }
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    run_case<C>(1, 1, 1, lo, hi, 1, coeff, "a single brick: every face at once");
  }
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {8, 4, 2};
    run_case<C>(8, 4, 2, lo, hi, 2, mpi7, "exact tiles");
  }
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {17, 9, 2};
    run_case<C>(17, 9, 2, lo, hi, 2, coeff, "several tiles each way");
  }
}

int main() {
  srand(12345);
  all_cases<Geo<4, 4, 4, 2>>("BK_FUSED_COMPOSED      Cfg<2,4,4,4,2,3,168,2,..,2>");
  all_cases<Geo<4, 8, 4, 2>>("BK_FUSED_COMPOSED_WIDE Cfg<2,4,8,4,2,3,255,4,..,232,40,2>");
  all_cases<Geo<2, 4, 4, 2>>("developer variant      Cfg<2,2,4,4,2,3,128,2,..,2>");
  bank_report<Geo<4, 4, 4, 2>>("4x4 tiles, 4 rows");
  bank_report<Geo<4, 8, 4, 2>>("8x4 tiles, 4 rows");
  bank_report<Geo<2, 4, 4, 2>>("4x4 tiles, 2 rows");
  if (failures) {
    printf("FAILED (%d)\n", failures);
    return 1;
  }
  printf("diamond emulation ok\n");
  return 0;
}
