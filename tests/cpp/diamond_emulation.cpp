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

// ---- Cfg<2, 4, 4, 4, 2, 3, ...> restated (bk_stencil_tiled.cu: struct Cfg) ---------------------------------------------
constexpr int R = 2, YT = 4, TI = 4, TJ = 4, G = 2, W = 2 * R + 1;
constexpr int RUP = ((R + G - 1) / G) * G;
constexpr int SW = TI + 2, SH = TJ + 1;
constexpr int SLOTP = G * 512 + 64;
constexpr int STAGE = ((SH * SW * SLOTP + 127) / 128) * 128;
constexpr int NCONS = TI * TJ * 32 / YT;
constexpr int NJH = TI + 2;  // DIAM: corner columns too
constexpr int NCOPY = TI * TJ + 2 * TJ + 2 * NJH;
constexpr int slotoff(int bi, int bj) { return (bj * SW + bi) * SLOTP; }

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
void fill_stage(std::vector<unsigned char> &stage, const Grid &g, const std::vector<double> &in, int i0, int j0, int kb0, int n) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (size_t b = 0; b + 8 <= stage.size(); b += 8) memcpy(&stage[b], &nan, 8);
  const int z_first = kb0 * 8 - RUP;
  const int zabs = z_first + n * G;
  const int kb = (zabs >= 0) ? zabs / 8 : -((7 - zabs) / 8);
  const int pz = zabs - kb * 8;
  CHECK(pz >= 0 && pz + G <= 8, "stage crosses a brick layer: pz %d", pz);
  for (int job0 = 0; job0 < NCOPY; ++job0) {
    int job = job0, kind = 0, sbi = 0, sbj = 0;
    if (job < TI * TJ) {
      kind = 1, sbi = 1 + job % TI, sbj = 1 + job / TI;
    } else if ((job -= TI * TJ) < 2 * TJ) {
      kind = 1, sbi = (job & 1) ? TI + 1 : 0, sbj = 1 + (job >> 1);
    } else if ((job -= 2 * TJ) < 2 * NJH) {
      kind = 2 + (job & 1), sbi = 0 + (job >> 1), sbj = (job & 1) ? TJ + 1 : 0;
    }
    const int dsto = slotoff(sbi, kind >= 2 ? 0 : sbj) + (kind == 2 ? (8 - R) * 64 : 0);
    const unsigned id = g.at(i0 + sbi - 1, j0 + sbj - 1, kb);
    const double *src = in.data() + (size_t) id * 512 + pz * 64;
    if (kind == 1) {
      CHECK(dsto >= 0 && dsto + G * 512 <= STAGE, "slot copy outside the stage");
      memcpy(&stage[dsto], src, G * 512);
    } else {
      src += (kind == 2 ? (8 - R) * 8 : 0);
      for (int q = 0; q < G; ++q) {
        CHECK(dsto + q * 512 >= 0 && dsto + q * 512 + R * 64 <= STAGE, "row copy outside the stage");
        memcpy(&stage[dsto + q * 512], src + q * 64, R * 64);
      }
    }
  }
}

// one CTA: tile (tx, ty), k segment [kb0, kb0 + nl), box [lo, hi) -- bk_stencil_tiled.cu: march_body, consumers
void run_cta(const Grid &g, const std::vector<double> &in, std::vector<double> &out, const bk::DiamondCoef &cf, const int lo[3],
             const int hi[3], int tx, int ty, int kb0, int nl) {
  const int i0 = lo[0] + tx * TI, j0 = lo[1] + ty * TJ;
  const int P = nl * 8 + 2 * RUP, NS = P / G;
  std::vector<std::vector<unsigned char>> stages(NS, std::vector<unsigned char>(STAGE));
  for (int n = 0; n < NS; ++n) fill_stage(stages[n], g, in, i0, j0, kb0, n);
  for (int tid = 0; tid < NCONS; ++tid) {
    const int c = tid & 3;
    const int e = (tid >> 2) & 1;
    int rest = tid >> 3;
    const int y0 = (rest % (8 / YT)) * YT;
    rest /= (8 / YT);
    const int bi = (rest % (TI / 2)) * 2 + e, bj = rest / (TI / 2);
    const int own_slot = slotoff(bi + 1, bj + 1);
    const int own_off = own_slot + y0 * 64 + c * 16;
    int joff[2 * R];
    for (int h = 0; h < 2 * R; ++h) {
      const int ya = (h < R) ? y0 - R + h : y0 + YT + (h - R);
      int base;
      if (ya < 0)
        base = slotoff(bi + 1, bj) + (8 + ya) * 64;
      else if (ya >= 8)
        base = slotoff(bi + 1, (bj == TJ - 1) ? 0 : bj + 2) + (ya - 8) * 64;
      else
        base = own_slot + ya * 64;
      joff[h] = base + c * 16;
    }
    const int dl = slotoff(bi, bj + 1) - own_slot, dr = slotoff(bi + 2, bj + 1) - own_slot;
    const int ioffL = -16 * 1 + ((c - 1 < 0) ? dl + 64 : 0);
    const int ioffR = 16 * 1 + ((c + 1 > 3) ? dr - 64 : 0);
    // every address diamond_plane forms must stay inside the stage (plane 0 and plane G-1)
    for (int pl = 0; pl < G; ++pl) {
      const int offs[] = {own_off, own_off + (YT - 1) * 64, joff[0], joff[1], joff[2], joff[3]};
      for (int o : offs)
        for (int d : {0, ioffL, ioffR}) CHECK(o + d + pl * 512 >= 0 && o + d + pl * 512 + 16 <= STAGE && (o + d) % 16 == 0, "offset");
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
      for (int r = 0; r < YT; ++r)
        if ((gj == 0 && y0 + r == 0) || (gj == g.gy - 1 && y0 + r == 7)) edge_ij |= 4u << r;
    }
    double2 acc[W][YT];
    for (int w = 0; w < W; ++w)
      for (int r = 0; r < YT; ++r) acc[w][r] = make_double2(0.0, 0.0);
    int orel = -R - RUP;
    const int nout = nl * 8;
    for (int tb = 0; tb < P; tb += W)
      for (int u = 0; u < W; ++u) {
        if (tb + u >= P) continue;
        const int t = tb + u;
        const unsigned char *pb = stages[t / G].data() + (t % G) * 512;
        const int sF = ((u - R) % W + W) % W;
        double2 v[YT];
        for (int r = 0; r < YT; ++r) v[r] = bk::diamond_ld2(pb + own_off + r * 64);
        const int zt = kb0 * 8 - RUP + tb + u;
        const unsigned edge = edge_ij | ((zt == 0 || zt == g.gz * 8 - 1) ? bk::kDiamondEdgeK : 0u);
        bk::diamond_plane<YT>(pb, own_off, joff[0], joff[1], joff[2 * R - 2], joff[2 * R - 1], ioffL, ioffR, cf, acc, u, edge, v);
        if (orel >= 0 && orel < nout) {  // store_plane
          const int oz = orel & 7;
          if (oz == 0) {
            outp = (size_t) id_next * 512 + y0 * 8 + c * 2;
            if (mine && (orel >> 3) + 1 < nl) id_next = g.id[gcol + ((orel >> 3) + 1) * glayer];
          }
          if (mine)
            for (int r = 0; r < YT; ++r) {
              out[outp + oz * 64 + r * 8] = acc[sF][r].x;
              out[outp + oz * 64 + r * 8 + 1] = acc[sF][r].y;
            }
        }
        ++orel;
      }
  }
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
  const int ntx = (hi[0] - lo[0] + TI - 1) / TI, nty = (hi[1] - lo[1] + TJ - 1) / TJ, nz = hi[2] - lo[2];
  for (int q = 0; q * kl < nz; ++q)
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx) run_cta(g, in, got, cf, lo, hi, tx, ty, lo[2] + q * kl, std::min(kl, nz - q * kl));
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

int main() {
  srand(12345);
  const Star mpi7 = {0.4, {0.1, 0.1, 0.1}, {0.1, 0.1, 0.1}};              // stencils/mpi7pt.py, fake.h:11-12
  const Star coeff = {0.31, {0.11, 0.23, 0.07}, {0.19, 0.05, 0.29}};      // stencils/7pt.py with seven different coeff[]
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {5, 6, 3};
    run_case(5, 6, 3, lo, hi, 3, coeff, "whole grid, partial tiles, one segment");
    run_case(5, 6, 3, lo, hi, 1, coeff, "whole grid, one layer per segment");
    run_case(5, 6, 3, lo, hi, 2, mpi7, "whole grid, mpi7pt, ragged last segment");
  }
  {
    const int lo[3] = {1, 1, 1}, hi[3] = {5, 5, 3};
    run_case(6, 6, 4, lo, hi, 2, coeff, "interior box (ghost shell skipped)");
  }
  {
    const int lo[3] = {0, 2, 1}, hi[3] = {3, 7, 2};
    run_case(9, 7, 3, lo, hi, 1, coeff, "box touching two faces, off the tile grid");
  }
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    run_case(1, 1, 1, lo, hi, 1, coeff, "a single brick: every face at once");
  }
  {
    const int lo[3] = {0, 0, 0}, hi[3] = {8, 4, 2};
    run_case(8, 4, 2, lo, hi, 2, mpi7, "exact tiles");
  }
  if (failures) {
    printf("FAILED (%d)\n", failures);
    return 1;
  }
  printf("diamond emulation ok\n");
  return 0;
}
