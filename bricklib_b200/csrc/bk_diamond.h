// bk_diamond.h -- two time steps of a radius-1 star stencil as ONE application of the composed operator.
//
// One step is  S u = c0 u + sum over the axes a of ( cp[a] u(+e_a) + cm[a] u(-e_a) )   (stencils/7pt.py, mpi7pt.py).
// Two steps are S(S u) = sum over pairs of taps: a 25-point "diamond" (|di| + |dj| + |dk| <= 2).  With S = S_xy + Z+ + Z-
// (the five in-plane taps and the two k taps) the square splits by k distance of the input plane t:
//
//     output t-2  +=  cpz^2                        * u_t
//     output t-1  +=  2 cpz                        * (S_xy u_t)
//     output t    +=  (S_xy^2 + 2 cpz cmz)           u_t          13 in-plane taps
//     output t+1  +=  2 cmz                        * (S_xy u_t)
//     output t+2   =  cmz^2                        * u_t
//
// i.e. 22 FMA per point for TWO time steps, one pass over shared memory, no intermediate plane, no CTA barrier and no
// halo strip: the marching kernel of bk_stencil_tiled.cu runs it exactly like a radius-2 star (five partial outputs per
// point in registers) plus four diagonal reads.
//
// Semantics = bk_stencil_advance(steps = 2): apply over the whole grid, then over [lo,hi), with the intermediate field
// ZERO outside the grid (null-brick semantics).  The composed operator would let an out-of-grid intermediate cell q feed
// its in-grid neighbour p with c(p<-q) * (S u)(q); of the inputs (S u)(q) reads only u(p) lies in the grid (everything
// else q touches is outside and reads the null brick), so the spurious term is cp[a] cm[a] u(p) for every grid face p
// lies on along axis a -- subtracted by the `fix` coefficients from the centre tap of exactly those cells.
//
// This header is compiled by nvcc into the kernel AND by g++ into tests/cpp/diamond_emulation.cpp, which replays the
// kernel's shared-memory layout and thread mapping on the host against a plain two-step reference.
#ifndef BK_DIAMOND_H
#define BK_DIAMOND_H

#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define BK_HD inline
#ifndef BK_HOST_DOUBLE2
#define BK_HOST_DOUBLE2
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) {
  double2 r = {x, y};
  return r;
}
#endif
#endif

namespace bk {

struct DiamondCoef {
  double s[5];    // one step in the plane: [0] centre, [1] +i, [2] -i, [3] +j, [4] -j
  double z[4];    // along k: [0] cpz^2 (output t-2), [1] 2 cpz (x S_xy, output t-1), [2] 2 cmz (t+1), [3] cmz^2 (t+2)
  double d[13];   // S_xy^2 + 2 cpz cmz: [0] (0,0); [1] (+1,0) [2] (-1,0) [3] (0,+1) [4] (0,-1); [5] (+2,0) [6] (-2,0)
                  // [7] (0,+2) [8] (0,-2); [9] (+1,+1) [10] (+1,-1) [11] (-1,+1) [12] (-1,-1)      (di, dj)
  double fix[3];  // a cell on a grid face along i / j / k must not receive cp*cm of that axis (see above)
};

// c0, then (cp, cm) per axis i, j, k
inline DiamondCoef diamond_coef(double c0, double cpx, double cmx, double cpy, double cmy, double cpz, double cmz) {
  DiamondCoef o;
  o.s[0] = c0, o.s[1] = cpx, o.s[2] = cmx, o.s[3] = cpy, o.s[4] = cmy;
  o.z[0] = cpz * cpz, o.z[1] = 2.0 * cpz, o.z[2] = 2.0 * cmz, o.z[3] = cmz * cmz;
  o.d[0] = c0 * c0 + 2.0 * cpx * cmx + 2.0 * cpy * cmy + 2.0 * cpz * cmz;
  o.d[1] = 2.0 * c0 * cpx, o.d[2] = 2.0 * c0 * cmx, o.d[3] = 2.0 * c0 * cpy, o.d[4] = 2.0 * c0 * cmy;
  o.d[5] = cpx * cpx, o.d[6] = cmx * cmx, o.d[7] = cpy * cpy, o.d[8] = cmy * cmy;
  o.d[9] = 2.0 * cpx * cpy, o.d[10] = 2.0 * cpx * cmy, o.d[11] = 2.0 * cmx * cpy, o.d[12] = 2.0 * cmx * cmy;
  o.fix[0] = cpx * cmx, o.fix[1] = cpy * cmy, o.fix[2] = cpz * cmz;
  return o;
}

// bits of `edge`: 0 = my cell .x lies on a grid face along i, 1 = my cell .y does, 2 + r = my row r lies on a face along j,
// 31 = this plane lies on a face along k
constexpr unsigned kDiamondEdgeK = 1u << 31;

BK_HD double2 diamond_ld2(const unsigned char *p) { return *reinterpret_cast<const double2 *>(p); }

// One input plane of the composed update on a patch of an x-pair times YT rows.
//   pb       base of the plane in the shared-memory stage
//   own_off  my x-pair in my first row;  jo0..jo3: my x-pair in rows y0-2, y0-1, y0+YT, y0+YT+1
//   iL, iR   from an x-pair to the 16-byte chunks left / right of it (cells x0-2,x0-1 / x0+2,x0+3) in the same row
//   swp      odd row group of the half warp (see below)
//   acc      five partial outputs per point; slot (u + 3) % 5 is COMPLETE on return (output t-2), the caller stores it
//   v        my own cells of this plane
// Rows y0-1 and y0+YT only feed the diagonal taps: one cell left (x0-1) and one cell right (x0+2) of my x-pair.  Those are
// 64-bit loads at a 16-byte lane stride -- all "x0-1" cells sit in the odd 8-byte bank positions, all "x0+2" cells in the
// even ones -- so the odd row groups of a half warp load the RIGHT cell first and the left one second: each instruction
// then covers all 32 banks (2 wavefronts instead of 4; tests/cpp/diamond_emulation.cpp has the bank model).
template <int YT>
BK_HD void diamond_plane(const unsigned char *pb, int own_off, int jo0, int jo1, int jo2, int jo3, int iL, int iR, const bool swp,
                         const DiamondCoef &cf, double2 (&acc)[5][YT], const int u, const unsigned edge,
                         const double2 (&v)[YT]) {
  const int sF = (u + 3) % 5, sM = (u + 4) % 5, s0 = u % 5, sP = (u + 1) % 5, sN = (u + 2) % 5;
#pragma unroll
  for (int r = 0; r < YT; ++r) {
    acc[sF][r].x = fma(cf.z[0], v[r].x, acc[sF][r].x);
    acc[sF][r].y = fma(cf.z[0], v[r].y, acc[sF][r].y);
  }
  double2 rows[YT + 4];  // my x-pair in rows y0-2 .. y0+YT+1
  rows[0] = diamond_ld2(pb + jo0), rows[1] = diamond_ld2(pb + jo1);
#pragma unroll
  for (int r = 0; r < YT; ++r) rows[2 + r] = v[r];
  rows[YT + 2] = diamond_ld2(pb + jo2), rows[YT + 3] = diamond_ld2(pb + jo3);
  // cells x0-2, x0-1 (lf) and x0+2, x0+3 (rg) of rows y0-1 .. y0+YT; of the first and the last row only lf.y and rg.x
  double2 lf[YT + 2], rg[YT + 2];
  {
    const int first = swp ? iR : iL + 8, second = swp ? iL + 8 : iR;
    const double a0 = *reinterpret_cast<const double *>(pb + jo1 + first), b0 = *reinterpret_cast<const double *>(pb + jo1 + second);
    const double a1 = *reinterpret_cast<const double *>(pb + jo2 + first), b1 = *reinterpret_cast<const double *>(pb + jo2 + second);
    lf[0] = make_double2(0.0, swp ? b0 : a0), rg[0] = make_double2(swp ? a0 : b0, 0.0);
    lf[YT + 1] = make_double2(0.0, swp ? b1 : a1), rg[YT + 1] = make_double2(swp ? a1 : b1, 0.0);
  }
#pragma unroll
  for (int q = 1; q <= YT; ++q) {
    const unsigned char *pr = pb + own_off + (q - 1) * 64;
    lf[q] = diamond_ld2(pr + iL);
    rg[q] = diamond_ld2(pr + iR);
  }
#pragma unroll
  for (int r = 0; r < YT; ++r) {
    const int q = r + 1;
    const double2 c = v[r], up = rows[r + 3], dn = rows[r + 1], up2 = rows[r + 4], dn2 = rows[r];
    const double xm2 = lf[q].x, xm1 = lf[q].y, xp2 = rg[q].x, xp3 = rg[q].y;
    // one step in the plane at my two cells -> the outputs one plane below and above
    double sa = cf.s[0] * c.x, sb = cf.s[0] * c.y;
    sa = fma(cf.s[1], c.y, sa), sb = fma(cf.s[1], xp2, sb);
    sa = fma(cf.s[2], xm1, sa), sb = fma(cf.s[2], c.x, sb);
    sa = fma(cf.s[3], up.x, sa), sb = fma(cf.s[3], up.y, sb);
    sa = fma(cf.s[4], dn.x, sa), sb = fma(cf.s[4], dn.y, sb);
    acc[sM][r].x = fma(cf.z[1], sa, acc[sM][r].x), acc[sM][r].y = fma(cf.z[1], sb, acc[sM][r].y);
    acc[sP][r].x = fma(cf.z[2], sa, acc[sP][r].x), acc[sP][r].y = fma(cf.z[2], sb, acc[sP][r].y);
    // two steps in the plane -> this plane's own output
    double da = fma(cf.d[0], c.x, acc[s0][r].x), db = fma(cf.d[0], c.y, acc[s0][r].y);
    da = fma(cf.d[1], c.y, da), db = fma(cf.d[1], xp2, db);
    da = fma(cf.d[2], xm1, da), db = fma(cf.d[2], c.x, db);
    da = fma(cf.d[3], up.x, da), db = fma(cf.d[3], up.y, db);
    da = fma(cf.d[4], dn.x, da), db = fma(cf.d[4], dn.y, db);
    da = fma(cf.d[5], xp2, da), db = fma(cf.d[5], xp3, db);
    da = fma(cf.d[6], xm2, da), db = fma(cf.d[6], xm1, db);
    da = fma(cf.d[7], up2.x, da), db = fma(cf.d[7], up2.y, db);
    da = fma(cf.d[8], dn2.x, da), db = fma(cf.d[8], dn2.y, db);
    da = fma(cf.d[9], up.y, da), db = fma(cf.d[9], rg[q + 1].x, db);
    da = fma(cf.d[10], dn.y, da), db = fma(cf.d[10], rg[q - 1].x, db);
    da = fma(cf.d[11], lf[q + 1].y, da), db = fma(cf.d[11], up.x, db);
    da = fma(cf.d[12], lf[q - 1].y, da), db = fma(cf.d[12], dn.x, db);
    acc[s0][r].x = da, acc[s0][r].y = db;
    acc[sN][r].x = cf.z[3] * c.x, acc[sN][r].y = cf.z[3] * c.y;
  }
  if (edge != 0u) {  // cells on a face of the grid: the intermediate value outside the grid is zero, not (S u)(outside)
    const double fk = (edge & kDiamondEdgeK) ? cf.fix[2] : 0.0;
#pragma unroll
    for (int r = 0; r < YT; ++r) {
      const double fj = ((edge >> (2 + r)) & 1u) ? cf.fix[1] : 0.0;
      const double fa = fk + fj + ((edge & 1u) ? cf.fix[0] : 0.0), fb = fk + fj + ((edge & 2u) ? cf.fix[0] : 0.0);
      acc[s0][r].x = fma(-fa, v[r].x, acc[s0][r].x);
      acc[s0][r].y = fma(-fb, v[r].y, acc[s0][r].y);
    }
  }
}

}  // namespace bk
#endif  // BK_DIAMOND_H
