/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (checker, never shipped, never on the product path).
 *
 * A thin extern "C" shell around the UNMODIFIED reference sources, compiled where they lie under /root/reference
 * (recipe: oracle/Makefile; outputs only into oracle/_ref/).  Everything numerical in here is reference code:
 *   - brick layout, accessors:      /root/reference/include/brick.h:53-395
 *   - init_grid / copyToBrick:      /root/reference/include/bricksetup.h:73-221
 *   - BrickDecomp numbering+tables: /root/reference/include/brick-mpi.h:178-513, src/brick-mpi.cpp:9-52
 *   - exchange():                   /root/reference/include/brick-mpi.h:466-495 (over oracle/mpi_stub/mpi.h)
 *   - stencil bodies:               the reference's own code generator (codegen/vecscatter) expands every
 *                                   brick("stencils/<x>.py", VSVEC, (8,8,8), (VFOLD), b) below, exactly as
 *                                   stencils/3axis.cpp:68-76 and weak/main.cpp:26-36 do.
 * This file only moves raw pointers in and out so Python (ctypes) and the parity tests can drive it.
 */
#include "stencils/fake.h"
#include "stencils/cpuvfold.h"
#include <brick.h>
#include <bricksetup.h>
#include <brickcompare.h>
#include <omp.h>
#include <algorithm>
#include <vector>
#include <cstdint>

typedef Brick<Dim<8, 8, 8>, Dim<VFOLD>> Brick3D;

/* required by 7pt.py's ConstRef("coeff[i]") */
static bElem *coeff = nullptr;

namespace {

BrickStorage borrow_storage(double *dat, long chunks, size_t step) {
  BrickStorage s;
  s.dat = std::shared_ptr<bElem>(dat, [](bElem *) {});
  s.chunks = chunks;
  s.step = step;
  return s;
}

struct InfoView {
  BrickInfo<3> info;
  InfoView(unsigned *adj, unsigned nbricks) : info(0) {
    free(info.adj);
    info.adj = (BrickInfo<3>::adjlist) adj;
    info.nbricks = nbricks;
  }
};

/* the brick(...) line must stand alone: vecscatter replaces that whole source line (codegen/vecscatter:172-175) */
void sweep_7pt(Brick3D &bIn, Brick3D &bOut, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/7pt.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}
void sweep_mpi7pt(Brick3D &in, Brick3D &out, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/mpi7pt.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}
void sweep_mpi13pt(Brick3D &in, Brick3D &out, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/mpi13pt.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}
void sweep_mpi25pt(Brick3D &in, Brick3D &out, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/mpi25pt.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}
void sweep_mpi125pt(Brick3D &in, Brick3D &out, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/mpi125pt.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}

/* stencils/cond.py: coeff[t] * max(in, 0.0) summed, then If(calc > 0, calc, -calc); its generated code calls max() */
using std::max;
void sweep_cond(Brick3D &bIn, Brick3D &bOut, const unsigned *grid, const long *sb, const long *lo, const long *hi) {
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        brick("/root/reference/stencils/cond.py", VSVEC, (8, 8, 8), (VFOLD), b);
      }
}

struct Decomp {
  BrickDecomp<3, 8, 8, 8> d;
  std::vector<unsigned> tdims;
  Decomp(const std::vector<unsigned> &dom, unsigned depth) : d(dom, depth) {}
};

}  // namespace

extern "C" {

const char *ref_isa() { return VSVEC; }
int ref_threads() { return omp_get_max_threads(); }

/* ---- init_grid (single/ drivers) ---------------------------------------------------------------------------- */
void ref_init_grid(const long *dimlist, unsigned *grid_out, unsigned *adj_out) {
  unsigned *g;
  std::vector<long> dl(dimlist, dimlist + 3);
  BrickInfo<3> info = init_grid<3>(g, dl);
  long n = dl[0] * dl[1] * dl[2];
  memcpy(grid_out, g, sizeof(unsigned) * n);
  memcpy(adj_out, info.adj, sizeof(unsigned) * 27 * n);
  free(g);
  free(info.adj);
}

/* ---- BrickDecomp ------------------------------------------------------------------------------------------- */
void *ref_decomp_new(const unsigned *dom, unsigned depth, const int *cart_dims, const int *cart_coo) {
  std::vector<unsigned> dv(dom, dom + 3);
  Decomp *D = new Decomp(dv, depth);
  D->d.comm = 0;
  mpistub_set_cart(cart_dims[0], cart_dims[1], cart_dims[2]);
  int coo[3] = {cart_coo[0], cart_coo[1], cart_coo[2]};
  MPI_Comm c = 0;
  populate(c, D->d, 0, 1, coo);
  D->d.initialize(skin3d_good);
  for (int i = 0; i < 3; ++i) D->tdims.push_back(dom[i] / 8 + 2 * (depth / 8));
  return D;
}
void ref_decomp_free(void *h) { delete (Decomp *) h; }
unsigned ref_decomp_nbricks(void *h) { return ((Decomp *) h)->d.getBrickInfo().nbricks; }
void ref_decomp_sep_pos(void *h, unsigned *out) {
  for (int i = 0; i < 3; ++i) out[i] = ((Decomp *) h)->d.sep_pos[i];
}
void ref_decomp_grid(void *h, unsigned *out) {
  Decomp *D = (Decomp *) h;
  long p = 0;
  for (unsigned k = 0; k < D->tdims[2]; ++k)
    for (unsigned j = 0; j < D->tdims[1]; ++j)
      for (unsigned i = 0; i < D->tdims[0]; ++i) out[p++] = D->d[k][j][i];
}
void ref_decomp_adj(void *h, unsigned *out) {
  BrickInfo<3> info = ((Decomp *) h)->d.getBrickInfo();
  memcpy(out, info.adj, sizeof(unsigned) * 27 * info.nbricks);
}
int ref_decomp_nregions(void *h) { return (int) ((Decomp *) h)->d.ghost.size(); }
/* which: 0 ghost, 1 skin.  out[n][6] = pos,len,skin_st,skin_ed,first_pad,last_pad; set[n] = neighbor BitSet; peer[n] = rank */
void ref_decomp_regions(void *h, int which, unsigned *out, uint64_t *set, int *peer) {
  Decomp *D = (Decomp *) h;
  auto &v = which ? D->d.skin : D->d.ghost;
  for (size_t n = 0; n < v.size(); ++n) {
    out[n * 6 + 0] = v[n].pos;
    out[n * 6 + 1] = v[n].len;
    out[n * 6 + 2] = v[n].skin_st;
    out[n * 6 + 3] = v[n].skin_ed;
    out[n * 6 + 4] = v[n].first_pad;
    out[n * 6 + 5] = v[n].last_pad;
    set[n] = v[n].neighbor.set;
    peer[n] = D->d.rank_map[v[n].neighbor.set];
  }
}
void ref_decomp_skin_size(void *h, long *out) {
  Decomp *D = (Decomp *) h;
  for (size_t i = 0; i < D->d.skin_size.size(); ++i) out[i] = D->d.skin_size[i];
}
/* post this rank's 42 Irecv + 42 Isend through the reference's own exchange(); deliver with ref_mpi_deliver() */
void ref_decomp_exchange(void *h, int rank, double *dat, long chunks, size_t step) {
  Decomp *D = (Decomp *) h;
  mpistub_set_rank(rank);
  BrickStorage s = borrow_storage(dat, chunks, step);
  D->d.exchange(s);
}
void ref_mpi_deliver() { mpistub_deliver(); }

/* ---- array <-> brick -------------------------------------------------------------------------------------- */
static void copy_common(int dir, const long *dimlist, const long *padding, const long *ghost, double *arr,
                        unsigned *grid, unsigned *adj, unsigned nbricks, double *dat, size_t step, unsigned offset) {
  InfoView iv(adj, nbricks);
  BrickStorage s = borrow_storage(dat, nbricks, step);
  Brick3D br(&iv.info, s, offset);
  std::vector<long> dl(dimlist, dimlist + 3), pd(padding, padding + 3), gz(ghost, ghost + 3);
  if (dir == 0)
    copyToBrick<3>(dl, pd, gz, arr, grid, br);
  else
    copyFromBrick<3>(dl, pd, gz, arr, grid, br);
  iv.info.adj = nullptr;
}
void ref_copy_to_brick(const long *dimlist, const long *padding, const long *ghost, double *arr, unsigned *grid,
                       unsigned *adj, unsigned nbricks, double *dat, size_t step, unsigned offset) {
  copy_common(0, dimlist, padding, ghost, arr, grid, adj, nbricks, dat, step, offset);
}
void ref_copy_from_brick(const long *dimlist, const long *padding, const long *ghost, double *arr, unsigned *grid,
                         unsigned *adj, unsigned nbricks, double *dat, size_t step, unsigned offset) {
  copy_common(1, dimlist, padding, ghost, arr, grid, adj, nbricks, dat, step, offset);
}
/* the reference comparator, tolerance BRICK_TOLERANCE = 1e-6 (cmpconst.h:9) */
int ref_compare_brick(const long *dimlist, const long *padding, const long *ghost, double *arr, unsigned *grid,
                      unsigned *adj, unsigned nbricks, double *dat, size_t step, unsigned offset) {
  InfoView iv(adj, nbricks);
  BrickStorage s = borrow_storage(dat, nbricks, step);
  Brick3D br(&iv.info, s, offset);
  std::vector<long> dl(dimlist, dimlist + 3), pd(padding, padding + 3), gz(ghost, ghost + 3);
  bool ok = compareBrick<3>(dl, pd, gz, arr, grid, br);
  iv.info.adj = nullptr;
  return ok ? 1 : 0;
}

/* ---- one sweep of the generated brick code over the brick box [lo,hi) of `grid` ----------------------------- */
/* stencil: 0 7pt.py (coeff[7]), 1 mpi7pt, 2 mpi13pt, 3 mpi25pt, 4 mpi125pt, 5 cond.py (coeff[7]) */
int ref_sweep_brick(int stencil, const unsigned *grid, const long *sb, const long *lo, const long *hi, unsigned *adj,
                    unsigned nbricks, double *dat_in, size_t step_in, unsigned off_in, double *dat_out,
                    size_t step_out, unsigned off_out, const double *cf) {
  InfoView iv(adj, nbricks);
  BrickStorage si = borrow_storage(dat_in, nbricks, step_in), so = borrow_storage(dat_out, nbricks, step_out);
  Brick3D in(&iv.info, si, off_in), out(&iv.info, so, off_out);
  coeff = const_cast<double *>(cf);
  int rc = 0;
  switch (stencil) {
    case 0: sweep_7pt(in, out, grid, sb, lo, hi); break;
    case 1: sweep_mpi7pt(in, out, grid, sb, lo, hi); break;
    case 2: sweep_mpi13pt(in, out, grid, sb, lo, hi); break;
    case 3: sweep_mpi25pt(in, out, grid, sb, lo, hi); break;
    case 4: sweep_mpi125pt(in, out, grid, sb, lo, hi); break;
    case 5: sweep_cond(in, out, grid, sb, lo, hi); break; /* cond.py (coeff[7]) */
    default: rc = -1;
  }
  iv.info.adj = nullptr;
  return rc;
}

}  // extern "C"
