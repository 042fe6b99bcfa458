/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C restatement of the reference hot path.
 * Every function cites the reference lines it follows; paths are relative to /root/reference.
 */
#include "oracle.h"
#include "taps.h"
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------
 * stencil specs: stencils/{7pt,mpi7pt,mpi13pt,mpi25pt,mpi125pt}.py, coefficient values stencils/fake.h:11-33
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  const orc_tap_t *taps;
  int ntaps;
  const double *coef; /* NULL: caller supplies coeff[] */
  int radius, st_iter;
} orc_spec_t;

static const orc_spec_t SPECS[5] = {
    {orc_taps_7pt, ORC_7PT_NTAPS, NULL, 1, 8},                        /* fake.h:344 (7pt family: ST_ITER 8) */
    {orc_taps_mpi7pt, ORC_MPI7PT_NTAPS, orc_coef_mpi7pt, 1, 8},       /* fake.h:343-344 */
    {orc_taps_mpi13pt, ORC_MPI13PT_NTAPS, orc_coef_mpi13pt, 2, 4},    /* fake.h:71-72 */
    {orc_taps_mpi25pt, ORC_MPI25PT_NTAPS, orc_coef_mpi25pt, 4, 2},    /* fake.h:42-43 */
    {orc_taps_mpi125pt, ORC_MPI125PT_NTAPS, orc_coef_mpi125pt, 2, 4}, /* fake.h:88-89 */
};

int orc_radius(int s) { return (s < 0 || s > 4) ? -1 : SPECS[s].radius; }
int orc_st_iter(int s) { return (s < 0 || s > 4) ? -1 : SPECS[s].st_iter; }

/* out = c0*in(t0) + c1*in(t1) + ... evaluated left to right, as the spec's `+` chain is written */
int orc_sweep_array(int stencil, const long *ext, const long *lo, const long *hi, const double *in, double *out,
                    const double *coeff) {
  if (stencil < 0 || stencil > 4) return -1;
  const orc_spec_t *sp = &SPECS[stencil];
  const double *cf = sp->coef ? sp->coef : coeff;
  if (!cf) return -2;
  const long sj = ext[0], sk = ext[0] * ext[1];
#pragma omp parallel for collapse(2)
  for (long k = lo[2]; k < hi[2]; ++k)
    for (long j = lo[1]; j < hi[1]; ++j)
      for (long i = lo[0]; i < hi[0]; ++i) {
        const long p = i + j * sj + k * sk;
        double acc = 0.0;
        for (int t = 0; t < sp->ntaps; ++t) {
          const orc_tap_t *tp = &sp->taps[t];
          double v = cf[tp->c] * in[p + tp->di + tp->dj * sj + tp->dk * sk];
          acc = t ? acc + v : v;
        }
        out[p] = acc;
      }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * init_grid<3>: bricksetup.h:73-90 (ids), :37-43 + :20-35 (adjacency fill: grid_ptr +- stride, only the LINEAR
 * range [low,high) is checked, so rows wrap into their neighbours at the i/j borders -- faithfully restated)
 * ------------------------------------------------------------------------------------------------------------- */
void orc_init_grid(const long *dl, unsigned *grid, unsigned *adj) {
  const long n = dl[0] * dl[1] * dl[2];
  const long st[3] = {1, dl[0], dl[0] * dl[1]};
  for (long p = 0; p < n; ++p) grid[p] = (unsigned) p;
  for (long p = 0; p < n; ++p)
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          long q = p + di * st[0] + dj * st[1] + dk * st[2];
          adj[p * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] = (q >= 0 && q < n) ? grid[q] : 0;
        }
}

/* ---------------------------------------------------------------------------------------------------------------
 * copyToBrick / copyFromBrick: bricksetup.h:139-159 (iter_grid strides), :103-127 (iter: brick index s runs over
 * [ghost/tile, (dim+ghost)/tile), array origin padding + s*tile), :92-101 (fill: element loop through the accessor)
 * ------------------------------------------------------------------------------------------------------------- */
void orc_copy_brick(int dir, const long *dl, const long *pad, const long *gz, double *arr, const unsigned *grid,
                    double *dat, size_t step, size_t off) {
  long sA[3], sB[3], a = 1, b = 1;
  for (int d = 0; d < 3; ++d) {
    sA[d] = a;
    sB[d] = b;
    a *= dl[d] + 2 * (pad[d] + gz[d]);
    b *= (dl[d] + 2 * gz[d]) / 8;
  }
#pragma omp parallel for collapse(2)
  for (long bk = gz[2] / 8; bk < (dl[2] + gz[2]) / 8; ++bk)
    for (long bj = gz[1] / 8; bj < (dl[1] + gz[1]) / 8; ++bj)
      for (long bi = gz[0] / 8; bi < (dl[0] + gz[0]) / 8; ++bi) {
        unsigned id = grid[bi * sB[0] + bj * sB[1] + bk * sB[2]];
        double *bp = dat + (size_t) id * step + off;
        double *ap = arr + (pad[0] + bi * 8) * sA[0] + (pad[1] + bj * 8) * sA[1] + (pad[2] + bk * 8) * sA[2];
        for (int k = 0; k < 8; ++k)
          for (int j = 0; j < 8; ++j)
            for (int i = 0; i < 8; ++i) {
              double *e = ap + i * sA[0] + j * sA[1] + k * sA[2];
              if (dir == 0)
                bp[64 * k + 8 * j + i] = *e;
              else
                *e = bp[64 * k + 8 * j + i];
            }
      }
}

/* ---------------------------------------------------------------------------------------------------------------
 * brick form: the accessor (brick.h:234-246) maps an out-of-brick index to adjacency slot pos*3 + (idx+D)/D and
 * local index (idx+D)%D per axis; slot = (ok+1)*9 + (oj+1)*3 + (oi+1).  Loop nest: weak/main.cpp:26-36.
 * ------------------------------------------------------------------------------------------------------------- */
int orc_sweep_brick(int stencil, const unsigned *grid, const long *sb, const long *lo, const long *hi,
                    const unsigned *adj, const double *din, size_t step_in, size_t off_in, double *dout,
                    size_t step_out, size_t off_out, const double *coeff) {
  if (stencil < 0 || stencil > 4) return -1;
  const orc_spec_t *sp = &SPECS[stencil];
  const double *cf = sp->coef ? sp->coef : coeff;
  if (!cf) return -2;
#pragma omp parallel for collapse(2)
  for (long tk = lo[2]; tk < hi[2]; ++tk)
    for (long tj = lo[1]; tj < hi[1]; ++tj)
      for (long ti = lo[0]; ti < hi[0]; ++ti) {
        const unsigned b = grid[(tk * sb[1] + tj) * sb[0] + ti];
        const unsigned *nb = adj + (size_t) b * 27;
        double *o = dout + (size_t) b * step_out + off_out;
        for (int k = 0; k < 8; ++k)
          for (int j = 0; j < 8; ++j)
            for (int i = 0; i < 8; ++i) {
              double acc = 0.0;
              for (int t = 0; t < sp->ntaps; ++t) {
                const orc_tap_t *tp = &sp->taps[t];
                int x = i + tp->di + 8, y = j + tp->dj + 8, z = k + tp->dk + 8;
                unsigned src = nb[(z / 8) * 9 + (y / 8) * 3 + (x / 8)];
                double v = cf[tp->c] * din[(size_t) src * step_in + off_in + 64 * (z % 8) + 8 * (y % 8) + (x % 8)];
                acc = t ? acc + v : v;
              }
              o[64 * k + 8 * j + i] = acc;
            }
      }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * BitSet: bitset.h:19-128.  element +a -> bit a, -a -> bit 31+a; operator! swaps the two 31-bit halves (:120-127)
 * ------------------------------------------------------------------------------------------------------------- */
static uint64_t bs_bit(int e) { return 1ull << (e < 0 ? 31 - e : e); }
static int bs_has(uint64_t s, int e) { return (s & bs_bit(e)) != 0; }
uint64_t orc_bitset_of(int a, int b, int c) {
  uint64_t s = 0;
  if (a) s ^= bs_bit(a);
  if (b) s ^= bs_bit(b);
  if (c) s ^= bs_bit(c);
  return s;
}
uint64_t orc_bitset_neg(uint64_t s) {
  const uint64_t mask = (1ull << 32) - 1ull;
  return ((s & mask) << 31) | (s >> 31);
}

/* skin3d_good: src/brick-mpi.cpp:25-52 (order matters: it fixes the brick numbering) */
static const int SKIN3D_GOOD[26][3] = {
    {1, 0, 0},   {1, -3, 0},   {1, 2, -3},  {1, 2, 0},    {1, 2, 3},   {2, 3, 0},  {2, 0, 0},   {2, -3, 0}, {-1, 2, -3},
    {-1, 2, 0},  {-1, 2, 3},   {-1, 3, 0},  {-1, 0, 0},   {-3, 0, 0},  {-1, -3, 0}, {-1, -2, -3}, {-1, -2, 0},
    {-1, -2, 3}, {-2, 3, 0},   {-2, 0, 0},  {-2, -3, 0},  {1, -2, -3}, {1, -2, 0}, {1, -2, 3},  {1, 3, 0},  {3, 0, 0}};

/* allneighbors(0,1,3,...): src/brick-mpi.cpp:9-23 -- per axis the order is (+, none, -), axis 1 outermost */
static void all_neighbors(uint64_t out[27]) {
  static const int pick[3] = {1, 0, -1};
  int n = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) out[n++] = orc_bitset_of(pick[a] * 1, pick[b] * 2, pick[c] * 3);
}

/* _populate: brick-mpi.h:218-239 (region -> coordinate ranges, axis dim-1 outermost, numbering in visit order) */
static void populate_rec(orc_decomp_t *D, const long *stride, uint64_t region, long ref, int d, unsigned *pos) {
  if (d < 0) {
    D->grid[ref] = (*pos)++;
    return;
  }
  long st, n;
  if (bs_has(region, -(d + 1))) {
    st = D->gdepth[d];
    n = D->gdepth[d];
  } else if (bs_has(region, d + 1)) {
    st = D->dims[d];
    n = D->gdepth[d];
  } else {
    st = 2 * D->gdepth[d];
    n = (long) D->dims[d] - 2 * D->gdepth[d];
  }
  for (long i = 0; i < n; ++i) populate_rec(D, stride, region, ref + (st + i) * stride[d], d - 1, pos);
}

/* populate: brick-mpi.h:241-254 (owner shifts the reference point by +-dims[d]*stride[d]) */
static void populate(orc_decomp_t *D, const long *stride, uint64_t owner, uint64_t region, unsigned *pos) {
  long ref = 0;
  for (int d = 0; d < 3; ++d) {
    if (bs_has(owner, d + 1)) ref += (long) D->dims[d] * stride[d];
    if (bs_has(owner, -(d + 1))) ref -= (long) D->dims[d] * stride[d];
  }
  populate_rec(D, stride, region, ref, 2, pos);
}

orc_decomp_t *orc_decomp_new(const unsigned *dom, unsigned depth) {
  orc_decomp_t *D = (orc_decomp_t *) calloc(1, sizeof(orc_decomp_t));
  long stride[3], gsz = 1;
  for (int d = 0; d < 3; ++d) { /* ctor brick-mpi.h:304-316, initialize :333-337 */
    if (depth % 8 || dom[d] % 8) {
      free(D);
      return NULL;
    }
    D->dims[d] = dom[d] / 8;
    D->gdepth[d] = depth / 8;
    D->tdims[d] = D->dims[d] + 2 * D->gdepth[d];
    stride[d] = gsz;
    gsz *= D->tdims[d];
  }
  D->grid = (unsigned *) calloc((size_t) gsz, sizeof(unsigned));
  uint64_t skinlist[26];
  for (int l = 0; l < 26; ++l) skinlist[l] = orc_bitset_of(SKIN3D_GOOD[l][0], SKIN3D_GOOD[l][1], SKIN3D_GOOD[l][2]);

  unsigned pos = 1; /* factor = 1 for 4 KiB bricks on 4 KiB pages (:342-353): id 0 is the reserved null brick */
  unsigned st_pos[27];
  populate(D, stride, 0, 0, &pos); /* inner region :383 */
  st_pos[0] = pos;
  D->sep_pos[0] = pos;
  for (int l = 0; l < 26; ++l) { /* skin regions :390-398 (pad = 0 under DECOMP_PAGEUNALIGN) */
    unsigned before = pos;
    populate(D, stride, 0, skinlist[l], &pos);
    st_pos[l + 1] = pos;
    D->skin_size[l] = (long) pos - before;
  }
  D->sep_pos[1] = pos;

  uint64_t nbrs[27];
  all_neighbors(nbrs);
  int nr = 0;
  for (int n = 0; n < 27; ++n) { /* ghost regions :403-451: one entry per maximal run of matching skin regions */
    if (!nbrs[n]) continue;
    uint64_t in = orc_bitset_neg(nbrs[n]);
    int last = -1;
    for (int l = 0; l <= 26; ++l) {
      int match = (l < 26) && ((in & skinlist[l]) == in);
      if (match) {
        if (last < 0) {
          last = l;
          D->ghost[nr].neighbor = nbrs[n];
          D->skin[nr].neighbor = in;
          D->ghost[nr].skin_st = D->skin[nr].skin_st = (unsigned) l;
          D->ghost[nr].pos = pos;
          D->skin[nr].pos = st_pos[l];
        }
        populate(D, stride, nbrs[n], skinlist[l], &pos);
      } else if (last >= 0) {
        D->ghost[nr].skin_ed = D->skin[nr].skin_ed = (unsigned) l;
        D->ghost[nr].len = pos - D->ghost[nr].pos;
        D->skin[nr].len = st_pos[l] - D->skin[nr].pos;
        ++nr;
        last = -1;
      }
    }
  }
  D->nregions = nr;
  D->sep_pos[2] = pos;
  D->nbricks = pos;

  /* adjacency :266-291, :458-459: 3x3x3 grid neighbourhood, 0 outside the (ghost-inclusive) grid */
  D->adj = (unsigned *) calloc((size_t) pos * 27, sizeof(unsigned));
  for (long k = 0; k < D->tdims[2]; ++k)
    for (long j = 0; j < D->tdims[1]; ++j)
      for (long i = 0; i < D->tdims[0]; ++i) {
        unsigned id = D->grid[i + j * stride[1] + k * stride[2]];
        for (int dk = -1; dk <= 1; ++dk)
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di) {
              long x = i + di, y = j + dj, z = k + dk;
              int inside = x >= 0 && x < D->tdims[0] && y >= 0 && y < D->tdims[1] && z >= 0 && z < D->tdims[2];
              D->adj[(size_t) id * 27 + (dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] =
                  inside ? D->grid[x + y * stride[1] + z * stride[2]] : 0;
            }
      }
  return D;
}

void orc_decomp_free(orc_decomp_t *D) {
  if (!D) return;
  free(D->grid);
  free(D->adj);
  free(D);
}

/* populate(comm, bDecomp, 0, 1, coo): brick-mpi.h:730-753.  Axis d (1 = i) uses cart coordinate index 3-d; the set
 * element +d is paired with coordinate c-1 and -d with c+1 (the map is keyed by the set the REGION tables use:
 * ghost[i].neighbor for receives, skin[i].neighbor for sends, :476-485). */
void orc_rank_map(const int *cart, const int *coo, uint64_t *sets, int *ranks) {
  int n = 0;
  static const int sgn[3] = {1, 0, -1};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) {
        const int s[3] = {sgn[a], sgn[b], sgn[c]}; /* sign of set element for axis d = 1,2,3 */
        int co[3];
        for (int d = 1; d <= 3; ++d) {
          int idx = 3 - d;
          int v = coo[idx] - s[d - 1]; /* +d <-> c-1, -d <-> c+1 */
          co[idx] = ((v % cart[idx]) + cart[idx]) % cart[idx];
        }
        sets[n] = orc_bitset_of(s[0] * 1, s[1] * 2, s[2] * 3);
        ranks[n] = (co[0] * cart[1] + co[1]) * cart[2] + co[2];
        ++n;
      }
}

void orc_exchange_region(const orc_decomp_t *D, int i, double *dst, const double *src, size_t step) {
  memcpy(dst + (size_t) D->ghost[i].pos * step, src + (size_t) D->skin[i].pos * step,
         (size_t) D->ghost[i].len * step * sizeof(double));
}
