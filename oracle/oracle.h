/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's hot path: the stencil specs, the 8x8x8 brick
 * layout, init_grid, the BrickDecomp numbering and the ghost<-skin exchange contract.  It is the CHECKER of the CUDA
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  It is never shipped,
 * never measured as the product and the product never falls back to it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks every function here against (a) the unmodified reference
 * compiled into oracle/_ref (when present) and (b) the committed fixtures in tests/golden/ that
 * oracle/gen_golden.py produced from that same reference build.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_7PT = 0, ORC_MPI7PT = 1, ORC_MPI13PT = 2, ORC_MPI25PT = 3, ORC_MPI125PT = 4 };

int orc_radius(int stencil);
int orc_st_iter(int stencil); /* sweeps per exchange, stencils/fake.h:39-344 */

/* array form: out[k][j][i] = sum_t c_t * in[k+dk][j+dj][i+di] over cells [lo,hi) of an array of extents ext (i,j,k) */
int orc_sweep_array(int stencil, const long *ext, const long *lo, const long *hi, const double *in, double *out,
                    const double *coeff);

/* init_grid<3> (bricksetup.h:73-90): grid[p]=p, adjacency by +-stride, ids outside [0,n) -> 0 */
void orc_init_grid(const long *dimlist, unsigned *grid, unsigned *adj);

/* copyToBrick / copyFromBrick (bricksetup.h:139-221); dir 0 = array->brick, 1 = brick->array.  Layout of one brick is
 * row-major [k][j][i] (fold 8 and fold 4,8: brick.h:234-246), element (k,j,i) of brick b at dat[b*step+off+64k+8j+i] */
void orc_copy_brick(int dir, const long *dimlist, const long *padding, const long *ghost, double *arr,
                    const unsigned *grid, double *dat, size_t step, size_t off);

/* brick form through the adjacency list, the way the accessor resolves bIn[b][k+dk][j+dj][i+di] (brick.h:234-246) */
int orc_sweep_brick(int stencil, const unsigned *grid, const long *sb, const long *lo, const long *hi,
                    const unsigned *adj, const double *dat_in, size_t step_in, size_t off_in, double *dat_out,
                    size_t step_out, size_t off_out, const double *coeff);

/* BrickDecomp<3,8,8,8>(dims, depth) + initialize(skin3d_good), DECOMP_PAGEUNALIGN (brick-mpi.h:304-460) */
typedef struct {
  uint64_t neighbor; /* BitSet.set of the neighbour this region talks to */
  unsigned skin_st, skin_ed, pos, len;
} orc_region_t;

typedef struct {
  unsigned dims[3], gdepth[3], tdims[3]; /* in bricks */
  unsigned nbricks, sep_pos[3];
  unsigned *grid; /* tdims[0]*tdims[1]*tdims[2] */
  unsigned *adj;  /* nbricks*27 */
  int nregions;
  orc_region_t ghost[64], skin[64];
  long skin_size[26];
} orc_decomp_t;

orc_decomp_t *orc_decomp_new(const unsigned *dom_cells, unsigned depth_cells);
void orc_decomp_free(orc_decomp_t *d);

/* BitSet helpers (bitset.h:19-128) */
uint64_t orc_bitset_of(int a, int b, int c);     /* signed axis ids, 0 = unused */
uint64_t orc_bitset_neg(uint64_t s);             /* operator! */

/* populate(): neighbour set -> rank for a periodic Cartesian grid (brick-mpi.h:730-753 over MPI_Cart_rank).
 * cart[0] is the slowest (k) extent.  out_sets/out_ranks have 27 entries in allneighbors order. */
void orc_rank_map(const int *cart, const int *coo, uint64_t *out_sets, int *out_ranks);

/* exchange contract (brick-mpi.h:466-495): ghost[i] of `dst` <- skin[i] of the peer storage */
void orc_exchange_region(const orc_decomp_t *d, int i, double *dst_dat, const double *src_dat, size_t step);

#ifdef __cplusplus
}
#endif
#endif
