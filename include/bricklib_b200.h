/*
 * bricklib_b200.h -- C ABI of libbrick_b200.so: the B200 (sm_100a) implementation of bricklib's hot path
 * (FP64 7/13/25/125-point stencils over 8x8x8 bricks + ghost-zone exchange).
 *
 * The reference (CtopCsUtahEdu/bricklib) has no runtime FFI: its seam is source level -- the brick(...) macro inside a
 * __global__ function, the kernel launch, BrickDecomp::exchange and the movBrick* helpers.
 * Each entry point below names the reference interface it stands in for (file:line under the reference tree).
 * The C++ headers next to this file (brick.h, brick-mpi.h, ...) re-create the reference's template surface on top of
 * these calls; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only.  "dev" pointers are CUDA device pointers on the current device,
 * "host" pointers are ordinary memory.  Every call returns 0 (BK_OK) or a negative code; bk_last_error() gives text.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Nothing is thread-hostile: state is
 * per call or per handle.  There is NO CPU fallback: without a CUDA device every compute call fails with BK_ECUDA.
 */
#ifndef BRICKLIB_B200_H
#define BRICKLIB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_OK 0
#define BK_EINVAL (-1)   /* bad argument */
#define BK_ECUDA (-2)    /* CUDA runtime error (text in bk_last_error) */
#define BK_ENOMEM (-3)
#define BK_EUNSUPPORTED (-4)

/* stencil ids = the reference's stencil specs, stencils/{7pt,mpi7pt,mpi13pt,mpi25pt,mpi125pt}.py */
enum { BK_ST_7PT = 0, BK_ST_MPI7PT = 1, BK_ST_MPI13PT = 2, BK_ST_MPI25PT = 3, BK_ST_MPI125PT = 4, BK_ST_COUNT = 5 };

/* kernel families (bk_stencil_apply picks; tests can force one through `flags`) */
#define BK_KERNEL_AUTO 0u
#define BK_KERNEL_BRICK 1u  /* one CTA per brick, neighbours through adj: works for any brick set */
#define BK_KERNEL_TILED 2u  /* CTA marches a column of bricks of a dense grid box: the fast path */

const char *bk_version(void);
const char *bk_last_error(void);

/* ---- stencil metadata (stencils/fake.h:11-33, :39-344) --------------------------------------------------------- */
int bk_stencil_radius(int stencil);  /* 1,1,2,4,2 */
int bk_stencil_st_iter(int stencil); /* sweeps per ghost exchange: 8,8,4,2,4 (= ghost depth 8 / radius) */
int bk_stencil_points(int stencil);  /* 7,7,13,25,125 */
int bk_stencil_fused_steps(int stencil); /* `steps` of bk_stencil_advance that is fastest: 2,2,1,1,1 (radius 2 can fuse, but its
                                            two-step kernel is slower than two sweeps) */
/* Which kernel runs bk_stencil_advance(steps = 2) for the radius-1 stars (7pt, mpi7pt and compiled radius-1 stars):
 *   BK_FUSED_STAGED   (default) k_star2: two stencil stages per plane, the intermediate plane in shared memory;
 *   BK_FUSED_COMPOSED the composed operator S(S u), a 25-point diamond, as ONE radius-2 marching update
 *                     (bricklib_b200/csrc/bk_diamond.h) -- same semantics incl. the zero intermediate outside the grid;
 *                     4x4-brick tiles, two CTAs per SM;
 *   BK_FUSED_COMPOSED_WIDE the same update on 8x4-brick tiles, one CTA per SM with register re-balancing.
 * Process-wide; the environment variable BK_FUSED_VARIANT=staged|composed|wide sets the initial value.  _set returns the
 * previous value, or BK_EINVAL. */
#define BK_FUSED_STAGED 0
#define BK_FUSED_COMPOSED 1
#define BK_FUSED_COMPOSED_WIDE 2
int bk_stencil_fused_variant_set(int variant);
int bk_stencil_fused_variant_get(void);

/* ---- device plumbing (stands in for include/brick-gpu.h:43-103 movBrickInfo/movBrickStorage, cudaarray.h:11-31) - */
int bk_device_count(int *n);
int bk_set_device(int dev);
/* bind the calling host thread to the CPUs local to the current device (NUMA: call before allocating pinned memory);
 * best effort, BK_EUNSUPPORTED when the platform does not say */
int bk_bind_host_to_device(void);
int bk_dev_alloc(void **dev, size_t bytes); /* cudaMalloc: 256-B aligned, exportable with bk_ipc_export */
int bk_dev_free(void *dev);
int bk_dev_memset(void *dev, int byte, size_t bytes, void *stream);
int bk_host_alloc(void **host, size_t bytes); /* pinned */
int bk_host_free(void *host);
int bk_memcpy_h2d(void *dev, const void *host, size_t bytes, void *stream);
int bk_memcpy_d2h(void *host, const void *dev, size_t bytes, void *stream);
int bk_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int bk_stream_create(void **stream);
int bk_stream_create_priority(void **stream, int high); /* high != 0: the device's highest priority (exchange stream) */
int bk_stream_destroy(void *stream);
int bk_stream_sync(void *stream);
int bk_device_sync(void);
/* events: timing on the launching stream (stencils/stencils_cu.h:13-28 cutime_func) and cross-stream ordering */
int bk_event_create(void **ev);
int bk_event_destroy(void *ev);
int bk_event_record(void *ev, void *stream);
int bk_event_sync(void *ev);
int bk_event_elapsed_ms(void *start, void *stop, float *ms);
int bk_stream_wait_event(void *stream, void *ev);

/* ---- grids and adjacency (host side, pure integer work) -------------------------------------------------------- */
/* init_grid<3> (include/bricksetup.h:73-90): grid[p] = p, adj by +-stride, out of linear range -> 0.
 * dimlist = bricks per axis, i first.  grid: prod(dimlist) unsigned; adj: prod(dimlist)*27 unsigned. */
int bk_init_grid(const long *dimlist, unsigned *grid_host, unsigned *adj_host);

/* BrickDecomp<3,8,8,8>(dims, depth).initialize(skin3d_good) (include/brick-mpi.h:304-460, src/brick-mpi.cpp:25-52),
 * DECOMP_PAGEUNALIGN numbering (the weak/strong CUDA drivers' configuration, weak/CMakeLists.txt:32). */
typedef struct bk_decomp bk_decomp_t;
typedef struct {
  uint64_t neighbor;          /* BitSet.set (include/bitset.h): bit a = +axis a, bit 31+a = -axis a, axis 1 = i */
  unsigned skin_st, skin_ed;  /* range in the skin list */
  unsigned pos, len;          /* first brick id and number of bricks (contiguous) */
  unsigned first_pad, last_pad; /* always 0 for 4 KiB bricks; kept for BrickDecomp::g_region layout parity */
} bk_region_t;

int bk_decomp_create(bk_decomp_t **d, const unsigned *dom_cells, unsigned depth_cells);
int bk_decomp_destroy(bk_decomp_t *d);
unsigned bk_decomp_nbricks(const bk_decomp_t *d);                /* BrickInfo::nbricks, includes null brick 0 */
int bk_decomp_sep_pos(const bk_decomp_t *d, unsigned *sep3);     /* BrickDecomp::sep_pos */
int bk_decomp_tdims(const bk_decomp_t *d, unsigned *tdims3);     /* bricks per axis incl. ghost shell */
const unsigned *bk_decomp_grid(const bk_decomp_t *d);            /* BrickDecomp::operator[]: [k][j][i] -> id */
const unsigned *bk_decomp_adj(const bk_decomp_t *d);             /* BrickInfo<3>::adj, nbricks*27 */
int bk_decomp_nregions(const bk_decomp_t *d);                    /* ghost.size() == skin.size() (42) */
int bk_decomp_region(const bk_decomp_t *d, int which /*0 ghost,1 skin*/, int i, bk_region_t *out);
int bk_decomp_skin_size(const bk_decomp_t *d, long *out26);      /* BrickDecomp::skin_size */
/* brick-id lists for overlap: which 0 = inner (reads no ghost), 1 = skin, 2 = ghost; returns count, fills if non-NULL */
long bk_decomp_list(const bk_decomp_t *d, int which, unsigned *ids_host);

/* populate(comm, bDecomp, 0, 1, coo) (include/brick-mpi.h:730-753) for a periodic Cartesian grid `cart` (cart[0]
 * slowest, MPI order) seen from coordinates `coo`: 27 (set, rank) pairs in allneighbors order. */
int bk_rank_map(const int *cart, const int *coo, uint64_t *sets27, int *ranks27);
/* Z-Morton helpers for the strong driver (include/zmort.h:18-105) */
unsigned long bk_zmort_encode(const unsigned long *coord3);
int bk_zmort_decode(unsigned long id, unsigned long *coord3);

/* ---- stitching a rank's subdomains into one brick grid (strong scaling) -------------------------------------------
 * The reference's CPU strong driver aliases same-rank ghost zones onto their owners' skins with mmap views
 * (strong/main.cpp:205-262); its CUDA driver copies them with cudaCopy links (strong/main.cu:76-83, :188-247).  Here a
 * rank's subdomains -- Z-Morton ids [first, first+count), strong/args.cpp:104-113 -- are swept as ONE dense brick grid
 * whose entries are global brick ids  q * nbricks + local id  (q = id - first) into a subdomain-major allocation:
 * interior positions name the owning subdomain's brick, the one-brick shell names the ghost brick of the nearest
 * boundary subdomain, and along an axis where the box spans the whole periodic arrangement the shell aliases the
 * interior bricks of the far side.  Same-GPU ghost regions are then never copied or recomputed; only regions on the
 * box surface (bk_stitch_region_needed) take part in the exchange. */
typedef struct {
  unsigned long first, count; /* the Z-Morton section */
  unsigned long lo[3];        /* lowest subdomain coordinate of the bounding box, i first */
  long n[3];                  /* subdomains per axis */
  int wrap[3];                /* the box spans the whole periodic arrangement along this axis */
  int is_box;                 /* the section fills its bounding box (always for power-of-two rank counts) */
} bk_stitch_box_t;
int bk_stitch_box(unsigned long first, unsigned long count, unsigned long subdim, bk_stitch_box_t *box);
/* extents of the stitched grid in bricks: n[a] * (bricks per subdomain edge) + 2 */
int bk_stitch_dims(const bk_decomp_t *d, const bk_stitch_box_t *box, unsigned *dims3);
/* fill grid_host[k][j][i] (bk_stitch_dims extents) with global brick ids; needs is_box and a one-brick ghost shell */
int bk_stitch_grid(const bk_decomp_t *d, const bk_stitch_box_t *box, unsigned *grid_host);
/* 1 if ghost region `region` (index into BrickDecomp::ghost) of subdomain `sub_id` lies on the box surface, 0 if the
 * stitched grid never reads it, negative on bad arguments */
int bk_stitch_region_needed(const bk_decomp_t *d, const bk_stitch_box_t *box, unsigned long sub_id, int region);

/* ---- array <-> brick on the device (include/bricksetup.h:139-221, include/brickcompare.h:30-57) ---------------- */
/* dimlist/padding/ghost in cells, i first, meaning exactly as in copyToBrick<3>(dimlist, padding, ghost, arr, grid, b) */
int bk_copy_to_brick(const long *dimlist, const long *padding, const long *ghost, const double *arr_dev,
                     const unsigned *grid_dev, double *dat_dev, size_t step, void *stream);
int bk_copy_from_brick(const long *dimlist, const long *padding, const long *ghost, double *arr_dev,
                       const unsigned *grid_dev, const double *dat_dev, size_t step, void *stream);
/* compareBrick with an explicit tolerance: counts cells with |a-b| >= tol && |a-b| >= (|a|+|b|)*tol; also returns the
 * largest relative difference.  Synchronises `stream`. */
int bk_compare_brick(const long *dimlist, const long *padding, const long *ghost, const double *arr_dev,
                     const unsigned *grid_dev, const double *dat_dev, size_t step, double tol,
                     unsigned long long *mismatches, double *max_rel, void *stream);

/* Synthetic field straight into the bricks (the reference fills host arrays with randomArray, src/multiarray.cpp:33-45,
 * then copyToBrick): every cell of every non-null brick named by the dense id array `grid_dev` (gdims = bricks per axis,
 * i first) becomes a counter-based hash of its GLOBAL periodic cell coordinate,
 *     value = bk_synthetic_value(seed, ((z * global[1] + y) * global[0] + x)),   (x,y,z) = (origin + cell) mod global,
 * U[0,1) with 53 random bits (splitmix64).  origin_cells = global coordinate of cell 0 of grid position (0,0,0), may be
 * negative (a ghost shell wraps periodically).  Position-addressable: any rank, and a CPU checker, can evaluate any cell. */
int bk_fill_synthetic(const unsigned *grid_dev, const unsigned *gdims, const long *origin_cells, const long *global_cells,
                      uint64_t seed, double *dat_dev, size_t step, void *stream);
double bk_synthetic_value(uint64_t seed, uint64_t linear_cell); /* the same hash on the host */
/* compareBrick between two brick storages over the bricks of box [lo,hi) of `grid_dev` (same tolerance rule as
 * bk_compare_brick).  Synchronises `stream`. */
int bk_compare_storage(const unsigned *grid_dev, const unsigned *gdims, const unsigned *lo, const unsigned *hi,
                       const double *a_dev, size_t a_step, const double *b_dev, size_t b_step, double tol,
                       unsigned long long *mismatches, double *max_rel, void *stream);

/* ---- the stencil sweep ---------------------------------------------------------------------------------------- */
/* One field pair = the two Brick<Dim<8,8,8>,Dim<4,8>> objects a reference kernel receives by value
 * (include/brick.h:353-395): dat pointers already include the field offset, steps are in elements. */
typedef struct {
  const unsigned *adj; /* dev, BrickInfo<3>::adj */
  const double *in;    /* dev, Brick::dat of the input  */
  size_t in_step;      /* Brick::step */
  double *out;         /* dev, Brick::dat of the output */
  size_t out_step;
} bk_field_t;

/* Replaces the launch  brick_kernel<<<dim3(strideb...),32>>>(grid, in, out, stride)  (weak/main.cu:35-43, :277-282;
 * stencils/3axis.cu:28-37, :163-169): out = stencil(in) for every brick of the half-open brick box [lo,hi) of the dense
 * id array `grid_dev` (gdims = bricks per axis, i first).  coeff_host: 7 doubles for BK_ST_7PT (coeff[0..6] of
 * single/cpu.cpp:11-17), ignored (may be NULL) otherwise.  Asynchronous on `stream`. */
int bk_stencil_apply(int stencil, const bk_field_t *f, const unsigned *grid_dev, const unsigned *gdims,
                     const unsigned *lo, const unsigned *hi, const double *coeff_host, unsigned flags, void *stream);
/* Split sweep for overlapping the ghost exchange (replaces the blocking exchange-then-compute order of
 * weak/main.cu:251-282).  The bricks inside the half-open box [ready_lo,ready_hi) are final already (the subdomain's own
 * bricks); the others (the ghost shell) are final once the exchange has finished.  part = BK_PART_READY launches the
 * CTAs of the box's tile decomposition that read only ready bricks, BK_PART_REST all the other CTAs; together they
 * cover [lo,hi) exactly once, with the same tiles as a whole-box launch (no thin slab launches).  Enqueue READY on the
 * compute stream at once and REST on a stream that waits for the exchange.  BK_EUNSUPPORTED when the storage layout
 * rules out the marching kernel (use bk_stencil_apply after the exchange instead). */
#define BK_PART_ALL 0
#define BK_PART_READY 1
#define BK_PART_REST 2
/* OR-ed into READY / REST (both launches of a split must agree): the brick layers whose CTAs read not-yet-ready bricks
 * get thin k segments of their own, which raises the READY share of the sweep from ~45 % to ~70 % at 512^3 at the price
 * of a few per cent more halo planes.  Worth it when the exchange is slow (crosses NVLink), not for self-exchanges. */
#define BK_PART_THIN 4
/* OR-ed into `part` of bk_stencil_apply_part / bk_stencil_advance / bk_stencil_def_advance: the dense grid IS the topology
 * of this launch -- skip the grid-vs-adjacency check described below.  For grids built on purpose to differ from the
 * per-subdomain adjacency: the stitched super grid of the strong driver (bk_stitch_grid) aliases ghost positions onto
 * other subdomains' bricks, which is exactly what the adjacency list of ONE subdomain cannot say. */
#define BK_PART_GRID_TOPOLOGY 8
int bk_stencil_apply_part(int stencil, const bk_field_t *f, const unsigned *grid_dev, const unsigned *gdims,
                          const unsigned *lo, const unsigned *hi, const double *coeff_host, const unsigned *ready_lo,
                          const unsigned *ready_hi, int part, void *stream);
/* Grid vs adjacency.  The marching kernels (the fast path of every call above and below) take neighbour ids from the
 * dense `grid_dev` array and read brick 0 outside it, where the reference's accessor follows BrickInfo::adj
 * (include/brick.h:234-246).  Before a (adj, grid, box) triple is first swept that way, a small kernel verifies
 * adj[grid[p]][s] == grid[p + delta_s] (0 outside the grid) for every position and neighbour slot the launch reads, and
 * the verdict is cached (one stream synchronisation, once).  On a mismatch -- periodic or otherwise hand-made adjacency
 * -- BK_KERNEL_AUTO uses the adjacency-following family instead, BK_KERNEL_TILED and the split / fused calls return
 * BK_EUNSUPPORTED: never silently different numbers.  bk_dev_free forgets the verdicts of the freed address; call
 * bk_adjacency_forget (NULL = all) after rewriting an adjacency list or grid in place. */
int bk_adjacency_forget(const void *adj_or_grid_dev);
/* `steps` time steps in ONE pass over HBM (temporal blocking; steps = 1 or 2).  Between two ghost exchanges the
 * reference applies ST_ITER sweeps that depend on nothing outside the subdomain's (ghost-inclusive) grid
 * (weak/main.cu:275-285), so consecutive sweeps can be fused.  steps = 2 is equivalent to
 *     bk_stencil_apply(stencil, {in -> tmp}, whole grid);  bk_stencil_apply(stencil, {tmp -> out}, [lo,hi));
 * except that tmp lives in shared memory: the intermediate is evaluated at every in-grid cell the second step reads and
 * is zero outside the grid (the null brick).  HBM traffic: 16 B per point per `steps` steps.  part/ready_* as in
 * bk_stencil_apply_part (BK_PART_ALL: ready_* may be NULL).  BK_EUNSUPPORTED for stencils or layouts without a fused
 * kernel (radius 4, cube): the caller then issues single steps. */
int bk_stencil_advance(int stencil, int steps, const bk_field_t *f, const unsigned *grid_dev, const unsigned *gdims,
                       const unsigned *lo, const unsigned *hi, const double *coeff_host, const unsigned *ready_lo,
                       const unsigned *ready_hi, int part, void *stream);
/* bk_stencil_advance with THE EXCHANGE INSIDE THE SWEEP: ghost brick id ghost_lo + g of the input field is read from
 * remap_dev[g] -- the address (peer access / CUDA IPC / own storage) of the neighbour's skin brick it mirrors, what
 * BrickDecomp::exchange (include/brick-mpi.h:466-495) would have copied into it -- instead of f->in + id * in_step.  Only
 * the launches that read freshly exchanged ghosts need it (the REST half of a period's first pass).  BK_EUNSUPPORTED for
 * the staged two-step kernel and the developer geometries. */
int bk_stencil_advance_remote(int stencil, int steps, const bk_field_t *f, const unsigned *grid, const unsigned *gdims,
                              const unsigned *lo, const unsigned *hi, const double *coeff, const unsigned *ready_lo,
                              const unsigned *ready_hi, int part, const double *const *remap_dev, unsigned ghost_lo,
                              unsigned ghost_n, void *stream);
/* same over an explicit list of brick ids (inner / skin / ghost lists for overlap; any adjacency-defined set) */
int bk_stencil_apply_list(int stencil, const bk_field_t *f, const unsigned *ids_dev, size_t n,
                          const double *coeff_host, void *stream);
/* The reference's array-layout baseline kernel (arr_kernel, weak/main.cu:27-33; d3pt7_arr, stencils/3axis.cu): the same
 * stencil over a plain padded array, out[p] = sum_t c_t in[p + offset_t] for the cell box lo <= (i,j,k) < hi; extent =
 * cells per axis of the allocation (i first).  The box must keep the stencil radius inside the allocation. */
int bk_array_stencil_apply(int stencil, const double *in_dev, double *out_dev, const long *extent, const long *lo,
                           const long *hi, const double *coeff_host, void *stream);
/* strong/main.cu:85-99 brick_kernel over `nsub` subdomains that share grid/adj: fields[s] per subdomain (host array) */
int bk_stencil_apply_multi(int stencil, const bk_field_t *fields_dev, unsigned nsub, const unsigned *grid_dev,
                           const unsigned *gdims, const unsigned *lo, const unsigned *hi, const double *coeff_host,
                           void *stream);
/* ---- stencils lowered from tap lists ------------------------------------------------------------------------------
 * The reference turns a stencil expression (the stencils/ scripts: Grid / Index / ConstRef arithmetic) into code at build time
 * with codegen/vecscatter (backend table vecscatter:82-109, CUDA backend codegen/st/codegen/backend/cuda.py).  Here the
 * expression is lowered to its tap list -- out(i,j,k) = sum_t c_t * in(i+di_t, j+dj_t, k+dk_t) -- by the host side
 * (bricklib_b200/dsl.py evaluates the same scripts) and bk_stencil_compile picks the kernel family:
 *   BK_KIND_STAR  taps on the axes only, radius <= 4, ANY coefficients  -> marching kernels (k_star / k_star2)
 *   BK_KIND_CUBE  radius <= 2, coefficient a function of the sorted (|di|,|dj|,|dk|) -> marching cube kernel
 *   BK_KIND_GENERATED  anything else with radius <= 4: the library EMITS CUDA source for a marching kernel specialised to
 *                 the tap pattern (tap loop unrolled into straight-line FMAs, coefficient values stay kernel
 *                 parameters), compiles it for sm_100a with NVRTC and launches it through the same CTA enumeration as
 *                 the built-in kernels -- the role codegen/vecscatter:145-175 + codegen/st/codegen/backend/cuda.py play
 *                 in the reference.  bk_stencil_def_source returns the text.
 *   BK_KIND_TAPS  fallback when NVRTC is unavailable (or BK_NO_CODEGEN is set): per-brick kernel walking a tap table
 * Repeated offsets are merged, zero coefficients dropped.  BK_EUNSUPPORTED for radius > 4. */
typedef struct {
  int di, dj, dk;
  double c;
} bk_tap_t;
typedef struct bk_stencil_def bk_stencil_def_t;
#define BK_KIND_STAR 0
#define BK_KIND_CUBE 1
#define BK_KIND_TAPS 2
#define BK_KIND_GENERATED 3 /* a marching kernel GENERATED for this tap pattern and compiled with NVRTC (see below) */
int bk_stencil_compile(bk_stencil_def_t **def, const bk_tap_t *taps_host, int ntaps);
/* The two non-linearities a stencil script may carry and a kernel applies for free (stencils/cond.py uses both):
 *     out = post( sum_t c_t * pre( in(. + d_t) ) )
 * pre clamps every value read (cond.py: max(in, 0.0)), post clamps the sum (cond.py: If(calc > 0, calc, -calc) = abs).
 * NULL or op BK_OP_NONE = identity.  Such stencils run on a generated kernel (BK_KIND_GENERATED) or the tap-table kernel. */
#define BK_OP_NONE 0
#define BK_OP_MAX 1 /* max(x, c) */
#define BK_OP_MIN 2 /* min(x, c) */
#define BK_OP_ABS 3 /* |x| */
typedef struct {
  int op;
  double c;
} bk_pointwise_t;
int bk_stencil_compile_pointwise(bk_stencil_def_t **def, const bk_tap_t *taps_host, int ntaps, const bk_pointwise_t *pre,
                                 const bk_pointwise_t *post);
int bk_stencil_def_destroy(bk_stencil_def_t *def);
/* the generated CUDA source of a BK_KIND_GENERATED stencil, NUL-terminated, truncated to cap; *len = full length.  For
 * other kinds: BK_EUNSUPPORTED and the reason no kernel was generated (empty for star / cube). */
int bk_stencil_def_source(const bk_stencil_def_t *def, char *buf, size_t cap, size_t *len);
/* any out pointer may be NULL; st_iter = 8 / radius (sweeps per exchange at ghost depth 8); fused_steps as bk_stencil_fused_steps */
int bk_stencil_def_info(const bk_stencil_def_t *def, int *kind, int *radius, int *ntaps, int *st_iter, int *fused_steps);
/* = bk_stencil_apply / bk_stencil_advance for a compiled stencil (flags: BK_KERNEL_*) */
int bk_stencil_def_apply(const bk_stencil_def_t *def, const bk_field_t *f, const unsigned *grid_dev, const unsigned *gdims,
                         const unsigned *lo, const unsigned *hi, unsigned flags, void *stream);
int bk_stencil_def_advance(const bk_stencil_def_t *def, int steps, const bk_field_t *f, const unsigned *grid_dev,
                           const unsigned *gdims, const unsigned *lo, const unsigned *hi, const unsigned *ready_lo,
                           const unsigned *ready_hi, int part, unsigned flags, void *stream);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
unsigned long long bk_launch_count(void);

/* ---- ghost exchange --------------------------------------------------------------------------------------------- */
/* A batch of contiguous copies executed by ONE kernel; replaces cudaCopy<<<links,64>>> (strong/main.cu:76-83) and,
 * with src pointing into a peer GPU's storage, the MPI_Isend/Irecv pairs of BrickDecomp::exchange
 * (include/brick-mpi.h:466-495).  bytes must be multiples of 16, pointers 16-B aligned. */
typedef struct {
  const void *src;
  void *dst;
  size_t bytes;
} bk_seg_t;
typedef struct bk_xplan bk_xplan_t;
int bk_xplan_create(bk_xplan_t **plan, const bk_seg_t *segs_host, int nseg);
/* Array-layout exchange -- the reference's exchangeArr (include/array-mpi.h:146-213: pack 26 regions, Isend/Irecv,
 * unpack) as the same single pull kernel over strided 3-D boxes: cell (i,j,k) of a box lives at
 * ptr[i + j*stride[0] + k*stride[1]] (elements); src may be a peer GPU's array.  No pack or unpack buffer exists.
 * Runs through bk_xplan_run / _run_sync / _run_gate like a range plan. */
typedef struct {
  const double *src;
  double *dst;
  long n[3];          /* cells per axis, i first */
  long src_stride[2]; /* elements between consecutive j rows / k planes of the source array */
  long dst_stride[2];
} bk_box_t;
int bk_xplan_create_boxes(bk_xplan_t **plan, const bk_box_t *boxes_host, int nbox);
int bk_xplan_destroy(bk_xplan_t *plan);
size_t bk_xplan_bytes(const bk_xplan_t *plan);
/* Launch shape of the pull kernel.  Default (0, 0): up to 8 CTAs of 256 threads per SM -- the fastest copy on an idle
 * GPU.  A narrow shape (e.g. 32 CTAs of 1024 threads) confines the pull to a few SMs: the marching sweep kernels
 * allocate an SM's whole register file, so pull CTAs never co-reside with them; a narrow pull leaves the other SMs to
 * the overlapped sweep instead of time-sharing all of them. */
int bk_xplan_set_shape(bk_xplan_t *plan, int ctas, int threads_per_cta);
/* run the plan on `stream`.  wait/signal (optional, may be NULL/0) implement the cross-process handshake:
 * before copying, spin until every wait_flags[i] >= epoch; after copying (and a system fence) store epoch to every
 * signal_flags[i] (peer-visible device memory). */
int bk_xplan_run(bk_xplan_t *plan, void *stream);
int bk_xplan_run_sync(bk_xplan_t *plan, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                      int nsignal, uint64_t epoch, void *stream);
/* same, but the LAST CTA of the pull kernel itself publishes completion: it stores epoch to the signal flags and to
 * `gate` (optional extra flag in local device memory) -- no separate signal launch. */
int bk_xplan_run_gate(bk_xplan_t *plan, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                      int nsignal, uint64_t *gate, uint64_t epoch, void *stream);
/* the same plan executed by the COPY ENGINES (one cudaMemcpyAsync per segment on a few internal lane streams, forked
 * from and joined to `stream`), bracketed by the flag wait / flag signal kernels.  Takes no SM from the sweep kernels --
 * they allocate an SM's whole register file, so a pull KERNEL cannot co-reside with them and the exchange would
 * time-share the SMs instead of overlapping.  Same handshake semantics as bk_xplan_run_sync. */
int bk_xplan_run_ce(bk_xplan_t *plan, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                    int nsignal, uint64_t epoch, void *stream);
/* tiny kernels for the handshake on a stream: store `value` to n flags / spin until n flags >= value */
int bk_flags_signal(uint64_t *const *flags, int n, uint64_t value, void *stream);
int bk_flags_wait(const uint64_t *const *flags, int n, uint64_t value, void *stream);

/* CUDA IPC so each process (one per GPU) can map its neighbours' brick storage and flags */
#define BK_IPC_HANDLE_BYTES 64
int bk_ipc_export(void *dev, unsigned char *handle64);
int bk_ipc_open(const unsigned char *handle64, void **dev);
int bk_ipc_close(void *dev);
int bk_peer_enable(int peer_dev); /* single-process multi-GPU: cudaDeviceEnablePeerAccess */

#ifdef __cplusplus
}
#endif
#endif
