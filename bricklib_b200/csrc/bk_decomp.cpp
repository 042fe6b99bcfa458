// bk_decomp.cpp -- host-side integer work of the hot path: brick numbering, adjacency, region tables, rank maps.
//
// Re-creates, from the published algorithm, what the reference computes in
//   include/bricksetup.h:73-90      init_grid<3>
//   include/brick-mpi.h:304-460     BrickDecomp<3,8,8,8>::BrickDecomp / initialize (DECOMP_PAGEUNALIGN numbering)
//   src/brick-mpi.cpp:9-52          allneighbors order, skin3d_good order
//   include/brick-mpi.h:730-753     populate() -> rank_map for a periodic Cartesian communicator
//   include/zmort.h:18-105          Z-Morton ids (strong driver)
// The numbering must be reproduced exactly (ghost[i]/skin[i] pairing is the exchange contract), so the tables are
// checked bit-for-bit against the reference in tests/test_decomp.py.  The formulation here is sign-vector based: a
// region is a vector s in {-1,0,+1}^3 (lower skin layer / middle / upper skin layer per axis) instead of a BitSet.
#include "bk_common.h"
#include <array>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

using Sign3 = std::array<int, 3>;  // per axis (0 = i): -1 lower, 0 middle, +1 upper

uint64_t to_bitset(const Sign3 &s) {  // include/bitset.h:22-27: +a -> bit a, -a -> bit 31+a (a = axis+1)
  uint64_t b = 0;
  for (int a = 0; a < 3; ++a) {
    if (s[a] > 0) b |= 1ull << (a + 1);
    if (s[a] < 0) b |= 1ull << (31 + a + 1);
  }
  return b;
}

// order of the 26 surface regions; fixes where each skin brick lives in memory (src/brick-mpi.cpp:25-52)
const Sign3 kSkinOrder[26] = {
    {+1, 0, 0},  {+1, 0, -1},  {+1, +1, -1}, {+1, +1, 0},  {+1, +1, +1}, {0, +1, +1}, {0, +1, 0},   {0, +1, -1}, {-1, +1, -1},
    {-1, +1, 0}, {-1, +1, +1}, {-1, 0, +1},  {-1, 0, 0},   {0, 0, -1},   {-1, 0, -1}, {-1, -1, -1}, {-1, -1, 0},
    {-1, -1, +1}, {0, -1, +1}, {0, -1, 0},   {0, -1, -1},  {+1, -1, -1}, {+1, -1, 0}, {+1, -1, +1}, {+1, 0, +1}, {0, 0, +1}};

}  // namespace

struct bk_decomp {
  unsigned D[3], G[3], T[3];
  unsigned nbricks = 0, sep[3] = {0, 0, 0};
  std::vector<unsigned> grid, adj;
  std::vector<bk_region_t> ghost, skin;
  long skin_size[26];

  size_t slot(long i, long j, long k) const { return (size_t) i + (size_t) T[0] * ((size_t) j + (size_t) T[1] * k); }

  // number every brick of `region` as seen from the subdomain shifted by `owner` whole subdomains
  void number_region(const Sign3 &owner, const Sign3 &region, unsigned &next) {
    long lo[3], n[3];
    for (int a = 0; a < 3; ++a) {
      long base = region[a] < 0 ? G[a] : region[a] > 0 ? D[a] : 2 * G[a];
      n[a] = region[a] != 0 ? (long) G[a] : (long) D[a] - 2 * (long) G[a];
      lo[a] = base + (long) owner[a] * D[a];
    }
    for (long k = 0; k < n[2]; ++k)
      for (long j = 0; j < n[1]; ++j)
        for (long i = 0; i < n[0]; ++i) grid[slot(lo[0] + i, lo[1] + j, lo[2] + k)] = next++;
  }
};

extern "C" {

int bk_init_grid(const long *dl, unsigned *grid, unsigned *adj) {
  BK_REQUIRE(dl && grid && adj, "null argument");
  const long n = dl[0] * dl[1] * dl[2];
  BK_REQUIRE(n > 0 && n < (1l << 32), "bad grid size");
  for (long p = 0; p < n; ++p) grid[p] = (unsigned) p;
  const long st[3] = {1, dl[0], dl[0] * dl[1]};
#pragma omp parallel for
  for (long p = 0; p < n; ++p) {
    unsigned *row = adj + p * 27;
    int s = 0;
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di, ++s) {
          // the reference tests only the linear range of the id array (bricksetup.h:37-43)
          const long q = p + di * st[0] + dj * st[1] + dk * st[2];
          row[s] = (q >= 0 && q < n) ? (unsigned) q : 0u;
        }
  }
  return BK_OK;
}

int bk_decomp_create(bk_decomp_t **out, const unsigned *dom, unsigned depth) {
  BK_REQUIRE(out && dom, "null argument");
  BK_REQUIRE(depth > 0 && depth % bk::BRICK_EDGE == 0, "ghost depth must be a positive multiple of 8");
  bk_decomp *d = new bk_decomp();
  size_t total = 1;
  for (int a = 0; a < 3; ++a) {
    if (dom[a] % bk::BRICK_EDGE || dom[a] / bk::BRICK_EDGE < 2 * (depth / bk::BRICK_EDGE)) {
      delete d;
      bk::set_error("bk_decomp_create: extent %u must be a multiple of 8 and at least twice the ghost depth", dom[a]);
      return BK_EINVAL;
    }
    d->D[a] = dom[a] / bk::BRICK_EDGE;
    d->G[a] = depth / bk::BRICK_EDGE;
    d->T[a] = d->D[a] + 2 * d->G[a];
    total *= d->T[a];
  }
  if (total >= (1ull << 32)) {
    delete d;
    bk::set_error("bk_decomp_create: too many bricks for 32-bit ids");
    return BK_EINVAL;
  }
  d->grid.assign(total, 0u);

  const Sign3 none = {0, 0, 0};
  unsigned next = 1;  // id 0 = the null brick every out-of-domain adjacency entry points at (brick-mpi.h:353,:275)
  d->number_region(none, none, next);
  d->sep[0] = next;
  unsigned start[27];
  for (int l = 0; l < 26; ++l) {
    start[l] = next;
    d->number_region(none, kSkinOrder[l], next);
    d->skin_size[l] = (long) next - (long) start[l];
  }
  start[26] = next;
  d->sep[1] = next;

  // ghost shells, neighbour by neighbour in (+,0,-) order with axis i outermost (src/brick-mpi.cpp:9-23)
  const int pick[3] = {+1, 0, -1};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) {
        const Sign3 nb = {pick[a], pick[b], pick[c]};
        if (nb == none) continue;
        const Sign3 mirror = {-nb[0], -nb[1], -nb[2]};
        int run = -1;
        for (int l = 0; l <= 26; ++l) {
          bool match = l < 26;
          for (int x = 0; match && x < 3; ++x)
            if (mirror[x] != 0 && kSkinOrder[l][x] != mirror[x]) match = false;
          if (match) {
            if (run < 0) {
              run = l;
              bk_region_t g = {to_bitset(nb), (unsigned) l, 0, next, 0, 0, 0};
              bk_region_t s = {to_bitset(mirror), (unsigned) l, 0, start[l], 0, 0, 0};
              d->ghost.push_back(g);
              d->skin.push_back(s);
            }
            d->number_region(nb, kSkinOrder[l], next);
          } else if (run >= 0) {
            bk_region_t &g = d->ghost.back(), &s = d->skin.back();
            g.skin_ed = s.skin_ed = (unsigned) l;
            g.len = next - g.pos;
            s.len = start[l] - s.pos;
            run = -1;
          }
        }
      }
  d->sep[2] = next;
  d->nbricks = next;

  d->adj.assign((size_t) next * 27, 0u);
  const long T0 = d->T[0], T1 = d->T[1], T2 = d->T[2];
#pragma omp parallel for collapse(2)
  for (long k = 0; k < T2; ++k)
    for (long j = 0; j < T1; ++j)
      for (long i = 0; i < T0; ++i) {
        unsigned *row = d->adj.data() + (size_t) d->grid[d->slot(i, j, k)] * 27;
        int s = 0;
        for (int dk = -1; dk <= 1; ++dk)
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di, ++s) {
              const long x = i + di, y = j + dj, z = k + dk;
              row[s] = (x >= 0 && x < T0 && y >= 0 && y < T1 && z >= 0 && z < T2) ? d->grid[d->slot(x, y, z)] : 0u;
            }
      }
  *out = d;
  return BK_OK;
}

int bk_decomp_destroy(bk_decomp_t *d) {
  delete d;
  return BK_OK;
}
unsigned bk_decomp_nbricks(const bk_decomp_t *d) { return d ? d->nbricks : 0; }
int bk_decomp_sep_pos(const bk_decomp_t *d, unsigned *sep3) {
  BK_REQUIRE(d && sep3, "null argument");
  memcpy(sep3, d->sep, sizeof(d->sep));
  return BK_OK;
}
int bk_decomp_tdims(const bk_decomp_t *d, unsigned *t) {
  BK_REQUIRE(d && t, "null argument");
  memcpy(t, d->T, sizeof(d->T));
  return BK_OK;
}
const unsigned *bk_decomp_grid(const bk_decomp_t *d) { return d ? d->grid.data() : nullptr; }
const unsigned *bk_decomp_adj(const bk_decomp_t *d) { return d ? d->adj.data() : nullptr; }
int bk_decomp_nregions(const bk_decomp_t *d) { return d ? (int) d->ghost.size() : 0; }
int bk_decomp_region(const bk_decomp_t *d, int which, int i, bk_region_t *out) {
  BK_REQUIRE(d && out, "null argument");
  BK_REQUIRE(i >= 0 && i < (int) d->ghost.size() && (which == 0 || which == 1), "index out of range");
  *out = which ? d->skin[i] : d->ghost[i];
  return BK_OK;
}
int bk_decomp_skin_size(const bk_decomp_t *d, long *out26) {
  BK_REQUIRE(d && out26, "null argument");
  memcpy(out26, d->skin_size, sizeof(d->skin_size));
  return BK_OK;
}
long bk_decomp_list(const bk_decomp_t *d, int which, unsigned *ids) {
  if (!d || which < 0 || which > 2) return BK_EINVAL;
  const unsigned lo = which == 0 ? 1u : d->sep[which - 1], hi = d->sep[which];
  if (ids)
    for (unsigned b = lo; b < hi; ++b) ids[b - lo] = b;
  return (long) hi - (long) lo;
}

int bk_rank_map(const int *cart, const int *coo, uint64_t *sets, int *ranks) {
  BK_REQUIRE(cart && coo && sets && ranks, "null argument");
  BK_REQUIRE(cart[0] > 0 && cart[1] > 0 && cart[2] > 0, "bad cartesian extents");
  const int pick[3] = {+1, 0, -1};
  int n = 0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c, ++n) {
        const Sign3 s = {pick[a], pick[b], pick[c]};
        // axis x (0 = i) lives in Cartesian coordinate 2-x; set element +axis pairs with coordinate-1
        int r = 0;
        for (int idx = 0; idx < 3; ++idx) {
          const int axis = 2 - idx;
          int v = (coo[idx] - s[axis]) % cart[idx];
          if (v < 0) v += cart[idx];
          r = r * cart[idx] + v;
        }
        sets[n] = to_bitset(s);
        ranks[n] = r;
      }
  return BK_OK;
}

unsigned long bk_zmort_encode(const unsigned long *c) {
  unsigned long id = 0;
  for (unsigned long bit = 0; bit < 21; ++bit)
    for (int a = 0; a < 3; ++a) id |= ((c[a] >> bit) & 1ul) << (bit * 3 + a);
  return id;
}
int bk_zmort_decode(unsigned long id, unsigned long *c) {
  BK_REQUIRE(c, "null argument");
  c[0] = c[1] = c[2] = 0;
  for (unsigned long bit = 0; bit < 21; ++bit)
    for (int a = 0; a < 3; ++a) c[a] |= ((id >> (bit * 3 + a)) & 1ul) << bit;
  return BK_OK;
}

// ---- stitching: a rank's subdomains as one brick grid (strong scaling) ---------------------------------------------
// GPU analogue of the reference's ghost aliasing through mmap views (strong/main.cpp:205-262): the subdomains with
// Z-Morton ids [first, first+count) (strong/args.cpp:104-113) form a box of the periodic subdim^3 arrangement; their
// bricks, addressed as  q * nbricks + local id  in one subdomain-major allocation, are presented as ONE dense grid.
int bk_stitch_box(unsigned long first, unsigned long count, unsigned long subdim, bk_stitch_box_t *b) {
  BK_REQUIRE(b && count > 0 && subdim > 0, "bad arguments");
  unsigned long lo[3] = {~0ul, ~0ul, ~0ul}, hi[3] = {0, 0, 0};
  for (unsigned long q = 0; q < count; ++q) {
    unsigned long c[3];
    bk_zmort_decode(first + q, c);
    for (int a = 0; a < 3; ++a) lo[a] = c[a] < lo[a] ? c[a] : lo[a], hi[a] = c[a] > hi[a] ? c[a] : hi[a];
  }
  unsigned long vol = 1;
  for (int a = 0; a < 3; ++a) {
    b->lo[a] = lo[a];
    b->n[a] = (long) (hi[a] - lo[a] + 1);
    b->wrap[a] = (unsigned long) b->n[a] == subdim;
    vol *= (unsigned long) b->n[a];
  }
  b->first = first, b->count = count;
  b->is_box = vol == count;
  return BK_OK;
}

int bk_stitch_dims(const bk_decomp_t *d, const bk_stitch_box_t *b, unsigned *dims3) {
  BK_REQUIRE(d && b && dims3, "null argument");
  for (int a = 0; a < 3; ++a) dims3[a] = (unsigned) (b->n[a] * (long) (d->T[a] - 2 * d->G[a]) + 2 * d->G[a]);
  return BK_OK;
}

int bk_stitch_grid(const bk_decomp_t *d, const bk_stitch_box_t *b, unsigned *grid) {
  BK_REQUIRE(d && b && grid, "null argument");
  BK_REQUIRE(b->is_box, "the section is not a box of subdomains");
  BK_REQUIRE((unsigned long long) b->count * d->nbricks < (1ull << 32), "global brick ids exceed 32 bits");
  for (int a = 0; a < 3; ++a) BK_REQUIRE(d->G[a] == 1, "stitching assumes a ghost shell of one brick");
  long B[3], S[3];
  for (int a = 0; a < 3; ++a) B[a] = (long) d->T[a] - 2, S[a] = b->n[a] * B[a] + 2;
  for (long K = 0; K < S[2]; ++K)
    for (long J = 0; J < S[1]; ++J)
      for (long I = 0; I < S[0]; ++I) {
        const long p[3] = {I - 1, J - 1, K - 1};  // brick position relative to the box: -1 .. n
        unsigned long cs[3];
        long lb[3];
        for (int a = 0; a < 3; ++a) {
          const long n = b->n[a] * B[a];
          long pw = p[a];
          if (b->wrap[a]) pw = (pw + n) % n;                 // periodic onto myself: the far side's interior brick
          long cl = pw < 0 ? 0 : pw >= n ? n - 1 : pw;       // nearest brick of the box along this axis ...
          cl /= B[a];                                        // ... and the subdomain it belongs to
          cs[a] = b->lo[a] + (unsigned long) cl;
          lb[a] = pw - cl * B[a] + 1;                        // position in that subdomain's ghost-inclusive grid
        }
        const unsigned long q = bk_zmort_encode(cs) - b->first;
        grid[((size_t) K * S[1] + J) * S[0] + I] = (unsigned) (q * d->nbricks + d->grid[d->slot(lb[0], lb[1], lb[2])]);
      }
  return BK_OK;
}

int bk_stitch_region_needed(const bk_decomp_t *d, const bk_stitch_box_t *b, unsigned long sub_id, int region) {
  if (!d || !b || region < 0 || region >= (int) d->ghost.size() || sub_id < b->first || sub_id >= b->first + b->count)
    return BK_EINVAL;
  unsigned long c[3];
  bk_zmort_decode(sub_id, c);
  const uint64_t set = d->ghost[region].neighbor;
  for (int a = 0; a < 3; ++a) {
    const int off = (set >> (a + 1)) & 1 ? 1 : (set >> (31 + a + 1)) & 1 ? -1 : 0;
    if (off == 0) continue;
    // the region lies beyond the subdomain along this axis: it is on the box surface only if the subdomain is the
    // outermost one on that side -- and the box must not wrap onto itself there
    if (b->wrap[a]) return 0;
    if (off > 0 ? c[a] != b->lo[a] + (unsigned long) b->n[a] - 1 : c[a] != b->lo[a]) return 0;
  }
  return 1;
}

}  // extern "C"
