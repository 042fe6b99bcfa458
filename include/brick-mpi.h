/*
 * brick-mpi.h -- domain decomposition and ghost-zone exchange for bricks on a multi-GPU node.
 *
 * Keeps the surface of the reference's include/brick-mpi.h (+ bitset.h): BitSet (bitset.h:19-128), BrickDecomp
 * (:178-713: ghost/skin region tables, sep_pos, skin_size, rank_map, operator[], getBrickInfo, exchange,
 * exchangeView), ExchangeView (:82-124), populate (:730-753), mpi_stats / mpi_statistics (:758-793), skin3d_good
 * (src/brick-mpi.cpp:25-52).  What changes is the transport: there is no MPI on a single 8xB200 node build -- a
 * "communicator" is a BrickComm (periodic Cartesian grid of GPUs, one rank per GPU), and an exchange is ONE kernel that
 * pulls every ghost region straight out of the neighbours' skin regions over NVLink (peer or CUDA-IPC pointers):
 * ghost[i] <- rank_map[ghost[i].neighbor].skin[i], no packing, no staging (C ABI: bk_xplan_*).
 * The numbering itself (which brick gets which id) is computed by libbrick_b200.so and is bit-identical to the
 * reference's DECOMP_PAGEUNALIGN numbering (tests/test_decomp.py).
 */
#ifndef BRICK_MPI_H
#define BRICK_MPI_H

#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "brick-b200.h"

/// Set of signed axis ids (+-1..+-31) in one 64-bit word: +a is bit a, -a is bit 31+a (same encoding as the reference)
struct BitSet {
  uint64_t set;
  BitSet() : set(0) {}
  BitSet(uint64_t s) : set(s) {}
  BitSet(std::initializer_list<int> l) : set(0) {
    for (int p : l) set ^= bit(p);
  }
  static uint64_t bit(long pos) { return 1ull << (uint64_t) (pos < 0 ? 31 - pos : pos); }
  BitSet &flip(long pos) {
    set ^= bit(pos);
    return *this;
  }
  long size() const { return __builtin_popcountll(set); }
  bool get(long pos) const { return (set & bit(pos)) != 0; }
  BitSet operator&(BitSet a) const { return BitSet(set & a.set); }
  BitSet operator|(BitSet a) const { return BitSet(set | a.set); }
  BitSet operator^(BitSet a) const { return BitSet(set ^ a.set); }
  explicit operator bool() const { return set != 0; }
  bool operator<=(BitSet a) const { return (set & a.set) == set; }
  bool operator>=(BitSet a) const { return (set & a.set) == a.set; }
  bool operator==(BitSet a) const { return set == a.set; }
  /// every element negated
  BitSet operator!() const {
    const uint64_t low = (1ull << 32) - 1;
    return BitSet(((set & low) << 31) | (set >> 31));
  }
};
inline std::ostream &operator<<(std::ostream &os, const BitSet &b) {
  os << "{";
  for (long a = 1; a < 32; ++a) {
    if (b.get(a)) os << a << "+";
    if (b.get(-a)) os << a << "-";
  }
  return os << "}";
}

/// surface-region order used by every driver; this build implements exactly this order (tag object)
inline std::vector<BitSet> make_skin3d_good() {
  return {{1}, {1, -3}, {1, 2, -3}, {1, 2}, {1, 2, 3}, {2, 3}, {2}, {2, -3}, {-1, 2, -3}, {-1, 2}, {-1, 2, 3}, {-1, 3}, {-1},
          {-3}, {-1, -3}, {-1, -2, -3}, {-1, -2}, {-1, -2, 3}, {-2, 3}, {-2}, {-2, -3}, {1, -2, -3}, {1, -2}, {1, -2, 3},
          {1, 3}, {3}};
}
static const std::vector<BitSet> skin3d_good = make_skin3d_good();

/// What the ranks of one job share when "ranks" are host threads, one per GPU: a barrier and one slot per rank for the
/// statistics reductions (the stand-in for MPI_COMM_WORLD's collective machinery).
struct BrickWorld {
  int size;
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0, gen = 0;
  std::vector<double> slot;
  explicit BrickWorld(int size) : size(size), slot(size, 0.0) {}
  void barrier() {
    std::unique_lock<std::mutex> l(m);
    const int g = gen;
    if (++waiting == size) {
      waiting = 0, ++gen;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return g != gen; });
    }
  }
};

/// Stand-in for the periodic Cartesian MPI communicator (weak/args.cpp:105-108): dims/coords in MPI order
/// (index 0 varies slowest and corresponds to axis k), one rank per GPU.  `world` is optional: set for jobs whose ranks
/// call the collectives below (mpi_statistics(double, comm), MPI_Barrier-like comm.barrier()).
struct BrickComm {
  int dims[3] = {1, 1, 1};
  int coords[3] = {0, 0, 0};
  int rank = 0, size = 1;
  std::shared_ptr<BrickWorld> world;
  static BrickComm cart(const int *d, int rank, std::shared_ptr<BrickWorld> world = nullptr) {
    BrickComm c;
    c.size = d[0] * d[1] * d[2];
    c.rank = rank;
    for (int i = 0; i < 3; ++i) c.dims[i] = d[i];
    c.coords[2] = rank % d[2];
    c.coords[1] = (rank / d[2]) % d[1];
    c.coords[0] = rank / (d[1] * d[2]);
    c.world = std::move(world);
    return c;
  }
  void barrier() const {
    if (world) world->barrier();
  }
};
#ifndef MPI_VERSION
/// there is no MPI in a one-node build: reference call sites that say MPI_Comm name the stand-in
typedef BrickComm MPI_Comm;
#endif

/// One fused pull of all ghost regions of one storage (reference: ExchangeView::exchange, brick-mpi.h:96-123)
class ExchangeView {
  bk_xplan_t *plan = nullptr;

 public:
  size_t bytes = 0;  ///< bytes received per exchange
  ExchangeView() = default;
  ExchangeView(const std::vector<bk_seg_t> &segs) {
    bkCheck(bk_xplan_create(&plan, segs.data(), (int) segs.size()));
    bytes = bk_xplan_bytes(plan);
  }
  ExchangeView(ExchangeView &&o) noexcept : plan(o.plan), bytes(o.bytes) { o.plan = nullptr; }
  ExchangeView &operator=(ExchangeView &&o) noexcept {
    std::swap(plan, o.plan);
    std::swap(bytes, o.bytes);
    return *this;
  }
  ExchangeView(const ExchangeView &) = delete;
  ~ExchangeView() { bk_xplan_destroy(plan); }
  /// asynchronous on `stream`; the caller orders it against the peers' sweeps (events / flags / a barrier)
  void exchange(void *stream = nullptr) { bkCheck(bk_xplan_run(plan, stream)); }
  /// launch shape of the pull kernel: (0, 0) = wide default, e.g. (32, 1024) = confined to a few SMs (bk_xplan_set_shape)
  void setShape(int ctas, int threads_per_cta) { bkCheck(bk_xplan_set_shape(plan, ctas, threads_per_cta)); }
  /// the same plan with the big ranges on the copy engines (bk_xplan_run_ce): takes no SM from the sweeps
  void exchangeCopyEngines(void *stream = nullptr) { bkCheck(bk_xplan_run_ce(plan, nullptr, 0, nullptr, 0, 0, stream)); }
  /// with the cross-process handshake of bk_xplan_run_sync
  void exchange(const std::vector<const uint64_t *> &wait, const std::vector<uint64_t *> &signal, uint64_t epoch,
                void *stream = nullptr) {
    bkCheck(bk_xplan_run_sync(plan, wait.data(), (int) wait.size(), signal.data(), (int) signal.size(), epoch, stream));
  }
};

template <unsigned dim, unsigned... BDims>
class BrickDecomp {
  static_assert(dim == 3 && sizeof...(BDims) == 3 && cal_size<BDims...>::value == 512,
                "libbrick_b200 decomposes 3-D domains of 8x8x8 bricks");
  bk_decomp_t *h = nullptr;
  std::vector<unsigned> dims_cells;
  unsigned depth;
  BrickInfo<dim> *bInfo = nullptr;
  unsigned tdims[3] = {0, 0, 0};
  const unsigned *grid = nullptr;

 public:
  typedef struct {
    BitSet neighbor;
    unsigned skin_st, skin_ed, pos, len, first_pad, last_pad;
  } g_region;
  std::vector<g_region> ghost, skin;
  unsigned sep_pos[3] = {0, 0, 0};
  std::vector<BitSet> skinlist;
  std::vector<long> skin_size;
  BrickComm comm;
  std::unordered_map<uint64_t, int> rank_map;

  /// numfield interleaved fields share one storage: a brick id then names a chunk of numfield bricks (allocate the storage
  /// with step = numfield * 512 and view field f as Brick(bInfo, storage, f * 512)); the numbering does not depend on it
  /// (brick-mpi.h:304-350 uses it for page padding only, which is 0 for 4 KiB bricks), the exchange moves whole chunks
  unsigned numfield;
  BrickDecomp(const std::vector<unsigned> &dims, const unsigned depth, unsigned numfield = 1)
      : dims_cells(dims), depth(depth), numfield(numfield) {
    if (numfield < 1) throw std::runtime_error("BrickDecomp: numfield must be at least 1");
  }
  BrickDecomp(const BrickDecomp &) = delete;
  ~BrickDecomp() {
    bk_decomp_destroy(h);
    if (bInfo) {
      free(bInfo->adj);
      delete bInfo;
    }
  }

  /// number the bricks; `skinlist` must be skin3d_good (the only order the drivers use)
  void initialize(const std::vector<BitSet> &skinlist) {
    if (skinlist.size() != skin3d_good.size()) throw std::runtime_error("BrickDecomp: only skin3d_good is supported");
    for (size_t i = 0; i < skinlist.size(); ++i)
      if (!(skinlist[i] == skin3d_good[i])) throw std::runtime_error("BrickDecomp: only skin3d_good is supported");
    this->skinlist = skinlist;
    bkCheck(bk_decomp_create(&h, dims_cells.data(), depth));
    bkCheck(bk_decomp_sep_pos(h, sep_pos));
    bkCheck(bk_decomp_tdims(h, tdims));
    grid = bk_decomp_grid(h);
    const int n = bk_decomp_nregions(h);
    for (int which = 0; which < 2; ++which)
      for (int i = 0; i < n; ++i) {
        bk_region_t r;
        bkCheck(bk_decomp_region(h, which, i, &r));
        g_region g = {BitSet(r.neighbor), r.skin_st, r.skin_ed, r.pos, r.len, r.first_pad, r.last_pad};
        (which == 0 ? ghost : skin).push_back(g);
      }
    skin_size.assign(26, 0);
    bkCheck(bk_decomp_skin_size(h, skin_size.data()));
    const unsigned nb = bk_decomp_nbricks(h);
    bInfo = new BrickInfo<dim>(nb);
    std::memcpy(bInfo->adj, bk_decomp_adj(h), (size_t) nb * 27 * sizeof(unsigned));
  }

  BrickInfo<dim> getBrickInfo() { return *bInfo; }  ///< shares the adjacency list with the decomposition (as the reference)
  unsigned nbricks() const { return bk_decomp_nbricks(h); }
  std::vector<long> gridDims() const { return {(long) tdims[0], (long) tdims[1], (long) tdims[2]}; }
  const unsigned *gridData() const { return grid; }
  const bk_decomp_t *handle() const { return h; }  ///< for the C-ABI calls that take a decomposition (bk_stitch_*)
  size_t exchangeSize() const {  ///< bricks received per exchange
    size_t n = 0;
    for (auto &g : ghost) n += g.len;
    return n;
  }

  /// bDecomp[k][j][i]: id of the brick at grid position (i,j,k), ghost shell included
  struct Row {
    const unsigned *p;
    unsigned operator[](long i) const { return p[i]; }
  };
  struct Plane {
    const unsigned *p;
    unsigned sx;
    Row operator[](long j) const { return Row{p + j * sx}; }
  };
  Plane operator[](long k) const { return Plane{grid + (size_t) k * tdims[0] * tdims[1], tdims[0]}; }

  /// brick-id lists for overlap: 0 inner (reads no ghost brick), 1 skin, 2 ghost
  std::vector<unsigned> idList(int which) const {
    std::vector<unsigned> ids((size_t) bk_decomp_list(h, which, nullptr));
    bk_decomp_list(h, which, ids.data());
    return ids;
  }

  /// Pull plan for one DEVICE storage.  peers[r] = base address of rank r's storage as seen from this GPU (own
  /// address for r == comm.rank; a peer-enabled or IPC-mapped address otherwise).
  ExchangeView exchangeView(BrickStorage &bStorage_dev, const std::vector<bElem *> &peers) {
    if (bStorage_dev.step % (512 * (size_t) numfield))
      throw std::runtime_error("exchangeView: storage step is not a multiple of numfield bricks");
    std::vector<bk_seg_t> segs(ghost.size());
    for (size_t i = 0; i < ghost.size(); ++i) {
      const int src_rank = rank_map.at(ghost[i].neighbor.set);
      segs[i].src = peers.at(src_rank) + (size_t) skin[i].pos * bStorage_dev.step;
      segs[i].dst = bStorage_dev.dat.get() + (size_t) ghost[i].pos * bStorage_dev.step;
      segs[i].bytes = (size_t) ghost[i].len * bStorage_dev.step * sizeof(bElem);
    }
    return ExchangeView(segs);
  }
  /// single rank, periodic in every direction: every neighbour is this storage itself
  ExchangeView exchangeView(BrickStorage &bStorage_dev) {
    return exchangeView(bStorage_dev, std::vector<bElem *>(comm.size, bStorage_dev.dat.get()));
  }
  /// one-shot form of the reference's bDecomp.exchange(storage): builds the plan, runs it, waits
  void exchange(BrickStorage &bStorage_dev) {
    ExchangeView ev = exchangeView(bStorage_dev);
    ev.exchange(nullptr);
    bkCheck(bk_stream_sync(nullptr));
  }
};

/// populate(comm, bDecomp, 0, 1, coo) -- same signature as the reference's (brick-mpi.h:730-753; there `neighbor` and `d`
/// drive its recursion over the axes, here the whole 27-entry map comes from one bk_rank_map call, so any start values
/// mean "all of it"): fill bDecomp.rank_map with the rank of each neighbour set on the periodic Cartesian grid `comm`
template <unsigned dim, unsigned... BDims>
void populate(MPI_Comm &comm, BrickDecomp<dim, BDims...> &bDecomp, BitSet neighbor = BitSet(), int d = 1, int *coo = nullptr) {
  (void) neighbor, (void) d;
  uint64_t sets[27];
  int ranks[27];
  bkCheck(bk_rank_map(comm.dims, coo ? coo : comm.coords, sets, ranks));
  for (int i = 0; i < 27; ++i) bDecomp.rank_map[sets[i]] = ranks[i];
  bDecomp.comm = comm;
}

/// min/avg/max/sigma over ranks (reference: 4x MPI_Reduce, brick-mpi.h:768-785) -- here over a vector of per-GPU values
typedef struct {
  double min, max, avg, sigma;
} mpi_stats;
inline mpi_stats mpi_statistics(const std::vector<double> &per_rank) {
  mpi_stats r = {0, 0, 0, 0};
  if (per_rank.empty()) return r;
  const int n = (int) per_rank.size();
  double sum = 0, sq = 0;
  r.min = r.max = per_rank[0];
  for (double v : per_rank) sum += v, sq += v * v, r.min = std::min(r.min, v), r.max = std::max(r.max, v);
  r.avg = sum / n;
  r.sigma = std::sqrt(std::max(0.0, (sq - sum * sum / n) / std::max(n - 1, 1)));
  return r;
}
/// mpi_statistics(stats, comm) -- the reference's collective form (brick-mpi.h:768-785: four MPI_Reduce calls): every
/// rank of `comm` calls it with its own value and every rank gets the statistics (the reference returns them on rank 0)
inline mpi_stats mpi_statistics(double stats, const MPI_Comm &comm) {
  if (!comm.world) return mpi_statistics(std::vector<double>{stats});
  BrickWorld &w = *comm.world;
  w.slot[comm.rank] = stats;
  w.barrier();
  const mpi_stats r = mpi_statistics(w.slot);
  w.barrier();  // nobody overwrites a slot before everybody has read it
  return r;
}
inline std::ostream &operator<<(std::ostream &os, const mpi_stats &s) {
  return os << "[" << s.min << ", " << s.avg << ", " << s.max << "] (σ: " << s.sigma << ")";
}

#endif  // BRICK_MPI_H
