/*
 * brick.h -- the brick data model of bricklib, re-created for the B200 build.
 *
 * Same type names, members and meaning as the reference's include/brick.h (BrickStorage :53-82, BrickInfo :96-127,
 * Dim :130, Brick :353-395, accessor math :214-327), so driver code written against the reference compiles against
 * this header.  What differs:
 *   - storage may live in host memory (BrickStorage::allocate) or in device memory (brick-b200.h: movBrickStorage);
 *   - the element accessor b[brick][k][j][i] is a small run-time proxy meant for set-up and validation on the HOST;
 *     the stencil sweep itself never goes through it -- it is one call into libbrick_b200.so (brick-b200.h);
 *   - only folds whose in-brick order is plain row-major are accepted: Dim<8> (the reference's AVX-512 fold) and
 *     Dim<4,8> (its CUDA fold, stencils/cudavfold.h:9).  That is the layout the CUDA kernels read.
 */
#ifndef BRICK_H
#define BRICK_H

#include <cstddef>
#include <cstdlib>
#include <memory>
#include <type_traits>

#ifndef bElem
#define bElem double /* include/vecscatter.h:12-14 */
#endif

#define ALIGN 2048 /* include/brick.h:15 */

template <unsigned base, unsigned exp>
struct static_power {
  static constexpr unsigned value = base * static_power<base, exp - 1>::value;
};
template <unsigned base>
struct static_power<base, 0> {
  static constexpr unsigned value = 1;
};

/// Owner of brick memory: `chunks` bricks (or groups of interleaved fields), `step` elements apart.
struct BrickStorage {
  std::shared_ptr<bElem> dat;
  long chunks = 0;
  size_t step = 0;
  void *mmap_info = nullptr;  ///< kept for layout compatibility; the memfd/mmap allocator has no GPU counterpart

  static BrickStorage allocate(long chunks, size_t step) {
    BrickStorage b;
    b.chunks = chunks;
    b.step = step;
    const size_t bytes = ((size_t) chunks * step * sizeof(bElem) + ALIGN - 1) / ALIGN * ALIGN;
    b.dat = std::shared_ptr<bElem>(static_cast<bElem *>(aligned_alloc(ALIGN, bytes)), [](bElem *p) { free(p); });
    return b;
  }
};

/// Adjacency list: adj[b][s] is the brick at offset s = sum (o_d + 1) * 3^d of brick b; 0 = the null brick.
template <unsigned dims>
struct BrickInfo {
  typedef unsigned (*adjlist)[static_power<3, dims>::value];
  adjlist adj;
  unsigned nbricks;

  explicit BrickInfo(unsigned nbricks) : nbricks(nbricks) {
    adj = (adjlist) malloc((size_t) nbricks * static_power<3, dims>::value * sizeof(unsigned));
  }
  BrickStorage allocate(long step) { return BrickStorage::allocate(nbricks, step); }
};

template <unsigned... Ds>
struct Dim {};

template <unsigned... xs>
struct cal_size;
template <unsigned x>
struct cal_size<x> {
  static constexpr unsigned value = x;
};
template <unsigned x, unsigned... xs>
struct cal_size<x, xs...> {
  static constexpr unsigned value = x * cal_size<xs...>::value;
};

namespace brick_detail {
// a fold is row-major when every fold extent but the slowest equals the brick extent of the same (fastest) axes
template <class B, class F>
struct row_major_fold;
template <unsigned... B, unsigned... F>
struct row_major_fold<Dim<B...>, Dim<F...>> {
  static constexpr bool check() {
    constexpr unsigned nb = sizeof...(B), nf = sizeof...(F);
    constexpr unsigned b[] = {B...}, f[] = {F...};
    if (nf > nb) return false;
    for (unsigned i = 1; i < nf; ++i)
      if (f[nf - i] != b[nb - i]) return false;
    return b[nb - nf] % f[0] == 0;
  }
  static constexpr bool value = check();
};

template <class BrickT, unsigned level>
struct Accessor {
  BrickT *br;
  unsigned b;          // brick the indexing started from
  int idx[BrickT::DIMS];
  inline auto operator[](int i) -> typename std::conditional<level + 1 == BrickT::DIMS, bElem &, Accessor<BrickT, level + 1>>::type {
    return step(i, std::integral_constant<bool, level + 1 == BrickT::DIMS>());
  }

 private:
  inline Accessor<BrickT, level + 1> step(int i, std::false_type) {
    Accessor<BrickT, level + 1> n{br, b, {}};
    for (unsigned d = 0; d < level; ++d) n.idx[d] = idx[d];
    n.idx[level] = i;
    return n;
  }
  inline bElem &step(int i, std::true_type) {
    idx[level] = i;
    return br->elem(b, idx);
  }
};
}  // namespace brick_detail

template <class BD, class F>
struct Brick;

/// A view of one field of a BrickStorage: bricks of extents BDims (slowest first), element (k,j,i) of brick b at
/// dat[b*step + (k*BJ + j)*BI + i]; indices in [-B, 2B) step into the neighbour through the adjacency list.
template <unsigned... BDims, unsigned... Folds>
struct Brick<Dim<BDims...>, Dim<Folds...>> {
  typedef Brick<Dim<BDims...>, Dim<Folds...>> mytype;
  typedef BrickInfo<sizeof...(BDims)> myBrickInfo;
  static constexpr unsigned DIMS = sizeof...(BDims);
  static constexpr unsigned VECLEN = cal_size<Folds...>::value;
  static constexpr unsigned BRICKSIZE = cal_size<BDims...>::value;
  /// is the in-brick order of this fold plain row-major (Dim<8>, Dim<4,8>)?  That is the layout the CUDA kernels read:
  /// the device wrappers of brick-b200.h accept only such bricks.  Other folds (the reference's AVX2 fold Dim<2,2>,
  /// include/brick.h:234-246, stencils/cpuvfold.h:12-40) are HOST views: the accessor below addresses them exactly like
  /// the reference, and refoldBrick() rewrites a field from one fold into another so that such data can be moved.
  static constexpr bool ROW_MAJOR = brick_detail::row_major_fold<Dim<BDims...>, Dim<Folds...>>::value;
  static_assert(sizeof...(Folds) <= sizeof...(BDims), "more fold dimensions than brick dimensions");

  myBrickInfo *bInfo;
  size_t step;
  bElem *dat;
  BrickStorage bStorage;

  Brick(myBrickInfo *bInfo, const BrickStorage &brickStorage, unsigned offset) : bInfo(bInfo) {
    bStorage = brickStorage;
    dat = bStorage.dat.get() + offset;
    step = bStorage.step;
  }

  /// brick extents with the contiguous axis first ({i,j,k} order, like every dimension vector of the library)
  static void extents_fast_first(long *out) {
    constexpr unsigned ext[] = {BDims...};
    for (unsigned d = 0; d < DIMS; ++d) out[d] = ext[DIMS - 1 - d];
  }

  inline brick_detail::Accessor<mytype, 0> operator[](unsigned b) { return brick_detail::Accessor<mytype, 0>{this, b, {}}; }

  /// element `idx` (slowest axis first, each in [-B_d, 2*B_d)) seen from brick b -- HOST memory only
  inline bElem &elem(unsigned b, const int *idx) {
    constexpr unsigned ext[] = {BDims...};
    constexpr unsigned nf = sizeof...(Folds);
    constexpr unsigned fold_raw[] = {Folds..., 1};  // (the trailing 1 keeps the array non-empty for Dim<>)
    unsigned slot = 0, nvec = 0, wvec = 0;
    for (unsigned d = 0; d < DIMS; ++d) {  // slowest axis first: slot digit weight 3^(DIMS-1-d)
      int i = idx[d], o = 1;
      if (i < 0) i += (int) ext[d], o = 0;
      else if (i >= (int) ext[d]) i -= (int) ext[d], o = 2;
      slot = slot * 3 + (unsigned) o;
      // the folds belong to the FASTEST nf axes; an axis without a fold has fold extent 1 (include/brick.h:234-246):
      // element = (which vector) * VECLEN + (position inside the vector), both mixed-radix over the axes
      const unsigned f = d + nf >= DIMS ? fold_raw[d + nf - DIMS] : 1u;
      wvec = wvec * f + (unsigned) i % f;
      nvec = nvec * (ext[d] / f) + (unsigned) i / f;
    }
    const unsigned off = nvec * VECLEN + wvec;
    return dat[(size_t) bInfo->adj[b][slot] * step + off];
  }

  /// first element of the neighbour brick at the given offsets (each -1, 0 or 1; fastest axis first, as in the reference)
  template <int... Offsets>
  inline bElem *neighbor(unsigned b) {
    constexpr int o[] = {Offsets...};
    unsigned slot = 0, w = 1;
    for (unsigned d = 0; d < sizeof...(Offsets); ++d) slot += (unsigned) (o[d] + 1) * w, w *= 3;
    return &dat[(size_t) bInfo->adj[b][slot] * step];
  }
};

/// Rewrite one field from one fold into another (same brick extents, same adjacency): dst[b][k][j][i] = src[b][k][j][i] for
/// every brick.  The way to bring host data kept in a vector fold (e.g. Dim<2,2>, the reference's AVX2 layout) into the
/// row-major fold the device kernels read, and back.
template <unsigned... BDims, unsigned... FA, unsigned... FB>
void refoldBrick(Brick<Dim<BDims...>, Dim<FA...>> &src, Brick<Dim<BDims...>, Dim<FB...>> &dst) {
  static_assert(sizeof...(BDims) == 3, "3-D bricks");
  constexpr unsigned ext[] = {BDims...};
  for (unsigned b = 0; b < src.bInfo->nbricks; ++b)
    for (unsigned k = 0; k < ext[0]; ++k)
      for (unsigned j = 0; j < ext[1]; ++j)
        for (unsigned i = 0; i < ext[2]; ++i) {
          const int idx[3] = {(int) k, (int) j, (int) i};
          dst.elem(b, idx) = src.elem(b, idx);  // indices inside the brick: slot 13 = the brick itself
        }
}

#endif  // BRICK_H
