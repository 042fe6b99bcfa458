/*
 * brick-b200.h -- device mirror + stencil launch for the B200 build, as inline C++ over the C ABI (bricklib_b200.h).
 *
 * Stands in for the reference's include/brick-gpu.h + brick-cuda.h (movBrickInfo :43-57, movBrickInfoDeep :65-74,
 * movBrickStorage :84-103, gpuCheck :16-33), stencils/cudaarray.h (copyToDevice/copyFromDevice :11-31) and for the
 * kernel launch itself (brick_kernel<<<strideb, 32>>>, weak/main.cu:35-43,:277-282).  No CUDA headers are needed:
 * host code is plain C++ linked against libbrick_b200.so, so the drivers build with g++.
 */
#ifndef BRICK_B200_H
#define BRICK_B200_H

#include <stdexcept>
#include <string>
#include <vector>
#include "brick.h"
#include "bricklib_b200.h"

/// the reference's gpuCheck: turn an error code of the C ABI into an exception
#define bkCheck(call)                                                                                          \
  do {                                                                                                         \
    int _rc = (call);                                                                                          \
    if (_rc != BK_OK)                                                                                          \
      throw std::runtime_error(std::string(#call) + " failed (" + std::to_string(_rc) + "): " + bk_last_error()); \
  } while (0)

enum brickMemcpyKind { brickMemcpyHostToDevice = 1, brickMemcpyDeviceToHost = 2 };

/// movBrickInfo: copy the adjacency list across; the returned BrickInfo owns device (or host) memory, free with
/// freeBrickInfo.  Sizes are size_t (the reference's unsigned arithmetic overflows beyond 4 GiB, brick-gpu.h:47).
template <unsigned dims>
BrickInfo<dims> movBrickInfo(BrickInfo<dims> &bInfo, brickMemcpyKind kind) {
  BrickInfo<dims> ret = bInfo;
  const size_t bytes = (size_t) bInfo.nbricks * static_power<3, dims>::value * sizeof(unsigned);
  if (kind == brickMemcpyHostToDevice) {
    void *p = nullptr;
    bkCheck(bk_dev_alloc(&p, bytes));
    bkCheck(bk_memcpy_h2d(p, bInfo.adj, bytes, nullptr));
    bkCheck(bk_stream_sync(nullptr));
    ret.adj = (typename BrickInfo<dims>::adjlist) p;
  } else {
    ret.adj = (typename BrickInfo<dims>::adjlist) malloc(bytes);
    bkCheck(bk_memcpy_d2h(ret.adj, bInfo.adj, bytes, nullptr));
    bkCheck(bk_stream_sync(nullptr));
  }
  return ret;
}
template <unsigned dims>
void freeBrickInfoDevice(BrickInfo<dims> &bInfo_dev) {
  bk_dev_free(bInfo_dev.adj);
  bInfo_dev.adj = nullptr;
}

/// movBrickStorage: a new storage on the other side with the same chunks/step, contents copied
inline BrickStorage movBrickStorage(BrickStorage &bStorage, brickMemcpyKind kind) {
  BrickStorage ret;
  ret.chunks = bStorage.chunks;
  ret.step = bStorage.step;
  const size_t bytes = (size_t) bStorage.chunks * bStorage.step * sizeof(bElem);
  if (kind == brickMemcpyHostToDevice) {
    void *p = nullptr;
    bkCheck(bk_dev_alloc(&p, bytes));
    bkCheck(bk_memcpy_h2d(p, bStorage.dat.get(), bytes, nullptr));
    bkCheck(bk_stream_sync(nullptr));
    ret.dat = std::shared_ptr<bElem>((bElem *) p, [](bElem *q) { bk_dev_free(q); });
  } else {
    ret = BrickStorage::allocate(bStorage.chunks, bStorage.step);
    bkCheck(bk_memcpy_d2h(ret.dat.get(), bStorage.dat.get(), bytes, nullptr));
    bkCheck(bk_stream_sync(nullptr));
  }
  return ret;
}

/// a zero-filled storage that lives on the device from the start (no host twin needed)
inline BrickStorage deviceBrickStorage(long chunks, size_t step) {
  BrickStorage ret;
  ret.chunks = chunks;
  ret.step = step;
  void *p = nullptr;
  const size_t bytes = (size_t) chunks * step * sizeof(bElem);
  bkCheck(bk_dev_alloc(&p, bytes));
  bkCheck(bk_dev_memset(p, 0, bytes, nullptr));
  ret.dat = std::shared_ptr<bElem>((bElem *) p, [](bElem *q) { bk_dev_free(q); });
  return ret;
}

/// copyToDevice / copyFromDevice of stencils/cudaarray.h: plain arrays described by their extents
template <typename T>
void copyToDevice(const std::vector<long> &extent, T *&dst_dev, const T *src_host) {
  size_t n = 1;
  for (long e : extent) n *= (size_t) e;
  void *p = nullptr;
  bkCheck(bk_dev_alloc(&p, n * sizeof(T)));
  bkCheck(bk_memcpy_h2d(p, src_host, n * sizeof(T), nullptr));
  bkCheck(bk_stream_sync(nullptr));
  dst_dev = (T *) p;
}
template <typename T>
void copyFromDevice(const std::vector<long> &extent, T *dst_host, const T *src_dev) {
  size_t n = 1;
  for (long e : extent) n *= (size_t) e;
  bkCheck(bk_memcpy_d2h(dst_host, src_dev, n * sizeof(T), nullptr));
  bkCheck(bk_stream_sync(nullptr));
}

/// device twins of copyToBrick / copyFromBrick / compareBrick (arr, grid and the brick all device resident)
template <typename T>
void copyToBrickDevice(const std::vector<long> &dimlist, const std::vector<long> &padding, const std::vector<long> &ghost,
                       const bElem *arr_dev, const unsigned *grid_dev, T &brick_dev, void *stream = nullptr) {
  static_assert(T::ROW_MAJOR, "device bricks are row-major (Dim<8> / Dim<4,8>): refoldBrick() converts host data in other folds");
  bkCheck(bk_copy_to_brick(dimlist.data(), padding.data(), ghost.data(), arr_dev, grid_dev, brick_dev.dat, brick_dev.step,
                           stream));
}
template <typename T>
void copyFromBrickDevice(const std::vector<long> &dimlist, const std::vector<long> &padding,
                         const std::vector<long> &ghost, bElem *arr_dev, const unsigned *grid_dev, T &brick_dev,
                         void *stream = nullptr) {
  bkCheck(bk_copy_from_brick(dimlist.data(), padding.data(), ghost.data(), arr_dev, grid_dev, brick_dev.dat,
                             brick_dev.step, stream));
}
template <typename T>
bool compareBrickDevice(const std::vector<long> &dimlist, const std::vector<long> &padding, const std::vector<long> &ghost,
                        const bElem *arr_dev, const unsigned *grid_dev, T &brick_dev, double tol = 1e-12,
                        double *max_rel = nullptr) {
  unsigned long long bad = 0;
  double worst = 0;
  bkCheck(bk_compare_brick(dimlist.data(), padding.data(), ghost.data(), arr_dev, grid_dev, brick_dev.dat, brick_dev.step,
                           tol, &bad, &worst, nullptr));
  if (max_rel) *max_rel = worst;
  return bad == 0;
}

/// The launch.  `bIn`/`bOut` are Brick views over DEVICE storage whose bInfo->adj is a DEVICE adjacency list
/// (movBrickInfo); grid_dev is the dense id array with extents gdims ({i,j,k}); the brick box [lo,hi) is swept.
/// Equivalent of   brick_kernel<<<dim3(strideb...), 32>>>(grid_dev, bIn, bOut, stride_dev)   for the stencil
/// `stencil` (BK_ST_*), which the reference selects at compile time through -DMPI_7PT... (stencils/fake.h:35-353).
template <typename T>
void brickStencil(int stencil, const unsigned *grid_dev, const std::vector<long> &gdims, T &bIn, T &bOut,
                  const std::vector<long> &lo, const std::vector<long> &hi, const bElem *coeff = nullptr,
                  void *stream = nullptr, unsigned kernel = BK_KERNEL_AUTO) {
  static_assert(T::ROW_MAJOR, "device bricks are row-major (Dim<8> / Dim<4,8>): refoldBrick() converts host data in other folds");
  bk_field_t f = {&bIn.bInfo->adj[0][0], bIn.dat, bIn.step, bOut.dat, bOut.step};
  const unsigned g[3] = {(unsigned) gdims[0], (unsigned) gdims[1], (unsigned) gdims[2]};
  const unsigned l[3] = {(unsigned) lo[0], (unsigned) lo[1], (unsigned) lo[2]};
  const unsigned h[3] = {(unsigned) hi[0], (unsigned) hi[1], (unsigned) hi[2]};
  bkCheck(bk_stencil_apply(stencil, &f, grid_dev, g, l, h, coeff, kernel, stream));
}
/// One half of a split sweep over [lo,hi) (bk_stencil_apply_part): `part` = BK_PART_READY launches the CTAs that read only
/// bricks of [ready_lo,ready_hi) (final before the ghost exchange), BK_PART_REST all the others.  Returns false when the
/// storage layout rules the marching kernel out (then sweep the whole box after the exchange instead).
template <typename T>
bool brickStencilPart(int stencil, const unsigned *grid_dev, const std::vector<long> &gdims, T &bIn, T &bOut,
                      const std::vector<long> &lo, const std::vector<long> &hi, const std::vector<long> &ready_lo,
                      const std::vector<long> &ready_hi, int part, const bElem *coeff = nullptr, void *stream = nullptr) {
  bk_field_t f = {&bIn.bInfo->adj[0][0], bIn.dat, bIn.step, bOut.dat, bOut.step};
  unsigned g[3], l[3], h[3], rl[3], rh[3];
  for (int d = 0; d < 3; ++d)
    g[d] = (unsigned) gdims[d], l[d] = (unsigned) lo[d], h[d] = (unsigned) hi[d], rl[d] = (unsigned) ready_lo[d],
    rh[d] = (unsigned) ready_hi[d];
  const int rc = bk_stencil_apply_part(stencil, &f, grid_dev, g, l, h, coeff, rl, rh, part, stream);
  if (rc == BK_EUNSUPPORTED) return false;
  bkCheck(rc);
  return true;
}
/// `steps` (1 or 2) time steps in one pass over HBM (bk_stencil_advance): bOut = stencil^steps(bIn) on [lo,hi), every
/// intermediate evaluated over the whole grid in shared memory.  part/ready_* as in brickStencilPart (BK_PART_ALL: the
/// ready box is ignored).  Returns false when no fused kernel exists for this stencil / storage layout.
template <typename T>
bool brickAdvance(int stencil, int steps, const unsigned *grid_dev, const std::vector<long> &gdims, T &bIn, T &bOut,
                  const std::vector<long> &lo, const std::vector<long> &hi, const std::vector<long> &ready_lo,
                  const std::vector<long> &ready_hi, int part = BK_PART_ALL, const bElem *coeff = nullptr,
                  void *stream = nullptr) {
  bk_field_t f = {&bIn.bInfo->adj[0][0], bIn.dat, bIn.step, bOut.dat, bOut.step};
  unsigned g[3], l[3], h[3], rl[3], rh[3];
  for (int d = 0; d < 3; ++d)
    g[d] = (unsigned) gdims[d], l[d] = (unsigned) lo[d], h[d] = (unsigned) hi[d], rl[d] = (unsigned) ready_lo[d],
    rh[d] = (unsigned) ready_hi[d];
  const int rc = bk_stencil_advance(stencil, steps, &f, grid_dev, g, l, h, coeff, rl, rh, part, stream);
  if (rc == BK_EUNSUPPORTED) return false;
  bkCheck(rc);
  return true;
}
template <typename T>
void brickStencil(int stencil, const unsigned *grid_dev, const std::vector<long> &gdims, T &bIn, T &bOut,
                  const bElem *coeff = nullptr, void *stream = nullptr) {
  brickStencil(stencil, grid_dev, gdims, bIn, bOut, {0, 0, 0}, gdims, coeff, stream);
}

/// A stencil lowered from its tap list (bk_stencil_compile) -- what `brick("script.py", "CUDA", ...)` is to the reference:
/// the expression of a stencils/*.py script, here as data.  `python -m bricklib_b200.dsl script.py --emit-c NAME` writes
/// the table for a script at build time.  launch() = brickStencil for this stencil.
class BrickStencilDef {
  bk_stencil_def_t *def = nullptr;

 public:
  int kind = 0, radius = 0, ntaps = 0, st_iter = 0, fused_steps = 1;
  BrickStencilDef(const bk_tap_t *taps, int n) {
    bkCheck(bk_stencil_compile(&def, taps, n));
    bkCheck(bk_stencil_def_info(def, &kind, &radius, &ntaps, &st_iter, &fused_steps));
  }
  explicit BrickStencilDef(const std::vector<bk_tap_t> &taps) : BrickStencilDef(taps.data(), (int) taps.size()) {}
  /// with the pointwise clamps of stencils/cond.py: out = post(sum c * pre(in)), e.g. pre = {BK_OP_MAX, 0.0}, post = {BK_OP_ABS, 0}
  BrickStencilDef(const std::vector<bk_tap_t> &taps, bk_pointwise_t pre, bk_pointwise_t post) {
    bkCheck(bk_stencil_compile_pointwise(&def, taps.data(), (int) taps.size(), &pre, &post));
    bkCheck(bk_stencil_def_info(def, &kind, &radius, &ntaps, &st_iter, &fused_steps));
  }
  BrickStencilDef(const BrickStencilDef &) = delete;
  ~BrickStencilDef() { bk_stencil_def_destroy(def); }
  template <typename T>
  void launch(const unsigned *grid_dev, const std::vector<long> &gdims, T &bIn, T &bOut, const std::vector<long> &lo,
              const std::vector<long> &hi, void *stream = nullptr, unsigned kernel = BK_KERNEL_AUTO) const {
    static_assert(T::ROW_MAJOR, "device bricks are row-major (Dim<8> / Dim<4,8>): refoldBrick() converts host data in other folds");
    bk_field_t f = {&bIn.bInfo->adj[0][0], bIn.dat, bIn.step, bOut.dat, bOut.step};
    unsigned g[3], l[3], h[3];
    for (int d = 0; d < 3; ++d) g[d] = (unsigned) gdims[d], l[d] = (unsigned) lo[d], h[d] = (unsigned) hi[d];
    bkCheck(bk_stencil_def_apply(def, &f, grid_dev, g, l, h, kernel, stream));
  }
};

/// cutime_func (stencils/stencils_cu.h:13-28): one warm-up launch, then `reps` launches between two events;
/// returns seconds per launch
template <typename F>
double cutime_func(F f, int reps = 100) {
  void *e0, *e1;
  bkCheck(bk_event_create(&e0));
  bkCheck(bk_event_create(&e1));
  f();
  bkCheck(bk_device_sync());
  bkCheck(bk_event_record(e0, nullptr));
  for (int i = 0; i < reps; ++i) f();
  bkCheck(bk_event_record(e1, nullptr));
  bkCheck(bk_event_sync(e1));
  float ms = 0;
  bkCheck(bk_event_elapsed_ms(e0, e1, &ms));
  bk_event_destroy(e0);
  bk_event_destroy(e1);
  return ms / 1e3 / reps;
}

#endif  // BRICK_B200_H
