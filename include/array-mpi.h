/*
 * array-mpi.h -- ghost-zone exchange and sweep for plain (array-layout) fields on the device: the baseline the
 * reference times next to bricks ("Arr:" block of weak/main.cu:161-213).
 *
 * Stands in for the reference's include/array-mpi.h: exchangeArr<dim> (:146-213) packs the 26 surface regions of a
 * padded array into buffers, posts Isend/Irecv pairs and unpacks.  Here the arrays stay on the GPU and the exchange is
 * ONE kernel that pulls every ghost box straight from the neighbour's array through NVLink peer (or CUDA-IPC)
 * pointers -- strided boxes, no pack / unpack buffers (C ABI: bk_xplan_create_boxes).  evalsize (:128-138) keeps its
 * meaning.  arrayStencil is arr_kernel (weak/main.cu:27-33).
 */
#ifndef ARRAY_MPI_H
#define ARRAY_MPI_H

#include <unordered_map>
#include <vector>
#include "brick-mpi.h"

/// cells of the region in direction `region` (ghost wide along its axes, the domain along the others); inner = true
/// counts the surface region of the INTERIOR that does not overlap other regions (array-mpi.h:128-138)
inline unsigned long evalsize(BitSet region, const std::vector<long> &dimlist, const std::vector<long> &ghost, bool inner = true) {
  unsigned long size = 1;
  for (int i = 1; i <= (int) dimlist.size(); ++i)
    size *= (region.get(i) || region.get(-i)) ? ghost[i - 1] : dimlist[i - 1] - (inner ? 2 * ghost[i - 1] : 0);
  return size;
}

/// the pull plan of one device array: 26 boxes, ghost box in direction v <- the neighbour's interior box facing it
class ArrayExchangeView {
  bk_xplan_t *plan = nullptr;

 public:
  size_t bytes = 0;  ///< bytes received per exchange
  ArrayExchangeView() = default;
  /// peers[r] = address of rank r's array as seen from this GPU; rank_map as filled by populate()
  ArrayExchangeView(bElem *arr_dev, const std::unordered_map<uint64_t, int> &rank_map, const std::vector<bElem *> &peers,
                    const std::vector<long> &dimlist, const std::vector<long> &padding, const std::vector<long> &ghost) {
    long ext[3], stride[3] = {1, 0, 0};
    for (int a = 0; a < 3; ++a) ext[a] = dimlist[a] + 2 * (padding[a] + ghost[a]);
    stride[1] = ext[0], stride[2] = ext[0] * ext[1];
    std::vector<bk_box_t> boxes;
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          const int v[3] = {di, dj, dk};
          if (!di && !dj && !dk) continue;
          BitSet set;
          bk_box_t b;
          long so = 0, dso = 0;
          for (int a = 0; a < 3; ++a) {
            const long p = padding[a], g = ghost[a], d = dimlist[a];
            long src, dst;
            if (v[a] > 0) set.flip(a + 1), b.n[a] = g, src = p + g, dst = p + g + d;  // upper ghost <- neighbour's lowest interior
            else if (v[a] < 0) set.flip(-(a + 1)), b.n[a] = g, src = p + d, dst = p;  // lower ghost <- neighbour's highest interior
            else b.n[a] = d, src = dst = p + g;
            so += src * stride[a], dso += dst * stride[a];
          }
          b.src = peers.at(rank_map.at(set.set)) + so;
          b.dst = arr_dev + dso;
          b.src_stride[0] = b.dst_stride[0] = stride[1];
          b.src_stride[1] = b.dst_stride[1] = stride[2];
          boxes.push_back(b);
          bytes += (size_t) b.n[0] * b.n[1] * b.n[2] * sizeof(bElem);
        }
    bkCheck(bk_xplan_create_boxes(&plan, boxes.data(), (int) boxes.size()));
  }
  ArrayExchangeView(ArrayExchangeView &&o) noexcept : plan(o.plan), bytes(o.bytes) { o.plan = nullptr; }
  ArrayExchangeView &operator=(ArrayExchangeView &&o) noexcept {
    std::swap(plan, o.plan);
    std::swap(bytes, o.bytes);
    return *this;
  }
  ArrayExchangeView(const ArrayExchangeView &) = delete;
  ~ArrayExchangeView() { bk_xplan_destroy(plan); }
  /// asynchronous on `stream`; the caller orders it against the neighbours' sweeps
  void exchange(void *stream = nullptr) { bkCheck(bk_xplan_run(plan, stream)); }
};

/// one-shot exchangeArr for a single rank that is its own neighbour in every direction (periodic): builds, runs, waits
template <unsigned dim>
void exchangeArr(bElem *arr_dev, BrickComm &comm, std::unordered_map<uint64_t, int> &rank_map, const std::vector<long> &dimlist,
                 const std::vector<long> &padding, const std::vector<long> &ghost) {
  static_assert(dim == 3, "3-D arrays");
  ArrayExchangeView ev(arr_dev, rank_map, std::vector<bElem *>(comm.size, arr_dev), dimlist, padding, ghost);
  ev.exchange(nullptr);
  bkCheck(bk_stream_sync(nullptr));
}

/// arr_kernel<<<strideb, TILE^3>>>: out = stencil(in) over the cell box [lo,hi) of arrays with `extent` cells per axis
inline void arrayStencil(int stencil, const bElem *in_dev, bElem *out_dev, const std::vector<long> &extent,
                         const std::vector<long> &lo, const std::vector<long> &hi, const bElem *coeff = nullptr, void *stream = nullptr) {
  bkCheck(bk_array_stencil_apply(stencil, in_dev, out_dev, extent.data(), lo.data(), hi.data(), coeff, stream));
}

#endif  // ARRAY_MPI_H
