/*
 * oracle/mpi_stub/mpi.h -- TEST INFRASTRUCTURE ONLY (part of the parity oracle, never shipped or measured).
 *
 * No MPI exists in the build image, yet the reference's decomposition header (include/brick-mpi.h) includes
 * <mpi.h> and its exchange (brick-mpi.h:466-495) is written against MPI_Isend/MPI_Irecv/MPI_Waitall.  This header
 * is an in-process stand-in: several "ranks" live in ONE process, the harness selects the calling rank with
 * mpistub_set_rank(), every Isend/Irecv is recorded, and mpistub_deliver() matches (src,dst,tag) pairs and
 * memcpy's the bytes.  Because skin (send) and ghost (receive) ranges never overlap, delivering after every rank
 * has posted is equivalent to a real two-sided exchange.  MPI_Cart_rank implements a periodic row-major Cartesian
 * grid exactly as MPI_Cart_create(periodic=1) would, which is all `populate()` (brick-mpi.h:730-753) needs.
 */
#ifndef BRICK_B200_MPI_STUB_H
#define BRICK_B200_MPI_STUB_H

#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Win;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_NULL (-1)
#define MPI_COMM_WORLD 0
#define MPI_CHAR 1
#define MPI_DOUBLE 8
#define MPI_SUM 0
#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUCCESS 0

namespace mpistub {
struct Msg { int src, dst, tag; void *buf; size_t len; };
struct State {
  int rank = 0;
  int dims[3] = {1, 1, 1};          // Cartesian extents, dims[0] slowest (MPI row-major order)
  std::vector<Msg> sends, recvs;
};
inline State &state() { static State s; return s; }
inline int world() { State &s = state(); return s.dims[0] * s.dims[1] * s.dims[2]; }
}  // namespace mpistub

inline void mpistub_set_cart(int d0, int d1, int d2) {
  mpistub::State &s = mpistub::state();
  s.dims[0] = d0; s.dims[1] = d1; s.dims[2] = d2;
}
inline void mpistub_set_rank(int r) { mpistub::state().rank = r; }

/* match every posted receive with the send carrying the same (src,dst,tag) and copy the payload */
inline void mpistub_deliver() {
  mpistub::State &s = mpistub::state();
  for (const mpistub::Msg &r : s.recvs) {
    const mpistub::Msg *hit = nullptr;
    for (const mpistub::Msg &m : s.sends)
      if (m.src == r.src && m.dst == r.dst && m.tag == r.tag) { hit = &m; break; }
    if (!hit || hit->len != r.len) throw std::runtime_error("mpistub: unmatched receive");
    std::memcpy(r.buf, hit->buf, r.len);
  }
  s.sends.clear();
  s.recvs.clear();
}

inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *size) { *size = mpistub::world(); return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = mpistub::state().rank; return MPI_SUCCESS; }

inline int MPI_Irecv(void *buf, size_t count, MPI_Datatype, int source, int tag, MPI_Comm, MPI_Request *req) {
  mpistub::State &s = mpistub::state();
  s.recvs.push_back({source, s.rank, tag, buf, count});
  *req = 0;
  return MPI_SUCCESS;
}
inline int MPI_Isend(const void *buf, size_t count, MPI_Datatype, int dest, int tag, MPI_Comm, MPI_Request *req) {
  mpistub::State &s = mpistub::state();
  s.sends.push_back({s.rank, dest, tag, const_cast<void *>(buf), count});
  *req = 0;
  return MPI_SUCCESS;
}
inline int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return MPI_SUCCESS; }

inline int MPI_Cart_rank(MPI_Comm, const int *coords, int *rank) {
  const int *d = mpistub::state().dims;
  int r = 0;
  for (int a = 0; a < 3; ++a) {
    int c = ((coords[a] % d[a]) + d[a]) % d[a];
    r = r * d[a] + c;
  }
  *rank = r;
  return MPI_SUCCESS;
}

/* one-sided + reductions: referenced by templates that the oracle never instantiates on a hot path */
inline int MPI_Win_fence(int, MPI_Win) { return MPI_SUCCESS; }
inline int MPI_Get(void *, size_t, MPI_Datatype, int, size_t, size_t, MPI_Datatype, MPI_Win) {
  throw std::runtime_error("mpistub: MPI_Get is not emulated");
}
inline int MPI_Reduce(const void *in, void *out, int count, MPI_Datatype, MPI_Op, int, MPI_Comm) {
  std::memcpy(out, in, sizeof(double) * (size_t) count);
  return MPI_SUCCESS;
}

#endif
