// bk_exchange.cu -- ghost-zone exchange as ONE kernel over a list of contiguous brick ranges.
//
// Because BrickDecomp lays every ghost and skin region out as a contiguous run of bricks, an exchange is a list of
// plain range copies ghost[i] <- peer.skin[i] (include/brick-mpi.h:466-495 posts them as 42 Irecv/Isend pairs;
// strong/main.cu:76-83 cudaCopy moves one range per CTA with 8-byte accesses).  Here a whole plan (all 42 ranges of a
// subdomain, or every link of a strong-scaling rank) runs as a single persistent kernel: the ranges are cut into
// 16 KiB chunks, CTAs stride over the chunk list, every thread moves 16 bytes per access with four loads in flight.
// A source pointer may live in a peer GPU (CUDA IPC / peer access): the loads then cross NVLink, which makes this a
// PULL exchange -- no pack, no unpack, no staging buffer.  The optional flags implement the cross-process handshake.
#include "bk_common.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <vector>

struct bk_xplan {
  bk_seg_t *segs_dev = nullptr;
  unsigned long long *chunk_first_dev = nullptr;  // prefix sum of chunks per segment, nseg+1 entries
  int nseg = 0;
  unsigned long long nchunks = 0;
  size_t bytes = 0;
  // scratch for flag pointer lists (64 wait + epoch + 64 signal): kFlagSlots of them, used round-robin, so that runs of
  // one plan enqueued on different streams (or back to back) never share a list that is still being read
  uint64_t **flagbuf_dev = nullptr;
  std::atomic<unsigned> flag_turn{0};
  // the list of the last upload: wait / signal addresses are the same in every period (only the epoch changes, and that
  // is a kernel argument), so a run re-uses the uploaded list unless the addresses or the stream changed
  uint64_t *last_list[130] = {nullptr};
  uint64_t **last_fb = nullptr;
  cudaStream_t last_stream = nullptr;
  unsigned *done_dev = nullptr;  // CTAs of the running gated launch that have finished; wraps to 0 with the last one
  int shape_ctas = 0, shape_threads = 0;   // bk_xplan_set_shape: 0 = default (many 256-thread CTAs)
  // array-layout plans (bk_xplan_create_boxes): strided 3-D boxes instead of contiguous ranges; `chunk_first_dev` then
  // counts chunks of kBoxChunk elements per box
  bk_box_t *boxes_dev = nullptr;
  int nbox = 0;
  // copy-engine transport (bk_xplan_run_ce): the segments as host data, dealt largest-first to a few lanes
  std::vector<bk_seg_t> segs_host;
  std::vector<int> lane_of;                // segment -> lane, in issue order `order`
  std::vector<int> order;
  cudaStream_t lanes[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  int nlanes = 0;
  bk_seg_t *small_segs_dev = nullptr;              // segments below kCeMinBytes: moved by a narrow pull kernel
  unsigned long long *small_first_dev = nullptr;
  int small_nseg = 0;
  unsigned long long small_nchunks = 0;
};

namespace {

constexpr unsigned kChunkBytes = 16384;
constexpr size_t kCeMinBytes = 1u << 20;  // copy-engine transport: smaller segments are cheaper through the kernel
constexpr int kThreads = 256;
constexpr int kFlagSlots = 16, kFlagSlotLen = 192;
constexpr unsigned kBoxChunk = 4096;  // elements of a strided box one thread group moves per iteration

__device__ __forceinline__ void spin_until(const uint64_t *flag, uint64_t value) {
  const volatile uint64_t *f = flag;
  while (*f < value) __nanosleep(64);
}

// A CTA is `blockDim.x / 256` groups of 256 threads; each group moves one 16 KiB chunk per iteration.  The default
// shape is many 1-group CTAs (fastest when the GPU is otherwise idle); a NARROW shape -- few 1024-thread CTAs -- keeps
// the pull on a handful of SMs so that sweep CTAs (which need an SM's whole register file) run beside it on the others.
__global__ void __launch_bounds__(1024) k_xplan(const bk_seg_t *__restrict__ segs,
                                                const unsigned long long *__restrict__ first, int nseg,
                                                unsigned long long nchunks, uint64_t *const *wait_flags, int nwait,
                                                int nsignal, uint64_t *gate, unsigned *done,
                                                const bk_box_t *__restrict__ boxes, uint64_t epoch) {
  if (nwait > 0) {
    if ((int) threadIdx.x < nwait) spin_until(wait_flags[threadIdx.x], epoch);
    __syncthreads();
  }
  const unsigned groups = blockDim.x / kThreads, grp = threadIdx.x / kThreads, tid = threadIdx.x % kThreads;
  for (unsigned long long c = (unsigned long long) blockIdx.x * groups + grp; c < nchunks;
       c += (unsigned long long) gridDim.x * groups) {
    int lo = 0, hi = nseg;  // segment s with first[s] <= c < first[s+1]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (first[mid] <= c) lo = mid; else hi = mid;
    }
    if (boxes) {
      // array-layout exchange (exchangeArr, array-mpi.h:146-213, without its pack / unpack buffers): element e of box b
      // is cell (e % n0, (e / n0) % n1, e / (n0 n1)); consecutive threads move consecutive cells of a row
      const bk_box_t bx = boxes[lo];
      const unsigned long long tot = (unsigned long long) bx.n[0] * bx.n[1] * bx.n[2];
      const unsigned long long e0 = (c - first[lo]) * kBoxChunk;
      for (unsigned long long e = e0 + tid; e < min(tot, e0 + kBoxChunk); e += kThreads) {
        const long i = (long) (e % bx.n[0]), r = (long) (e / bx.n[0]);
        const long j = r % bx.n[1], k = r / bx.n[1];
        bx.dst[i + j * bx.dst_stride[0] + k * bx.dst_stride[1]] = bx.src[i + j * bx.src_stride[0] + k * bx.src_stride[1]];
      }
      continue;
    }
    const bk_seg_t sg = segs[lo];
    const size_t off = (size_t) (c - first[lo]) * kChunkBytes;
    const size_t n16 = (min((size_t) kChunkBytes, sg.bytes - off)) >> 4;
    const int4 *src = reinterpret_cast<const int4 *>(static_cast<const char *>(sg.src) + off);
    int4 *dst = reinterpret_cast<int4 *>(static_cast<char *>(sg.dst) + off);
    if (n16 == kChunkBytes / 16) {
      int4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = src[tid + u * kThreads];
#pragma unroll
      for (int u = 0; u < 4; ++u) dst[tid + u * kThreads] = v[u];
    } else {
      for (size_t x = tid; x < n16; x += kThreads) dst[x] = src[x];
    }
  }
  if (done) {
    // the last CTA to finish publishes completion: the local gate (read by the gated sweep's producer warps) and the
    // peers' "done reading your skin" flags.  flag list layout: [0, nwait) wait pointers, [65, 65+nsignal) = signal pointers
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    // self-resetting counter (atomicInc wraps to 0 at gridDim.x - 1): correct whatever grid the previous launch used
    if (threadIdx.x == 0) last = atomicInc(done, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (last) {
      __threadfence_system();
      if (threadIdx.x == 0 && gate) *reinterpret_cast<volatile uint64_t *>(gate) = epoch;
      if ((int) threadIdx.x < nsignal) *reinterpret_cast<volatile uint64_t *>(wait_flags[65 + threadIdx.x]) = epoch;
      __threadfence_system();
    }
  }
}

// after the copy kernel: make the writes visible system-wide, then raise the flags (possibly in peer memory)
__global__ void k_signal(uint64_t *const *flags, int n, uint64_t value) {
  __threadfence_system();
  if ((int) threadIdx.x < n) {
    volatile uint64_t *f = flags[threadIdx.x];
    *f = value;
  }
  __threadfence_system();
}
__global__ void k_wait(const uint64_t *const *flags, int n, uint64_t value) {
  if ((int) threadIdx.x < n) spin_until(flags[threadIdx.x], value);
}

struct FlagList {
  uint64_t *p[64];
};
__global__ void k_signal_list(const __grid_constant__ FlagList list, int n, uint64_t value) {
  __threadfence_system();
  if ((int) threadIdx.x < n) {
    volatile uint64_t *f = list.p[threadIdx.x];
    *f = value;
  }
  __threadfence_system();
}
__global__ void k_wait_list(const __grid_constant__ FlagList list, int n, uint64_t value) {
  if ((int) threadIdx.x < n) spin_until(list.p[threadIdx.x], value);
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

uint64_t **flag_slot(bk_xplan *p) { return p->flagbuf_dev + (size_t) (p->flag_turn++ % kFlagSlots) * kFlagSlotLen; }

// device copy of a flag pointer list ([0,64) wait, [65,129) signal addresses): uploaded on `s` when it differs from the
// last upload or the stream changed (the upload is ordered on ITS stream only), else the resident copy
int flag_list(bk_xplan *p, uint64_t *const (&host)[130], cudaStream_t s, uint64_t ***out) {
  if (p->last_fb && p->last_stream == s && memcmp(p->last_list, host, sizeof(host)) == 0) {
    *out = p->last_fb;
    return BK_OK;
  }
  uint64_t **fb = flag_slot(p);
  BK_CUDA(cudaMemcpyAsync(fb, host, sizeof(host), cudaMemcpyHostToDevice, s));
  memcpy(p->last_list, host, sizeof(host));
  p->last_fb = fb, p->last_stream = s;
  *out = fb;
  return BK_OK;
}

// The wait for the peers' "my skin is final" flags is a ONE-CTA kernel (k_wait) launched just ahead of the pull on the same
// stream, not a spin inside the pull itself: a wide pull whose every CTA spins holds all the CTA slots of the GPU for as
// long as the slowest peer takes.  That (a) keeps the READY half of the split sweep off the SMs exactly when it should
// overlap the exchange, and (b) can deadlock two processes that each keep several independent domains in flight (the
// end-to-end leg of bench.py: domain 1's spinning pull fills GPU A and waits for B, while B's signal for domain 1 sits
// behind B's spinning pull of domain 0, which waits for a signal A cannot schedule).  A spinner that holds one small CTA
// never keeps another kernel from being scheduled.
int launch_copy(bk_xplan *p, uint64_t *const *wait_dev, int nwait, cudaStream_t s, int nsignal = 0,
                uint64_t *gate = nullptr, bool publish = false, uint64_t epoch = 0) {
  if (p->nchunks == 0 && nwait == 0 && !publish) return BK_OK;
  if (nwait > 0) {
    k_wait<<<1, 64, 0, s>>>(wait_dev, nwait, epoch);
    BK_LAUNCHED();
    nwait = 0;
    if (p->nchunks == 0 && !publish) return BK_OK;
  }
  unsigned long long want = p->nchunks ? p->nchunks : 1;
  unsigned threads = kThreads;
  unsigned long long cap = (unsigned long long) sm_count() * 8;
  if (p->shape_ctas > 0) {
    threads = (unsigned) p->shape_threads;
    cap = (unsigned long long) p->shape_ctas;
    want = (want + threads / kThreads - 1) / (threads / kThreads);
  }
  const unsigned grid = (unsigned) (want < cap ? want : cap);
  k_xplan<<<grid, threads, 0, s>>>(p->segs_dev, p->chunk_first_dev, p->nbox ? p->nbox : p->nseg, p->nchunks, wait_dev, nwait,
                                   nsignal, gate, publish ? p->done_dev : nullptr, p->boxes_dev, epoch);
  BK_LAUNCHED();
  return BK_OK;
}

}  // namespace

extern "C" {

int bk_xplan_create(bk_xplan_t **out, const bk_seg_t *segs, int nseg) {
  BK_REQUIRE(out && (segs || nseg == 0) && nseg >= 0, "bad arguments");
  std::vector<unsigned long long> first(nseg + 1, 0);
  size_t total = 0;
  for (int i = 0; i < nseg; ++i) {
    BK_REQUIRE(segs[i].bytes % 16 == 0 && ((size_t) segs[i].src % 16) == 0 && ((size_t) segs[i].dst % 16) == 0,
               "segments must be 16-byte aligned and sized");
    first[i + 1] = first[i] + (segs[i].bytes + kChunkBytes - 1) / kChunkBytes;
    total += segs[i].bytes;
  }
  bk_xplan *p = new bk_xplan();
  p->segs_host.assign(segs, segs + nseg);
  p->nseg = nseg;
  p->nchunks = first[nseg];
  p->bytes = total;
  if (nseg) {
    BK_CUDA(cudaMalloc(&p->segs_dev, sizeof(bk_seg_t) * nseg));
    BK_CUDA(cudaMemcpy(p->segs_dev, segs, sizeof(bk_seg_t) * nseg, cudaMemcpyHostToDevice));
  }
  BK_CUDA(cudaMalloc(&p->chunk_first_dev, sizeof(unsigned long long) * (nseg + 1)));
  BK_CUDA(cudaMemcpy(p->chunk_first_dev, first.data(), sizeof(unsigned long long) * (nseg + 1), cudaMemcpyHostToDevice));
  BK_CUDA(cudaMalloc(&p->flagbuf_dev, sizeof(uint64_t *) * kFlagSlots * kFlagSlotLen));
  BK_CUDA(cudaMalloc(&p->done_dev, sizeof(unsigned)));
  BK_CUDA(cudaMemset(p->done_dev, 0, sizeof(unsigned)));
  *out = p;
  return BK_OK;
}

int bk_xplan_create_boxes(bk_xplan_t **out, const bk_box_t *boxes, int nbox) {
  BK_REQUIRE(out && boxes && nbox > 0, "bad arguments");
  std::vector<unsigned long long> first(nbox + 1, 0);
  size_t total = 0;
  for (int i = 0; i < nbox; ++i) {
    BK_REQUIRE(boxes[i].src && boxes[i].dst && boxes[i].n[0] > 0 && boxes[i].n[1] > 0 && boxes[i].n[2] > 0, "empty box");
    const unsigned long long cells = (unsigned long long) boxes[i].n[0] * boxes[i].n[1] * boxes[i].n[2];
    first[i + 1] = first[i] + (cells + kBoxChunk - 1) / kBoxChunk;
    total += cells * sizeof(double);
  }
  bk_xplan *p = new bk_xplan();
  p->nbox = nbox;
  p->nchunks = first[nbox];
  p->bytes = total;
  BK_CUDA(cudaMalloc(&p->boxes_dev, sizeof(bk_box_t) * nbox));
  BK_CUDA(cudaMemcpy(p->boxes_dev, boxes, sizeof(bk_box_t) * nbox, cudaMemcpyHostToDevice));
  BK_CUDA(cudaMalloc(&p->chunk_first_dev, sizeof(unsigned long long) * (nbox + 1)));
  BK_CUDA(cudaMemcpy(p->chunk_first_dev, first.data(), sizeof(unsigned long long) * (nbox + 1), cudaMemcpyHostToDevice));
  BK_CUDA(cudaMalloc(&p->flagbuf_dev, sizeof(uint64_t *) * kFlagSlots * kFlagSlotLen));
  BK_CUDA(cudaMalloc(&p->done_dev, sizeof(unsigned)));
  BK_CUDA(cudaMemset(p->done_dev, 0, sizeof(unsigned)));
  *out = p;
  return BK_OK;
}

int bk_xplan_destroy(bk_xplan_t *p) {
  if (!p) return BK_OK;
  cudaFree(p->boxes_dev);
  cudaFree(p->segs_dev);
  cudaFree(p->chunk_first_dev);
  cudaFree(p->flagbuf_dev);
  cudaFree(p->done_dev);
  cudaFree(p->small_segs_dev);
  cudaFree(p->small_first_dev);
  for (int l = 0; l < p->nlanes; ++l) {
    cudaStreamDestroy(p->lanes[l]);
    cudaEventDestroy(p->ev_join[l]);
  }
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  delete p;
  return BK_OK;
}

size_t bk_xplan_bytes(const bk_xplan_t *p) { return p ? p->bytes : 0; }

int bk_xplan_set_shape(bk_xplan_t *p, int ctas, int threads_per_cta) {
  BK_REQUIRE(p, "null plan");
  BK_REQUIRE((ctas == 0 && threads_per_cta == 0) ||
                 (ctas > 0 && threads_per_cta >= 256 && threads_per_cta <= 1024 && threads_per_cta % 256 == 0),
             "shape: ctas > 0 and threads a multiple of 256 up to 1024 (or 0, 0 for the default)");
  p->shape_ctas = ctas, p->shape_threads = threads_per_cta;
  return BK_OK;
}

int bk_xplan_run(bk_xplan_t *p, void *stream) {
  BK_REQUIRE(p, "null plan");
  return launch_copy(p, nullptr, 0, (cudaStream_t) stream);
}

int bk_xplan_run_sync(bk_xplan_t *p, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                      int nsignal, uint64_t epoch, void *stream) {
  BK_REQUIRE(p, "null plan");
  BK_REQUIRE(nwait >= 0 && nwait <= 64 && nsignal >= 0 && nsignal <= 64, "at most 64 flags each");
  cudaStream_t s = (cudaStream_t) stream;
  // flag pointer lists travel through a small device scratch: [0,64) wait pointers, [64] = epoch, [65,129) signals
  uint64_t *host[130] = {nullptr};
  for (int i = 0; i < nwait; ++i) host[i] = const_cast<uint64_t *>(wait_flags[i]);
  for (int i = 0; i < nsignal; ++i) host[65 + i] = signal_flags[i];
  uint64_t **fb = nullptr;
  int rc = flag_list(p, host, s, &fb);
  if (rc != BK_OK) return rc;
  rc = launch_copy(p, fb, nwait, s, 0, nullptr, false, epoch);
  if (rc != BK_OK) return rc;
  if (nsignal > 0) {
    k_signal<<<1, 64, 0, s>>>(fb + 65, nsignal, epoch);
    BK_LAUNCHED();
  }
  return BK_OK;
}

int bk_xplan_run_gate(bk_xplan_t *p, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                      int nsignal, uint64_t *gate, uint64_t epoch, void *stream) {
  BK_REQUIRE(p, "null plan");
  BK_REQUIRE(nwait >= 0 && nwait <= 64 && nsignal >= 0 && nsignal <= 64, "at most 64 flags each");
  cudaStream_t s = (cudaStream_t) stream;
  uint64_t *host[130] = {nullptr};
  for (int i = 0; i < nwait; ++i) host[i] = const_cast<uint64_t *>(wait_flags[i]);
  for (int i = 0; i < nsignal; ++i) host[65 + i] = signal_flags[i];
  uint64_t **fb = nullptr;
  const int rc = flag_list(p, host, s, &fb);
  if (rc != BK_OK) return rc;
  return launch_copy(p, fb, nwait, s, nsignal, gate, true, epoch);
}

// The same plan with the big segments on the COPY ENGINES: no SM is taken from the sweep kernels, which matters because
// the marching kernels allocate the whole register file of an SM (a pull CTA cannot co-reside with them: a kernel-driven
// exchange and the sweep time-share SMs instead of overlapping).  On `stream`: a NARROW pull kernel (<= 32 CTAs) waits
// for the peers' ready flags and moves the small segments (edges, corners: a few MB, not worth 36 API calls); then the
// big segments (the six faces) fan out over a few lane streams as one cudaMemcpyAsync each (peer sources cross NVLink
// through the IPC / peer mappings); join; one tiny kernel raises the done flags.
int bk_xplan_run_ce(bk_xplan_t *p, const uint64_t *const *wait_flags, int nwait, uint64_t *const *signal_flags,
                    int nsignal, uint64_t epoch, void *stream) {
  BK_REQUIRE(p, "null plan");
  BK_REQUIRE(p->nbox == 0, "the copy-engine transport moves contiguous ranges, not array boxes");
  BK_REQUIRE(nwait >= 0 && nwait <= 64 && nsignal >= 0 && nsignal <= 64, "at most 64 flags each");
  cudaStream_t s = (cudaStream_t) stream;
  if (p->nlanes == 0) {
    int lo = 0, hi = 0;
    BK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    for (int l = 0; l < 4; ++l) {
      BK_CUDA(cudaStreamCreateWithPriority(&p->lanes[l], cudaStreamNonBlocking, hi));
      BK_CUDA(cudaEventCreateWithFlags(&p->ev_join[l], cudaEventDisableTiming));
    }
    BK_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    std::vector<bk_seg_t> small;
    size_t ce_min = kCeMinBytes;
    if (const char *e = getenv("BK_CE_MIN_BYTES")) ce_min = (size_t) atoll(e);  // tests: exercise both halves on small domains
    for (int i = 0; i < p->nseg; ++i) {
      if (p->segs_host[i].bytes >= ce_min) p->order.push_back(i);
      else if (p->segs_host[i].bytes) small.push_back(p->segs_host[i]);
    }
    std::stable_sort(p->order.begin(), p->order.end(),
                     [&](int a, int b) { return p->segs_host[a].bytes > p->segs_host[b].bytes; });
    size_t load[4] = {0, 0, 0, 0};
    p->lane_of.assign(p->nseg, 0);
    for (int i : p->order) {
      int best = 0;
      for (int l = 1; l < 4; ++l)
        if (load[l] < load[best]) best = l;
      p->lane_of[i] = best;
      load[best] += p->segs_host[i].bytes;
    }
    std::vector<unsigned long long> first(small.size() + 1, 0);
    for (size_t i = 0; i < small.size(); ++i) first[i + 1] = first[i] + (small[i].bytes + kChunkBytes - 1) / kChunkBytes;
    p->small_nseg = (int) small.size();
    p->small_nchunks = first[small.size()];
    if (!small.empty()) {
      BK_CUDA(cudaMalloc(&p->small_segs_dev, sizeof(bk_seg_t) * small.size()));
      BK_CUDA(cudaMemcpy(p->small_segs_dev, small.data(), sizeof(bk_seg_t) * small.size(), cudaMemcpyHostToDevice));
    }
    BK_CUDA(cudaMalloc(&p->small_first_dev, sizeof(unsigned long long) * first.size()));
    BK_CUDA(cudaMemcpy(p->small_first_dev, first.data(), sizeof(unsigned long long) * first.size(), cudaMemcpyHostToDevice));
    p->nlanes = 4;
  }
  uint64_t *host[130] = {nullptr};
  for (int i = 0; i < nwait; ++i) host[i] = const_cast<uint64_t *>(wait_flags[i]);
  for (int i = 0; i < nsignal; ++i) host[65 + i] = signal_flags[i];
  uint64_t **fb = nullptr;
  if (nwait > 0 || nsignal > 0) {
    const int rc = flag_list(p, host, s, &fb);
    if (rc != BK_OK) return rc;
  }
  if (p->small_nchunks > 0 || nwait > 0) {
    const unsigned grid = (unsigned) std::max(1ull, std::min(p->small_nchunks, 32ull));
    k_xplan<<<grid, kThreads, 0, s>>>(p->small_segs_dev, p->small_first_dev, p->small_nseg, p->small_nchunks,
                                      fb, nwait, 0, nullptr, nullptr, nullptr, epoch);
    BK_LAUNCHED();
  }
  if (!p->order.empty()) {
    BK_CUDA(cudaEventRecord(p->ev_fork, s));
    bool used[4] = {false, false, false, false};
    for (int i : p->order) {
      const bk_seg_t &g = p->segs_host[i];
      const int l = p->lane_of[i];
      if (!used[l]) BK_CUDA(cudaStreamWaitEvent(p->lanes[l], p->ev_fork, 0));
      used[l] = true;
      BK_CUDA(cudaMemcpyAsync(g.dst, g.src, g.bytes, cudaMemcpyDefault, p->lanes[l]));
    }
    for (int l = 0; l < 4; ++l)
      if (used[l]) {
        BK_CUDA(cudaEventRecord(p->ev_join[l], p->lanes[l]));
        BK_CUDA(cudaStreamWaitEvent(s, p->ev_join[l], 0));
      }
  }
  if (nsignal > 0) {
    k_signal<<<1, 64, 0, s>>>(fb + 65, nsignal, epoch);
    BK_LAUNCHED();
  }
  return BK_OK;
}

// The flag addresses travel as a kernel argument (64 pointers by value = 512 B of the 4 KiB parameter space): no device
// scratch, no stream-ordered allocation, no copy -- one launch per call.  (Until round 2 each call did cudaMallocAsync +
// cudaMemcpyAsync + launch + cudaFreeAsync on the critical path of every period.)
int bk_flags_signal(uint64_t *const *flags, int n, uint64_t value, void *stream) {
  BK_REQUIRE(flags && n > 0 && n <= 64, "1..64 flags");
  FlagList list = {};
  for (int i = 0; i < n; ++i) list.p[i] = flags[i];
  k_signal_list<<<1, 64, 0, (cudaStream_t) stream>>>(list, n, value);
  BK_LAUNCHED();
  return BK_OK;
}

int bk_flags_wait(const uint64_t *const *flags, int n, uint64_t value, void *stream) {
  BK_REQUIRE(flags && n > 0 && n <= 64, "1..64 flags");
  FlagList list = {};
  for (int i = 0; i < n; ++i) list.p[i] = const_cast<uint64_t *>(flags[i]);
  k_wait_list<<<1, 64, 0, (cudaStream_t) stream>>>(list, n, value);
  BK_LAUNCHED();
  return BK_OK;
}

}  // extern "C"
