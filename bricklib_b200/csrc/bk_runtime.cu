// bk_runtime.cu -- device plumbing behind the C ABI: memory, streams, events, IPC.  Replaces the reference's
// movBrickInfo / movBrickStorage / copyToDevice helpers (include/brick-gpu.h:43-103, stencils/cudaarray.h:11-31),
// which cudaMalloc + blocking cudaMemcpy with 32-bit sizes; here sizes are size_t and copies are stream-ordered.
#include "bk_common.h"
#include <sched.h>
#include <cctype>
#include <cstring>
#include <mutex>

namespace bk {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace bk

extern "C" {

const char *bk_version(void) { return "bricklib_b200 0.1 (sm_100a)"; }
const char *bk_last_error(void) { return bk::g_err; }
unsigned long long bk_launch_count(void) { return bk::g_launches.load(); }

int bk_device_count(int *n) {
  BK_REQUIRE(n, "null");
  BK_CUDA(cudaGetDeviceCount(n));
  return BK_OK;
}
int bk_set_device(int dev) {
  BK_CUDA(cudaSetDevice(dev));
  return BK_OK;
}
// NUMA placement: pinned buffers are first-touch pages, and an 8-GPU box hangs its GPUs off two sockets.  Bind the
// calling thread (and so the pages it touches next) to the CPUs local to the current device's PCIe root, as listed by
// /sys/bus/pci/devices/<bus id>/local_cpulist.  Best effort: BK_EUNSUPPORTED when sysfs has no answer; nothing changes.
int bk_bind_host_to_device(void) {
  int dev = 0;
  BK_CUDA(cudaGetDevice(&dev));
  char bus[32] = "";
  BK_CUDA(cudaDeviceGetPCIBusId(bus, sizeof(bus), dev));
  for (char *c = bus; *c; ++c) *c = (char) tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE *f = fopen(path, "r");
  if (!f) {
    bk::set_error("bk_bind_host_to_device: %s is not readable", path);
    return BK_EUNSUPPORTED;
  }
  char list[1024] = "";
  const bool got = fgets(list, sizeof(list), f) != nullptr;
  fclose(f);
  cpu_set_t want, have;
  CPU_ZERO(&want);
  int n = 0;
  for (char *tok = got ? strtok(list, ",\n") : nullptr; tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    if (sscanf(tok, "%d-%d", &a, &b) == 2) {
    } else if (sscanf(tok, "%d", &a) == 1) {
      b = a;
    } else {
      continue;
    }
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET(c, &want), ++n;
  }
  // stay inside the mask we were given (containers, taskset)
  if (n == 0 || sched_getaffinity(0, sizeof(have), &have) != 0) return BK_EUNSUPPORTED;
  CPU_AND(&want, &want, &have);
  if (CPU_COUNT(&want) == 0 || sched_setaffinity(0, sizeof(want), &want) != 0) return BK_EUNSUPPORTED;
  return BK_OK;
}
int bk_dev_alloc(void **dev, size_t bytes) {
  BK_REQUIRE(dev, "null");
  BK_CUDA(cudaMalloc(dev, bytes ? bytes : 1));
  return BK_OK;
}
int bk_dev_free(void *dev) {
  if (dev) bk_adjacency_forget(dev);  // verdicts about an adjacency list / grid at this address die with the allocation
  BK_CUDA(cudaFree(dev));
  return BK_OK;
}
int bk_dev_memset(void *dev, int byte, size_t bytes, void *stream) {
  BK_CUDA(cudaMemsetAsync(dev, byte, bytes, (cudaStream_t) stream));
  return BK_OK;
}
int bk_host_alloc(void **host, size_t bytes) {
  BK_REQUIRE(host, "null");
  BK_CUDA(cudaHostAlloc(host, bytes ? bytes : 1, cudaHostAllocDefault));
  return BK_OK;
}
int bk_host_free(void *host) {
  BK_CUDA(cudaFreeHost(host));
  return BK_OK;
}
int bk_memcpy_h2d(void *dev, const void *host, size_t bytes, void *stream) {
  BK_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream));
  return BK_OK;
}
int bk_memcpy_d2h(void *host, const void *dev, size_t bytes, void *stream) {
  BK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
  return BK_OK;
}
int bk_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  BK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t) stream));
  return BK_OK;
}
int bk_stream_create(void **stream) {
  BK_REQUIRE(stream, "null");
  cudaStream_t s;
  BK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void *) s;
  return BK_OK;
}
int bk_stream_create_priority(void **stream, int high) {
  BK_REQUIRE(stream, "null");
  int least = 0, greatest = 0;  // numerically lower = higher priority
  BK_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  cudaStream_t s;
  BK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? greatest : least));
  *stream = (void *) s;
  return BK_OK;
}
int bk_stream_destroy(void *stream) {
  BK_CUDA(cudaStreamDestroy((cudaStream_t) stream));
  return BK_OK;
}
int bk_stream_sync(void *stream) {
  BK_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
  return BK_OK;
}
int bk_device_sync(void) {
  BK_CUDA(cudaDeviceSynchronize());
  return BK_OK;
}
int bk_event_create(void **ev) {
  BK_REQUIRE(ev, "null");
  cudaEvent_t e;
  BK_CUDA(cudaEventCreate(&e));
  *ev = (void *) e;
  return BK_OK;
}
int bk_event_destroy(void *ev) {
  BK_CUDA(cudaEventDestroy((cudaEvent_t) ev));
  return BK_OK;
}
int bk_event_record(void *ev, void *stream) {
  BK_CUDA(cudaEventRecord((cudaEvent_t) ev, (cudaStream_t) stream));
  return BK_OK;
}
int bk_event_sync(void *ev) {
  BK_CUDA(cudaEventSynchronize((cudaEvent_t) ev));
  return BK_OK;
}
int bk_event_elapsed_ms(void *start, void *stop, float *ms) {
  BK_REQUIRE(ms, "null");
  BK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t) start, (cudaEvent_t) stop));
  return BK_OK;
}
int bk_stream_wait_event(void *stream, void *ev) {
  BK_CUDA(cudaStreamWaitEvent((cudaStream_t) stream, (cudaEvent_t) ev, 0));
  return BK_OK;
}

int bk_ipc_export(void *dev, unsigned char *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == BK_IPC_HANDLE_BYTES, "handle size");
  BK_REQUIRE(dev && handle64, "null");
  cudaIpcMemHandle_t h;
  BK_CUDA(cudaIpcGetMemHandle(&h, dev));
  memcpy(handle64, &h, sizeof(h));
  return BK_OK;
}
int bk_ipc_open(const unsigned char *handle64, void **dev) {
  BK_REQUIRE(dev && handle64, "null");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  BK_CUDA(cudaIpcOpenMemHandle(dev, h, cudaIpcMemLazyEnablePeerAccess));
  return BK_OK;
}
int bk_ipc_close(void *dev) {
  BK_CUDA(cudaIpcCloseMemHandle(dev));
  return BK_OK;
}
int bk_peer_enable(int peer_dev) {
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return BK_OK;
  }
  BK_CUDA(e);
  return BK_OK;
}

}  // extern "C"
