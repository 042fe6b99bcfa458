// micro-benchmark: do warp shuffles share the shared-memory crossbar with LDS?  (decides whether moving the stencils'
// i-halo from LDS.64 to SHFL would relieve the shared-memory pipe)    nvcc -arch=sm_100a -O3 shfl_vs_lds.cu -o shfl_vs_lds
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>  // 0: LDS.64 only, 1: SHFL only, 2: both interleaved
__global__ void k(double *out, int iters) {
  __shared__ double s[1024];
  s[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double a = 0, b = threadIdx.x;
  const int l = threadIdx.x & 31, base = (threadIdx.x >> 5) * 32;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE != 1) a += *(volatile double *) &s[base + ((l + u + i) & 31)];
      if (MODE != 0) b = __shfl_down_sync(0xffffffffu, b, 1) + 1.0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
}
int main() {
  double *out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int iters = 20000;
  float ms[3];
  for (int m = 0; m < 3; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148 * 2, 1024>>>(out, iters);
      if (m == 1) k<1><<<148 * 2, 1024>>>(out, iters);
      if (m == 2) k<2><<<148 * 2, 1024>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[m], e0, e1);
    }
  }
  // per SM: 2 CTAs x 32 warps x iters x 8 ops; LDS.64 = 2 wavefronts, SHFL of a double = 2 SHFL.32
  const double ops = 2.0 * 32 * iters * 8;
  printf("LDS.64 only : %.3f ms  -> %.2f cycles per warp-LDS.64 per SM (at 1.9 GHz)\n", ms[0], ms[0] * 1.9e6 / ops);
  printf("SHFL64 only : %.3f ms  -> %.2f cycles per warp-shuffle-of-double per SM\n", ms[1], ms[1] * 1.9e6 / ops);
  printf("both        : %.3f ms  (sum of the two alone: %.3f ms, max: %.3f ms)\n", ms[2], ms[0] + ms[1], ms[0] > ms[1] ? ms[0] : ms[1]);
  return 0;
}
