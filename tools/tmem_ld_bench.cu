// micro-benchmark: tcgen05.ld.32x32b.x32 latency / throughput with 1, 4, 8 concurrent warps, waiting after every load or
// after groups of 4 loads
#include <cstdio>
#include <cuda_runtime.h>
#include "../tensorflow-implementation-of-triple-gan_b200/csrc/tc_common.cuh"
using namespace tgan;
__global__ void __launch_bounds__(256, 1) k(int nwarps, int group, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  float acc = 0.f;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    for (int i = 0; i < iters; ++i) {
      uint32_t r[4][32];
      for (int g = 0; g < 4; ++g) {
        if (g < group) tmem_ld32(base + ((i * 4 + g) & 7) * 32, r[g]);
      }
      tmem_ld_wait();
      for (int g = 0; g < 4; ++g) if (g < group) for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[g][j]);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 1.2345f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
int main() {
  long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  for (int nw : {1, 4, 8}) for (int group : {1, 2, 4}) {
    const int iters = 2000;
    k<<<148, 256>>>(nw, group, iters, d, s);
    k<<<148, 256>>>(nw, group, iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("warps %d, loads per wait %d: %7.1f cycles per x32 load per warp; %6.1f B/clk/SM  %s\n", nw, group,
           (double)c / (iters * group), (double)nw * iters * group * 4096 / (double)c, cudaGetErrorString(e));
  }
  return 0;
}
