// micro-benchmark: back-to-back tcgen05.mma (M=128, K=16, bf16, smem operands) issue rate vs N, one CTA per SM
#include <cstdio>
#include <cuda_runtime.h>
#include "../tensorflow-implementation-of-triple-gan_b200/csrc/tc_common.cuh"
using namespace tgan;
__global__ void __launch_bounds__(128, 1) k(int N, int iters, int kmajor, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t a = smem_u32(smem) >> 4, b = a + (16384 >> 4);
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        umma_bf16(tm, hi | (a + 0), hi | (b + 0), idesc, 1u);
        umma_bf16(tm, hi | (a + 2), hi | (b + 2), idesc, 1u);
        umma_bf16(tm, hi | (a + 4), hi | (b + 4), idesc, 1u);
        umma_bf16(tm, hi | (a + 6), hi | (b + 6), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int Ns[] = {32, 64, 96, 128, 144, 160, 192, 208, 224, 240, 256};
  for (int grid : {1, 148}) for (int N : Ns) {
    const int iters = 2000;
    k<<<grid, 128, 64 * 1024>>>(N, iters, 1, d);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<grid, 128, 64 * 1024>>>(N, iters, 1, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d N=%3d: %6.1f cycles/MMA (clock64)  %.1f ns/MMA (events)  -> %.0f TFLOP/s at 148 SMs  %s\n", grid, N,
           (double)c / (4.0 * iters), ms * 1e6 / (4.0 * iters), 2.0 * 128 * N * 16 * 148 / (ms * 1e6 / (4.0 * iters)) / 1e3,
           cudaGetErrorString(e));
  }
  return 0;
}
