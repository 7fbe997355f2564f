// Does a K-major SWIZZLE_128B UMMA operand descriptor whose start address is shifted by s rows (s*128 B, not a multiple
// of the 1024 B swizzle atom) address rows [s, s+N) of a tile written with the address-based TMA swizzle pattern?
// Tested with base_offset = 0 and base_offset = (start >> 7) & 7.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../tensorflow-implementation-of-triple-gan_b200/csrc/tc_common.cuh"
using namespace tgan;
// element (r, k) of a [rows][64] bf16 K-major tile in the canonical 128B-swizzled layout at a 1024B-aligned base
__device__ __forceinline__ uint32_t sw_off(int r, int k) { return r * 128 + ((((k >> 3) ^ (r & 7)) << 4) | ((k & 7) << 1)); }
__global__ void __launch_bounds__(128, 1) k(int shift, int use_bo, int mode, int kstep, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem; uint8_t* sB = smem + 16384;
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    int r = i / 64, kk = i % 64;
    *(__nv_bfloat16*)(sA + sw_off(r, kk)) = __float2bfloat16((kk % 16) == (r % 16) ? 1.f : 0.f);
  }
  for (int i = threadIdx.x; i < 288 * 64; i += 128) {
    int r = i / 64, kk = i % 64;
    *(__nv_bfloat16*)(sB + sw_off(r, kk)) = __float2bfloat16(mode == 0 ? (float)(r % 256) : (float)kk);
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
      const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t a = (smem_u32(sA) >> 4) + 2 * kstep;
      const uint32_t bsa = smem_u32(sB) + shift * 128 + 32 * kstep;
      uint64_t bd = hi | (uint64_t)((bsa >> 4) & 0x3FFF);
      if (use_bo) bd |= (uint64_t)((bsa >> 7) & 7) << 49;
      umma_bf16(tm, hi | a, bd, idesc, 0u);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // thread t of warp w reads TMEM lane 32*w + t, columns 0..255
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (threadIdx.x & 31)) * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}
int main() {
  float* d; cudaMalloc(&d, 128 * 256 * 4);
  static float h[128 * 256];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int shift = 0; shift <= 19; ++shift) {
      int bad[2] = {0, 0};
      for (int mode = 0; mode < 2; ++mode)
        for (int kstep = 0; kstep < 4; ++kstep) {
          k<<<1, 128, 80 * 1024>>>(shift, use_bo, mode, kstep, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 256; ++n) {
              // D[m][n] = sum_k A[m][k] B[n+shift][k] over k in [16*kstep, 16*kstep+16) = B[n+shift][16*kstep + m%16]
              float want = mode == 0 ? (float)((n + shift) % 256) : (float)(16 * kstep + m % 16);
              if (h[m * 256 + n] != want) ++bad[mode];
            }
        }
      printf("base_offset %s shift %2d rows: row-mapping mismatches %d, k-chunk mismatches %d\n", use_bo ? "set " : "zero", shift,
             bad[0], bad[1]);
    }
  return 0;
}
