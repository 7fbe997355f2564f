// colreduce.cuh -- deterministic per-channel (column) reductions over a [rows, C] matrix with an
// optional fused elementwise output.  Stage 1: (32 channel lanes x 8 row lanes) CTAs, grid sized to
// ~4 CTAs per SM, coalesced (vectorised x4) along the NHWC channel axis, shared-memory reduce across the
// row lanes, one partial per CTA row-group.  Stage 2: one thread per channel sums the partials in a
// fixed order (no atomics -> bit-reproducible statistics).
#pragma once
#include "common.cuh"

namespace tgan {

// ------------------------------------------------------------------------------------------------
constexpr int RY = 8;  // row lanes per block

template <int NACC, int VEC, typename F>
__global__ void __launch_bounds__(32 * RY) colreduce_kernel(F f, int64_t rows, int C, float* __restrict__ partials) {
  __shared__ float sm[RY][NACC][32 * VEC + 1];
  const int c0 = (blockIdx.x * 32 + threadIdx.x) * VEC;
  float acc[NACC][VEC];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[a][j] = 0.f;
  if (c0 < C) {
    for (int64_t r = (int64_t)blockIdx.y * RY + threadIdx.y; r < rows; r += (int64_t)gridDim.y * RY) {
      float v[NACC][VEC];
      f(r, c0, v);
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[a][j] += v[a][j];
    }
  }
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < VEC; ++j) sm[threadIdx.y][a][threadIdx.x * VEC + j] = acc[a][j];
  __syncthreads();
  const int t = threadIdx.y * 32 + threadIdx.x;
  for (int e = t; e < NACC * 32 * VEC; e += 32 * RY) {
    int a = e / (32 * VEC), cc = e % (32 * VEC);
    int c = blockIdx.x * 32 * VEC + cc;
    if (c >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < RY; ++y) s += sm[y][a][cc];
    partials[((int64_t)blockIdx.y * NACC + a) * C + c] = s;
  }
}

// out[a][c] = beta*out[a][c] + sum_p partials[p][a][c].  (32 channels x 8 part-lanes) per CTA: the partial
// loads of one channel are spread over 8 threads and issued back to back (MLP), then folded in a fixed order.
template <int NACC>
__global__ void __launch_bounds__(32 * RY) colreduce_final_kernel(const float* __restrict__ partials, int parts, int C,
                                                                  float* o0, float* o1, float beta) {
  __shared__ float sm[RY][NACC][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
  if (c < C) {
    for (int p = threadIdx.y; p < parts; p += RY) {
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] += partials[((int64_t)p * NACC + a) * C + c];
    }
  }
#pragma unroll
  for (int a = 0; a < NACC; ++a) sm[threadIdx.y][a][threadIdx.x] = acc[a];
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      float* o = a == 0 ? o0 : o1;
      if (!o) continue;
      float s = 0.f;
#pragma unroll
      for (int y = 0; y < RY; ++y) s += sm[y][a][threadIdx.x];
      o[c] = (beta != 0.f ? beta * o[c] : 0.f) + s;
    }
  }
}

static inline int pick_parts(int64_t rows, int C, int vec) {
  int xblocks = ceil_div(C, 32 * vec);
  int64_t want = (148 * 4 + xblocks - 1) / xblocks;           // ~4 CTAs per SM over the whole grid
  int64_t maxp = (rows + RY * 4 - 1) / (RY * 4);              // >= 4 rows per thread
  int64_t p = want < maxp ? want : maxp;
  if (p < 1) p = 1;
  if (p > TGAN_STATS_MAX_PARTS) p = TGAN_STATS_MAX_PARTS;
  return (int)p;
}

template <int NACC, typename F1, typename F4>
static int run_colreduce(F1 f1, F4 f4, bool vec_ok, int64_t rows, int C, float* o0, float* o1, float beta, float* ws,
                         cudaStream_t st) {
  int vec = vec_ok ? 4 : 1;
  int parts = pick_parts(rows, C, vec);
  dim3 grid(ceil_div(C, 32 * vec), parts), block(32, RY);
  if (vec_ok) colreduce_kernel<NACC, 4, F4><<<grid, block, 0, st>>>(f4, rows, C, ws);
  else colreduce_kernel<NACC, 1, F1><<<grid, block, 0, st>>>(f1, rows, C, ws);
  TGAN_LAUNCHED();
  colreduce_final_kernel<NACC><<<ceil_div(C, 32), dim3(32, RY), 0, st>>>(ws, parts, C, o0, o1, beta);
  TGAN_LAUNCHED();
  return 0;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


}  // namespace tgan
