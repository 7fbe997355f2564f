// colreduce.cuh -- deterministic per-channel (column) reductions over a [rows, C] matrix with an optional fused
// elementwise output, in ONE launch.  Each CTA owns 32 channels x a strided subset of the rows (256 threads:
// 8 channel lanes x 4-wide vectors x 32 row lanes when C % 4 == 0, else 32 x 8), coalesced along the NHWC channel
// axis, and publishes one partial per accumulator.  The last CTA of a channel block to finish (device-scope ticket)
// folds the partials in a fixed order, so the statistics are bit-reproducible without a second kernel and without
// floating-point atomics.  Accumulation is in fp64 end to end (registers, shared memory, partials): bias / beta
// gradients behind a batch norm are sums of large terms that cancel almost exactly (sum of BN input-gradients is 0),
// so fp32 summation noise would otherwise depend on the fold order at the 1e-2 level.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace tgan {

// Ticket counters of the "last CTA folds" protocol, one set of 1024 per WORKSPACE: launches that may run concurrently (on
// different streams) use different workspaces, hence different counters.  Zero at load; the last CTA resets its slot.
constexpr int COLREDUCE_WS_SLOTS = 8;
__device__ unsigned int g_colreduce_ticket[COLREDUCE_WS_SLOTS][1024];

template <typename F, typename = void>
struct has_split_load : std::false_type {};
template <typename F>
struct has_split_load<F, std::void_t<typename F::Regs>> : std::true_type {};

template <int NACC, int VEC, typename F>
__global__ void __launch_bounds__(256) colreduce_kernel(F f, int64_t rows, int C, double* __restrict__ partials,
                                                        float* o0, float* o1, float beta, float* acc0, int segmode, int slot) {
  pdl_entry();
  constexpr int CL = VEC == 4 ? 8 : 32, RL = 256 / CL;
  __shared__ double sm[RL][NACC][33];
  __shared__ bool is_last;
  const int t = threadIdx.x, tx = t % CL, ty = t / CL;
  const int c0 = blockIdx.x * 32 + tx * VEC;
  double acc[NACC][VEC];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[a][j] = 0.0;
  if (c0 < C) {
    const int64_t step = (int64_t)gridDim.y * RL;
    int64_t r = (int64_t)blockIdx.y * RL + ty;
    if constexpr (has_split_load<F>::value) {
      // functors with a load / finish split: the loads of four rows are issued before the first result is stored (the
      // one-row-at-a-time loop keeps only two 8-byte loads per thread in flight and runs at a third of the HBM rate)
      for (; r + 3 * step < rows; r += 4 * step) {
        typename F::Regs q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) f.load(r + u * step, c0, q[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float v[NACC][VEC];
          f.finish(r + u * step, c0, q[u], v);
#pragma unroll
          for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int j = 0; j < VEC; ++j) acc[a][j] += (double)v[a][j];
        }
      }
    }
    for (; r < rows; r += step) {
      float v[NACC][VEC];
      f(r, c0, v);
#pragma unroll
      for (int a = 0; a < NACC; ++a)
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[a][j] += (double)v[a][j];
    }
  }
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < VEC; ++j) sm[ty][a][tx * VEC + j] = acc[a][j];
  __syncthreads();
  if (t < 32 * NACC) {
    const int a = t / 32, cc = t % 32, c = blockIdx.x * 32 + cc;
    if (c < C) {
      double s = 0.0;
#pragma unroll
      for (int y = 0; y < RL; ++y) s += sm[y][a][cc];
      partials[((int64_t)blockIdx.y * NACC + a) * C + c] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) is_last = (atomicAdd(&g_colreduce_ticket[slot][blockIdx.x], 1u) == gridDim.y - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ---- last CTA of this channel block: fold gridDim.y partials (32 channels x 8 part lanes), fixed order ----
  const int fc = t % 32, fl = t / 32, c = blockIdx.x * 32 + fc;
  const int parts = gridDim.y;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    double s = 0.0;
    if (c < C)
      for (int p = fl; p < parts; p += 8) s += __ldcg(&partials[((int64_t)p * NACC + a) * C + c]);
    sm[fl][a][fc] = s;
  }
  __syncthreads();
  if (t < 32 && c < C) {
    double tot = 0.0;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      double s = 0.0;
#pragma unroll
      for (int y = 0; y < 8; ++y) s += sm[y][a][fc];
      tot += s;
      if (segmode) {            // accumulator a = batch segment a: o0 is [NACC][C]; acc0 += the sum over all segments
        o0[a * C + c] = (float)s;
      } else {
        float* o = a == 0 ? o0 : o1;
        if (o) o[c] = (float)((beta != 0.f ? (double)beta * o[c] : 0.0) + s);
        if (a == 0 && acc0) acc0[c] = (float)((double)acc0[c] + s);
      }
    }
    if (segmode && acc0) acc0[c] = (float)((double)acc0[c] + tot);
  }
  if (t == 0) g_colreduce_ticket[slot][blockIdx.x] = 0;
}

// workspace pointer -> ticket slot (first come, first served; the step uses two workspaces: main and side stream)
static inline int colreduce_slot(const void* ws) {
  static const void* seen[COLREDUCE_WS_SLOTS] = {nullptr};
  static int next = 0;
  for (int i = 0; i < COLREDUCE_WS_SLOTS; ++i)
    if (seen[i] == ws) return i;
  // a new workspace takes the oldest slot: its previous owner belongs to a context that was torn down (the counters
  // are left at zero by every launch)
  const int i = next;
  next = (next + 1) % COLREDUCE_WS_SLOTS;
  seen[i] = ws;
  return i;
}

static inline int pick_parts(int64_t rows, int C, int vec) {
  const int rl = vec == 4 ? 32 : 8;
  int xblocks = ceil_div(C, 32);
  int64_t want = (148 * 4 + xblocks - 1) / xblocks;           // ~4 CTAs per SM over the whole grid
  int64_t maxp = (rows + rl * 2 - 1) / (rl * 2);              // >= 2 rows per thread
  int64_t p = want < maxp ? want : maxp;
  if (p < 1) p = 1;
  if (p > TGAN_STATS_MAX_PARTS) p = TGAN_STATS_MAX_PARTS;
  return (int)p;
}

// out0/out1 = beta*out + column sums of accumulator 0/1 (either may be NULL); acc0 (optional) += sums of accumulator 0
template <int NACC, typename F1, typename F4>
static int run_colreduce(F1 f1, F4 f4, bool vec_ok, int64_t rows, int C, float* o0, float* o1, float beta, float* ws,
                         cudaStream_t st, float* acc0 = nullptr, int segmode = 0) {
  if (ceil_div(C, 32) > 1024) { set_error("colreduce: more than 32768 channels"); return 1; }
  int vec = vec_ok ? 4 : 1;
  const int slot = colreduce_slot(ws);
  int parts = pick_parts(rows, C, vec);
  dim3 grid(ceil_div(C, 32), parts);
  double* wsd = reinterpret_cast<double*>(ws);       // fp64 partials: the first 4*MAX_PARTS*C floats of ws
  if (vec_ok) pdl_launch(colreduce_kernel<NACC, 4, F4>, grid, 256, 0, (cudaStream_t)(st), f4, rows, C, wsd, o0, o1, beta, acc0, segmode, slot);
  else pdl_launch(colreduce_kernel<NACC, 1, F1>, grid, 256, 0, (cudaStream_t)(st), f1, rows, C, wsd, o0, o1, beta, acc0, segmode, slot);
  TGAN_LAUNCHED();
  return 0;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tgan
