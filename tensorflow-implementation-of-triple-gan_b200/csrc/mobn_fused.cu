// mobn_fused.cu -- mean-only batch normalisation (nn.py:147-187) WITHOUT a pass of its own over the activation.
//
// z = conv(x, W) is linear in x, so its batch mean is known before the contraction runs:
//     mean_s[co] = sum_{t, ci} W[t][co][ci] * S_s[t][ci] / count_s ,
// where S_s[t][ci] is the sum of x over the input positions tap t reads -- all positions minus the border rows / columns
// that tap never touches.  The producer of x (the previous layer's GEMM epilogue, the pooling kernel, or
// tgan_class_sums for the network input) emits nine border-class sums per channel and batch segment; this file turns them
// into the per-segment epilogue bias  shift_s = b - mean_s  and the pop_mean update.  The contraction then applies
// `- mean + b` and the leaky ReLU in its own epilogue (csrc/igemm_tc.cu) and the standalone apply pass disappears.
//
// Backward (ops.py): dz = du - mean_s(du) is linear as well; the producer's leaky-ReLU derivative is applied by the
// consumer's input-gradient epilogue through a 1-bit-per-element mask the forward epilogue wrote.
//
// Roofline: these kernels touch O(weights) bytes (<= 2.4 MB) once per layer: latency bound, a few microseconds.
#include "common.cuh"

namespace tgan {

__device__ __forceinline__ float q24f(const long long* p, int64_t i) { return (float)((double)p[i] * (1.0 / 16777216.0)); }

// one warp per output channel: mean over every segment, pop_mean chain in call order, shift[s][co] = b - mean
struct InvCount { float v[4]; };

__global__ void __launch_bounds__(256) mobn_mean_kernel(const long long* __restrict__ clsum, int nseg, const InvCount inv_count,
                                                        const bf16* __restrict__ wp, int T, int Cout, int Cin, int64_t w_ts,
                                                        int64_t w_cs, const float* __restrict__ b, float* __restrict__ pop_mean,
                                                        float decay, float* __restrict__ shift) {
  pdl_entry();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= Cout) return;
  const int co = warp;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < T; ++t) {
    // 3x3 / stride 1 / SAME, taps row-major: tap row r reads input rows [r-1, H-2+r] -> r = 0 skips the last row class,
    // r = 2 the first; same for columns.  T == 1 (1x1 / dense): everything.
    const int r = T == 9 ? t / 3 : 1, c = T == 9 ? t % 3 : 1;
    const int rc0 = r == 2 ? 1 : 0, rc1 = r == 0 ? 1 : 2, cc0 = c == 2 ? 1 : 0, cc1 = c == 0 ? 1 : 2;
    const bf16* wrow = wp + (int64_t)t * w_ts + (int64_t)co * w_cs;
    for (int ci = lane; ci < Cin; ci += 32) {
      const float w = __bfloat162float(wrow[ci]);
      for (int s = 0; s < nseg; ++s) {
        float S = 0.f;
        for (int rc = rc0; rc <= rc1; ++rc)
          for (int cc = cc0; cc <= cc1; ++cc) S += q24f(clsum, ((int64_t)s * 9 + rc * 3 + cc) * Cin + ci);
        acc[s] += w * S;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
  if (lane == 0) {
    float pm = pop_mean ? pop_mean[co] : 0.f;
    for (int s = 0; s < nseg; ++s) {
      const float m = acc[s] * inv_count.v[s];
      shift[s * Cout + co] = b[co] - m;
      pm = pm * decay + m * (1.f - decay);      // one update per call, in call order (nn.py:181)
    }
    if (pop_mean) pop_mean[co] = pm;
  }
}

struct ClsSegs { int end[4]; int n; float inv_count[4]; };

// border-class sums of a small-channel tensor (the classifier's 3-channel input): one CTA per image
__global__ void __launch_bounds__(256) class_sums_kernel(const void* __restrict__ x, int xdt, int H, int W, int C, int ld,
                                                         ClsSegs sg, long long* __restrict__ clsum) {
  pdl_entry();
  __shared__ unsigned long long part[9 * 16];      // Q24 integers: the sums do not depend on the order of the atomics
  for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) part[i] = 0ull;
  __syncthreads();
  const int n = blockIdx.x;
  const int s = (n >= sg.end[0]) + (n >= sg.end[1]) + (n >= sg.end[2]);
  for (int i = threadIdx.x; i < H * W * C; i += blockDim.x) {
    const int c = i % C, px = i / C, xx = px % W, yy = px / W;
    const int64_t off = ((int64_t)n * H * W + px) * ld + c;
    float v = xdt == TGAN_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(x)[off]) : reinterpret_cast<const float*>(x)[off];
    v = __bfloat162float(__float2bfloat16_rn(v));      // the contraction reads the bf16-rounded value
    const int rc = yy == 0 ? 0 : yy == H - 1 ? 2 : 1, cc = xx == 0 ? 0 : xx == W - 1 ? 2 : 1;
    atomicAdd(&part[(rc * 3 + cc) * 16 + c], (unsigned long long)__float2ll_rn(v * 16777216.f));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) {
    const int k = i / 16, c = i % 16;
    if (c < C && part[i] != 0ull)
      atomicAdd(reinterpret_cast<unsigned long long*>(&clsum[((int64_t)s * 9 + k) * C + c]), part[i]);
  }
}

// Backward bookkeeping of a fused layer: the consumer's input-gradient epilogue left per-segment channel sums of
// du = dy * lrelu'(y) in Q24; this turns them into the fp32 sums the mean-subtraction reads and adds db = sum_s.
__global__ void seg_sums_finalize_kernel(const long long* __restrict__ q, int nseg, int C, float* __restrict__ colsums,
                                         float* __restrict__ grad_acc) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float tot = 0.f;
  for (int s = 0; s < 4; ++s) {
    const float v = s < nseg ? q24f(q, (int64_t)s * C + c) : 0.f;
    colsums[s * C + c] = v;
    tot += v;
  }
  if (grad_acc) grad_acc[c] += tot;
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_mobn_mean_from_sums(const void* clsum, int nseg, const int64_t* count, const void* wp, int T, int Cout,
                                        int Cin, int64_t w_tap_stride, int64_t w_co_stride, const float* b, float* pop_mean,
                                        float decay, float* shift, void* stream) {
  TGAN_CHECK_ARG(clsum && count && wp && b && shift && nseg >= 1 && nseg <= 4 && (T == 1 || T == 9) && Cout > 0 && Cin > 0 &&
                     w_co_stride >= 1, "mobn_mean_from_sums: bad args");
  InvCount ic;      // by value in the kernel arguments: nothing to stage, CUDA-graph capturable
  for (int s = 0; s < 4; ++s) ic.v[s] = s < nseg ? (float)(1.0 / (double)count[s]) : 0.f;
  pdl_launch(mobn_mean_kernel, ceil_div(Cout * 32, 256), 256, 0, (cudaStream_t)stream, (const long long*)clsum, nseg, ic,
             (const bf16*)wp, T, Cout, Cin, w_tap_stride, w_co_stride, b, pop_mean, decay, shift);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_class_sums(const void* x, int xdt, int N, int H, int W, int C, int ld, int nseg,
                               const int* seg_end_images, void* clsum, void* stream) {
  TGAN_CHECK_ARG(x && clsum && N > 0 && H > 1 && W > 1 && C >= 1 && C <= 16 && ld >= C && nseg >= 1 && nseg <= 4,
                 "class_sums: bad args (C <= 16)");
  ClsSegs sg;
  for (int i = 0; i < 4; ++i) sg.end[i] = (nseg > 1 && i < nseg - 1 && seg_end_images) ? seg_end_images[i] : 0x7fffffff;
  sg.n = nseg;
  pdl_launch(class_sums_kernel, N, 256, 0, (cudaStream_t)stream, x, xdt, H, W, C, ld, sg, (long long*)clsum);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_seg_sums_finalize(const void* q24, int nseg, int C, float* colsums, float* grad_acc, void* stream) {
  TGAN_CHECK_ARG(q24 && colsums && nseg >= 1 && nseg <= 4 && C > 0, "seg_sums_finalize: bad args");
  pdl_launch(seg_sums_finalize_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)stream, (const long long*)q24, nseg, C, colsums, grad_acc);
  TGAN_LAUNCHED();
  return 0;
}
