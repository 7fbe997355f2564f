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

// One warp per output channel.  A lane owns the input channels ci = lane, lane + 32, ...: per (segment, ci) it loads
// the nine class sums, forms the nine tap sums in registers (a tap at row offset -1 never reads the LAST input row, +1
// never the first; same for columns) and multiplies them with the nine tap weights -- every load of an iteration is
// independent of the others, so the kernel is not bound by one memory latency per multiply.  The warp then reduces,
// updates pop_mean in call order and writes shift[s][co] = b - mean_s.
struct InvCount { float v[4]; };
struct ClsSegs { int end[4]; int n; };
typedef ClsSegs ClsSegsFwd;

__global__ void __launch_bounds__(128) mobn_mean_kernel(const long long* __restrict__ clsum, int nseg, const InvCount inv_count,
                                                        const bf16* __restrict__ wp, int T, int Cout, int Cin, int64_t w_ts,
                                                        int64_t w_cs, const float* __restrict__ b, float* __restrict__ pop_mean,
                                                        float decay, float* __restrict__ shift) {
  pdl_entry();
  // one CTA (4 warps) per output channel: the 128 threads split the input channels, so even Cin = 256 is two
  // iterations per thread and Cout CTAs keep every SM busy
  __shared__ float red[4][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ci = threadIdx.x; ci < Cin; ci += 128) {
    float w[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t] = t < T ? __bfloat162float(wp[(int64_t)t * w_ts + (int64_t)co * w_cs + ci]) : 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s >= nseg) break;
      long long a[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) a[k] = clsum[((int64_t)s * 9 + k) * Cin + ci];
      if (T == 9) {
        // rows: R[r][cc], r = 0 -> classes {first, interior}, 1 -> all, 2 -> {interior, last}
        long long R[3][3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          R[0][cc] = a[cc] + a[3 + cc];
          R[2][cc] = a[3 + cc] + a[6 + cc];
          R[1][cc] = R[0][cc] + a[6 + cc];
        }
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const long long c0 = R[r][0] + R[r][1], c2 = R[r][1] + R[r][2], c1 = c0 + R[r][2];
          sum += w[r * 3 + 0] * (float)((double)c0 * (1.0 / 16777216.0));
          sum += w[r * 3 + 1] * (float)((double)c1 * (1.0 / 16777216.0));
          sum += w[r * 3 + 2] * (float)((double)c2 * (1.0 / 16777216.0));
        }
        acc[s] += sum;
      } else {
        long long tot = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) tot += a[k];
        acc[s] += w[0] * (float)((double)tot * (1.0 / 16777216.0));
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], o);
    if (lane == 0) red[warp][s] = acc[s];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float pm = pop_mean ? pop_mean[co] : 0.f;
    for (int s = 0; s < nseg; ++s) {
      const float m = (((red[0][s] + red[1][s]) + red[2][s]) + red[3][s]) * inv_count.v[s];
      shift[s * Cout + co] = b[co] - m;
      pm = pm * decay + m * (1.f - decay);      // one update per call, in call order (nn.py:181)
    }
    if (pop_mean) pop_mean[co] = pm;
  }
}

// Border-class sums of a wide bf16 tensor [N, H, W, C] (C % 8 == 0): one CTA per image, thread = (8 channels, one image
// row).  A thread folds its row into (first / interior / last column) x 8 channels in registers; the rows then meet in
// shared memory in a FIXED order (first row, interior rows top to bottom, last row), so the per-image sums are bit
// reproducible, and leave as Q24 integer atomics (one per class and channel and image).
// border-class sums of a small-channel tensor (the classifier's 3-channel input), one CTA per image.  Interior pixels
// (88 % of a 32x32 image) accumulate in registers and meet in a warp shuffle; border pixels and the per-warp interior
// totals go through Q24 integer atomics in shared memory, so the result does not depend on the order of the adds.
constexpr int CLS_IMGS = 1;      // images per CTA (4 was measured 3x slower: 256 threads walk the images one after the other)
__global__ void __launch_bounds__(256) class_sums_kernel(const void* __restrict__ x, int xdt, int N, int H, int W, int C, int ld,
                                                         ClsSegs sg, long long* __restrict__ clsum) {
  pdl_entry();
  __shared__ unsigned long long part[9 * 16];
  for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) part[i] = 0ull;
  __syncthreads();
  const int n0 = blockIdx.x * CLS_IMGS, n1 = min(N, n0 + CLS_IMGS);
  int cur = (n0 >= sg.end[0]) + (n0 >= sg.end[1]) + (n0 >= sg.end[2]);
  float mid[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) mid[c] = 0.f;
  // interior sums of the thread (registers) -> shared memory -> this segment's global Q24 sums; leaves part[] zeroed
  auto flush = [&](int s) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c >= C) break;
      float v = mid[c];
      mid[c] = 0.f;
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&part[4 * 16 + c], (unsigned long long)__float2ll_rn(v * 16777216.f));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 16; i += blockDim.x) {
      const int k = i / 16, c = i % 16;
      if (c < C && part[i] != 0ull)
        atomicAdd(reinterpret_cast<unsigned long long*>(&clsum[((int64_t)s * 9 + k) * C + c]), part[i]);
      part[i] = 0ull;
    }
    __syncthreads();
  };
  for (int n = n0; n < n1; ++n) {
    const int s = (n >= sg.end[0]) + (n >= sg.end[1]) + (n >= sg.end[2]);
    if (s != cur) { flush(cur); cur = s; }      // (block-uniform)
    for (int px = threadIdx.x; px < H * W; px += blockDim.x) {
      const int xx = px % W, yy = px / W;
      const int k = (yy == 0 ? 0 : yy == H - 1 ? 2 : 1) * 3 + (xx == 0 ? 0 : xx == W - 1 ? 2 : 1);
      const int64_t off = ((int64_t)n * H * W + px) * ld;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (c >= C) break;
        float v = xdt == TGAN_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(x)[off + c])
                                   : reinterpret_cast<const float*>(x)[off + c];
        v = __bfloat162float(__float2bfloat16_rn(v));      // the contraction reads the bf16-rounded value
        if (k == 4) mid[c] += v;
        else atomicAdd(&part[k * 16 + c], (unsigned long long)__float2ll_rn(v * 16777216.f));
      }
    }
  }
  flush(cur);
}

__global__ void __launch_bounds__(1024) class_sums_wide_kernel(const bf16* __restrict__ x, int H, int W, int C, ClsSegs sg,
                                                               long long* __restrict__ clsum) {
  pdl_entry();
  extern __shared__ float rows_sm[];      // [H][3][C]
  const int cv = C / 8, n = blockIdx.x;
  const int s = (n >= sg.end[0]) + (n >= sg.end[1]) + (n >= sg.end[2]);
  for (int i = threadIdx.x; i < cv * H; i += blockDim.x) {
    const int cg = i % cv, y = i / cv;
    const bf16* row = x + (((int64_t)n * H + y) * W) * C + cg * 8;
    float fi[8], mid[8], la[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mid[j] = 0.f;
    {
      const uint4 u = *reinterpret_cast<const uint4*>(row);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { fi[2 * j] = __low2float(h[j]); fi[2 * j + 1] = __high2float(h[j]); }
    }
    {
      const uint4 u = *reinterpret_cast<const uint4*>(row + (int64_t)(W - 1) * C);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { la[2 * j] = __low2float(h[j]); la[2 * j + 1] = __high2float(h[j]); }
    }
    for (int xx = 1; xx < W - 1; ++xx) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + (int64_t)xx * C);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(h[j]); v[2 * j + 1] = __high2float(h[j]); }
#pragma unroll
      for (int j = 0; j < 8; ++j) mid[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rows_sm[((size_t)y * 3 + 0) * C + cg * 8 + j] = fi[j];
      rows_sm[((size_t)y * 3 + 1) * C + cg * 8 + j] = mid[j];
      rows_sm[((size_t)y * 3 + 2) * C + cg * 8 + j] = la[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {      // i = cc * C + channel
    const float first = rows_sm[i], last = rows_sm[(size_t)(H - 1) * 3 * C + i];
    float inner = 0.f;
    for (int y = 1; y < H - 1; ++y) inner += rows_sm[(size_t)y * 3 * C + i];
    const int cc = i / C, ch = i - cc * C;
    const float val[3] = {first, inner, last};
#pragma unroll
    for (int rc = 0; rc < 3; ++rc)
      if (val[rc] != 0.f)
        atomicAdd(reinterpret_cast<unsigned long long*>(&clsum[((int64_t)s * 9 + rc * 3 + cc) * C + ch]),
                  (unsigned long long)__float2ll_rn(val[rc] * 16777216.f));
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
  pdl_launch(mobn_mean_kernel, Cout, 128, 0, (cudaStream_t)stream, (const long long*)clsum, nseg, ic,
             (const bf16*)wp, T, Cout, Cin, w_tap_stride, w_co_stride, b, pop_mean, decay, shift);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_class_sums(const void* x, int xdt, int N, int H, int W, int C, int ld, int nseg,
                               const int* seg_end_images, void* clsum, void* stream) {
  TGAN_CHECK_ARG(x && clsum && N > 0 && H > 2 && W > 2 && C >= 1 && ld >= C && nseg >= 1 && nseg <= 4, "class_sums: bad args");
  ClsSegs sg;
  for (int i = 0; i < 4; ++i) sg.end[i] = (nseg > 1 && i < nseg - 1 && seg_end_images) ? seg_end_images[i] : 0x7fffffff;
  sg.n = nseg;
  if (C > 16) {      // wide bf16 activation (the pooled tensor in front of conv2_1)
    const size_t smem = (size_t)H * 3 * C * sizeof(float);
    TGAN_CHECK_ARG(xdt == TGAN_BF16 && C % 8 == 0 && ld == C && ((uintptr_t)x & 15) == 0 && smem <= 96 * 1024,
                   "class_sums: wide tensors must be contiguous bf16 with C %% 8 == 0 and H * C <= 8192");
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(class_sums_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      attr_set = true;
    }
    int threads = (C / 8) * H;
    threads = threads > 1024 ? 1024 : (threads + 31) / 32 * 32;
    pdl_launch(class_sums_wide_kernel, N, threads, smem, (cudaStream_t)stream, (const bf16*)x, H, W, C, sg, (long long*)clsum);
    TGAN_LAUNCHED();
    return 0;
  }
  pdl_launch(class_sums_kernel, ceil_div(N, CLS_IMGS), 256, 0, (cudaStream_t)stream, x, xdt, N, H, W, C, ld, sg, (long long*)clsum);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_seg_sums_finalize(const void* q24, int nseg, int C, float* colsums, float* grad_acc, void* stream) {
  TGAN_CHECK_ARG(q24 && colsums && nseg >= 1 && nseg <= 4 && C > 0, "seg_sums_finalize: bad args");
  pdl_launch(seg_sums_finalize_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)stream, (const long long*)q24, nseg, C, colsums, grad_acc);
  TGAN_LAUNCHED();
  return 0;
}
