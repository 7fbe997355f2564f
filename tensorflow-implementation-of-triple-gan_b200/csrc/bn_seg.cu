// bn_seg.cu -- training-mode tf.contrib.layers.batch_norm (modle_base.py:229-237) of a GROUPED batch: the calls of one
// network inside a phase run as one batch (ops.group_batch) and every call keeps its own batch statistics (its rows are
// a contiguous segment).  The per-call version costs 3 launches forward and 4 backward PER SEGMENT AND LAYER (21 for the
// three-call groups of the Good_GAN classifier, Good_GAN.py:220-299, which normalises after every convolution); here a
// layer is 3 launches forward and 3 backward for all segments:
//
//   forward : partial sums (x, x^2) per (segment, row part)  ->  fold + finalize (mean, rstd, scale, shift per segment;
//             moving statistics updated once per segment in call order)  ->  y = x * scale[seg] + shift[seg]
//   backward: partial sums (dy, dy * xhat)  ->  fold (s1, s2 per segment; dbeta += sum s1, dgamma += sum s2)  ->
//             dx = gamma * rstd * (dy - s1/n - xhat * s2/n)
//
// A thread sums its few dozen rows in fp32; everything across threads and CTAs is fp64 and folded in a fixed order
// (bit-reproducible; the backward sums cancel almost exactly, see colreduce.cuh).  HBM-bound: forward reads x twice and
// writes y once, backward reads dy and x twice and writes dx.  Measured on the Good_GAN classifier's largest layer
// (52 MB bf16): statistics 4.0 TB/s forward, apply 5.5 TB/s, backward apply 3.1 TB/s.
#include <type_traits>
#include "common.cuh"

namespace tgan {

constexpr int BNS_MAX_PARTS = 64;
struct SegRows {
  int n;
  int64_t end[4];      // exclusive end row of every segment (end[n - 1] = rows)
  __device__ __forceinline__ int of(int64_t r) const { return (r >= end[0]) + (n > 2 && r >= end[1]) + (n > 3 && r >= end[2]); }
  __device__ __forceinline__ int64_t begin(int s) const { return s ? end[s - 1] : 0; }
};

// VEC consecutive channels per thread: 8 (16-byte bf16 / 2 x 16-byte fp32 accesses) or 1
template <typename T, int VEC> __device__ __forceinline__ void ldv(const T* p, int64_t i, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = ldf<T>(p, i);
  } else if constexpr (std::is_same<T, float>::value) {
    const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(p + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 t = *reinterpret_cast<const uint4*>(p + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(h[j]); v[2 * j + 1] = __high2float(h[j]); }
  }
}
template <typename T, int VEC> __device__ __forceinline__ void stv(T* p, int64_t i, const float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    stf<T>(p, i, v[0]);
  } else if constexpr (std::is_same<T, float>::value) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p + i) = t;
  }
}

template <typename T, int VEC>
struct BnFwdSegF {
  const T* x; int C;
  __device__ __forceinline__ void prepare(int, int) {}
  __device__ __forceinline__ void operator()(int64_t r, int c, int, float (&a)[VEC], float (&b)[VEC]) const {
    ldv<T, VEC>(x, r * C + c, a);
#pragma unroll
    for (int j = 0; j < VEC; ++j) b[j] = a[j] * a[j];
  }
};
template <typename TDY, typename TX, int VEC>
struct BnBwdSegF {
  const TDY* dy; const TX* x; const float* mean; const float* rstd; int C;
  float m[VEC], rs[VEC];      // this thread's (segment, channels): loaded once by prepare()
  __device__ __forceinline__ void prepare(int c, int s) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) { m[j] = mean[s * C + c + j]; rs[j] = rstd[s * C + c + j]; }
  }
  __device__ __forceinline__ void operator()(int64_t r, int c, int, float (&a)[VEC], float (&b)[VEC]) const {
    float xv[VEC];
    ldv<TDY, VEC>(dy, r * C + c, a);
    ldv<TX, VEC>(x, r * C + c, xv);
#pragma unroll
    for (int j = 0; j < VEC; ++j) b[j] = a[j] * ((xv[j] - m[j]) * rs[j]);
  }
};

// grid (ceil(C / CH), parts, nseg); 256 threads = CL channel lanes (VEC channels each) x RL row lanes, CH = CL * VEC
// channels per CTA; partials [seg][part][2][C] (fp64).  Four rows per thread are in flight (64 B per thread with VEC = 8).
template <typename F, int VEC>
__global__ void __launch_bounds__(256) segreduce_kernel(const F f_in, SegRows sg, int C, double* __restrict__ partials) {
  pdl_entry();
  F f = f_in;
  constexpr int CL = VEC == 8 ? 8 : 32, RL = 256 / CL, CH = CL * VEC;
  __shared__ double sm[RL][2][CH + 1];
  const int s = blockIdx.z, tx = threadIdx.x % CL, ty = threadIdx.x / CL, c = blockIdx.x * CH + tx * VEC;
  const int64_t r1 = sg.end[s], step = (int64_t)gridDim.y * RL;
  // per-thread sums (a few dozen rows) in fp32, everything across threads and CTAs in fp64: sixteen fp64 accumulators per
  // thread cost 32 registers and the compiler then serialised the four row loads (84 registers, 1.8 TB/s in the backward)
  float a0[VEC], a1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { a0[j] = 0.f; a1[j] = 0.f; }
  if (c < C) {
    f.prepare(c, s);
    int64_t r = sg.begin(s) + (int64_t)blockIdx.y * RL + ty;
    for (; r + 3 * step < r1; r += 4 * step) {
      float u[4][VEC], v[4][VEC];
#pragma unroll
      for (int k = 0; k < 4; ++k) f(r + k * step, c, s, u[k], v[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { a0[j] += u[k][j]; a1[j] += v[k][j]; }
    }
    for (; r < r1; r += step) {
      float u[VEC], v[VEC];
      f(r, c, s, u, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) { a0[j] += u[j]; a1[j] += v[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) { sm[ty][0][tx * VEC + j] = (double)a0[j]; sm[ty][1][tx * VEC + j] = (double)a1[j]; }
  __syncthreads();
  if (threadIdx.x < 2 * CH) {
    const int a = threadIdx.x / CH, cc = threadIdx.x % CH, co = blockIdx.x * CH + cc;
    if (co < C) {
      double t = 0.0;
#pragma unroll 8
      for (int y = 0; y < RL; ++y) t += sm[y][a][cc];
      partials[(((int64_t)s * gridDim.y + blockIdx.y) * 2 + a) * C + co] = t;
    }
  }
}

// Fold of the partials: 256 threads = 32 channels x 8 part lanes; lane l sums parts l, l + 8, ... of every (segment,
// statistic), the eight lane sums are added in a fixed order.  out[seg * 2 + stat][channel] in shared memory.
__device__ __forceinline__ void fold_partials(const double* __restrict__ partials, int parts, int nseg, int C, int c,
                                              double (*lanes)[8][33], double (*out)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // constant trip counts (parts <= 64 = 8 lanes x 8 rounds, 8 (segment, statistic) pairs): fully unrolled, so the
  // up to 64 loads of a thread are independent and in flight together (a runtime loop ran at one L2 latency per round)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = ty + 8 * i;
      v[i] = (c < C && q < nseg * 2 && p < parts) ? partials[(((int64_t)(q >> 1) * parts + p) * 2 + (q & 1)) * C + c] : 0.0;
    }
    lanes[ty][q][tx] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  }
  __syncthreads();
  {
    const int q = threadIdx.x >> 5;      // 8 warps = up to 4 segments x 2 statistics
    if (q < nseg * 2) {
      double t = 0.0;
#pragma unroll
      for (int l = 0; l < 8; ++l) t += lanes[l][q][tx];
      out[q][tx] = t;
    }
  }
  __syncthreads();
}

// finalize (mean, rstd, scale, shift per segment) and update the moving statistics once per segment in call order
// (tf.contrib batch_norm with updates_collections=None runs its update inside every call).  grid ceil(C / 32) x 256 threads
__global__ void __launch_bounds__(256) bn_finalize_seg_kernel(const double* __restrict__ partials, int parts, SegRows sg, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float decay,
                                       int unbiased, float* mm, float* mv, float* __restrict__ mean, float* __restrict__ rstd,
                                       float* __restrict__ scale, float* __restrict__ shift) {
  pdl_entry();
  __shared__ double lanes[8][8][33];
  __shared__ double tot[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  fold_partials(partials, parts, sg.n, C, c, lanes, tot);
  if (threadIdx.x >= 32 || c >= C) return;
  float m_run = mm ? mm[c] : 0.f, v_run = mv ? mv[c] : 0.f;
  for (int s = 0; s < sg.n; ++s) {
    const float rows = (float)(sg.end[s] - sg.begin(s));
    // (the same fp32 arithmetic as bn_finalize_kernel on the fp32-rounded totals: the grouped and the per-call paths agree)
    const float mu = (float)tot[2 * s][threadIdx.x] / rows;
    const float var = fmaxf((float)tot[2 * s + 1][threadIdx.x] / rows - mu * mu, 0.f);
    const float rs = rsqrtf(var + eps);
    mean[s * C + c] = mu; rstd[s * C + c] = rs;
    const float sc = gamma[c] * rs;
    scale[s * C + c] = sc; shift[s * C + c] = beta[c] - mu * sc;
    m_run = m_run * decay + mu * (1.f - decay);
    v_run = v_run * decay + var * (unbiased ? rows / fmaxf(rows - 1.f, 1.f) : 1.f) * (1.f - decay);
  }
  if (mm) mm[c] = m_run;
  if (mv) mv[c] = v_run;
}

// (32-bit index arithmetic: the host checks rows * C < 2^31; a 64-bit division per vector costs more than the math)
struct SegRows32 {
  int n; unsigned end[4]; float inv_rows[4];
  __device__ __forceinline__ int of(unsigned r) const { return (r >= end[0]) + (n > 2 && r >= end[1]) + (n > 3 && r >= end[2]); }
};
// per-(segment, channel) coefficient vector: 16-byte loads when VEC == 8 (tables are 16-byte aligned, C % 8 == 0)
template <int VEC> __device__ __forceinline__ void ldc(const float* __restrict__ t, int i, float (&v)[VEC]) {
  if constexpr (VEC == 8) {
    const float4 a = *reinterpret_cast<const float4*>(t + i), b = *reinterpret_cast<const float4*>(t + i + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    v[0] = t[i];
  }
}
template <typename TX, typename TY, int VEC>
__global__ void bn_apply_seg_kernel(const TX* __restrict__ x, TY* __restrict__ y, unsigned nvec, unsigned cv, int C, SegRows32 sg,
                                    const float* __restrict__ scale, const float* __restrict__ shift) {
  pdl_entry();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const unsigned r = i / cv, c = (i - r * cv) * VEC;
    const int k = sg.of(r) * C + (int)c;
    float v[VEC], sc[VEC], sh[VEC];
    ldv<TX, VEC>(x, (int64_t)i * VEC, v);
    ldc<VEC>(scale, k, sc); ldc<VEC>(shift, k, sh);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = v[j] * sc[j] + sh[j];
    stv<TY, VEC>(y, (int64_t)i * VEC, v);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_fold_seg_kernel(const double* __restrict__ partials, int parts, SegRows sg, int C,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, float* __restrict__ ca,
                                                              float* __restrict__ cb, float* __restrict__ cc, float* dgamma,
                                                              float* dbeta, float beta_acc) {
  const int nseg = sg.n;
  pdl_entry();
  __shared__ double lanes[8][8][33];
  __shared__ double tot[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  fold_partials(partials, parts, nseg, C, c, lanes, tot);
  if (threadIdx.x >= 32 || c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int s = 0; s < nseg; ++s) {
    const float a = (float)tot[2 * s][threadIdx.x], b = (float)tot[2 * s + 1][threadIdx.x];
    t1 += (double)a; t2 += (double)b;      // (per-call path: every call adds its fp32-rounded sums)
    // dx = g rs (dy - s1/n - xhat s2/n), xhat = (x - mu) rs   ==   ca * dy + cb * x + cc
    const float inv_n = 1.0f / (float)(sg.end[s] - sg.begin(s));
    const float g = gamma[c], rs = rstd[s * C + c], mu = mean[s * C + c];
    const float ca_ = g * rs, cb_ = -g * rs * rs * b * inv_n;
    ca[s * C + c] = ca_; cb[s * C + c] = cb_; cc[s * C + c] = -ca_ * a * inv_n - cb_ * mu;
  }
  if (dbeta) dbeta[c] = (float)((beta_acc != 0.f ? (double)beta_acc * dbeta[c] : 0.0) + t1);
  if (dgamma) dgamma[c] = (float)((beta_acc != 0.f ? (double)beta_acc * dgamma[c] : 0.0) + t2);
}

template <typename TDY, typename TX, typename TDX, int VEC>
__global__ void bn_bwd_apply_seg_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, TDX* __restrict__ dx, unsigned nvec,
                                        unsigned cv, int C, SegRows32 sg, const float* __restrict__ ca,
                                        const float* __restrict__ cb, const float* __restrict__ cc) {
  pdl_entry();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
    const unsigned r = i / cv, c = (i - r * cv) * VEC;
    const int k = sg.of(r) * C + (int)c;
    float d[VEC], xv[VEC], a[VEC], b[VEC], e[VEC];
    ldv<TDY, VEC>(dy, (int64_t)i * VEC, d);
    ldv<TX, VEC>(x, (int64_t)i * VEC, xv);
    ldc<VEC>(ca, k, a); ldc<VEC>(cb, k, b); ldc<VEC>(cc, k, e);
#pragma unroll
    for (int j = 0; j < VEC; ++j) d[j] = a[j] * d[j] + (b[j] * xv[j] + e[j]);
    stv<TDX, VEC>(dx, (int64_t)i * VEC, d);
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int make_seg_rows(SegRows& sg, int64_t rows, int nseg, int64_t r0, int64_t r1, int64_t r2) {
  if (nseg < 1 || nseg > 4) { set_error("bn_seg: nseg %d out of range [1,4]", nseg); return 1; }
  const int64_t e[4] = {r0, r1, r2, rows};
  int64_t prev = 0;
  sg.n = nseg;
  for (int i = 0; i < 4; ++i) {
    const int64_t end = i < nseg - 1 ? e[i] : rows;
    if (i < nseg && end <= prev) { set_error("bn_seg: segment boundaries must be increasing and non-empty"); return 1; }
    sg.end[i] = end;
    if (i < nseg) prev = end;
  }
  return 0;
}
static int make_seg_rows32(SegRows32& s32, const SegRows& sg, int64_t rows, int C) {
  if (rows * C >= ((int64_t)1 << 31)) { set_error("bn_seg: rows * C must be < 2^31"); return 1; }
  s32.n = sg.n;
  for (int i = 0; i < 4; ++i) {
    s32.end[i] = (unsigned)sg.end[i];
    s32.inv_rows[i] = i < sg.n ? 1.0f / (float)(sg.end[i] - (i ? sg.end[i - 1] : 0)) : 0.f;
  }
  return 0;
}
static int pick_seg_parts(int64_t rows, int nseg, int C, int vec) {
  const int ch = vec == 8 ? 64 : 32, rl = vec == 8 ? 32 : 8;
  const int xb = ceil_div(C, ch);
  int64_t want = (148 * 6 + xb * nseg - 1) / (xb * nseg);      // ~6 CTAs per SM over the whole grid
  const int64_t maxp = (rows / nseg + rl * 2 - 1) / (rl * 2);    // >= 2 rows per thread
  int64_t p = want < maxp ? want : maxp;
  if (p < 1) p = 1;
  if (p > BNS_MAX_PARTS) p = BNS_MAX_PARTS;
  return (int)p;
}
static inline int grid1d(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

}  // namespace tgan

using namespace tgan;

// ws layout (floats): [0, 1024 C) fp64 partials [nseg][parts <= 64][2][C]; [1024 C, 1036 C) up to three [4][C] fp32 tables
extern "C" int tgan_bn_fwd_seg(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                               int64_t r2, const float* gamma, const float* beta, float eps, float decay,
                               int unbiased_moving_var, float* moving_mean, float* moving_var, float* mean, float* rstd,
                               float* ws, void* stream) {
  TGAN_CHECK_ARG(x && y && gamma && beta && mean && rstd && ws && rows > 0 && C > 0 && aligned16(ws), "bn_fwd_seg: bad args");
  SegRows sg;
  SegRows32 s32;
  if (make_seg_rows(sg, rows, nseg, r0, r1, r2) || make_seg_rows32(s32, sg, rows, C)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v8 = C % 8 == 0 && aligned16(x) && aligned16(y);
  const int parts = pick_seg_parts(rows, nseg, C, v8 ? 8 : 1);
  double* partials = reinterpret_cast<double*>(ws);
  float* scale = ws + (int64_t)1024 * C;
  float* shift = scale + (int64_t)4 * C;
  TGAN_DISPATCH_1(xdt, TX, {
    if (v8) {
      BnFwdSegF<TX, 8> f{(const TX*)x, C};
      pdl_launch(segreduce_kernel<BnFwdSegF<TX, 8>, 8>, dim3(ceil_div(C, 64), parts, nseg), 256, 0, st, f, sg, C, partials);
    } else {
      BnFwdSegF<TX, 1> f{(const TX*)x, C};
      pdl_launch(segreduce_kernel<BnFwdSegF<TX, 1>, 1>, dim3(ceil_div(C, 32), parts, nseg), 256, 0, st, f, sg, C, partials);
    }
  });
  TGAN_LAUNCHED();
  pdl_launch(bn_finalize_seg_kernel, ceil_div(C, 32), 256, 0, st, (const double*)partials, parts, sg, C, gamma, beta, eps, decay,
             unbiased_moving_var, moving_mean, moving_var, mean, rstd, scale, shift);
  TGAN_LAUNCHED();
  const int64_t n = rows * C;
  TGAN_DISPATCH_1(xdt, TX, TGAN_DISPATCH_1(ydt, TY, {
    if (v8) pdl_launch(bn_apply_seg_kernel<TX, TY, 8>, grid1d(n / 8), 256, 0, st, (const TX*)x, (TY*)y, (unsigned)(n / 8),
                       (unsigned)(C / 8), C, s32, (const float*)scale, (const float*)shift);
    else pdl_launch(bn_apply_seg_kernel<TX, TY, 1>, grid1d(n), 256, 0, st, (const TX*)x, (TY*)y, (unsigned)n, (unsigned)C, C, s32,
                    (const float*)scale, (const float*)shift);
  }));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_bn_bwd_seg(const void* dy, int dydt, const void* x, int xdt, void* dx, int dxdt, int64_t rows, int C,
                               int nseg, int64_t r0, int64_t r1, int64_t r2, const float* mean, const float* rstd,
                               const float* gamma, float* dgamma, float* dbeta, float beta_acc, float* ws, void* stream) {
  TGAN_CHECK_ARG(dy && x && dx && mean && rstd && gamma && ws && rows > 0 && C > 0 && aligned16(ws), "bn_bwd_seg: bad args");
  SegRows sg;
  SegRows32 s32;
  if (make_seg_rows(sg, rows, nseg, r0, r1, r2) || make_seg_rows32(s32, sg, rows, C)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool v8 = C % 8 == 0 && aligned16(dy) && aligned16(x) && aligned16(dx);
  const int parts = pick_seg_parts(rows, nseg, C, v8 ? 8 : 1);
  double* partials = reinterpret_cast<double*>(ws);
  float* ca = ws + (int64_t)1024 * C;       // three [4][C] coefficient tables
  float* cb = ca + (int64_t)4 * C;
  float* cc = cb + (int64_t)4 * C;
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(xdt, TX, {
    if (v8) {
      BnBwdSegF<TDY, TX, 8> f{(const TDY*)dy, (const TX*)x, mean, rstd, C};
      pdl_launch(segreduce_kernel<BnBwdSegF<TDY, TX, 8>, 8>, dim3(ceil_div(C, 64), parts, nseg), 256, 0, st, f, sg, C, partials);
    } else {
      BnBwdSegF<TDY, TX, 1> f{(const TDY*)dy, (const TX*)x, mean, rstd, C};
      pdl_launch(segreduce_kernel<BnBwdSegF<TDY, TX, 1>, 1>, dim3(ceil_div(C, 32), parts, nseg), 256, 0, st, f, sg, C, partials);
    }
  }));
  TGAN_LAUNCHED();
  pdl_launch(bn_bwd_fold_seg_kernel, ceil_div(C, 32), 256, 0, st, (const double*)partials, parts, sg, C, mean, rstd, gamma, ca, cb, cc,
             dgamma, dbeta, beta_acc);
  TGAN_LAUNCHED();
  const int64_t n = rows * C;
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(xdt, TX, TGAN_DISPATCH_1(dxdt, TDX, {
    if (v8) pdl_launch(bn_bwd_apply_seg_kernel<TDY, TX, TDX, 8>, grid1d(n / 8), 256, 0, st, (const TDY*)dy, (const TX*)x, (TDX*)dx,
                       (unsigned)(n / 8), (unsigned)(C / 8), C, s32, (const float*)ca, (const float*)cb, (const float*)cc);
    else pdl_launch(bn_bwd_apply_seg_kernel<TDY, TX, TDX, 1>, grid1d(n), 256, 0, st, (const TDY*)dy, (const TX*)x, (TDX*)dx,
                    (unsigned)n, (unsigned)C, C, s32, (const float*)ca, (const float*)cb, (const float*)cc);
  })));
  TGAN_LAUNCHED();
  return 0;
}
