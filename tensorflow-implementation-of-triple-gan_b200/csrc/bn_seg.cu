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
// Partials are fp64 and folded in a fixed order (bit-reproducible; the backward sums cancel almost exactly, see
// colreduce.cuh).  HBM-bound: forward reads x twice and writes y once, backward reads dy and x twice and writes dx.
#include "common.cuh"

namespace tgan {

constexpr int BNS_MAX_PARTS = 32;
struct SegRows {
  int n;
  int64_t end[4];      // exclusive end row of every segment (end[n - 1] = rows)
  __device__ __forceinline__ int of(int64_t r) const { return (r >= end[0]) + (n > 2 && r >= end[1]) + (n > 3 && r >= end[2]); }
  __device__ __forceinline__ int64_t begin(int s) const { return s ? end[s - 1] : 0; }
};

template <typename T>
struct BnFwdSegF {
  const T* x; int C;
  __device__ __forceinline__ void operator()(int64_t r, int c, int, float& a, float& b) const {
    const float v = ldf<T>(x, r * C + c);
    a = v; b = v * v;
  }
};
template <typename TDY, typename TX>
struct BnBwdSegF {
  const TDY* dy; const TX* x; const float* mean; const float* rstd; int C;
  __device__ __forceinline__ void operator()(int64_t r, int c, int s, float& a, float& b) const {
    const float d = ldf<TDY>(dy, r * C + c);
    const float xh = (ldf<TX>(x, r * C + c) - mean[s * C + c]) * rstd[s * C + c];
    a = d; b = d * xh;
  }
};

// grid (ceil(C / 32), parts, nseg); 256 threads = 32 channels x 8 row lanes; partials [seg][part][2][C] (fp64)
template <typename F>
__global__ void __launch_bounds__(256) segreduce_kernel(F f, SegRows sg, int C, double* __restrict__ partials) {
  pdl_entry();
  __shared__ double sm[8][2][33];
  const int s = blockIdx.z, tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
  const int64_t r1 = sg.end[s], step = (int64_t)gridDim.y * 8;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    int64_t r = sg.begin(s) + (int64_t)blockIdx.y * 8 + ty;
    for (; r + 3 * step < r1; r += 4 * step) {      // four rows in flight
      float u[4], v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) f(r + k * step, c, s, u[k], v[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) { a0 += (double)u[k]; a1 += (double)v[k]; }
    }
    for (; r < r1; r += step) {
      float u, v;
      f(r, c, s, u, v);
      a0 += (double)u; a1 += (double)v;
    }
  }
  sm[ty][0][tx] = a0; sm[ty][1][tx] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int a = threadIdx.x >> 5, cc = threadIdx.x & 31, co = blockIdx.x * 32 + cc;
    if (co < C) {
      double t = 0.0;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += sm[y][a][cc];
      partials[(((int64_t)s * gridDim.y + blockIdx.y) * 2 + a) * C + co] = t;
    }
  }
}

// one thread per channel: fold the partials of every segment in order, finalize, update the moving statistics once per
// segment in call order (tf.contrib batch_norm with updates_collections=None runs its update inside every call)
__global__ void bn_finalize_seg_kernel(const double* __restrict__ partials, int parts, SegRows sg, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float decay,
                                       int unbiased, float* mm, float* mv, float* __restrict__ mean, float* __restrict__ rstd,
                                       float* __restrict__ scale, float* __restrict__ shift) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float m_run = mm ? mm[c] : 0.f, v_run = mv ? mv[c] : 0.f;
  for (int s = 0; s < sg.n; ++s) {
    double s0 = 0.0, s1 = 0.0;
    for (int p = 0; p < parts; ++p) {
      s0 += partials[(((int64_t)s * parts + p) * 2 + 0) * C + c];
      s1 += partials[(((int64_t)s * parts + p) * 2 + 1) * C + c];
    }
    const float rows = (float)(sg.end[s] - sg.begin(s));
    // (the same fp32 arithmetic as bn_finalize_kernel on the fp32-rounded totals: the grouped and the per-call paths agree)
    const float mu = (float)s0 / rows;
    const float var = fmaxf((float)s1 / rows - mu * mu, 0.f);
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

template <typename TX, typename TY, int VEC>
__global__ void bn_apply_seg_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t nvec, int C, SegRows sg,
                                    const float* __restrict__ scale, const float* __restrict__ shift) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * VEC, r = e / C;
    const int c = (int)(e - r * C), s = sg.of(r);
    float v[VEC];
    if constexpr (VEC == 4) ld4<TX>(x, e, *reinterpret_cast<float(*)[4]>(v));
    else v[0] = ldf<TX>(x, e);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] = v[j] * scale[s * C + c + j] + shift[s * C + c + j];
    if constexpr (VEC == 4) st4<TY>(y, e, *reinterpret_cast<float(*)[4]>(v));
    else stf<TY>(y, e, v[0]);
  }
}

__global__ void bn_bwd_fold_seg_kernel(const double* __restrict__ partials, int parts, int nseg, int C, float* __restrict__ s1,
                                       float* __restrict__ s2, float* dgamma, float* dbeta, float beta_acc) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int s = 0; s < nseg; ++s) {
    double a = 0.0, b = 0.0;
    for (int p = 0; p < parts; ++p) {
      a += partials[(((int64_t)s * parts + p) * 2 + 0) * C + c];
      b += partials[(((int64_t)s * parts + p) * 2 + 1) * C + c];
    }
    s1[s * C + c] = (float)a; s2[s * C + c] = (float)b;
    t1 += (double)(float)a; t2 += (double)(float)b;      // (per-call path: every call adds its fp32-rounded sums)
  }
  if (dbeta) dbeta[c] = (float)((beta_acc != 0.f ? (double)beta_acc * dbeta[c] : 0.0) + t1);
  if (dgamma) dgamma[c] = (float)((beta_acc != 0.f ? (double)beta_acc * dgamma[c] : 0.0) + t2);
}

template <typename TDY, typename TX, typename TDX, int VEC>
__global__ void bn_bwd_apply_seg_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, TDX* __restrict__ dx, int64_t nvec,
                                        int C, SegRows sg, const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ s1,
                                        const float* __restrict__ s2) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * VEC, r = e / C;
    const int c = (int)(e - r * C), s = sg.of(r);
    const float inv_rows = 1.0f / (float)(sg.end[s] - sg.begin(s));
    float d[VEC], xv[VEC], o[VEC];
    if constexpr (VEC == 4) {
      ld4<TDY>(dy, e, *reinterpret_cast<float(*)[4]>(d)); ld4<TX>(x, e, *reinterpret_cast<float(*)[4]>(xv));
    } else {
      d[0] = ldf<TDY>(dy, e); xv[0] = ldf<TX>(x, e);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int k = s * C + c + j;
      const float xh = (xv[j] - mean[k]) * rstd[k];
      o[j] = gamma[c + j] * rstd[k] * (d[j] - s1[k] * inv_rows - xh * s2[k] * inv_rows);
    }
    if constexpr (VEC == 4) st4<TDX>(dx, e, *reinterpret_cast<float(*)[4]>(o));
    else stf<TDX>(dx, e, o[0]);
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
static int pick_seg_parts(int64_t rows, int nseg, int C) {
  const int xb = ceil_div(C, 32);
  int64_t want = (148 * 4 + xb * nseg - 1) / (xb * nseg);      // ~4 CTAs per SM over the whole grid
  const int64_t maxp = (rows / nseg + 15) / 16;                  // >= 2 rows per thread
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

extern "C" int tgan_bn_fwd_seg(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                               int64_t r2, const float* gamma, const float* beta, float eps, float decay,
                               int unbiased_moving_var, float* moving_mean, float* moving_var, float* mean, float* rstd,
                               float* ws, void* stream) {
  TGAN_CHECK_ARG(x && y && gamma && beta && mean && rstd && ws && rows > 0 && C > 0 && aligned16(ws), "bn_fwd_seg: bad args");
  SegRows sg;
  if (make_seg_rows(sg, rows, nseg, r0, r1, r2)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = pick_seg_parts(rows, nseg, C);
  double* partials = reinterpret_cast<double*>(ws);                    // [nseg][parts][2][C] fp64: <= 512 * C floats
  float* scale = ws + (int64_t)4 * BNS_MAX_PARTS * 4 * C;              // [nseg][C], then shift [nseg][C]
  float* shift = scale + (int64_t)4 * C;
  TGAN_DISPATCH_1(xdt, TX, {
    BnFwdSegF<TX> f{(const TX*)x, C};
    pdl_launch(segreduce_kernel<BnFwdSegF<TX>>, dim3(ceil_div(C, 32), parts, nseg), 256, 0, st, f, sg, C, partials);
  });
  TGAN_LAUNCHED();
  pdl_launch(bn_finalize_seg_kernel, ceil_div(C, 128), 128, 0, st, (const double*)partials, parts, sg, C, gamma, beta, eps, decay,
             unbiased_moving_var, moving_mean, moving_var, mean, rstd, scale, shift);
  TGAN_LAUNCHED();
  const int64_t n = rows * C;
  const bool v4 = C % 4 == 0 && aligned16(x) && aligned16(y);
  TGAN_DISPATCH_1(xdt, TX, TGAN_DISPATCH_1(ydt, TY, {
    if (v4) pdl_launch(bn_apply_seg_kernel<TX, TY, 4>, grid1d(n / 4), 256, 0, st, (const TX*)x, (TY*)y, n / 4, C, sg,
                       (const float*)scale, (const float*)shift);
    else pdl_launch(bn_apply_seg_kernel<TX, TY, 1>, grid1d(n), 256, 0, st, (const TX*)x, (TY*)y, n, C, sg, (const float*)scale,
                    (const float*)shift);
  }));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_bn_bwd_seg(const void* dy, int dydt, const void* x, int xdt, void* dx, int dxdt, int64_t rows, int C,
                               int nseg, int64_t r0, int64_t r1, int64_t r2, const float* mean, const float* rstd,
                               const float* gamma, float* dgamma, float* dbeta, float beta_acc, float* ws, void* stream) {
  TGAN_CHECK_ARG(dy && x && dx && mean && rstd && gamma && ws && rows > 0 && C > 0 && aligned16(ws), "bn_bwd_seg: bad args");
  SegRows sg;
  if (make_seg_rows(sg, rows, nseg, r0, r1, r2)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int parts = pick_seg_parts(rows, nseg, C);
  double* partials = reinterpret_cast<double*>(ws);
  float* s1 = ws + (int64_t)4 * BNS_MAX_PARTS * 4 * C;
  float* s2 = s1 + (int64_t)4 * C;
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(xdt, TX, {
    BnBwdSegF<TDY, TX> f{(const TDY*)dy, (const TX*)x, mean, rstd, C};
    pdl_launch(segreduce_kernel<BnBwdSegF<TDY, TX>>, dim3(ceil_div(C, 32), parts, nseg), 256, 0, st, f, sg, C, partials);
  }));
  TGAN_LAUNCHED();
  pdl_launch(bn_bwd_fold_seg_kernel, ceil_div(C, 128), 128, 0, st, (const double*)partials, parts, nseg, C, s1, s2, dgamma, dbeta,
             beta_acc);
  TGAN_LAUNCHED();
  const int64_t n = rows * C;
  const bool v4 = C % 4 == 0 && aligned16(dy) && aligned16(x) && aligned16(dx);
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(xdt, TX, TGAN_DISPATCH_1(dxdt, TDX, {
    if (v4) pdl_launch(bn_bwd_apply_seg_kernel<TDY, TX, TDX, 4>, grid1d(n / 4), 256, 0, st, (const TDY*)dy, (const TX*)x, (TDX*)dx,
                       n / 4, C, sg, mean, rstd, gamma, (const float*)s1, (const float*)s2);
    else pdl_launch(bn_bwd_apply_seg_kernel<TDY, TX, TDX, 1>, grid1d(n), 256, 0, st, (const TDY*)dy, (const TX*)x, (TDX*)dx, n, C, sg,
                    mean, rstd, gamma, (const float*)s1, (const float*)s2);
  })));
  TGAN_LAUNCHED();
  return 0;
}
