// pointwise.cu -- HBM-bound elementwise / per-channel-reduction kernels of the Triple-GAN step:
// mean-only BN and BN statistics + apply (nn.py:147-187; modle_base.py:229-237), fused
// bias/activation epilogues, Gaussian noise, dropout, pooling, label concat, argmax/one-hot.
// All kernels are vectorised (4 elements / thread) when the channel count allows, coalesced along the
// NHWC channel axis, and reduce with shared memory + a deterministic second stage (no atomics).
#include "common.cuh"
#include "colreduce.cuh"

namespace tgan {

template <typename T, int VEC>
struct StatsF {
  const T* x; int C;
  struct Regs { float t[VEC]; };
  __device__ __forceinline__ void load(int64_t r, int c0, Regs& q) const {
    if constexpr (VEC == 4) ld4<T>(x, r * C + c0, q.t);
    else q.t[0] = ldf<T>(x, r * C + c0);
  }
  __device__ __forceinline__ void finish(int64_t, int, const Regs& q, float (&v)[2][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) { v[0][j] = q.t[j]; v[1][j] = q.t[j] * q.t[j]; }
  }
  __device__ void operator()(int64_t r, int c0, float (&v)[2][VEC]) const {
    Regs q;
    load(r, c0, q);
    finish(r, c0, q, v);
  }
};

template <typename TDY, typename TY, typename TDU, int VEC, int ACT>
struct ActBwdF {
  const TDY* dy; const TY* y; TDU* du; int C; float alpha; int ldy;      // y may live in a channel-padded buffer
  struct Regs { float a[VEC], b[VEC]; };     // load / finish split (colreduce_kernel keeps four rows in flight)
  __device__ __forceinline__ void load(int64_t r, int c0, Regs& q) const {
    if constexpr (VEC == 4) {
      ld4<TDY>(dy, r * C + c0, q.a); ld4<TY>(y, r * ldy + c0, q.b);
    } else {
      q.a[0] = ldf<TDY>(dy, r * C + c0); q.b[0] = ldf<TY>(y, r * ldy + c0);
    }
  }
  __device__ __forceinline__ void finish(int64_t r, int c0, const Regs& q, float (&v)[1][VEC]) const {
    if constexpr (VEC == 4) {
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { o[j] = q.a[j] * act_grad_from_y_t<ACT>(q.b[j], alpha); v[0][j] = o[j]; }
      st4<TDU>(du, r * C + c0, o);
    } else {
      float o = q.a[0] * act_grad_from_y_t<ACT>(q.b[0], alpha);
      v[0][0] = o; stf<TDU>(du, r * C + c0, o);
    }
  }
  __device__ void operator()(int64_t r, int c0, float (&v)[1][VEC]) const {
    Regs q;
    load(r, c0, q);
    finish(r, c0, q, v);
  }
};

// batch segments: several calls of one network grouped along the batch axis keep their own batch statistics
struct Segs {
  int n;
  int64_t end[3];      // exclusive row index where segments 0..2 end (unused entries = INT64_MAX)
  float inv_rows[4];
  __device__ __forceinline__ int of(int64_t r) const { return (r >= end[0]) + (r >= end[1]) + (r >= end[2]); }
};

// du = dy * act'(y) with per-SEGMENT column sums (accumulator a = segment a)
template <typename TDY, typename TY, typename TDU, int VEC, int ACT>
struct ActBwdSegF {
  const TDY* dy; const TY* y; TDU* du; int C; float alpha; Segs sg;
  __device__ void operator()(int64_t r, int c0, float (&v)[4][VEC]) const {
    const int s = sg.of(r);
    float o[VEC];
    if constexpr (VEC == 4) {
      float a[4], b[4];
      ld4<TDY>(dy, r * C + c0, a); ld4<TY>(y, r * C + c0, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = a[j] * act_grad_from_y_t<ACT>(b[j], alpha);
      st4<TDU>(du, r * C + c0, *reinterpret_cast<float(*)[4]>(o));
    } else {
      o[0] = ldf<TDY>(dy, r * C + c0) * act_grad_from_y_t<ACT>(ldf<TY>(y, r * C + c0), alpha);
      stf<TDU>(du, r * C + c0, o[0]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[a][j] = (s == a) ? o[j] : 0.f;
  }
};

template <typename TDY, typename TX, int VEC>
struct BnBwdStatsF {
  const TDY* dy; const TX* x; const float* mean; const float* rstd; int C;
  struct Regs { float d[VEC], x[VEC]; };
  __device__ __forceinline__ void load(int64_t r, int c0, Regs& q) const {
    if constexpr (VEC == 4) {
      ld4<TDY>(dy, r * C + c0, q.d); ld4<TX>(x, r * C + c0, q.x);
    } else {
      q.d[0] = ldf<TDY>(dy, r * C + c0); q.x[0] = ldf<TX>(x, r * C + c0);
    }
  }
  __device__ __forceinline__ void finish(int64_t, int c0, const Regs& q, float (&v)[2][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float xh = (q.x[j] - mean[c0 + j]) * rstd[c0 + j];
      v[0][j] = q.d[j]; v[1][j] = q.d[j] * xh;
    }
  }
  __device__ void operator()(int64_t r, int c0, float (&v)[2][VEC]) const {
    Regs q;
    load(r, c0, q);
    finish(r, c0, q, v);
  }
};

// ------------------------------------------------------------------------------------------------
// finalize kernels (tiny, one thread per channel)
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, float rows, int C, const float* gamma,
                                   const float* beta, float eps, float decay, int unbiased, float* mm, float* mv,
                                   float* mean, float* rstd, float* scale, float* shift) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mu = sum[c] / rows;
  float var = fmaxf(sumsq[c] / rows - mu * mu, 0.f);
  float rs = rsqrtf(var + eps);
  mean[c] = mu; rstd[c] = rs;
  float sc = gamma[c] * rs;
  scale[c] = sc; shift[c] = beta[c] - mu * sc;
  if (mm) mm[c] = mm[c] * decay + mu * (1.f - decay);
  // tf.contrib batch_norm (fused path) feeds the UNBIASED variance to the moving average; nn.batch_norm_impl
  // (nn.py:207-214) feeds tf.nn.moments' biased one
  if (mv) mv[c] = mv[c] * decay + var * (unbiased ? rows / fmaxf(rows - 1.f, 1.f) : 1.f) * (1.f - decay);
}
__global__ void bn_eval_kernel(const float* gamma, const float* beta, const float* mm, const float* mv, float eps,
                               int C, float* scale, float* shift) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = gamma[c] * rsqrtf(mv[c] + eps);
  scale[c] = sc; shift[c] = beta[c] - mm[c] * sc;
}

// ------------------------------------------------------------------------------------------------
// elementwise kernels over [rows, C]
// ------------------------------------------------------------------------------------------------
template <typename TX, typename TY, int VEC, int ACT>
__global__ void affine_act_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t nvec, int C,
                                  const float* __restrict__ scale, const float* __restrict__ shift, float alpha) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = i * VEC;
    int c = (int)(e % C);
    if constexpr (VEC == 4) {
      float v[4]; ld4<TX>(x, e, v);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        v[j] = act_fwd_t<ACT>(v[j] * (scale ? scale[c + j] : 1.f) + (shift ? shift[c + j] : 0.f), alpha);
      st4<TY>(y, e, v);
    } else {
      float v = ldf<TX>(x, e);
      stf<TY>(y, e, act_fwd_t<ACT>(v * (scale ? scale[c] : 1.f) + (shift ? shift[c] : 0.f), alpha));
    }
  }
}

// mean-only batch norm apply (nn.py:170-187) fused with the nonlinearity: shift = b - mean, mean = sum/rows in
// training (and pop_mean <- decay*pop_mean + (1-decay)*mean, done by CTA 0) or pop_mean at test time.
template <typename TX, typename TY, int VEC, int ACT>
__global__ void mobn_apply_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t nvec, int C,
                                  const float* __restrict__ sum, float inv_rows, const float* __restrict__ b,
                                  float* __restrict__ pop_mean, float decay, int train, float alpha) {
  pdl_entry();
  if (train && pop_mean && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x)
      pop_mean[c] = pop_mean[c] * decay + sum[c] * inv_rows * (1.f - decay);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = i * VEC;
    int c = (int)(e % C);
    float v[VEC];
    if constexpr (VEC == 4) ld4<TX>(x, e, *reinterpret_cast<float(*)[4]>(v));
    else v[0] = ldf<TX>(x, e);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      // in training pop_mean is being rewritten by CTA 0, but it is not read here (sum is)
      const float m = train ? sum[c + j] * inv_rows : pop_mean[c + j];
      v[j] = act_fwd_t<ACT>(v[j] + (b ? b[c + j] : 0.f) - m, alpha);
    }
    if constexpr (VEC == 4) st4<TY>(y, e, *reinterpret_cast<float(*)[4]>(v));
    else stf<TY>(y, e, v[0]);
  }
}

// per-segment channel sum: fp32, or the Q24 fixed-point value accumulated by the GEMM epilogue's integer atomics
// Q24 -> float without fp64: integer part and 24-bit fraction are each exact in fp32 (|sum| < 2^24), one rounding in the fma
__device__ __forceinline__ float q24_to_float(long long v) {
  return __fmaf_rn((float)(int)(v & 0xffffff), 1.f / 16777216.f, (float)(int)(v >> 24));
}
__device__ __forceinline__ float seg_sum(const void* sums, int q24, int64_t i) {
  return q24 ? q24_to_float(reinterpret_cast<const long long*>(sums)[i]) : reinterpret_cast<const float*>(sums)[i];
}

// 8 bf16 per thread (16-byte accesses)
__device__ __forceinline__ void ld8(const bf16* p, int64_t i, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p + i);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(h[j]); v[2 * j + 1] = __high2float(h[j]); }
}
__device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(h[j]); v[2 * j + 1] = __high2float(h[j]); }
}
__device__ __forceinline__ void st8(bf16* p, int64_t i, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p + i) = t;
}

// bf16 round trip of two values: one F2FP.PACK_AB + a shift and a mask (same round-to-nearest-even as __float2bfloat16_rn)
__device__ __forceinline__ void round_bf16_pair(float a, float b, float& ra, float& rb) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
  ra = __uint_as_float(u << 16);
  rb = __uint_as_float(u & 0xffff0000u);
}

// Segment-aware mean-only batch norm apply + nonlinearity, whole grouped batch in one launch (bf16, C % 8 == 0):
//   y = act(z - mean_seg + b),  mean_seg = sums[seg] / rows_seg   (training)   |   mean = pop_mean (test)
// pop_mean is updated once per segment IN CALL ORDER (the reference runs the calls one after the other).
template <int ACT>
__global__ void mobn_apply_seg_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t nvec, int C,
                                      const void* __restrict__ sums, int q24, Segs sg, const float* __restrict__ b,
                                      float* __restrict__ pop_mean, float decay, int train, float alpha) {
  pdl_entry();
  if (train && pop_mean && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float pm = pop_mean[c];
      for (int s = 0; s < sg.n; ++s) pm = pm * decay + seg_sum(sums, q24, (int64_t)s * C + c) * sg.inv_rows[s] * (1.f - decay);
      pop_mean[c] = pm;
    }
  }
  // the first vector of x is requested before the table is built, so its HBM latency covers the table's L2 latency
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 cur = make_uint4(0, 0, 0, 0);
  if (i0 < nvec) cur = *reinterpret_cast<const uint4*>(x + i0 * 8);
  // per-block table shift[seg][c] = b - mean (the fixed-point -> float conversion is done once per block, not per element)
  extern __shared__ float seg_shift[];
  for (int i = threadIdx.x; i < sg.n * C; i += blockDim.x) {
    const int s = i / C, c = i - s * C;
    // (in training pop_mean is being rewritten by CTA 0, but it is not read here: sums is)
    const float m = train ? seg_sum(sums, q24, i) * sg.inv_rows[s] : pop_mean[c];
    seg_shift[i] = b[c] - m;
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i0; i < nvec; i += stride) {
    const int64_t e = i * 8;
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    const float* sh = seg_shift + sg.of(r) * C + c;
    float v[8];
    unpack8(cur, v);
    if (i + stride < nvec) cur = *reinterpret_cast<const uint4*>(x + (i + stride) * 8);     // next vector in flight
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_fwd_t<ACT>(v[j] + sh[j], alpha);
    st8(y, e, v);
  }
}

// du = dy * act'(y) with per-segment column sums, bf16 with 16-byte accesses (the generic colreduce path moves 8 bytes per
// thread and accumulates in fp64; this one is the backward of every mean-only-BN layer of the classifier and runs at HBM
// speed).  256 threads = (C/8 channel groups) x (2048/C row lanes); each CTA walks a strided set of row blocks, folds
// its row lanes through shared memory and writes one fp32 partial per (segment, channel); act_bwd_seg_fold_kernel adds
// the partials of all CTAs in a fixed order in fp64 (deterministic).
template <int ACT>
__global__ void __launch_bounds__(256) act_bwd_seg_v8_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ y,
                                                             bf16* __restrict__ du, int64_t rows, int C, Segs sg,
                                                             float alpha, float* __restrict__ partials, int ldy) {
  pdl_entry();
  extern __shared__ float abs_sm[];              // [row lanes][4][C]; every thread owns its (lane, segment, 8 channels) slots
  const int cg = C / 8, rl = 256 / cg;           // channel groups per row, row lanes per CTA
  const int tc = threadIdx.x % cg, tr = threadIdx.x / cg;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) abs_sm[((size_t)tr * 4 + a) * C + tc * 8 + j] = 0.f;
  // a CONTIGUOUS range of rows per CTA: the segment changes at most three times, so one accumulator set in registers is
  // enough (flushed to the thread's shared-memory slot when the segment changes)
  const int64_t per = ((rows + gridDim.x - 1) / gridDim.x + rl - 1) / rl * rl;
  const int64_t rbeg = (int64_t)blockIdx.x * per, rend = rbeg + per < rows ? rbeg + per : rows;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int cur = -1;
  for (int64_t r0 = rbeg + tr; r0 < rend; r0 += 2 * rl) {
    const int64_t r1 = r0 + rl;
    float a0[8], b0[8], a1[8], b1[8];
    ld8(dy, r0 * C + tc * 8, a0); ld8(y, r0 * ldy + tc * 8, b0);
    if (r1 < rend) { ld8(dy, r1 * C + tc * 8, a1); ld8(y, r1 * ldy + tc * 8, b1); }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = u ? r1 : r0;
      if (r >= rend) break;
      const int s = sg.of(r);
      if (s != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { abs_sm[((size_t)tr * 4 + cur) * C + tc * 8 + j] += acc[j]; acc[j] = 0.f; }
        }
        cur = s;
      }
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (u ? a1[j] : a0[j]) * act_grad_from_y_t<ACT>(u ? b1[j] : b0[j], alpha);
        acc[j] += o[j];
      }
      st8(du, r * C + tc * 8, o);
    }
  }
  if (cur >= 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) abs_sm[((size_t)tr * 4 + cur) * C + tc * 8 + j] += acc[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += 256) {
    float t = 0.f;
    for (int l = 0; l < rl; ++l) t += abs_sm[(size_t)l * 4 * C + i];
    partials[(size_t)blockIdx.x * 4 * C + i] = t;
  }
}
// 256 threads = 8 channels x 32 part lanes (C/8 CTAs); fixed summation order -> deterministic
__global__ void __launch_bounds__(256) act_bwd_seg_fold_kernel(const float* __restrict__ partials, int nparts, int C,
                                                               float* __restrict__ colsums, float* __restrict__ grad_acc,
                                                               int nseg_out) {
  pdl_entry();
  __shared__ double sm[32][4][9];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3, c = blockIdx.x * 8 + tx;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  if (c < C)
    for (int p = ty; p < nparts; p += 32)
#pragma unroll
      for (int a = 0; a < 4; ++a) t[a] += (double)partials[((size_t)p * 4 + a) * C + c];
#pragma unroll
  for (int a = 0; a < 4; ++a) sm[ty][a][tx] = t[a];
  __syncthreads();
  if (ty == 0 && c < C) {
    double tot = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      double u = 0.0;
#pragma unroll
      for (int l = 0; l < 32; ++l) u += sm[l][a][tx];
      if (a < nseg_out) colsums[a * C + c] = (float)u;
      tot += u;
    }
    if (grad_acc) grad_acc[c] = (float)((double)grad_acc[c] + tot);
  }
}

// dz = du - colsums[seg] / rows_seg  (bf16, C % 8 == 0), in place allowed
__global__ void sub_mean_seg_kernel(const bf16* __restrict__ du, bf16* __restrict__ dz, int64_t nvec, int C,
                                    const float* __restrict__ colsums, Segs sg) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8;
    const int64_t r = e / C;
    const int c = (int)(e - r * C);
    const int s = sg.of(r);
    float v[8];
    ld8(du, e, v);
    const float4* mp = reinterpret_cast<const float4*>(colsums + (int64_t)s * C + c);
    float4 m0 = mp[0], m1 = mp[1];
    const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    const float sc = sg.inv_rows[s];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] -= mm[j] * sc;
    st8(dz, e, v);
  }
}

template <typename TA, typename TB, int VEC>
__global__ void sub_mean_kernel(const TA* __restrict__ du, TB* __restrict__ dz, int64_t nvec, int C,
                                const float* __restrict__ colsum, float inv_rows) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t e = i * VEC;
    int c = (int)(e % C);
    if constexpr (VEC == 4) {
      float v[4]; ld4<TA>(du, e, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] -= colsum[c + j] * inv_rows;
      st4<TB>(dz, e, v);
    } else {
      stf<TB>(dz, e, ldf<TA>(du, e) - colsum[c] * inv_rows);
    }
  }
}

template <typename TDY, typename TX, typename TDX>
__global__ void bn_bwd_apply_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x, TDX* __restrict__ dx,
                                    int64_t n, int C, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ s1,
                                    const float* __restrict__ s2, float inv_rows) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float xh = (ldf<TX>(x, i) - mean[c]) * rstd[c];
    float d = ldf<TDY>(dy, i);
    stf<TDX>(dx, i, gamma[c] * rstd[c] * (d - s1[c] * inv_rows - xh * s2[c] * inv_rows));
  }
}
__global__ void add_to_kernel(float* dst, const float* src, int n, float beta) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (beta != 0.f ? beta * dst[i] : 0.f) + src[i];
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float u1 = ((a >> 8) + 1) * (1.0f / 16777216.0f);  // (0,1]
  float u2 = u32_to_unit(b);
  float r = sqrtf(-2.f * logf(u1));
  float s, c;
  sincospif(2.f * u2, &s, &c);
  n0 = r * c; n1 = r * s;
}

template <typename TX, typename TY>
__global__ void add_noise_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t n, float std,
                                 const float* __restrict__ noise, uint64_t seed, uint64_t stream_id,
                                 const uint64_t* __restrict__ counter) {
  pdl_entry();
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 elements
  int64_t e = q * 4;
  if (e >= n) return;
  float z[4];
  if (noise) {
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = (e + j < n) ? noise[e + j] : 0.f;
  } else {
    uint64_t ctr = counter ? *counter : 0;
    uint4 r = Philox(seed)((uint64_t)q, stream_id + (ctr << 20));
    box_muller(r.x, r.y, z[0], z[1]);
    box_muller(r.z, r.w, z[2], z[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (e + j < n) stf<TY>(y, e + j, ldf<TX>(x, e + j) + std * z[j]);
}

template <typename TX, typename TY>
__global__ void dropout_kernel(const TX* __restrict__ x, TY* __restrict__ y, uint8_t* __restrict__ mask, int64_t n,
                               float rate, float scale, int gen, uint64_t seed, uint64_t stream_id,
                               const uint64_t* __restrict__ counter) {
  pdl_entry();
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t e = q * 4;
  if (e >= n) return;
  uint8_t k[4];
  if (gen) {
    uint64_t ctr = counter ? *counter : 0;
    uint4 r = Philox(seed)((uint64_t)q, stream_id + (ctr << 20));
    k[0] = u32_to_unit(r.x) >= rate; k[1] = u32_to_unit(r.y) >= rate;
    k[2] = u32_to_unit(r.z) >= rate; k[3] = u32_to_unit(r.w) >= rate;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = (e + j < n) ? mask[e + j] : 0;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (e + j < n) {
      if (gen) mask[e + j] = k[j];
      stf<TY>(y, e + j, k[j] ? ldf<TX>(x, e + j) * scale : 0.f);
    }
  }
}

template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx, int N,
                                    int H, int W, int C, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int64_t t = i / C;
  int Wo = W / 2, Ho = H / 2;
  int wo = (int)(t % Wo); t /= Wo;
  int ho = (int)(t % Ho);
  int n = (int)(t / Ho);
  const T* p = x + ((int64_t)(n * H + 2 * ho) * W + 2 * wo) * C + c;
  float v0 = ldf<T>(p, 0), v1 = ldf<T>(p, C), v2 = ldf<T>(p, (int64_t)W * C), v3 = ldf<T>(p, (int64_t)W * C + C);
  float m = v0; int k = 0;
  if (v1 > m) { m = v1; k = 1; }
  if (v2 > m) { m = v2; k = 2; }
  if (v3 > m) { m = v3; k = 3; }
  stf<T>(y, i, m);
  idx[i] = (uint8_t)k;
}
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx, T* __restrict__ dx,
                                    int N, int H, int W, int C, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int64_t t = i / C;
  int Wo = W / 2, Ho = H / 2;
  int wo = (int)(t % Wo); t /= Wo;
  int ho = (int)(t % Ho);
  int n = (int)(t / Ho);
  T* p = dx + ((int64_t)(n * H + 2 * ho) * W + 2 * wo) * C + c;
  float d = ldf<T>(dy, i);
  int k = idx[i];
  stf<T>(p, 0, k == 0 ? d : 0.f);
  stf<T>(p, C, k == 1 ? d : 0.f);
  stf<T>(p, (int64_t)W * C, k == 2 ? d : 0.f);
  stf<T>(p, (int64_t)W * C + C, k == 3 ? d : 0.f);
}

// 2x2/s2 max pool fused with the inverted dropout that follows it in the classifier (Good_GAN_cifar10.py:123-124,
// 142-143): bf16, C % 8 == 0, one thread = 8 channels of one output pixel (16-byte accesses).  code = winner index
// (2 bits, first maximum wins like tf.nn.max_pool's gradient) | keep << 2.  The Philox stream is the one the
// stand-alone dropout kernel draws for the same tag (element groups of 4).
__global__ void maxpool2_dropout_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ code,
                                            int H, int W, int C, int64_t nvec, float rate, float scale,
                                            const uint8_t* __restrict__ mask, uint64_t seed, uint64_t stream_id,
                                            const uint64_t* __restrict__ counter) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int cv = C / 8, Wo = W / 2, Ho = H / 2;
  const int c = (int)(i % cv) * 8;
  int64_t t = i / cv;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int64_t n = t / Ho;
  const int64_t p0 = ((n * H + 2 * ho) * W + 2 * wo) * C + c;
  float a[4][8];
  ld8(x, p0, a[0]); ld8(x, p0 + C, a[1]); ld8(x, p0 + (int64_t)W * C, a[2]); ld8(x, p0 + (int64_t)W * C + C, a[3]);
  uint8_t keep[8];
  const int64_t e = i * 8;                      // first output element of this thread
  if (rate <= 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = 1;
  } else if (mask) {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = mask[e + j];
  } else {
    const uint64_t ctr = counter ? *counter : 0;
    Philox ph(seed);
    const uint4 r0 = ph((uint64_t)(e / 4), stream_id + (ctr << 20)), r1 = ph((uint64_t)(e / 4 + 1), stream_id + (ctr << 20));
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = u32_to_unit(rr[j]) >= rate;
  }
  float o[8];
  uint8_t cd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m = a[0][j]; int k = 0;
    if (a[1][j] > m) { m = a[1][j]; k = 1; }
    if (a[2][j] > m) { m = a[2][j]; k = 2; }
    if (a[3][j] > m) { m = a[3][j]; k = 3; }
    o[j] = keep[j] ? m * scale : 0.f;
    cd[j] = (uint8_t)(k | (keep[j] << 2));
  }
  st8(y, e, o);
  uint2 cw;
  cw.x = cd[0] | (cd[1] << 8) | (cd[2] << 16) | ((uint32_t)cd[3] << 24);
  cw.y = cd[4] | (cd[5] << 8) | (cd[6] << 16) | ((uint32_t)cd[7] << 24);
  *reinterpret_cast<uint2*>(code + e) = cw;
}
__global__ void maxpool2_dropout_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ code,
                                            bf16* __restrict__ dx, int H, int W, int C, int64_t nvec, float scale) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const int cv = C / 8, Wo = W / 2, Ho = H / 2;
  const int c = (int)(i % cv) * 8;
  int64_t t = i / cv;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int64_t n = t / Ho;
  const int64_t p0 = ((n * H + 2 * ho) * W + 2 * wo) * C + c;
  const int64_t e = i * 8;
  float d[8];
  ld8(dy, e, d);
  const uint2 cw = *reinterpret_cast<const uint2*>(code + e);
  const uint32_t cws[2] = {cw.x, cw.y};
  float o[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t cd = (cws[j >> 2] >> (8 * (j & 3))) & 0xff;
    const float g = (cd & 4) ? d[j] * scale : 0.f;
    const int k = cd & 3;
    o[0][j] = k == 0 ? g : 0.f; o[1][j] = k == 1 ? g : 0.f; o[2][j] = k == 2 ? g : 0.f; o[3][j] = k == 3 ? g : 0.f;
  }
  st8(dx, p0, o[0]); st8(dx, p0 + C, o[1]); st8(dx, p0 + (int64_t)W * C, o[2]); st8(dx, p0 + (int64_t)W * C + C, o[3]);
}

// ---- mean-only BN apply + nonlinearity + 2x2 max pool + dropout in ONE pass (conv1_3 / conv2_3 of the classifier,
// Good_GAN_cifar10.py:118-124, 137-143).  The full-resolution activation y is never written: the pool's gradient only
// reaches the window winners, and the winner of y is the winner of its pre-activation (monotonic nonlinearity), so the
// backward needs y at the winners only -- which is the pooled output itself.  bf16, C % 8 == 0, one thread = 8 channels
// of one pooled pixel.  Same code byte and Philox stream as maxpool2_dropout_fwd_kernel.
template <int ACT>
__global__ void mobn_pool_dropout_fwd_kernel(const bf16* __restrict__ z, bf16* __restrict__ y, uint8_t* __restrict__ code,
                                             int H, int W, int C, int64_t nvec, const void* __restrict__ sums, int q24, Segs sg,
                                             int rows_per_img, const float* __restrict__ b, float* __restrict__ pop_mean,
                                             float decay, int train, float alpha, float rate, float scale,
                                             const uint8_t* __restrict__ mask, uint64_t seed, uint64_t stream_id,
                                             const uint64_t* __restrict__ counter) {
  pdl_entry();
  if (train && pop_mean && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float pm = pop_mean[c];
      for (int s = 0; s < sg.n; ++s) pm = pm * decay + seg_sum(sums, q24, (int64_t)s * C + c) * sg.inv_rows[s] * (1.f - decay);
      pop_mean[c] = pm;
    }
  }
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  // 32-bit index arithmetic (the host checks nvec < 2^31): 64-bit divisions were a third of this kernel's instructions
  const unsigned cv = C / 8, Wo = W / 2, Ho = H / 2;
  const unsigned iu = (unsigned)i;
  const int c = (int)(iu % cv) * 8;
  unsigned t = iu / cv;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int64_t n = t / Ho;
  const int s = sg.of(n * rows_per_img);                    // segments are whole images
  const int64_t p0 = ((n * H + 2 * ho) * W + 2 * wo) * C + c;
  float a[4][8];
  ld8(z, p0, a[0]); ld8(z, p0 + C, a[1]); ld8(z, p0 + (int64_t)W * C, a[2]); ld8(z, p0 + (int64_t)W * C + C, a[3]);
  // shift = b - mean of this thread's (segment, 8 channels): requested after the z loads so the latencies overlap
  float sh[8];
  {
    const float4* bp = reinterpret_cast<const float4*>(b + c);
    const float4 b0 = bp[0], b1 = bp[1];
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const float msc = train ? sg.inv_rows[s] : 1.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) sh[j] = bb[j] - (train ? seg_sum(sums, q24, (int64_t)s * C + c + j) : pop_mean[c + j]) * msc;
  }
  uint8_t keep[8];
  const int64_t e = i * 8;
  if (rate <= 0.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = 1;
  } else if (mask) {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = mask[e + j];
  } else {
    const uint64_t ctr = counter ? *counter : 0;
    Philox ph(seed);
    const uint4 r0 = ph((uint64_t)(e / 4), stream_id + (ctr << 20)), r1 = ph((uint64_t)(e / 4 + 1), stream_id + (ctr << 20));
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = u32_to_unit(rr[j]) >= rate;
  }
  float o[8];
  uint8_t cd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // the unfused path stores y in bf16 before pooling: round each candidate the same way
    float q[4];
#pragma unroll
    for (int k = 0; k < 4; k += 2)      // rounded in pairs on the ALU pipe (element-wise F2F goes through the quarter-rate unit)
      round_bf16_pair(act_fwd_t<ACT>(a[k][j] + sh[j], alpha), act_fwd_t<ACT>(a[k + 1][j] + sh[j], alpha), q[k], q[k + 1]);
    float m = q[0]; int k = 0;
    if (q[1] > m) { m = q[1]; k = 1; }
    if (q[2] > m) { m = q[2]; k = 2; }
    if (q[3] > m) { m = q[3]; k = 3; }
    o[j] = keep[j] ? m * scale : 0.f;
    cd[j] = (uint8_t)(k | (keep[j] << 2));
  }
  st8(y, e, o);
  uint2 cw;
  cw.x = cd[0] | (cd[1] << 8) | (cd[2] << 16) | ((uint32_t)cd[3] << 24);
  cw.y = cd[4] | (cd[5] << 8) | (cd[6] << 16) | ((uint32_t)cd[7] << 24);
  *reinterpret_cast<uint2*>(code + e) = cw;
}

// backward of the fused pass: du (full resolution) = winner ? keep * dy/(1-rate) * act'(y_winner) : 0, with per-segment
// column sums of du (partials per CTA, folded by act_bwd_seg_fold_kernel).  y_winner = pooled output * (1-rate) where kept.
// 256 threads = (C/8 channel groups) x (2048/C pooled-pixel lanes); contiguous pooled-pixel ranges per CTA.
template <int ACT>
__global__ void __launch_bounds__(256) mobn_pool_dropout_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ yp,
                                                                    const uint8_t* __restrict__ code, bf16* __restrict__ du,
                                                                    int H, int W, int C, int64_t prow_total, Segs sg,
                                                                    int rows_per_img, float alpha, float scale,
                                                                    float* __restrict__ partials) {
  pdl_entry();
  extern __shared__ float abs_sm[];              // [lanes][4][C]
  const int cg = C / 8, rl = 256 / cg;
  const int tc = threadIdx.x % cg, tr = threadIdx.x / cg;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) abs_sm[((size_t)tr * 4 + a) * C + tc * 8 + j] = 0.f;
  const int Wo = W / 2, Ho = H / 2;
  const int64_t per = ((prow_total + gridDim.x - 1) / gridDim.x + rl - 1) / rl * rl;
  const int64_t rbeg = (int64_t)blockIdx.x * per, rend = rbeg + per < prow_total ? rbeg + per : prow_total;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int cur = -1;
  const float inv_scale = 1.f / scale;
  for (int64_t r = rbeg + tr; r < rend; r += rl) {          // r = pooled pixel index (n, ho, wo)
    const unsigned ru = (unsigned)r, q = ru / (unsigned)Wo;   // 32-bit arithmetic (host checks the pixel count < 2^31)
    const int wo = (int)(ru - q * (unsigned)Wo);
    const unsigned nn = q / (unsigned)Ho;
    const int ho = (int)(q - nn * (unsigned)Ho);
    const int64_t n = nn;
    const int s = sg.of(n * rows_per_img);
    if (s != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { abs_sm[((size_t)tr * 4 + cur) * C + tc * 8 + j] += acc[j]; acc[j] = 0.f; }
      }
      cur = s;
    }
    const int64_t e = r * C + tc * 8;
    float d[8], yv[8];
    ld8(dy, e, d); ld8(yp, e, yv);
    const uint2 cw = *reinterpret_cast<const uint2*>(code + e);
    const uint32_t cws[2] = {cw.x, cw.y};
    float o[4][8], gqv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t cd = (cws[j >> 2] >> (8 * (j & 3))) & 0xff;
      gqv[j] = (cd & 4) ? d[j] * scale * act_grad_from_y_t<ACT>(yv[j] * inv_scale, alpha) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) round_bf16_pair(gqv[j], gqv[j + 1], gqv[j], gqv[j + 1]);   // du is stored in bf16: sum what is stored
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t cd = (cws[j >> 2] >> (8 * (j & 3))) & 0xff;
      const float gq = gqv[j];
      const int k = cd & 3;
      o[0][j] = k == 0 ? gq : 0.f; o[1][j] = k == 1 ? gq : 0.f; o[2][j] = k == 2 ? gq : 0.f; o[3][j] = k == 3 ? gq : 0.f;
      acc[j] += gq;
    }
    const int64_t p0 = ((n * H + 2 * ho) * W + 2 * wo) * C + tc * 8;
    st8(du, p0, o[0]); st8(du, p0 + C, o[1]); st8(du, p0 + (int64_t)W * C, o[2]); st8(du, p0 + (int64_t)W * C + C, o[3]);
  }
  if (cur >= 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) abs_sm[((size_t)tr * 4 + cur) * C + tc * 8 + j] += acc[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * C; i += 256) {
    float t = 0.f;
    for (int l = 0; l < rl; ++l) t += abs_sm[(size_t)l * 4 * C + i];
    partials[(size_t)blockIdx.x * 4 * C + i] = t;
  }
}

template <typename TX, typename TY>
__global__ void global_pool_fwd_kernel(const TX* __restrict__ x, TY* __restrict__ y, uint8_t* __restrict__ idx,
                                       int N, int HW, int C, int mode) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  int c = i % C, n = i / C;
  const TX* p = x + (int64_t)n * HW * C + c;
  if (mode == 0) {
    float m = ldf<TX>(p, 0); int k = 0;
    for (int j = 1; j < HW; ++j) { float v = ldf<TX>(p, (int64_t)j * C); if (v > m) { m = v; k = j; } }
    stf<TY>(y, i, m);
    if (idx) idx[i] = (uint8_t)k;
  } else {
    float s = 0.f;
    for (int j = 0; j < HW; ++j) s += ldf<TX>(p, (int64_t)j * C);
    stf<TY>(y, i, s / HW);
  }
}
template <typename TDY, typename TDX>
__global__ void global_pool_bwd_kernel(const TDY* __restrict__ dy, const uint8_t* __restrict__ idx,
                                       TDX* __restrict__ dx, int N, int HW, int C, int mode, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int64_t t = i / C;
  int j = (int)(t % HW);
  int n = (int)(t / HW);
  float d = ldf<TDY>(dy, (int64_t)n * C + c);
  float v = mode == 0 ? (idx[n * C + c] == j ? d : 0.f) : d / HW;
  stf<TDX>(dx, i, v);
}

// mean pooling, bf16, 8 channels per thread (D's average_pooling2d(8, 1), Good_GAN_cifar10.py:94): the scalar kernel
// spends two 64-bit divisions on every element
__global__ void global_pool_bwd_mean_v8_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, unsigned HW, unsigned cv,
                                               unsigned nvec, float inv) {
  pdl_entry();
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec) return;
  const unsigned t = i / cv, c8 = i - t * cv, n = t / HW;
  float v[8];
  ld8(dy, ((int64_t)n * cv + c8) * 8, v);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] *= inv;
  st8(dx, (int64_t)i * 8, v);
}

template <typename TX, typename TO>
__global__ void concat_label_kernel(const TX* __restrict__ x, int64_t rows, int C, int ldx,
                                    const float* __restrict__ lab, int K, int rps, TO* __restrict__ out, int ldo,
                                    int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = (int)(i % ldo);
  int64_t r = i / ldo;
  float v = 0.f;
  if (j < C) v = ldf<TX>(x, r * ldx + j);
  else if (j < C + K) v = lab[(r / rps) * K + (j - C)];
  stf<TO>(out, i, v);
}
// Dropout / per-channel affine written straight into a label-concatenated tensor (row stride ldy, label planes and zero pad
// filled by the same launch): `dropout -> _conv_cond_concat` of the discriminator (Good_GAN_cifar10.py:73-75, 83-85) and
// `batch_norm -> _conv_cond_concat` of the generator (:43-46, 49-51) as one pass instead of two.  The first `ngroups`
// threads own 4 consecutive x elements (the Philox grouping of dropout_kernel), the rest own one label / pad element.
template <typename TX, typename TY>
__global__ void dropout_concat_kernel(const TX* __restrict__ x, TY* __restrict__ y, uint8_t* __restrict__ mask, int64_t n,
                                      int C, int ldy, float rate, float scale, int gen, uint64_t seed, uint64_t stream_id,
                                      const uint64_t* __restrict__ counter, const float* __restrict__ lab, int K, int rps,
                                      int64_t ngroups, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (i < ngroups) {
    const int64_t e = i * 4;
    uint8_t k[4];
    if (gen) {
      uint64_t ctr = counter ? *counter : 0;
      uint4 r = Philox(seed)((uint64_t)i, stream_id + (ctr << 20));
      k[0] = u32_to_unit(r.x) >= rate; k[1] = u32_to_unit(r.y) >= rate;
      k[2] = u32_to_unit(r.z) >= rate; k[3] = u32_to_unit(r.w) >= rate;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = (e + j < n) ? mask[e + j] : 0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (e + j < n) {
        if (gen) mask[e + j] = k[j];
        const int64_t r = (e + j) / C;
        const int c = (int)((e + j) - r * C);
        stf<TY>(y, r * ldy + c, k[j] ? ldf<TX>(x, e + j) * scale : 0.f);
      }
    }
  } else {
    const int w = ldy - C;
    const int64_t t = i - ngroups;
    const int j = (int)(t % w);
    const int64_t r = t / w;
    stf<TY>(y, r * ldy + C + j, j < K ? lab[(r / rps) * K + j] : 0.f);
  }
}
// narrow rows (the discriminator's input: 3 image channels + 10 label planes in a 16-channel bf16 row): one thread = one
// output row, written with 16-byte stores; same Philox stream / element mapping as dropout_concat_kernel.
template <typename TX>
__global__ void dropout_concat_narrow_kernel(const TX* __restrict__ x, bf16* __restrict__ y, uint8_t* __restrict__ mask,
                                             int64_t rows, int C, int ldy, float rate, float scale, int gen, uint64_t seed,
                                             uint64_t stream_id, const uint64_t* __restrict__ counter,
                                             const float* __restrict__ lab, int K, int rps) {
  pdl_entry();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const uint64_t ctr = (gen && counter) ? *counter : 0;
  const int64_t e0 = r * C;
  const float* l = lab + (r / rps) * K;
  int64_t gcur = -1;
  uint4 rnd = make_uint4(0, 0, 0, 0);
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    v[j] = 0.f;
    if (j < C) {
      const int64_t e = e0 + j;
      uint8_t k;
      if (gen) {
        const int64_t g = e >> 2;
        if (g != gcur) { rnd = Philox(seed)((uint64_t)g, stream_id + (ctr << 20)); gcur = g; }
        const int q = (int)(e & 3);
        const uint32_t w = q == 0 ? rnd.x : (q == 1 ? rnd.y : (q == 2 ? rnd.z : rnd.w));
        k = u32_to_unit(w) >= rate;
        mask[e] = k;
      } else {
        k = mask[e];
      }
      v[j] = k ? ldf<TX>(x, e) * scale : 0.f;
    } else if (j - C < K && j < ldy) {
      v[j] = l[j - C];
    }
  }
  float lo[8], hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
  st8(y, r * ldy, lo);
  if (ldy > 8) st8(y, r * ldy + 8, hi);
}

template <typename TX, typename TY>
__global__ void affine_concat_kernel(const TX* __restrict__ x, TY* __restrict__ y, int C, int ldy,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ lab, int K, int rps, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = (int)(i % ldy);
  const int64_t r = i / ldy;
  float v = 0.f;
  if (j < C) v = ldf<TX>(x, r * C + j) * (scale ? scale[j] : 1.f) + (shift ? shift[j] : 0.f);
  else if (j < C + K) v = lab[(r / rps) * K + (j - C)];
  stf<TY>(y, i, v);
}

// bf16 in and out, C % 8 == 0, ldy % 8 == 0: one thread = one 16-byte chunk of an output row
__global__ void affine_concat_v8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int C, int ldy,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ lab, int K, int rps, int64_t nvec) {
  pdl_entry();
  const int cw = ldy / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cw;
    const int j0 = (int)(i - r * cw) * 8;
    float v[8];
    if (j0 < C) {
      ld8(x, r * C + j0, v);
      if (scale) {
        const float4 s0 = *reinterpret_cast<const float4*>(scale + j0), s1 = *reinterpret_cast<const float4*>(scale + j0 + 4);
        v[0] *= s0.x; v[1] *= s0.y; v[2] *= s0.z; v[3] *= s0.w; v[4] *= s1.x; v[5] *= s1.y; v[6] *= s1.z; v[7] *= s1.w;
      }
      if (shift) {
        const float4 s0 = *reinterpret_cast<const float4*>(shift + j0), s1 = *reinterpret_cast<const float4*>(shift + j0 + 4);
        v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
      }
    } else {
      const float* l = lab + (r / rps) * K;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (j0 - C + j < K) ? l[j0 - C + j] : 0.f;
    }
    st8(y, r * ldy + j0, v);
  }
}

// out[r, C + j] = j < K ? lab[(r / rps) * K + j] : 0 for j in [0, ldo - C): the label planes (+ zero pad) of a tensor whose
// first C channels were written in place by the producing GEMM epilogue
template <typename TO>
__global__ void fill_label_kernel(const float* __restrict__ lab, int K, int rps, TO* __restrict__ out, int C, int ldo,
                                  int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = ldo - C;
  const int j = (int)(i % w);
  const int64_t r = i / w;
  stf<TO>(out, r * ldo + C + j, j < K ? lab[(r / rps) * K + j] : 0.f);
}
// bf16, C % 8 == 0 and ldo % 8 == 0: one thread = one 16-byte chunk of a row's label / pad tail
__global__ void fill_label_v8_kernel(const float* __restrict__ lab, int K, int rps, bf16* __restrict__ out, int C, int ldo,
                                     int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cw = (ldo - C) / 8;
  const int j0 = (int)(i % cw) * 8;
  const int64_t r = i / cw;
  const float* l = lab + (r / rps) * K;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (j0 + j < K) ? l[j0 + j] : 0.f;
  st8(out, r * ldo + C + j0, v);
}
template <typename TS, typename TD>
__global__ void copy_channels_kernel(const TS* __restrict__ src, int lds, TD* __restrict__ dst, int ldd, int C,
                                     int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int j = (int)(i % C);
  int64_t r = i / C;
  stf<TD>(dst, r * ldd + j, ldf<TS>(src, r * lds + j));
}
// Concatenation of up to 8 contiguous sources (fp32 or bf16 each) into one destination: the grouped batch of several
// network calls (ops.group_batch) in ONE launch instead of one copy per source.
struct GatherSrcs {
  const void* src[8];
  int64_t end[8];      // exclusive prefix ends, in elements
  int dt[8];
  int n;
};
template <typename TD>
__global__ void gather_rows_kernel(const __grid_constant__ GatherSrcs g, TD* __restrict__ dst, int64_t total) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (k + 1 < g.n && i >= g.end[k]) ++k;
    const int64_t j = i - (k ? g.end[k - 1] : 0);
    const float v = g.dt[k] == TGAN_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(g.src[k])[j])
                                         : reinterpret_cast<const float*>(g.src[k])[j];
    stf<TD>(dst, i, v);
  }
}
// y = sum of `parts` bf16 slices laid out back to back ([parts][n]): the split-K partial outputs of one contraction
// (tap groups run as output classes of ONE igemm launch, tgan/tc.py), 16-byte accesses, fp32 sum in a fixed order
__global__ void sum_slices_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t nvec, int parts) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < parts; ++p) {
      float v[8];
      ld8(x, ((int64_t)p * nvec + i) * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    st8(y, i * 8, acc);
  }
}
template <typename T>
__global__ void accumulate_kernel(T* __restrict__ y, const T* __restrict__ x, int64_t n) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(y, i, ldf<T>(y, i) + ldf<T>(x, i));
}
__global__ void fill_kernel(float* p, float v, int64_t n) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void argmax_onehot_kernel(const float* __restrict__ logits, int N, int K, int64_t* __restrict__ idx,
                                     float* __restrict__ onehot) {
  pdl_entry();
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* p = logits + (int64_t)n * K;
  // tf.argmax = Eigen's ArgMaxTupleReducer: the accumulator starts at (0, -FLT_MAX) and is replaced only by a
  // strictly GREATER element -> the lowest index wins ties, a NaN never wins and never blocks a later number,
  // rows holding only NaN / -inf give index 0.
  int best = 0;
  float m = -3.402823466e+38f;
  for (int k = 0; k < K; ++k) {
    float v = p[k];
    if (v > m) { m = v; best = k; }
  }
  if (idx) idx[n] = best;
  if (onehot)
    for (int k = 0; k < K; ++k) onehot[(int64_t)n * K + k] = (k == best) ? 1.f : 0.f;
}

static inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = 148 * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// Mean-only batch norm of a SMALL fp32 tensor (the classifier's 10 logits, Good_GAN_cifar10.py:166; nn.py:147-187) for
// all segments of a grouped batch in ONE single-CTA launch: per-segment column means (fixed summation order), pop_mean
// updated once per segment in call order, apply + nonlinearity.  Replaces 2 launches per segment (8 for the grouped
// phase-C batch) of a few microseconds each.  256 threads = 32 columns x 8 row lanes; C <= 32.
struct SmallSegs { int n; int end[4]; };      // exclusive end row of every segment (unused entries = rows)
__device__ __forceinline__ int small_seg_of(const SmallSegs& sg, int r) { return (r >= sg.end[0]) + (r >= sg.end[1]) + (r >= sg.end[2]); }
// ONE pass over the rows for all segments: thread (column tx, row lane ty) adds row r into the accumulator of r's segment
// (four independent accumulators, selected without branches), so the loads of all segments are in flight together.
template <typename Fn>
__device__ __forceinline__ void small_seg_colsums(const SmallSegs& sg, int C, float (*part)[4][33], float (*out)[32], Fn val) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rows = sg.end[sg.n - 1];
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  if (tx < C) {
    // eight rows per round with unpredicated loads (full rounds), then a checked tail: predicated loads with per-load
    // address arithmetic were compiled to one load in flight per L2 round trip (5-14 us for a 10 KB tensor)
    int r0 = ty;
    for (; r0 + 56 < rows; r0 += 64) {
      float v[8];
      const int base = r0 * C + tx;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = val(base + 8 * k * C);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int s = small_seg_of(sg, r0 + 8 * k);
        a[0] += s == 0 ? v[k] : 0.f; a[1] += s == 1 ? v[k] : 0.f; a[2] += s == 2 ? v[k] : 0.f; a[3] += s == 3 ? v[k] : 0.f;
      }
    }
    for (; r0 < rows; r0 += 8) {
      const float v = val(r0 * C + tx);
      const int s = small_seg_of(sg, r0);
      a[0] += s == 0 ? v : 0.f; a[1] += s == 1 ? v : 0.f; a[2] += s == 2 ? v : 0.f; a[3] += s == 3 ? v : 0.f;
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) part[ty][s][tx] = a[s];
  __syncthreads();
  if (threadIdx.x < 128) {      // 4 segments x 32 columns, fixed order over the 8 row lanes
    const int s = threadIdx.x >> 5, c = threadIdx.x & 31;
    float t = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) t += part[l][s][c];
    out[s][c] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) mobn_small_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, int C, SmallSegs sg,
                                                             const float* __restrict__ b, float* __restrict__ pop_mean,
                                                             float decay, int train, int act, float alpha) {
  pdl_entry();
  __shared__ float part[8][4][33];
  __shared__ float mean[4][32];
  if (train) {
    small_seg_colsums(sg, C, part, mean, [&](int i) { return z[i]; });
    if (threadIdx.x < 128) {
      const int s = threadIdx.x >> 5, c = threadIdx.x & 31;
      const int n = s < sg.n ? sg.end[s] - (s ? sg.end[s - 1] : 0) : 1;
      mean[s][c] *= 1.0f / (float)n;
    }
  } else if (threadIdx.x < 128) {
    mean[threadIdx.x >> 5][threadIdx.x & 31] = (int)(threadIdx.x & 31) < C ? pop_mean[threadIdx.x & 31] : 0.f;
  }
  __syncthreads();
  if (train && pop_mean && threadIdx.x < C) {
    float pm = pop_mean[threadIdx.x];
    for (int s = 0; s < sg.n; ++s) pm = pm * decay + mean[s][threadIdx.x] * (1.f - decay);
    pop_mean[threadIdx.x] = pm;
  }
  const int total = sg.end[sg.n - 1] * C;
  for (int i0 = threadIdx.x; i0 < total; i0 += 1024) {      // four elements per round, loads first
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + 256 * k < total ? z[i0 + 256 * k] : 0.f;
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + 256 * k;
      if (i >= total) break;
      const int r = i / C, c = i - r * C;
      y[i] = act_fwd(v[k] + (b ? b[c] : 0.f) - mean[small_seg_of(sg, r)][c], act, alpha);
    }
  }
}

// backward of the same: du = dy * act'(y); grad_acc[c] += sum over all rows of du (the bias gradient);
// dz = du - mean_{rows of the segment}(du) when the batch mean was subtracted (training), else du.
__global__ void __launch_bounds__(256) mobn_small_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                             float* __restrict__ dz, int C, SmallSegs sg, int act, float alpha,
                                                             int subtract_mean, float* __restrict__ grad_acc) {
  pdl_entry();
  __shared__ float part[8][4][33];
  __shared__ float sums[4][32];
  small_seg_colsums(sg, C, part, sums, [&](int i) { return dy[i] * act_grad_from_y(y[i], act, alpha); });
  if (grad_acc && threadIdx.x < C) {
    float t = 0.f;
    for (int s = 0; s < sg.n; ++s) t += sums[s][threadIdx.x];
    grad_acc[threadIdx.x] += t;
  }
  if (dz) {
    const int total = sg.end[sg.n - 1] * C;
    for (int i0 = threadIdx.x; i0 < total; i0 += 1024) {
      float d[4], yy[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = i0 + 256 * k < total;
        d[k] = ok ? dy[i0 + 256 * k] : 0.f; yy[k] = ok ? y[i0 + 256 * k] : 0.f;
      }
      asm volatile("" : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]), "+f"(yy[0]), "+f"(yy[1]), "+f"(yy[2]), "+f"(yy[3]));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + 256 * k;
        if (i >= total) break;
        const int r = i / C, c = i - r * C;
        const int s = small_seg_of(sg, r);
        const int n = sg.end[s] - (s ? sg.end[s - 1] : 0);
        dz[i] = d[k] * act_grad_from_y(yy[k], act, alpha) - (subtract_mean ? sums[s][c] / (float)n : 0.f);
      }
    }
  }
}

}  // namespace tgan

using namespace tgan;

#define DISPATCH_2(d1, T1, d2, T2, ...) TGAN_DISPATCH_1(d1, T1, TGAN_DISPATCH_1(d2, T2, __VA_ARGS__))

extern "C" int tgan_channel_stats(const void* x, int xdt, int64_t rows, int C, float* sum, float* sumsq, float beta,
                                  float* ws, void* stream) {
  TGAN_CHECK_ARG(x && sum && ws && rows > 0 && C > 0, "channel_stats: bad args");
  bool v = (C % 4 == 0) && aligned16(x);
  TGAN_DISPATCH_1(xdt, T, {
    StatsF<T, 1> f1{(const T*)x, C};
    StatsF<T, 4> f4{(const T*)x, C};
    return run_colreduce<2>(f1, f4, v, rows, C, sum, sumsq, beta, ws, (cudaStream_t)stream);
  });
  return 0;
}

extern "C" int tgan_mobn_apply(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, const float* sum,
                               const float* b, float* pop_mean, float decay, int train, int act, float alpha,
                               void* stream) {
  TGAN_CHECK_ARG(x && y && rows > 0 && C > 0, "mobn_apply: bad args");
  TGAN_CHECK_ARG(train ? sum != nullptr : pop_mean != nullptr, "mobn_apply: needs sum (train) or pop_mean (test)");
  int64_t n = rows * C;
  bool v = (C % 4 == 0) && aligned16(x) && aligned16(y);
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = 1.0f / (float)rows;
  DISPATCH_2(xdt, TX, ydt, TY, {
    TGAN_DISPATCH_ACT(act, A, {
      if (v) pdl_launch(mobn_apply_kernel<TX, TY, 4, A>, grid_for(n / 4), 256, 0, (cudaStream_t)(st), (const TX*)x, (TY*)y, n / 4, C, sum, inv, b, pop_mean, decay, train, alpha);
      else pdl_launch(mobn_apply_kernel<TX, TY, 1, A>, grid_for(n), 256, 0, (cudaStream_t)(st), (const TX*)x, (TY*)y, n, C, sum, inv, b, pop_mean, decay, train, alpha);
    });
  });
  TGAN_LAUNCHED();
  return 0;
}
static int make_segs(Segs& sg, int64_t rows, int nseg, int64_t r0, int64_t r1, int64_t r2) {
  if (nseg < 1 || nseg > 4) { set_error("segments: nseg %d out of range [1,4]", nseg); return 1; }
  const int64_t e[4] = {r0, r1, r2, rows};
  int64_t prev = 0;
  sg.n = nseg;
  for (int i = 0; i < 4; ++i) {
    const int64_t end = i < nseg - 1 ? e[i] : rows;
    if (i < nseg && end <= prev) { set_error("segments: boundaries must be increasing and non-empty"); return 1; }
    if (i < 3) sg.end[i] = i < nseg - 1 ? end : INT64_MAX;
    sg.inv_rows[i] = i < nseg ? 1.0f / (float)(end - prev) : 0.f;
    if (i < nseg) prev = end;
  }
  return 0;
}

extern "C" int tgan_mobn_apply_seg(const void* x, void* y, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                                   int64_t r2, const void* sums, int sums_q24, const float* b, float* pop_mean,
                                   float decay, int train, int act, float alpha, void* stream) {
  TGAN_CHECK_ARG(x && y && b && rows > 0 && C > 0 && C % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(b),
                 "mobn_apply_seg: bf16 tensors with C %% 8 == 0 and 16-byte alignment only");
  TGAN_CHECK_ARG(train ? (sums != nullptr && aligned16(sums)) : (pop_mean != nullptr && aligned16(pop_mean)),
                 "mobn_apply_seg: needs sums (train) or pop_mean (test)");
  Segs sg;
  if (make_segs(sg, rows, nseg, r0, r1, r2)) return 1;
  const int64_t nvec = rows * C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  TGAN_DISPATCH_ACT(act, A, (pdl_launch(mobn_apply_seg_kernel<A>, std::min(grid_for(nvec), 148 * 8), 256, (size_t)nseg * C * sizeof(float), (cudaStream_t)(st), (const bf16*)x, (bf16*)y, nvec, C, sums, sums_q24, sg, b, pop_mean, decay, train, alpha)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_act_bwd_seg(const void* dy, int dydt, const void* y, int ydt, void* du, int dudt, int64_t rows,
                                int C, int nseg, int64_t r0, int64_t r1, int64_t r2, int act, float alpha,
                                float* colsums, float* grad_acc, float* ws, void* stream) {
  TGAN_CHECK_ARG(dy && y && du && ws && colsums && rows > 0 && C > 0, "act_bwd_seg: bad args");
  Segs sg;
  if (make_segs(sg, rows, nseg, r0, r1, r2)) return 1;
  bool v = (C % 4 == 0) && aligned16(dy) && aligned16(y) && aligned16(du);
  if (dydt == TGAN_BF16 && ydt == TGAN_BF16 && dudt == TGAN_BF16 && v && C % 8 == 0 && 2048 % C == 0 && C >= 64 &&
      (act == TGAN_ACT_LRELU || act == TGAN_ACT_RELU || act == TGAN_ACT_NONE)) {
    const int nparts = TGAN_ACT_BWD_SEG_PARTS;    // 3 CTAs per SM; ws holds nparts * 4 * C floats
    const size_t smem = (size_t)(2048 / C) * 4 * C * sizeof(float);      // 32 KB
    cudaStream_t st = (cudaStream_t)stream;
    TGAN_DISPATCH_ACT(act, A, (pdl_launch(act_bwd_seg_v8_kernel<A>, nparts, 256, smem, (cudaStream_t)(st), (const bf16*)dy, (const bf16*)y, (bf16*)du, rows, C, sg, alpha, ws, C)));
    TGAN_LAUNCHED();
    pdl_launch(act_bwd_seg_fold_kernel, ceil_div(C, 8), 256, 0, (cudaStream_t)(st), ws, nparts, C, colsums, grad_acc, 4);
    TGAN_LAUNCHED();
    return 0;
  }
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(ydt, TY, TGAN_DISPATCH_1(dudt, TDU, {
    TGAN_DISPATCH_ACT(act, A, {
      ActBwdSegF<TDY, TY, TDU, 1, A> f1{(const TDY*)dy, (const TY*)y, (TDU*)du, C, alpha, sg};
      ActBwdSegF<TDY, TY, TDU, 4, A> f4{(const TDY*)dy, (const TY*)y, (TDU*)du, C, alpha, sg};
      return run_colreduce<4>(f1, f4, v, rows, C, colsums, nullptr, 0.f, ws, (cudaStream_t)stream, grad_acc, 1);
    });
  })));
  return 0;
}

extern "C" int tgan_sub_channel_mean_seg(const void* du, void* dz, int64_t rows, int C, int nseg, int64_t r0,
                                         int64_t r1, int64_t r2, const float* colsums, void* stream) {
  TGAN_CHECK_ARG(du && dz && colsums && rows > 0 && C > 0 && C % 8 == 0 && aligned16(du) && aligned16(dz) &&
                     aligned16(colsums), "sub_channel_mean_seg: bf16 tensors with C %% 8 == 0 and 16-byte alignment only");
  Segs sg;
  if (make_segs(sg, rows, nseg, r0, r1, r2)) return 1;
  const int64_t nvec = rows * C / 8;
  pdl_launch(sub_mean_seg_kernel, grid_for(nvec), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const bf16*)du, (bf16*)dz, nvec, C, colsums, sg);
  TGAN_LAUNCHED();
  return 0;
}

static int make_small_segs(SmallSegs& sg, int64_t rows, int C, int nseg, int64_t r0, int64_t r1, int64_t r2) {
  if (nseg < 1 || nseg > 4 || C < 1 || C > 32 || rows < 1 || rows * C > (1 << 20)) {
    set_error("mobn_small: needs 1 <= nseg <= 4, C <= 32, rows * C <= 2^20 (got nseg %d, C %d, rows %lld)", nseg, C, (long long)rows);
    return 1;
  }
  const int64_t e[4] = {r0, r1, r2, rows};
  int64_t prev = 0;
  sg.n = nseg;
  for (int i = 0; i < 4; ++i) {
    const int64_t end = i < nseg - 1 ? e[i] : rows;
    if (i < nseg && end <= prev) { set_error("mobn_small: segment boundaries must be increasing and non-empty"); return 1; }
    sg.end[i] = (int)end;
    if (i < nseg) prev = end;
  }
  return 0;
}
extern "C" int tgan_mobn_small_fwd(const float* z, float* y, int64_t rows, int C, int nseg, int64_t r0, int64_t r1, int64_t r2,
                                   const float* b, float* pop_mean, float decay, int train, int act, float alpha, void* stream) {
  TGAN_CHECK_ARG(z && y && (train || pop_mean), "mobn_small_fwd: bad args (test mode needs pop_mean)");
  SmallSegs sg;
  if (make_small_segs(sg, rows, C, nseg, r0, r1, r2)) return 1;
  pdl_launch(mobn_small_fwd_kernel, 1, 256, 0, (cudaStream_t)stream, z, y, C, sg, b, pop_mean, decay, train, act, alpha);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_mobn_small_bwd(const float* dy, const float* y, float* dz, int64_t rows, int C, int nseg, int64_t r0,
                                   int64_t r1, int64_t r2, int act, float alpha, int subtract_mean, float* grad_acc,
                                   void* stream) {
  TGAN_CHECK_ARG(dy && y && (dz || grad_acc), "mobn_small_bwd: bad args");
  SmallSegs sg;
  if (make_small_segs(sg, rows, C, nseg, r0, r1, r2)) return 1;
  pdl_launch(mobn_small_bwd_kernel, 1, 256, 0, (cudaStream_t)stream, dy, y, dz, C, sg, act, alpha, subtract_mean, grad_acc);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_bn_finalize(const float* sum, const float* sumsq, int64_t rows, int C, const float* gamma,
                                const float* beta, float eps, float decay, int unbiased_moving_var, float* moving_mean,
                                float* moving_var, float* mean, float* rstd, float* scale, float* shift, void* stream) {
  TGAN_CHECK_ARG(sum && sumsq && gamma && beta && mean && rstd && scale && shift, "bn_finalize: bad args");
  pdl_launch(bn_finalize_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)((cudaStream_t)stream), sum, sumsq, (float)rows, C, gamma, beta, eps,
                                                                         decay, unbiased_moving_var, moving_mean, moving_var, mean, rstd,
                                                                         scale, shift);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_bn_eval_affine(const float* gamma, const float* beta, const float* moving_mean,
                                   const float* moving_var, float eps, int C, float* scale, float* shift,
                                   void* stream) {
  pdl_launch(bn_eval_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)((cudaStream_t)stream), gamma, beta, moving_mean, moving_var, eps, C,
                                                                     scale, shift);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_affine_act(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, const float* scale,
                               const float* shift, int act, float alpha, void* stream) {
  TGAN_CHECK_ARG(x && y && rows > 0 && C > 0, "affine_act: bad args");
  int64_t n = rows * C;
  bool v = (C % 4 == 0) && aligned16(x) && aligned16(y);
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_2(xdt, TX, ydt, TY, {
    TGAN_DISPATCH_ACT(act, A, {
      if (v) pdl_launch(affine_act_kernel<TX, TY, 4, A>, grid_for(n / 4), 256, 0, (cudaStream_t)(st), (const TX*)x, (TY*)y, n / 4, C, scale, shift, alpha);
      else pdl_launch(affine_act_kernel<TX, TY, 1, A>, grid_for(n), 256, 0, (cudaStream_t)(st), (const TX*)x, (TY*)y, n, C, scale, shift, alpha);
    });
  });
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_act_bwd(const void* dy, int dydt, const void* y, int ydt, void* du, int dudt, int64_t rows, int C,
                            int act, float alpha, float* colsum, float* grad_acc, float* ws, void* stream) {
  return tgan_act_bwd_ld(dy, dydt, y, ydt, C, du, dudt, rows, C, act, alpha, colsum, grad_acc, ws, stream);
}
extern "C" int tgan_act_bwd_ld(const void* dy, int dydt, const void* y, int ydt, int ldy, void* du, int dudt,
                               int64_t rows, int C, int act, float alpha, float* colsum, float* grad_acc, float* ws,
                               void* stream) {
  TGAN_CHECK_ARG(dy && y && du && ws && rows > 0 && C > 0 && ldy >= C, "act_bwd: bad args");
  bool v = (C % 4 == 0) && (ldy % 4 == 0) && aligned16(dy) && aligned16(y) && aligned16(du);
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(ydt, TY, TGAN_DISPATCH_1(dudt, TDU, {
    TGAN_DISPATCH_ACT(act, A, {
      ActBwdF<TDY, TY, TDU, 1, A> f1{(const TDY*)dy, (const TY*)y, (TDU*)du, C, alpha, ldy};
      ActBwdF<TDY, TY, TDU, 4, A> f4{(const TDY*)dy, (const TY*)y, (TDU*)du, C, alpha, ldy};
      return run_colreduce<1>(f1, f4, v, rows, C, colsum, nullptr, 0.f, ws, (cudaStream_t)stream, grad_acc);
    });
  })));
  return 0;
}

extern "C" int tgan_sub_channel_mean(const void* du, int dudt, void* dz, int dzdt, int64_t rows, int C,
                                     const float* colsum, void* stream) {
  TGAN_CHECK_ARG(du && dz && colsum && rows > 0, "sub_channel_mean: bad args");
  int64_t n = rows * C;
  bool v = (C % 4 == 0) && aligned16(du) && aligned16(dz);
  cudaStream_t st = (cudaStream_t)stream;
  float inv = 1.0f / (float)rows;
  DISPATCH_2(dudt, TA, dzdt, TB, {
    if (v) pdl_launch(sub_mean_kernel<TA, TB, 4>, grid_for(n / 4), 256, 0, (cudaStream_t)(st), (const TA*)du, (TB*)dz, n / 4, C, colsum, inv);
    else pdl_launch(sub_mean_kernel<TA, TB, 1>, grid_for(n), 256, 0, (cudaStream_t)(st), (const TA*)du, (TB*)dz, n, C, colsum, inv);
  });
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_bn_bwd(const void* dy, int dydt, const void* x, int xdt, void* dx, int dxdt, int64_t rows, int C,
                           const float* mean, const float* rstd, const float* gamma, float* dgamma, float* dbeta,
                           float beta_acc, float* ws, void* stream) {
  TGAN_CHECK_ARG(dy && x && dx && mean && rstd && gamma && ws && rows > 0, "bn_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  // ws layout: [partials: 2*MAX_PARTS*C][s1: C][s2: C]
  float* s1 = ws + (int64_t)4 * TGAN_STATS_MAX_PARTS * C;
  float* s2 = s1 + C;
  DISPATCH_2(dydt, TDY, xdt, TX, {
    BnBwdStatsF<TDY, TX, 1> f1{(const TDY*)dy, (const TX*)x, mean, rstd, C};
    BnBwdStatsF<TDY, TX, 4> f4{(const TDY*)dy, (const TX*)x, mean, rstd, C};
    int rc = run_colreduce<2>(f1, f4, C % 4 == 0 && aligned16(dy) && aligned16(x), rows, C, s1, s2, 0.f, ws, st);
    if (rc) return rc;
  });
  int64_t n = rows * C;
  TGAN_DISPATCH_1(dydt, TDY, TGAN_DISPATCH_1(xdt, TX, TGAN_DISPATCH_1(dxdt, TDX, {
    pdl_launch(bn_bwd_apply_kernel<TDY, TX, TDX>, grid_for(n), 256, 0, (cudaStream_t)(st), (const TDY*)dy, (const TX*)x, (TDX*)dx, n, C, mean,
                                                                   rstd, gamma, s1, s2, 1.0f / (float)rows);
  })));
  TGAN_LAUNCHED();
  if (dbeta) { pdl_launch(add_to_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)(st), dbeta, s1, C, beta_acc); TGAN_LAUNCHED(); }
  if (dgamma) { pdl_launch(add_to_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)(st), dgamma, s2, C, beta_acc); TGAN_LAUNCHED(); }
  return 0;
}

extern "C" int tgan_add_noise(const void* x, int xdt, void* y, int ydt, int64_t n, float std, const float* noise,
                              uint64_t seed, uint64_t stream_id, const uint64_t* counter, void* stream) {
  TGAN_CHECK_ARG(x && y && n > 0, "add_noise: bad args");
  int64_t q = (n + 3) / 4;
  DISPATCH_2(xdt, TX, ydt, TY, (pdl_launch(add_noise_kernel<TX, TY>, ceil_div(q, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const TX*)x, (TY*)y, n, std, noise, seed, stream_id, counter)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_dropout(const void* x, int xdt, void* y, int ydt, uint8_t* mask, int64_t n, float rate, int gen,
                            uint64_t seed, uint64_t stream_id, const uint64_t* counter, void* stream) {
  TGAN_CHECK_ARG(x && y && mask && n > 0 && rate >= 0.f && rate < 1.f, "dropout: bad args");
  int64_t q = (n + 3) / 4;
  DISPATCH_2(xdt, TX, ydt, TY, (pdl_launch(dropout_kernel<TX, TY>, ceil_div(q, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const TX*)x, (TY*)y, mask, n, rate, 1.0f / (1.0f - rate), gen, seed, stream_id,
                                   counter)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_maxpool2_fwd(const void* x, int dt, void* y, uint8_t* idx, int N, int H, int W, int C,
                                 void* stream) {
  TGAN_CHECK_ARG(x && y && idx && H % 2 == 0 && W % 2 == 0, "maxpool2_fwd: bad args (even extents only)");
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
  TGAN_DISPATCH_1(dt, T, (pdl_launch(maxpool2_fwd_kernel<T>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const T*)x, (T*)y, idx, N, H, W, C, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_maxpool2_bwd(const void* dy, int dt, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                                 void* stream) {
  TGAN_CHECK_ARG(dy && dx && idx && H % 2 == 0 && W % 2 == 0, "maxpool2_bwd: bad args");
  int64_t total = (int64_t)N * (H / 2) * (W / 2) * C;
  TGAN_DISPATCH_1(dt, T, (pdl_launch(maxpool2_bwd_kernel<T>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const T*)dy, idx, (T*)dx, N, H, W, C, total)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_maxpool2_dropout_fwd(const void* x, void* y, uint8_t* code, int N, int H, int W, int C, float rate,
                                         const uint8_t* mask, uint64_t seed, uint64_t stream_id,
                                         const uint64_t* counter, void* stream) {
  TGAN_CHECK_ARG(x && y && code && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && rate >= 0.f && rate < 1.f && aligned16(x) &&
                     aligned16(y) && ((uintptr_t)code & 7) == 0,
                 "maxpool2_dropout_fwd: bf16, even extents, C %% 8 == 0, aligned buffers");
  const int64_t nvec = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  pdl_launch(maxpool2_dropout_fwd_kernel, ceil_div(nvec, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const bf16*)x, (bf16*)y, code, H, W, C, nvec, rate, 1.0f / (1.0f - rate), mask, seed, stream_id, counter);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_maxpool2_dropout_bwd(const void* dy, const uint8_t* code, void* dx, int N, int H, int W, int C,
                                         float rate, void* stream) {
  TGAN_CHECK_ARG(dy && dx && code && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && aligned16(dy) && aligned16(dx),
                 "maxpool2_dropout_bwd: bad args");
  const int64_t nvec = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  pdl_launch(maxpool2_dropout_bwd_kernel, ceil_div(nvec, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const bf16*)dy, code, (bf16*)dx, H, W, C, nvec, 1.0f / (1.0f - rate));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_mobn_pool_dropout_fwd(const void* z, void* y, uint8_t* code, int N, int H, int W, int C, int nseg,
                                          int64_t n0, int64_t n1, int64_t n2, const void* sums, int sums_q24,
                                          const float* b, float* pop_mean, float decay, int train, int act, float alpha,
                                          float rate,
                                          const uint8_t* mask, uint64_t seed, uint64_t stream_id, const uint64_t* counter,
                                          void* stream) {
  TGAN_CHECK_ARG(z && y && code && b && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && rate >= 0.f && rate < 1.f &&
                     aligned16(z) && aligned16(y) && aligned16(b) && ((uintptr_t)code & 7) == 0,
                 "mobn_pool_dropout_fwd: bf16, even extents, C %% 8 == 0, aligned buffers");
  TGAN_CHECK_ARG(train ? (sums != nullptr && aligned16(sums)) : (pop_mean != nullptr && aligned16(pop_mean)),
                 "mobn_pool_dropout_fwd: needs sums (train) or pop_mean (test)");
  TGAN_CHECK_ARG(act == TGAN_ACT_NONE || act == TGAN_ACT_RELU || act == TGAN_ACT_LRELU || act == TGAN_ACT_TANH ||
                     act == TGAN_ACT_SIGMOID || act == TGAN_ACT_SOFTPLUS, "mobn_pool_dropout_fwd: monotonic activations only");
  const int rpi = H * W;
  Segs sg;      // segment boundaries are given in images; Segs works on full-resolution rows
  if (make_segs(sg, (int64_t)N * rpi, nseg, n0 * rpi, n1 * rpi, n2 * rpi)) return 1;
  const int64_t nvec = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  TGAN_CHECK_ARG(nvec < (1ll << 31), "mobn_pool_dropout_fwd: more than 2^31 output vectors");
  cudaStream_t st = (cudaStream_t)stream;
  TGAN_DISPATCH_ACT(act, A, (pdl_launch(mobn_pool_dropout_fwd_kernel<A>, ceil_div(nvec, 256), 256, 0, st, (const bf16*)z,
                                        (bf16*)y, code, H, W, C, nvec, sums, sums_q24, sg, rpi, b, pop_mean, decay, train, alpha, rate,
                                        1.0f / (1.0f - rate), mask, seed, stream_id, counter)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_mobn_pool_dropout_bwd(const void* dy, const void* y, const uint8_t* code, void* du, int N, int H, int W,
                                          int C, int nseg, int64_t n0, int64_t n1, int64_t n2, int act, float alpha,
                                          float rate, float* colsums, float* grad_acc, float* ws, void* stream) {
  TGAN_CHECK_ARG(dy && y && code && du && colsums && ws && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && 2048 % C == 0 &&
                     C >= 64 && aligned16(dy) && aligned16(y) && aligned16(du),
                 "mobn_pool_dropout_bwd: bf16, even extents, C in {64,128,256,512,1024,2048}, aligned buffers");
  const int rpi = H * W;
  Segs sg;
  if (make_segs(sg, (int64_t)N * rpi, nseg, n0 * rpi, n1 * rpi, n2 * rpi)) return 1;
  const int nparts = TGAN_ACT_BWD_SEG_PARTS;
  const size_t smem = (size_t)(2048 / C) * 4 * C * sizeof(float);
  const int64_t prows = (int64_t)N * (H / 2) * (W / 2);
  TGAN_CHECK_ARG(prows < (1ll << 31), "mobn_pool_dropout_bwd: more than 2^31 pooled pixels");
  cudaStream_t st = (cudaStream_t)stream;
  TGAN_DISPATCH_ACT(act, A, (pdl_launch(mobn_pool_dropout_bwd_kernel<A>, nparts, 256, smem, st, (const bf16*)dy,
                                        (const bf16*)y, code, (bf16*)du, H, W, C, prows, sg, rpi, alpha, 1.0f / (1.0f - rate),
                                        ws)));
  TGAN_LAUNCHED();
  pdl_launch(act_bwd_seg_fold_kernel, ceil_div(C, 8), 256, 0, st, ws, nparts, C, colsums, grad_acc, 4);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_global_pool_fwd(const void* x, int xdt, void* y, int ydt, uint8_t* idx, int N, int HW, int C,
                                    int mode, void* stream) {
  TGAN_CHECK_ARG(x && y && HW > 0 && HW <= 256, "global_pool_fwd: bad args");
  TGAN_CHECK_ARG(mode == 1 || idx, "global_pool_fwd: max mode needs idx");
  DISPATCH_2(xdt, TX, ydt, TY, (pdl_launch(global_pool_fwd_kernel<TX, TY>, ceil_div((int64_t)N * C, 128), 128, 0, (cudaStream_t)((cudaStream_t)stream), (const TX*)x, (TY*)y, idx, N,
                                                                                         HW, C, mode)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_global_pool_bwd(const void* dy, int dydt, const uint8_t* idx, void* dx, int dxdt, int N, int HW,
                                    int C, int mode, void* stream) {
  TGAN_CHECK_ARG(dy && dx, "global_pool_bwd: bad args");
  int64_t total = (int64_t)N * HW * C;
  if (mode != 0 && dydt == TGAN_BF16 && dxdt == TGAN_BF16 && C % 8 == 0 && aligned16(dy) && aligned16(dx) &&
      total / 8 < ((int64_t)1 << 31)) {
    pdl_launch(global_pool_bwd_mean_v8_kernel, ceil_div(total / 8, 256), 256, 0, (cudaStream_t)stream, (const bf16*)dy, (bf16*)dx,
               (unsigned)HW, (unsigned)(C / 8), (unsigned)(total / 8), 1.0f / (float)HW);
    TGAN_LAUNCHED();
    return 0;
  }
  DISPATCH_2(dydt, TDY, dxdt, TDX, (pdl_launch(global_pool_bwd_kernel<TDY, TDX>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const TDY*)dy, idx, (TDX*)dx, N, HW, C, mode, total)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_concat_label(const void* x, int xdt, int64_t rows, int C, int ldx, const float* lab, int K,
                                 int rows_per_sample, void* out, int odt, int ldo, void* stream) {
  TGAN_CHECK_ARG(x && lab && out && ldo >= C + K && rows_per_sample > 0, "concat_label: bad args");
  int64_t total = rows * ldo;
  DISPATCH_2(xdt, TX, odt, TO, (pdl_launch(concat_label_kernel<TX, TO>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const TX*)x, rows, C, ldx, lab, K, rows_per_sample, (TO*)out, ldo, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_dropout_concat(const void* x, int xdt, void* y, int ydt, int ldy, uint8_t* mask, int64_t rows, int C,
                                   float rate, int gen, uint64_t seed, uint64_t stream_id, const uint64_t* counter,
                                   const float* lab, int K, int rows_per_sample, void* stream) {
  TGAN_CHECK_ARG(x && y && mask && lab && rows > 0 && ldy >= C + K && rate >= 0.f && rate < 1.f && rows_per_sample > 0,
                 "dropout_concat: bad args");
  const int64_t n = rows * C, ngroups = (n + 3) / 4, total = ngroups + rows * (ldy - C);
  if (ydt == TGAN_BF16 && (ldy == 8 || ldy == 16) && C <= 8 && aligned16(y)) {
    TGAN_DISPATCH_1(xdt, TX, (pdl_launch(dropout_concat_narrow_kernel<TX>, ceil_div(rows, 256), 256, 0, (cudaStream_t)stream,
                                         (const TX*)x, (bf16*)y, mask, rows, C, ldy, rate, 1.0f / (1.0f - rate), gen, seed, stream_id,
                                         counter, lab, K, rows_per_sample)));
    TGAN_LAUNCHED();
    return 0;
  }
  DISPATCH_2(xdt, TX, ydt, TY, (pdl_launch(dropout_concat_kernel<TX, TY>, ceil_div(total, 256), 256, 0, (cudaStream_t)stream,
                                           (const TX*)x, (TY*)y, mask, n, C, ldy, rate, 1.0f / (1.0f - rate), gen, seed,
                                           stream_id, counter, lab, K, rows_per_sample, ngroups, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_affine_concat(const void* x, int xdt, void* y, int ydt, int ldy, int64_t rows, int C, const float* scale,
                                  const float* shift, const float* lab, int K, int rows_per_sample, void* stream) {
  TGAN_CHECK_ARG(x && y && lab && rows > 0 && ldy >= C + K && rows_per_sample > 0, "affine_concat: bad args");
  const int64_t total = rows * ldy;
  if (xdt == TGAN_BF16 && ydt == TGAN_BF16 && C % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) &&
      (!scale || aligned16(scale)) && (!shift || aligned16(shift))) {
    pdl_launch(affine_concat_v8_kernel, grid_for(total / 8), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, C, ldy, scale, shift,
               lab, K, rows_per_sample, total / 8);
    TGAN_LAUNCHED();
    return 0;
  }
  DISPATCH_2(xdt, TX, ydt, TY, (pdl_launch(affine_concat_kernel<TX, TY>, ceil_div(total, 256), 256, 0, (cudaStream_t)stream,
                                           (const TX*)x, (TY*)y, C, ldy, scale, shift, lab, K, rows_per_sample, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_fill_label(const float* lab, int K, int rows_per_sample, void* out, int odt, int64_t rows, int C,
                               int ldo, void* stream) {
  TGAN_CHECK_ARG(lab && out && ldo >= C + K && rows_per_sample > 0 && rows > 0, "fill_label: bad args");
  if (odt == TGAN_BF16 && C % 8 == 0 && ldo % 8 == 0 && aligned16(out)) {
    const int64_t tv = rows * ((ldo - C) / 8);
    pdl_launch(fill_label_v8_kernel, ceil_div(tv, 256), 256, 0, (cudaStream_t)stream, lab, K, rows_per_sample, (bf16*)out, C,
               ldo, tv);
    TGAN_LAUNCHED();
    return 0;
  }
  int64_t total = rows * (ldo - C);
  TGAN_DISPATCH_1(odt, TO, (pdl_launch(fill_label_kernel<TO>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), lab, K, rows_per_sample, (TO*)out, C, ldo, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_copy_channels(const void* src, int sdt, int lds, void* dst, int ddt, int ldd, int64_t rows, int C,
                                  void* stream) {
  TGAN_CHECK_ARG(src && dst && lds >= C && ldd >= C, "copy_channels: bad args");
  int64_t total = rows * C;
  DISPATCH_2(sdt, TS, ddt, TD, (pdl_launch(copy_channels_kernel<TS, TD>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const TS*)src, lds, (TD*)dst, ldd, C, total)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_gather_rows(const void* const* srcs, const int* dts, const int64_t* counts, int n, void* dst, int ddt,
                                void* stream) {
  TGAN_CHECK_ARG(srcs && dts && counts && dst && n >= 1 && n <= 8, "gather_rows: 1..8 sources");
  GatherSrcs g;
  memset(&g, 0, sizeof(g));
  int64_t tot = 0;
  for (int i = 0; i < n; ++i) {
    TGAN_CHECK_ARG(srcs[i] && counts[i] > 0 && (dts[i] == TGAN_F32 || dts[i] == TGAN_BF16), "gather_rows: bad source %d", i);
    tot += counts[i];
    g.src[i] = srcs[i]; g.end[i] = tot; g.dt[i] = dts[i];
  }
  g.n = n;
  TGAN_DISPATCH_1(ddt, TD, (pdl_launch(gather_rows_kernel<TD>, grid_for(tot), 256, 0, (cudaStream_t)stream, g, (TD*)dst, tot)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_sum_slices_bf16(const void* x, void* y, int64_t n, int parts, void* stream) {
  TGAN_CHECK_ARG(x && y && n > 0 && n % 8 == 0 && parts >= 1 && aligned16(x) && aligned16(y), "sum_slices_bf16: bad args");
  pdl_launch(sum_slices_bf16_kernel, grid_for(n / 8), 256, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)y, n / 8, parts);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_accumulate(void* y, const void* x, int dt, int64_t n, void* stream) {
  TGAN_CHECK_ARG(y && x && n > 0, "accumulate: bad args");
  TGAN_DISPATCH_1(dt, T, (pdl_launch(accumulate_kernel<T>, grid_for(n), 256, 0, (cudaStream_t)((cudaStream_t)stream), (T*)y, (const T*)x, n)));
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_fill_f32(float* p, float v, int64_t n, void* stream) {
  TGAN_CHECK_ARG(p && n > 0, "fill: bad args");
  pdl_launch(fill_kernel, grid_for(n), 256, 0, (cudaStream_t)((cudaStream_t)stream), p, v, n);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_argmax_onehot(const float* logits, int N, int K, int64_t* idx, float* onehot, void* stream) {
  TGAN_CHECK_ARG(logits && N > 0 && K > 0, "argmax_onehot: bad args");
  pdl_launch(argmax_onehot_kernel, ceil_div(N, 128), 128, 0, (cudaStream_t)((cudaStream_t)stream), logits, N, K, idx, onehot);
  TGAN_LAUNCHED();
  return 0;
}
