// gemm_simt.cu -- fp32 CUDA-core GEMM + im2col / col2im.
// The fp32 parity path of every dense contraction (tf.nn.conv2d nn.py:504, tf.nn.conv2d_transpose
// modle_base.py:149,250, tf.matmul nn.py:553) and the production path of the skinny layers
// (Cout in {1,3,10}) where a 128-wide UMMA tile would be >90% padding.
#include "common.cuh"

namespace tgan {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A,
                                                    int lda, const float* __restrict__ B, int ldb, float beta,
                                                    float* __restrict__ C, int ldc, int kper, float* __restrict__ ws) {
  pdl_entry();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (!TA) { k = tid % 16; m = tid / 16 + i * 16; } else { m = tid % 64; k = tid / 64 + i * 4; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk];
      As[k][m] = v;
      int n, kb;
      if (!TB) { n = tid % 64; kb = tid / 64 + i * 4; } else { kb = tid % 16; n = tid / 16 + i * 16; }
      int gn = n0 + n, gkb = k0 + kb;
      float w = 0.f;
      if (gn < N && gkb < kend) w = TB ? B[(int64_t)gn * ldb + gkb] : B[(int64_t)gkb * ldb + gn];
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
      float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (ws) {
        ws[((int64_t)blockIdx.z * M + gm) * N + gn] = acc[i][j];
      } else {
        float* c = C + (int64_t)gm * ldc + gn;
        *c = alpha * acc[i][j] + (beta != 0.f ? beta * (*c) : 0.f);
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float alpha, float beta,
                                     float* __restrict__ C, int ldc) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[(int64_t)z * M * N + i];
  int m = (int)(i / N), n = (int)(i % N);
  float* c = C + (int64_t)m * ldc + n;
  *c = alpha * s + (beta != 0.f ? beta * (*c) : 0.f);
}

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, int N, int H, int W, int C, int ldx, int kh, int kw, int sh,
                              int sw, int pt, int pl, int Ho, int Wo, float* __restrict__ col, int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int64_t t = i / C;
  int s = (int)(t % kw); t /= kw;
  int r = (int)(t % kh); t /= kh;
  int wo = (int)(t % Wo); t /= Wo;
  int ho = (int)(t % Ho);
  int n = (int)(t / Ho);
  int hi = ho * sh + r - pt, wi = wo * sw + s - pl;
  float v = 0.f;
  if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = ldf<T>(x, ((int64_t)(n * H + hi) * W + wi) * ldx + c);
  col[i] = v;
}

// bf16 variant feeding the tensor-core GEMM of small-Cin convolutions: col[row, k] for k = (r*kw + s)*C + c < K, zero for
// K <= k < ldc (row stride padded to a multiple of 8 elements for TMA)
// 128 output pixels per CTA: thread = pixel gathers its kh*kw*C window into a shared-memory row (odd word pitch: no bank
// conflicts), then the CTA streams the [128, ldc] tile to global memory with 16-byte coalesced stores.
template <typename T>
__global__ void __launch_bounds__(128) im2col_bf16_kernel(const T* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                          int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                                                          bf16* __restrict__ col, int ldc, int64_t rows) {
  pdl_entry();
  extern __shared__ uint32_t im2col_sm[];
  const int pitch = ldc / 2 + 1;                       // 32-bit words per row, odd (ldc % 8 == 0)
  bf16* srow = reinterpret_cast<bf16*>(im2col_sm + (size_t)threadIdx.x * pitch);
  const int64_t row0 = (int64_t)blockIdx.x * 128, row = row0 + threadIdx.x;
  const int K = kh * kw * C;
  if (row < rows) {
    const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho), n = (int)(row / ((int64_t)Wo * Ho));
    int k = 0;
    for (int r = 0; r < kh; ++r) {
      const int hi = ho * sh + r - pt;
      for (int s2 = 0; s2 < kw; ++s2) {
        const int wi = wo * sw + s2 - pl;
        const bool in = hi >= 0 && hi < H && wi >= 0 && wi < W;
        const T* px = x + ((int64_t)(n * H + hi) * W + wi) * ldx;
        for (int c = 0; c < C; ++c, ++k) srow[k] = __float2bfloat16_rn(in ? ldf<T>(px, c) : 0.f);
      }
    }
    for (; k < ldc; ++k) srow[k] = __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int cpr = ldc / 8;                             // 16-byte chunks per row
  const int64_t nrows = rows - row0 < 128 ? rows - row0 : 128;
  for (int j = threadIdx.x; j < (int)nrows * cpr; j += 128) {
    const int rl = j / cpr, kk = (j - rl * cpr) * 4;   // word offset inside the row
    const uint32_t* sp = im2col_sm + (size_t)rl * pitch + kk;
    uint4 v = make_uint4(sp[0], sp[1], sp[2], sp[3]);
    *reinterpret_cast<uint4*>(col + (row0 + rl) * ldc + kk * 2) = v;
  }
}

// bf16 input whose pixels are 16-byte aligned (ldx % 8 == 0): work item = (tap, row); a pixel's channels arrive with 16-byte
// loads (consecutive threads -> consecutive pixels), the row is assembled in shared memory as above.  256 threads.
__global__ void __launch_bounds__(256) im2col_bf16_vec_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
                                                              int kh, int kw, int sh, int sw, int pt, int pl, int Ho, int Wo,
                                                              bf16* __restrict__ col, int ldc, int64_t rows) {
  pdl_entry();
  extern __shared__ uint32_t im2col_sm[];
  const int pitch = ldc / 2 + 1;
  const int64_t row0 = (int64_t)blockIdx.x * 128;
  const int nrows = (int)(rows - row0 < 128 ? rows - row0 : 128);
  const int taps = kh * kw, K = taps * C, nv = (C + 7) / 8;
  for (int item = threadIdx.x; item < taps * 128; item += 256) {
    const int t = item >> 7, rl = item & 127;
    if (rl >= nrows) continue;
    const int64_t row = row0 + rl;
    const int wo = (int)(row % Wo), ho = (int)((row / Wo) % Ho), n = (int)(row / ((int64_t)Wo * Ho));
    const int r = t / kw, s2 = t - r * kw;
    const int hi = ho * sh + r - pt, wi = wo * sw + s2 - pl;
    const bool in = hi >= 0 && hi < H && wi >= 0 && wi < W;
    bf16* dst = reinterpret_cast<bf16*>(im2col_sm + (size_t)rl * pitch) + t * C;
    const uint4* px = reinterpret_cast<const uint4*>(x + ((int64_t)(n * H + hi) * W + wi) * ldx);
    for (int v = 0; v < nv; ++v) {
      uint4 q = make_uint4(0, 0, 0, 0);
      if (in) q = px[v];
      const bf16* e = reinterpret_cast<const bf16*>(&q);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v * 8 + j < C) dst[v * 8 + j] = e[j];
    }
  }
  for (int i = threadIdx.x; i < nrows * (ldc - K); i += 256) {                 // zero the K..ldc tail of every row
    const int rl = i / (ldc - K), k = K + i - rl * (ldc - K);
    reinterpret_cast<bf16*>(im2col_sm + (size_t)rl * pitch)[k] = __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  const int cpr = ldc / 8;
  for (int j = threadIdx.x; j < nrows * cpr; j += 256) {
    const int rl = j / cpr, kk = (j - rl * cpr) * 4;
    const uint32_t* sp = im2col_sm + (size_t)rl * pitch + kk;
    *reinterpret_cast<uint4*>(col + (row0 + rl) * ldc + kk * 2) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
  }
}

template <typename T>
__global__ void col2im_kernel(const float* __restrict__ col, int N, int H, int W, int C, int kh, int kw, int sh,
                              int sw, int pt, int pl, int Ho, int Wo, T* __restrict__ x, int Cx, int ldx,
                              int64_t total) {
  pdl_entry();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % Cx);
  int64_t t = i / Cx;
  int w = (int)(t % W); t /= W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  float acc = 0.f;
  for (int r = 0; r < kh; ++r) {
    int hh = h + pt - r;
    if (hh < 0 || hh % sh) continue;
    int ho = hh / sh;
    if (ho >= Ho) continue;
    for (int s = 0; s < kw; ++s) {
      int ww = w + pl - s;
      if (ww < 0 || ww % sw) continue;
      int wo = ww / sw;
      if (wo >= Wo) continue;
      acc += col[(((int64_t)(n * Ho + ho) * Wo + wo) * (kh * kw) + r * kw + s) * C + c];
    }
  }
  stf<T>(x, ((int64_t)(n * H + h) * W + w) * ldx + c, acc);
}

}  // namespace tgan

namespace tgan {

// Skinny outputs (N <= 16: the discriminator head 138 -> 1, the classifier logits 128 -> 10; SURVEY 2.2 'warp-reduce
// GEMV for N in {1, 10}'): a 64x64-tile SIMT GEMM leaves 63 of 64 columns idle and took 11-19 us per call.
// C[m, :] = A[m, :] * B (+ beta * C): one warp per row, lanes stride K, N accumulators per lane, shuffle reduction.
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_nn_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, float beta,
                                                        float* __restrict__ C, int ldc) {
  pdl_entry();
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float a = A[(int64_t)m * lda + k];
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) acc[n] += a * B[(int64_t)k * ldb + n];
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0)
    for (int n = 0; n < N; ++n) C[(int64_t)m * ldc + n] = acc[n] + (beta != 0.f ? beta * C[(int64_t)m * ldc + n] : 0.f);
}

// C[i, :] = beta * C[i, :] + sum_r A[r, i] * B[r, :]  (the filter gradient of a skinny layer: A = layer input [rows, M],
// B = dlogits [rows, N]).  Thread (i, part): the row range is cut into PARTS fixed slices summed in a fixed order.
template <int NMAX>
__global__ void __launch_bounds__(512) skinny_tn_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                        const float* __restrict__ B, int ldb, float beta,
                                                        float* __restrict__ C, int ldc) {
  pdl_entry();
  constexpr int PARTS = 16;      // 512 threads: 32 output rows x 16 slices of the reduction axis
  __shared__ float red[PARTS][32][NMAX];
  const int il = threadIdx.x & 31, part = threadIdx.x >> 5, i = blockIdx.x * 32 + il;
  float acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
  const int per = (K + PARTS - 1) / PARTS, r0 = part * per, r1 = min(K, r0 + per);
  if (i < M)
#pragma unroll 4
    for (int r = r0; r < r1; ++r) {      // (independent loads: four rows in flight per thread)
      const float a = A[(int64_t)r * lda + i];
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) acc[n] += a * B[(int64_t)r * ldb + n];
    }
#pragma unroll
  for (int n = 0; n < NMAX; ++n) red[part][il][n] = acc[n];
  __syncthreads();
  if (part == 0 && i < M)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < PARTS; ++q) s += red[q][il][n];
      C[(int64_t)i * ldc + n] = s + (beta != 0.f ? beta * C[(int64_t)i * ldc + n] : 0.f);
    }
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                          const float* B, int ldb, float beta, float* C, int ldc, int splits, float* ws,
                          void* stream) {
  TGAN_CHECK_ARG(M > 0 && N > 0 && K > 0, "sgemm: empty problem %d %d %d", M, N, K);
  TGAN_CHECK_ARG(A && B && C, "sgemm: null pointer");
  if (N <= 16 && alpha == 1.f && !transB) {      // skinny outputs: dedicated kernels (fixed summation order, no split-K)
    cudaStream_t sst = (cudaStream_t)stream;
    if (!transA) pdl_launch(skinny_nn_kernel<16>, ceil_div((int64_t)M * 32, 256), 256, 0, sst, M, N, K, A, lda, B, ldb, beta, C, ldc);
    else pdl_launch(skinny_tn_kernel<16>, ceil_div(M, 32), 512, 0, sst, M, N, K, A, lda, B, ldb, beta, C, ldc);
    TGAN_LAUNCHED();
    return 0;
  }
  if (splits < 1) splits = 1;
  TGAN_CHECK_ARG(splits == 1 || ws, "sgemm: split-K needs a workspace");
  int kper = ((ceil_div(K, splits) + BK - 1) / BK) * BK;
  splits = ceil_div(K, kper);
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), splits);
  cudaStream_t st = (cudaStream_t)stream;
  float* w = splits > 1 ? ws : nullptr;
#define L(TA, TB) pdl_launch(sgemm_kernel<TA, TB>, grid, 256, 0, (cudaStream_t)(st), M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kper, w)
  if (!transA && !transB) L(false, false);
  else if (!transA && transB) L(false, true);
  else if (transA && !transB) L(true, false);
  else L(true, true);
#undef L
  TGAN_LAUNCHED();
  if (splits > 1) {
    int64_t tot = (int64_t)M * N;
    pdl_launch(splitk_reduce_kernel, ceil_div(tot, 256), 256, 0, (cudaStream_t)(st), ws, splits, M, N, alpha, beta, C, ldc);
    TGAN_LAUNCHED();
  }
  return 0;
}

extern "C" int tgan_im2col(const void* x, int xdt, int N, int H, int W, int C, int ldx, int kh, int kw, int sh,
                           int sw, int pt, int pl, int Ho, int Wo, float* col, void* stream) {
  TGAN_CHECK_ARG(x && col, "im2col: null pointer");
  int64_t total = (int64_t)N * Ho * Wo * kh * kw * C;
  TGAN_CHECK_ARG(total > 0, "im2col: empty");
  TGAN_DISPATCH_1(xdt, T, (pdl_launch(im2col_kernel<T>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), (const T*)x, N, H, W, C, ldx, kh, kw, sh, sw, pt, pl, Ho, Wo, col, total)));
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_im2col_bf16(const void* x, int xdt, int N, int H, int W, int C, int ldx, int kh, int kw, int sh,
                                int sw, int pt, int pl, int Ho, int Wo, void* col, int ldc, void* stream) {
  TGAN_CHECK_ARG(x && col && ldc >= kh * kw * C && ldc % 8 == 0, "im2col_bf16: bad args");
  const int64_t rows = (int64_t)N * Ho * Wo;
  TGAN_CHECK_ARG(rows > 0 && ldc <= 512 && ((uintptr_t)col & 15) == 0, "im2col_bf16: empty / ldc > 512 / unaligned col");
  const size_t smem = (size_t)128 * (ldc / 2 + 1) * 4;
  if (xdt == TGAN_BF16 && ldx % 8 == 0 && (C + 7) / 8 * 8 <= ldx && ((uintptr_t)x & 15) == 0) {
    static bool vattr = false;
    if (!vattr && smem > 48 * 1024) {
      cudaFuncSetAttribute(im2col_bf16_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
      vattr = true;
    }
    pdl_launch(im2col_bf16_vec_kernel, ceil_div(rows, 128), 256, smem, (cudaStream_t)stream, (const bf16*)x, N, H, W, C, ldx, kh, kw, sh,
               sw, pt, pl, Ho, Wo, (bf16*)col, ldc, rows);
    TGAN_LAUNCHED();
    return 0;
  }
  TGAN_DISPATCH_1(xdt, T, {
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
      cudaFuncSetAttribute(im2col_bf16_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
      attr = true;
    }
    pdl_launch(im2col_bf16_kernel<T>, ceil_div(rows, 128), 128, smem, (cudaStream_t)((cudaStream_t)stream), (const T*)x, N, H, W, C, ldx, kh, kw, sh, sw, pt, pl, Ho, Wo, (bf16*)col, ldc, rows);
  });
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_col2im(const float* col, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int pt,
                           int pl, int Ho, int Wo, void* x, int xdt, int Cx, int ldx, void* stream) {
  TGAN_CHECK_ARG(x && col, "col2im: null pointer");
  TGAN_CHECK_ARG(Cx <= C && Cx <= ldx, "col2im: bad channel slice");
  int64_t total = (int64_t)N * H * W * Cx;
  TGAN_CHECK_ARG(total > 0, "col2im: empty");
  TGAN_DISPATCH_1(xdt, T, (pdl_launch(col2im_kernel<T>, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), col, N, H, W, C, kh, kw, sh, sw, pt, pl, Ho, Wo, (T*)x, Cx, ldx, total)));
  TGAN_LAUNCHED();
  return 0;
}
