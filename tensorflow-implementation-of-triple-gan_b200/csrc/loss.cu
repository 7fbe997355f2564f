// loss.cu -- the three Triple-GAN losses with their closed-form dlogits, one fused kernel each
// (Training/train_base.py:113-154, :43-57, :75-84; gradients: SURVEY.md Appendix B).
// Logit tensors are tiny ([N,1] / [N,10], N <= a few hundred) so each loss is ONE single-CTA kernel:
// warp-shuffle + shared-memory tree reductions, deterministic summation order, loss scalar left on device.
#include "common.cuh"

namespace tgan {

constexpr int LT = 256;   // threads
constexpr int MAXK = 16;  // max classes held in registers

__device__ float block_sum(float v, float* sm) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LT / 32; ++i) s += sm[i];
  return s;
}

// N block sums behind ONE pair of barriers (loss_c needs 5 scalars + K class means: fifteen separate block_sum calls were
// thirty barriers, most of the kernel's 10 us); same fixed order as block_sum
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float (*sm)[LT / 32]) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) sm[k][w] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LT / 32; ++i) s += sm[k][i];
    v[k] = s;
  }
}

// tf.nn.sigmoid_cross_entropy_with_logits
__device__ __forceinline__ float sig_ce(float x, float z) { return fmaxf(x, 0.f) - x * z + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(LT) loss_d_kernel(const float* dr, int nr, const float* df, int nf, const float* du,
                                                    int nu, float* loss, float* g_dr, float* g_df, float* g_du) {
  pdl_entry();
  __shared__ float sm[32];
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < nr; i += LT) { float x = dr[i]; a += sig_ce(x, 1.f); if (g_dr) g_dr[i] = (sigmoidf_(x) - 1.f) / nr; }
  for (int i = threadIdx.x; i < nf; i += LT) { float x = df[i]; b += sig_ce(x, 0.f); if (g_df) g_df[i] = 0.5f * sigmoidf_(x) / nf; }
  for (int i = threadIdx.x; i < nu; i += LT) { float x = du[i]; c += sig_ce(x, 0.f); if (g_du) g_du[i] = 0.5f * sigmoidf_(x) / nu; }
  a = block_sum(a, sm); b = block_sum(b, sm); c = block_sum(c, sm);
  if (threadIdx.x == 0) *loss = a / nr + 0.5f * b / nf + 0.5f * c / nu;
}

__global__ void __launch_bounds__(LT) loss_g_kernel(const float* df, int nf, float* loss, float* g_df) {
  pdl_entry();
  __shared__ float sm[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < nf; i += LT) { float x = df[i]; a += sig_ce(x, 1.f); if (g_df) g_df[i] = 0.5f * (sigmoidf_(x) - 1.f) / nf; }
  a = block_sum(a, sm);
  if (threadIdx.x == 0) *loss = 0.5f * a / nf;
}

// softmax of one row into registers; returns log-sum-exp and the max probability.  KT > 0 = compile-time class count: the
// loops unroll and p[] stays in registers (with a runtime K every p[k] is a local-memory access).
template <int KT>
__device__ __forceinline__ float row_softmax(const float* x, int Krt, float (&p)[MAXK], int& am, float& pmax) {
  const int K = KT > 0 ? KT : Krt;
  float m = x[0]; am = 0;
#pragma unroll
  for (int k = 1; k < K; ++k) if (x[k] > m) { m = x[k]; am = k; }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = expf(x[k] - m); s += p[k]; }
  float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < K; ++k) p[k] *= inv;
  pmax = inv;                       // p[am] = exp(0) * inv
  return m + logf(s);
}

template <int KT>
__global__ void __launch_bounds__(LT)
loss_c_kernel(const float* __restrict__ c_real, const float* __restrict__ y_l_c, int n_real,
              const float* __restrict__ c_unl, const float* __restrict__ c_rep, const float* __restrict__ d_unl,
              int n_unl, const float* __restrict__ c_fake, const float* __restrict__ y_g, int n_fake, int Krt,
              const float* __restrict__ lambdas, float* loss, float* g_real, float* g_unl, float* g_rep,
              float* g_fake) {
  pdl_entry();
  const int K = KT > 0 ? KT : Krt;
  __shared__ float smn[5 + MAXK][LT / 32];
  __shared__ float q[MAXK];
  const float lambda_1 = lambdas[0], lambda_2 = lambdas[1];
  float p[MAXK];
  int am;
  float pam;
  // ---- supervised CE on labelled and generated images (train_base.py:130-131) ----
  float l_real = 0.f, l_fake = 0.f;
  for (int n = threadIdx.x; n < n_real; n += LT) {
    float lse = row_softmax<KT>(c_real + n * K, K, p, am, pam);
    float ys = 0.f, l = 0.f;
    _Pragma("unroll") for (int k = 0; k < K; ++k) { float y = y_l_c[n * K + k]; ys += y; l -= y * (c_real[n * K + k] - lse); }
    l_real += l;
    if (g_real) for (int k = 0; k < K; ++k) g_real[n * K + k] = (p[k] * ys - y_l_c[n * K + k]) / n_real;
  }
  for (int n = threadIdx.x; n < n_fake; n += LT) {
    float lse = row_softmax<KT>(c_fake + n * K, K, p, am, pam);
    float ys = 0.f, l = 0.f;
    _Pragma("unroll") for (int k = 0; k < K; ++k) { float y = y_g[n * K + k]; ys += y; l -= y * (c_fake[n * K + k] - lse); }
    l_fake += l;
    if (g_fake) for (int k = 0; k < K; ++k) g_fake[n * K + k] = lambda_1 * (p[k] * ys - y_g[n * K + k]) / n_fake;
  }
  // ---- unlabelled terms, pass 1: scalars + batch-mean softmax q (train_base.py:134-143, :52) ----
  float l_unl = 0.f, l_ent = 0.f, l_mse = 0.f;
  float qk[MAXK];
  for (int k = 0; k < MAXK; ++k) qk[k] = 0.f;
  for (int n = threadIdx.x; n < n_unl; n += LT) {
    const float* x = c_unl + n * K;
    float lse = row_softmax<KT>(x, K, p, am, pam);
    float s = sig_ce(d_unl[n], 1.f);
    l_unl += pam * s;
    float px = 0.f;
    _Pragma("unroll") for (int k = 0; k < K; ++k) { px += p[k] * x[k]; qk[k] += p[k]; }
    l_ent += lse - px;
    if (c_rep) for (int k = 0; k < K; ++k) { float d = x[k] - c_rep[n * K + k]; l_mse += d * d; }
  }
  {
    float red[5 + MAXK];
    red[0] = l_real; red[1] = l_fake; red[2] = l_unl; red[3] = l_ent; red[4] = l_mse;
    _Pragma("unroll") for (int k = 0; k < MAXK; ++k) red[5 + k] = qk[k];
    block_sum_n<5 + MAXK>(red, smn);
    l_real = red[0]; l_fake = red[1]; l_unl = red[2]; l_ent = red[3]; l_mse = red[4];
    if (threadIdx.x == 0) {
      _Pragma("unroll") for (int k = 0; k < MAXK; ++k) if (k < K) q[k] = red[5 + k] / n_unl;
    }
  }
  __syncthreads();
  float l_bal = 0.f;
  float r[MAXK];
  _Pragma("unroll") for (int k = 0; k < K; ++k) { r[k] = 1.f / (q[k] + 1e-12f); l_bal -= logf(q[k] + 1e-12f) / K; }
  if (threadIdx.x == 0) {
    float c_real_tot = l_real / n_real + 1e-6f * (l_ent / n_unl) + 1e-3f * l_bal;
    float v = 0.01f * 0.5f * (l_unl / n_unl) + c_real_tot + lambda_1 * (l_fake / n_fake);
    if (c_rep) v += lambda_2 * l_mse / (float)(n_unl * K);
    *loss = v;
  }
  // ---- pass 2: dlogits of the unlabelled terms ----
  if (g_unl) {
    for (int n = threadIdx.x; n < n_unl; n += LT) {
      const float* x = c_unl + n * K;
      row_softmax<KT>(x, K, p, am, pam);
      float s = sig_ce(d_unl[n], 1.f);
      float px = 0.f, pr = 0.f;
      _Pragma("unroll") for (int k = 0; k < K; ++k) { px += p[k] * x[k]; pr += p[k] * r[k]; }
      _Pragma("unroll") for (int k = 0; k < K; ++k) {
        float g = 0.005f * (s / n_unl) * pam * ((k == am ? 1.f : 0.f) - p[k]);
        g += 1e-6f * (-p[k] * (x[k] - px) / n_unl);
        g += 1e-3f * (-(1.f / (K * (float)n_unl)) * p[k] * (r[k] - pr));
        if (c_rep) {
          float d = lambda_2 * 2.f * (x[k] - c_rep[n * K + k]) / (float)(n_unl * K);
          g += d;
          if (g_rep) g_rep[n * K + k] = -d;
        }
        g_unl[n * K + k] = g;
      }
    }
  }
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_loss_d(const float* dr, int nr, const float* df, int nf, const float* du, int nu, float* loss,
                           float* g_dr, float* g_df, float* g_du, void* stream) {
  TGAN_CHECK_ARG(dr && df && du && loss && nr > 0 && nf > 0 && nu > 0, "loss_d: bad args");
  pdl_launch(loss_d_kernel, 1, LT, 0, (cudaStream_t)((cudaStream_t)stream), dr, nr, df, nf, du, nu, loss, g_dr, g_df, g_du);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_loss_g(const float* df, int nf, float* loss, float* g_df, void* stream) {
  TGAN_CHECK_ARG(df && loss && nf > 0, "loss_g: bad args");
  pdl_launch(loss_g_kernel, 1, LT, 0, (cudaStream_t)((cudaStream_t)stream), df, nf, loss, g_df);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_loss_c(const float* c_real, const float* y_l_c, int n_real, const float* c_unl, const float* c_rep,
                           const float* d_unl_logits, int n_unl, const float* c_fake, const float* y_g, int n_fake,
                           int K, const float* lambdas, float* loss, float* g_real, float* g_unl, float* g_rep,
                           float* g_fake, void* stream) {
  TGAN_CHECK_ARG(c_real && y_l_c && c_unl && d_unl_logits && c_fake && y_g && lambdas && loss, "loss_c: null pointer");
  TGAN_CHECK_ARG(K > 0 && K <= MAXK && n_real > 0 && n_unl > 0 && n_fake > 0, "loss_c: bad sizes (K <= %d)", MAXK);
  if (K == 10)
    pdl_launch(loss_c_kernel<10>, 1, LT, 0, (cudaStream_t)stream, c_real, y_l_c, n_real, c_unl, c_rep, d_unl_logits, n_unl, c_fake,
               y_g, n_fake, K, lambdas, loss, g_real, g_unl, g_rep, g_fake);
  else
    pdl_launch(loss_c_kernel<0>, 1, LT, 0, (cudaStream_t)stream, c_real, y_l_c, n_real, c_unl, c_rep, d_unl_logits, n_unl, c_fake,
               y_g, n_fake, K, lambdas, loss, g_real, g_unl, g_rep, g_fake);
  TGAN_LAUNCHED();
  return 0;
}
