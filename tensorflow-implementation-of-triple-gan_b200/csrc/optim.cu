// optim.cu -- weight-norm reparameterisation (fwd/bwd), fused Adam(+EMA) over flat parameter buffers,
// device-side step counters.  All HBM-bound, vectorised, no host round trips (graph capturable).
//   weight norm : nn.py:502,554; modle_base.py:66,101,148  (backward: SURVEY.md Appendix B)
//   Adam / EMA  : train_base.py:91-97; Train_goodGAN.py:85-103
#include "common.cuh"
#include "colreduce.cuh"

namespace tgan {

template <int VEC>
struct SumSqF {
  const float* v; int C;
  __device__ void operator()(int64_t r, int c0, float (&o)[1][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) { float t = v[r * C + c0 + j]; o[0][j] = t * t; }
  }
};
template <int VEC>
struct DotF {
  const float* a; const float* b; int C;
  __device__ void operator()(int64_t r, int c0, float (&o)[1][VEC]) const {
#pragma unroll
    for (int j = 0; j < VEC; ++j) o[0][j] = a[r * C + c0 + j] * b[r * C + c0 + j];
  }
};

__global__ void wn_finalize_kernel(const float* ss, const float* g, int Co, int eps_mode, float* inv_norm,
                                   float* scale) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Co) return;
  float s = ss[c];
  float inv = eps_mode ? rsqrtf(fmaxf(s, 1e-12f)) : 1.0f / sqrtf(s);
  inv_norm[c] = inv;
  scale[c] = g[c] * inv;
}
// W[a,co,b] = V[a,co,b] * scale[co]
__global__ void wn_scale_kernel(const float* __restrict__ V, const float* __restrict__ scale, float* __restrict__ W,
                                int64_t n, int Co, int B) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    W[i] = V[i] * scale[(i / B) % Co];
}
// generic (B > 1) sum of squares / dot: one CTA per output channel
__global__ void wn_reduce_generic_kernel(const float* __restrict__ a, const float* __restrict__ b, int A, int Co, int B,
                                         float* __restrict__ out) {
  pdl_entry();
  __shared__ float sm[32];
  int co = blockIdx.x;
  float s = 0.f;
  for (int e = threadIdx.x; e < A * B; e += blockDim.x) {
    int64_t idx = ((int64_t)(e / B) * Co + co) * B + (e % B);
    s += a[idx] * (b ? b[idx] : a[idx]);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)blockDim.x / 32; ++i) t += sm[i];
    out[co] = t;
  }
}
__global__ void wn_bwd_dg_kernel(const float* dot, const float* inv_norm, float* dg, int Co, float beta) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < Co) dg[c] = (beta != 0.f ? beta * dg[c] : 0.f) + dot[c] * inv_norm[c];
}
// dV = g*inv*(dW - V*inv^2*dot)
__global__ void wn_bwd_dv_kernel(const float* __restrict__ V, const float* __restrict__ g,
                                 const float* __restrict__ inv_norm, const float* __restrict__ dW,
                                 const float* __restrict__ dot, float* __restrict__ dV, int64_t n, int Co, int B,
                                 float beta) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)((i / B) % Co);
    float inv = inv_norm[c];
    float v = g[c] * inv * (dW[i] - V[i] * inv * inv * dot[c]);
    dV[i] = (beta != 0.f ? beta * dV[i] : 0.f) + v;
  }
}

__global__ void adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ grad, int64_t n, const float* __restrict__ state, float beta1,
                            float beta2, float eps, float gscale, float* __restrict__ ema, float ema_decay) {
  pdl_entry();
  const float lr = state[0], b1p = state[1], b2p = state[2];
  const float a = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 t = *reinterpret_cast<float4*>(theta + i), mm = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i), g = *reinterpret_cast<const float4*>(grad + i);
    float tt[4] = {t.x, t.y, t.z, t.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
    float ga[4] = {g.x * gscale, g.y * gscale, g.z * gscale, g.w * gscale};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ma[j] += (ga[j] - ma[j]) * (1.f - beta1);
      va[j] += (ga[j] * ga[j] - va[j]) * (1.f - beta2);
      tt[j] -= a * ma[j] / (sqrtf(va[j]) + eps);
    }
    *reinterpret_cast<float4*>(theta + i) = make_float4(tt[0], tt[1], tt[2], tt[3]);
    *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    if (ema) {
      float4 e = *reinterpret_cast<float4*>(ema + i);
      e.x -= (e.x - tt[0]) * (1.f - ema_decay); e.y -= (e.y - tt[1]) * (1.f - ema_decay);
      e.z -= (e.z - tt[2]) * (1.f - ema_decay); e.w -= (e.w - tt[3]) * (1.f - ema_decay);
      *reinterpret_cast<float4*>(ema + i) = e;
    }
  }
  // scalar tail (n % 4), handled by the first threads of the grid once
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t k = (n & ~(int64_t)3) + threadIdx.x;
    float g = grad[k] * gscale;
    float mk = m[k] + (g - m[k]) * (1.f - beta1);
    float vk = v[k] + (g * g - v[k]) * (1.f - beta2);
    float tk = theta[k] - a * mk / (sqrtf(vk) + eps);
    m[k] = mk; v[k] = vk; theta[k] = tk;
    if (ema) ema[k] -= (ema[k] - tk) * (1.f - ema_decay);
  }
}
__global__ void adam_advance_kernel(float* state, float beta1, float beta2) {
  pdl_entry();
  state[1] *= beta1;
  state[2] *= beta2;
}
__global__ void counter_advance_kernel(uint64_t* c, uint64_t inc) {
  pdl_entry(); *c += inc; }

static inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = 148 * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_weightnorm_fwd(const float* V, const float* g, float* W, float* inv_norm, float* scale, int A,
                                   int Co, int B, int eps_mode, float* ws, void* stream) {
  TGAN_CHECK_ARG(V && g && inv_norm && scale && ws && A > 0 && Co > 0 && B > 0, "weightnorm_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  float* ss = ws + (int64_t)4 * TGAN_STATS_MAX_PARTS * Co;
  if (B == 1) {
    SumSqF<1> f1{V, Co};
    SumSqF<4> f4{V, Co};
    int rc = run_colreduce<1>(f1, f4, Co % 4 == 0, A, Co, ss, nullptr, 0.f, ws, st);
    if (rc) return rc;
  } else {
    pdl_launch(wn_reduce_generic_kernel, Co, 256, 0, (cudaStream_t)(st), V, nullptr, A, Co, B, ss);
    TGAN_LAUNCHED();
  }
  pdl_launch(wn_finalize_kernel, ceil_div(Co, 128), 128, 0, (cudaStream_t)(st), ss, g, Co, eps_mode, inv_norm, scale);
  TGAN_LAUNCHED();
  if (W) {
    int64_t n = (int64_t)A * Co * B;
    pdl_launch(wn_scale_kernel, grid_for(n), 256, 0, (cudaStream_t)(st), V, scale, W, n, Co, B);
    TGAN_LAUNCHED();
  }
  return 0;
}

extern "C" int tgan_weightnorm_bwd(const float* V, const float* g, const float* inv_norm, const float* dW, float* dV,
                                   float* dg, int A, int Co, int B, float beta, float* ws, void* stream) {
  TGAN_CHECK_ARG(V && g && inv_norm && dW && dV && dg && ws, "weightnorm_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  float* dot = ws + (int64_t)4 * TGAN_STATS_MAX_PARTS * Co;
  if (B == 1) {
    DotF<1> f1{dW, V, Co};
    DotF<4> f4{dW, V, Co};
    int rc = run_colreduce<1>(f1, f4, Co % 4 == 0, A, Co, dot, nullptr, 0.f, ws, st);
    if (rc) return rc;
  } else {
    pdl_launch(wn_reduce_generic_kernel, Co, 256, 0, (cudaStream_t)(st), dW, V, A, Co, B, dot);
    TGAN_LAUNCHED();
  }
  pdl_launch(wn_bwd_dg_kernel, ceil_div(Co, 128), 128, 0, (cudaStream_t)(st), dot, inv_norm, dg, Co, beta);
  TGAN_LAUNCHED();
  int64_t n = (int64_t)A * Co * B;
  pdl_launch(wn_bwd_dv_kernel, grid_for(n), 256, 0, (cudaStream_t)(st), V, g, inv_norm, dW, dot, dV, n, Co, B, beta);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_adam(float* theta, float* m, float* v, const float* grad, int64_t n, const float* state,
                         float beta1, float beta2, float eps, float grad_scale, float* ema, float ema_decay,
                         void* stream) {
  TGAN_CHECK_ARG(theta && m && v && grad && state && n > 0, "adam: bad args");
  TGAN_CHECK_ARG(((uintptr_t)theta & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0 &&
                     ((uintptr_t)grad & 15) == 0 && (!ema || ((uintptr_t)ema & 15) == 0),
                 "adam: buffers must be 16-byte aligned");
  pdl_launch(adam_kernel, grid_for((n + 3) / 4), 256, 0, (cudaStream_t)((cudaStream_t)stream), theta, m, v, grad, n, state, beta1, beta2, eps,
                                                                       grad_scale, ema, ema_decay);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_adam_advance(float* state, float beta1, float beta2, void* stream) {
  TGAN_CHECK_ARG(state, "adam_advance: null state");
  pdl_launch(adam_advance_kernel, 1, 1, 0, (cudaStream_t)((cudaStream_t)stream), state, beta1, beta2);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int tgan_counter_advance(uint64_t* counter, uint64_t inc, void* stream) {
  TGAN_CHECK_ARG(counter, "counter_advance: null counter");
  pdl_launch(counter_advance_kernel, 1, 1, 0, (cudaStream_t)((cudaStream_t)stream), counter, inc);
  TGAN_LAUNCHED();
  return 0;
}
