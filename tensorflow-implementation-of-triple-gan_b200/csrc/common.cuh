// common.cuh -- shared helpers for libtgan (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/tgan.h"

namespace tgan {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define TGAN_CHECK_ARG(cond, ...)      \
  do {                                 \
    if (!(cond)) {                     \
      tgan::set_error(__VA_ARGS__);    \
      return 1;                        \
    }                                  \
  } while (0)

// every kernel launch goes through this so that bench.py can report gpu_launches
#define TGAN_LAUNCHED()                                                  \
  do {                                                                   \
    tgan::g_launches.fetch_add(1, std::memory_order_relaxed);            \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      tgan::set_error("%s:%d launch failed: %s", __FILE__, __LINE__,     \
                      cudaGetErrorString(e__));                          \
      return 2;                                                          \
    }                                                                    \
  } while (0)

typedef __nv_bfloat16 bf16;

// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_entry(): it lets its successor begin launching (block scheduling, prologue) while this grid
// is still running, then waits until its own predecessor has completed and flushed before touching global memory.  The
// step is ~350 back-to-back launches, many of them a few microseconds long, so launch latency is a visible share of it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() { pdl_trigger(); pdl_wait(); }
extern int g_use_pdl;      // TGAN_NO_PDL=1 disables the attribute (plain stream order)
template <typename... KArgs, typename... Args>
inline cudaError_t pdl_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <typename T> __device__ __forceinline__ float ldf(const T* p, int64_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, int64_t i) { return p[i]; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p, int64_t i) { return __bfloat162float(p[i]); }
template <typename T> __device__ __forceinline__ void stf(T* p, int64_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, int64_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, int64_t i, float v) { p[i] = __float2bfloat16_rn(v); }

// 4-wide vector access (caller guarantees 4-element alignment of the index and of the base pointer)
template <typename T> __device__ __forceinline__ void ld4(const T* p, int64_t i, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, int64_t i, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p + i);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16* p, int64_t i, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p + i);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void st4(T* p, int64_t i, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, int64_t i, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, int64_t i, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p + i) = t;
}

// compile-time activation (elementwise kernels are instantiated per activation so that no predicated-off
// transcendental code is issued)
template <int ACT> __device__ __forceinline__ float act_fwd_t(float u, float alpha) {
  if constexpr (ACT == TGAN_ACT_RELU) return u > 0.f ? u : 0.f;
  else if constexpr (ACT == TGAN_ACT_LRELU) return u > 0.f ? u : alpha * u;
  else if constexpr (ACT == TGAN_ACT_TANH) return tanhf(u);
  else if constexpr (ACT == TGAN_ACT_SIGMOID) return 1.f / (1.f + expf(-u));
  else if constexpr (ACT == TGAN_ACT_SOFTPLUS) return u > 20.f ? u : log1pf(expf(u));
  else return u;
}
template <int ACT> __device__ __forceinline__ float act_grad_from_y_t(float y, float alpha) {
  if constexpr (ACT == TGAN_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  else if constexpr (ACT == TGAN_ACT_LRELU) return y > 0.f ? 1.f : (y < 0.f ? alpha : 0.f);
  else if constexpr (ACT == TGAN_ACT_TANH) return 1.f - y * y;
  else if constexpr (ACT == TGAN_ACT_SIGMOID) return y * (1.f - y);
  else if constexpr (ACT == TGAN_ACT_SOFTPLUS) return 1.f - expf(-y);
  else return 1.f;
}
#define TGAN_DISPATCH_ACT(act, A, ...)                                        \
  switch (act) {                                                               \
    case TGAN_ACT_NONE: { constexpr int A = TGAN_ACT_NONE; __VA_ARGS__; } break;         \
    case TGAN_ACT_RELU: { constexpr int A = TGAN_ACT_RELU; __VA_ARGS__; } break;         \
    case TGAN_ACT_LRELU: { constexpr int A = TGAN_ACT_LRELU; __VA_ARGS__; } break;       \
    case TGAN_ACT_TANH: { constexpr int A = TGAN_ACT_TANH; __VA_ARGS__; } break;         \
    case TGAN_ACT_SIGMOID: { constexpr int A = TGAN_ACT_SIGMOID; __VA_ARGS__; } break;   \
    case TGAN_ACT_SOFTPLUS: { constexpr int A = TGAN_ACT_SOFTPLUS; __VA_ARGS__; } break; \
    default: tgan::set_error("bad activation %d", (int)(act)); return 1;      \
  }

__device__ __forceinline__ float act_fwd(float u, int act, float alpha) {
  switch (act) {
    case TGAN_ACT_RELU: return u > 0.f ? u : 0.f;
    case TGAN_ACT_LRELU: return u > 0.f ? u : alpha * u;
    case TGAN_ACT_TANH: return tanhf(u);
    case TGAN_ACT_SIGMOID: return 1.f / (1.f + expf(-u));
    case TGAN_ACT_SOFTPLUS: return u > 20.f ? u : log1pf(expf(u));
    default: return u;
  }
}
// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_grad_from_y(float y, int act, float alpha) {
  switch (act) {
    case TGAN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case TGAN_ACT_LRELU: return y > 0.f ? 1.f : (y < 0.f ? alpha : 0.f);
    case TGAN_ACT_TANH: return 1.f - y * y;
    case TGAN_ACT_SIGMOID: return y * (1.f - y);
    case TGAN_ACT_SOFTPLUS: return 1.f - expf(-y);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 (Salmon et al. 2011), counter = (i_lo, i_hi, stream_lo, stream_hi), key = seed.
struct Philox {
  uint32_t k0, k1;
  __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ uint4 operator()(uint64_t idx, uint64_t stream) const {
    uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f); }  // [0,1)

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#define TGAN_DISPATCH_1(dt, T, ...)                                   \
  if ((dt) == TGAN_F32) { typedef float T; __VA_ARGS__; }             \
  else if ((dt) == TGAN_BF16) { typedef tgan::bf16 T; __VA_ARGS__; }  \
  else { tgan::set_error("bad dtype %d", (int)(dt)); return 1; }

}  // namespace tgan
