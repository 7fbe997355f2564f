// dp_fused.cu -- the data-parallel parameter update as ONE kernel over NVLink / NVSwitch peer memory:
//
//     reduce-scatter (peer loads)  ->  Adam on the owned shard  ->  all-gather of the new parameters (peer stores)
//
// The reference is single-GPU (Train_goodGAN.py:48, :724); data parallelism is new here.  The baseline is three
// ncclAllReduce calls per step (D 1.3 MB, G 20.5 MB, C 12.5 MB fp32) each followed by a fused Adam launch: at 8 GPUs that
// is ~0.24 ms of exposed, mostly latency-bound collective time per 3.5 ms step.  Here every rank owns 1/W of each flat
// parameter buffer: it sums that slice of all W gradient buffers straight out of the peers' HBM (fixed rank order ->
// bit-identical on every rank), applies tf.train.AdamOptimizer's update (train_base.py:91-97) to its slice with its
// LOCAL slice of the m / v slots (Adam state is sharded: 1/W of the slot traffic per GPU), and writes the new parameters
// into every replica.  Gradient and parameter buffers live in symmetric memory (torch.distributed._symmetric_memory
// provides allocation + pointer exchange: plumbing); the kernels, the flag protocol and the barriers are ours.
//
// Ordering: tgan_dp_barrier(slot 0) before the update (every rank's gradients are complete and visible system-wide),
// tgan_dp_barrier(slot 1) after it (every replica holds every shard before anything reads the parameters).  A barrier is
// a one-CTA kernel: thread p stores this rank's monotonically increasing epoch into peer p's flag array (release, system
// scope) and spins (acquire, bounded: __trap after ~20 s instead of hanging the GPU) until peer p's epoch arrived here.
// Epochs live on the device, so CUDA-graph replays keep counting.
//
// Roofline: NVLink-bound.  Per rank (W-1)/W * 4n bytes in (gradients) and out (parameters): G at W = 8 -> 18 MB each way.
#include "common.cuh"

namespace tgan {

struct DpPeers {
  void* ptr[8];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;      // peer memory: never through the (non-coherent) L1
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// NVSwitch multicast (NVLS): ONE load returns the sum of the same address on every GPU (the switch adds), ONE store
// reaches every GPU -- the owner of a shard moves 1/W of the bytes the peer-pointer version moves.
__device__ __forceinline__ float4 ld_reduce_mc_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void st_mc_f4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flags: int [2 slots][8 ranks] per rank (symmetric); epoch: int [2] local counters
__global__ void dp_barrier_kernel(DpPeers flags, int rank, int world, int slot, int* __restrict__ epoch) {
  pdl_entry();
  __shared__ int e_sh;
  if (threadIdx.x == 0) {
    e_sh = epoch[slot] + 1;
    epoch[slot] = e_sh;
    __threadfence_system();      // everything this GPU wrote before the barrier is visible before the flag is
  }
  __syncthreads();
  const int e = e_sh, p = threadIdx.x;
  if (p < world) {
    st_release_sys(reinterpret_cast<int*>(flags.ptr[p]) + slot * 8 + rank, e);
    const int* mine = reinterpret_cast<const int*>(flags.ptr[rank]) + slot * 8 + p;
    long long spins = 0;
    while (ld_acquire_sys(mine) < e) {
      __nanosleep(64);
      if (++spins > (1ll << 24)) __trap();      // (~20 s) a peer never arrived: fail loudly instead of hanging the device
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

// one launch per network: rank r owns elements [r * per, min(n, (r + 1) * per))
__global__ void __launch_bounds__(256) dp_adam_kernel(DpPeers grads, DpPeers thetas, const float* __restrict__ mc_grad,
                                                      float* __restrict__ mc_theta, float* __restrict__ m, float* __restrict__ v,
                                                      int64_t n, int64_t per, int rank, int world,
                                                      const float* __restrict__ state, float beta1, float beta2, float eps) {
  pdl_entry();
  const float lr = state[0], b1p = state[1], b2p = state[2];
  const float a = lr * sqrtf(1.f - b2p) / (1.f - b1p);
  const float inv_w = 1.f / (float)world;
  const int64_t lo = (int64_t)rank * per, hi = lo + per < n ? lo + per : n;
  const float* th_local = reinterpret_cast<const float*>(thetas.ptr[rank]);
  // (lo, hi and n are multiples of 4: the flat buffers pad every tensor to 4 elements)
  for (int64_t i = lo + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hi; i += (int64_t)gridDim.x * blockDim.x * 4) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mc_grad) {
      g = ld_reduce_mc_f4(mc_grad + i);      // summed inside the switch; every shard is reduced exactly once, by its owner
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {      // fixed order
        if (p >= world) break;
        const float4 t = ld_peer_f4(reinterpret_cast<const float*>(grads.ptr[p]) + i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    const float4 t = *reinterpret_cast<const float4*>(th_local + i);
    const float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
    float tt[4] = {t.x, t.y, t.z, t.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
    const float ga[4] = {g.x * inv_w, g.y * inv_w, g.z * inv_w, g.w * inv_w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {      // identical arithmetic to adam_kernel (optim.cu)
      ma[j] += (ga[j] - ma[j]) * (1.f - beta1);
      va[j] += (ga[j] * ga[j] - va[j]) * (1.f - beta2);
      tt[j] -= a * ma[j] / (sqrtf(va[j]) + eps);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(ma[0], ma[1], ma[2], ma[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(va[0], va[1], va[2], va[3]);
    const float4 nt = make_float4(tt[0], tt[1], tt[2], tt[3]);
    if (mc_theta) {
      st_mc_f4(mc_theta + i, nt);
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        if (p >= world) break;
        st_peer_f4(reinterpret_cast<float*>(thetas.ptr[p]) + i, nt);
      }
    }
  }
}

// ExponentialMovingAverage(0.9999).apply(c_vars) (Train_goodGAN.py:101-103) over the complete, gathered parameters
__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ theta, int64_t n4, float decay) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 e = reinterpret_cast<float4*>(ema)[i];
    const float4 t = reinterpret_cast<const float4*>(theta)[i];
    e.x -= (e.x - t.x) * (1.f - decay); e.y -= (e.y - t.y) * (1.f - decay);
    e.z -= (e.z - t.z) * (1.f - decay); e.w -= (e.w - t.w) * (1.f - decay);
    reinterpret_cast<float4*>(ema)[i] = e;
  }
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_dp_barrier(const uint64_t* flag_ptrs, int rank, int world, int slot, int* epoch, void* stream) {
  TGAN_CHECK_ARG(flag_ptrs && epoch && world >= 2 && world <= 8 && rank >= 0 && rank < world && (slot == 0 || slot == 1),
                 "dp_barrier: bad args (2 <= world <= 8)");
  DpPeers f;
  for (int p = 0; p < 8; ++p) f.ptr[p] = p < world ? reinterpret_cast<void*>(flag_ptrs[p]) : nullptr;
  pdl_launch(dp_barrier_kernel, 1, 32, 0, (cudaStream_t)stream, f, rank, world, slot, epoch);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_dp_adam(const uint64_t* grad_ptrs, const uint64_t* theta_ptrs, const float* mc_grad, float* mc_theta, float* m,
                            float* v, int64_t n, int rank, int world, const float* state, float beta1, float beta2, float eps,
                            void* stream) {
  TGAN_CHECK_ARG(grad_ptrs && theta_ptrs && m && v && state && n > 0 && n % 4 == 0 && world >= 2 && world <= 8 && rank >= 0 &&
                     rank < world, "dp_adam: bad args (n %% 4 == 0, 2 <= world <= 8)");
  DpPeers g, t;
  for (int p = 0; p < 8; ++p) {
    g.ptr[p] = p < world ? reinterpret_cast<void*>(grad_ptrs[p]) : nullptr;
    t.ptr[p] = p < world ? reinterpret_cast<void*>(theta_ptrs[p]) : nullptr;
    TGAN_CHECK_ARG(p >= world || (g.ptr[p] && t.ptr[p] && ((uintptr_t)g.ptr[p] & 15) == 0 && ((uintptr_t)t.ptr[p] & 15) == 0),
                   "dp_adam: peer buffers must be 16-byte aligned");
  }
  const int64_t per = ((n / 4 + world - 1) / world) * 4;      // shard length, a multiple of 4 elements
  int grid = (int)((per / 4 + 255) / 256);
  if (grid > 148 * 4) grid = 148 * 4;
  if (grid < 1) grid = 1;
  pdl_launch(dp_adam_kernel, grid, 256, 0, (cudaStream_t)stream, g, t, mc_grad, mc_theta, m, v, n, per, rank, world, state, beta1,
             beta2, eps);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_ema(float* ema, const float* theta, int64_t n, float decay, void* stream) {
  TGAN_CHECK_ARG(ema && theta && n > 0 && n % 4 == 0, "ema: bad args (n %% 4 == 0)");
  int64_t g = (n / 4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  pdl_launch(ema_kernel, (int)g, 256, 0, (cudaStream_t)stream, ema, theta, n / 4, decay);
  TGAN_LAUNCHED();
  return 0;
}
