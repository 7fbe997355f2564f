// multi.cu -- multi-tensor weight preparation: ONE launch serves every layer of a network.
//   * weight-norm forward  (inv_norm, scale = g * inv_norm per output channel)  nn.py:502,554; modle_base.py:66,101,148
//   * weight-norm backward (dg, dV from the accumulated dW)                     SURVEY.md Appendix B
//   * bf16 K-major operand packing for the tcgen05 path, weight-norm scale folded in
// The per-layer descriptors live in a device table built once by the host (parameter pointers are views into the flat
// per-network buffers, so they never change); blockIdx.y selects the tensor.  The step launches these three kernels
// once per network and optimiser version instead of 3-5 small kernels per layer.
#include "common.cuh"

namespace tgan {

__device__ __forceinline__ int64_t wn_idx(int a, int co, int b, int Co, int B) { return ((int64_t)a * Co + co) * B + b; }

// 256 threads = 32 channels x 8 row lanes; deterministic shared-memory fold over the lanes
__device__ __forceinline__ float fold8(float (*sm)[33], float v, int tx, int ty) {
  sm[ty][tx] = v;
  __syncthreads();
  float s = 0.f;
  if (ty == 0) {
#pragma unroll
    for (int y = 0; y < 8; ++y) s += sm[y][tx];
    sm[0][tx] = s;
  }
  __syncthreads();
  s = sm[0][tx];
  __syncthreads();
  return s;
}

constexpr int WN_SPLITS = 16;      // row ranges per tensor (CTAs along blockIdx.z)

__global__ void __launch_bounds__(256) wn_fwd_part_kernel(const tgan_wn_desc* __restrict__ descs, float* __restrict__ part,
                                                          int max_co) {
  pdl_entry();
  const tgan_wn_desc d = descs[blockIdx.y];
  const int c0 = blockIdx.x * 32;
  if (c0 >= d.Co) return;
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, co = c0 + tx;
  const int AB = d.A * d.B, per = (AB + WN_SPLITS - 1) / WN_SPLITS;
  const int e0 = blockIdx.z * per, e1 = min(AB, e0 + per);
  float s = 0.f;
  if (co < d.Co) {
    if (d.B == 1) {      // [A, Co]: conv HWIO / dense kernels -- no index division, four rows in flight
      const float* vp = d.V + co;
      const int64_t Co = d.Co;
      int e = e0 + ty;
      for (; e + 24 < e1; e += 32) {
        const float v0 = vp[e * Co], v1 = vp[(e + 8) * Co], v2 = vp[(e + 16) * Co], v3 = vp[(e + 24) * Co];
        s += v0 * v0; s += v1 * v1; s += v2 * v2; s += v3 * v3;
      }
      for (; e < e1; e += 8) { const float v = vp[e * Co]; s += v * v; }
    } else {
      for (int e = e0 + ty; e < e1; e += 8) {
        const float v = d.V[wn_idx(e / d.B, co, e % d.B, d.Co, d.B)];
        s += v * v;
      }
    }
  }
  s = fold8(sm, s, tx, ty);
  if (ty == 0 && co < d.Co) part[((int64_t)blockIdx.y * WN_SPLITS + blockIdx.z) * max_co + co] = s;
}
__global__ void wn_fwd_final_kernel(const tgan_wn_desc* __restrict__ descs, const float* __restrict__ part, int max_co) {
  pdl_entry();
  const tgan_wn_desc d = descs[blockIdx.y];
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co >= d.Co) return;
  float s = 0.f;
#pragma unroll
  for (int z = 0; z < WN_SPLITS; ++z) s += part[((int64_t)blockIdx.y * WN_SPLITS + z) * max_co + co];
  const float inv = d.eps_mode ? rsqrtf(fmaxf(s, 1e-12f)) : 1.0f / sqrtf(s);
  d.inv_norm[co] = inv;
  d.scale[co] = d.g[co] * inv;
}

// dg[co] += <dW,V>*inv ; dV += g*inv*(dW - V*inv^2*<dW,V>).  Two kernels so that the rows of a tensor are spread over
// WN_SPLITS CTAs: (1) partial dots per row range, (2) fold the partials in a fixed order and apply over the same range.
__global__ void __launch_bounds__(256) wn_bwd_dot_kernel(const tgan_wn_desc* __restrict__ descs, float* __restrict__ part,
                                                         int max_co) {
  pdl_entry();
  const tgan_wn_desc d = descs[blockIdx.y];
  const int c0 = blockIdx.x * 32;
  if (c0 >= d.Co) return;
  __shared__ float sm[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, co = c0 + tx;
  const int AB = d.A * d.B, per = (AB + WN_SPLITS - 1) / WN_SPLITS;
  const int e0 = blockIdx.z * per, e1 = min(AB, e0 + per);
  float s = 0.f;
  if (co < d.Co) {
    if (d.B == 1) {
      const float* vp = d.V + co;
      const float* wp = d.dW + co;
      const int64_t Co = d.Co;
      int e = e0 + ty;
      for (; e + 24 < e1; e += 32) {
        const float a0 = wp[e * Co], a1 = wp[(e + 8) * Co], a2 = wp[(e + 16) * Co], a3 = wp[(e + 24) * Co];
        const float b0 = vp[e * Co], b1 = vp[(e + 8) * Co], b2 = vp[(e + 16) * Co], b3 = vp[(e + 24) * Co];
        s += a0 * b0; s += a1 * b1; s += a2 * b2; s += a3 * b3;
      }
      for (; e < e1; e += 8) s += wp[e * Co] * vp[e * Co];
    } else {
      for (int e = e0 + ty; e < e1; e += 8) {
        const int64_t i = wn_idx(e / d.B, co, e % d.B, d.Co, d.B);
        s += d.dW[i] * d.V[i];
      }
    }
  }
  s = fold8(sm, s, tx, ty);
  if (ty == 0 && co < d.Co) part[((int64_t)blockIdx.y * WN_SPLITS + blockIdx.z) * max_co + co] = s;
}

__global__ void __launch_bounds__(256) wn_bwd_apply_kernel(const tgan_wn_desc* __restrict__ descs,
                                                           const float* __restrict__ part, int max_co) {
  pdl_entry();
  const tgan_wn_desc d = descs[blockIdx.y];
  const int c0 = blockIdx.x * 32;
  if (c0 >= d.Co) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, co = c0 + tx;
  if (co >= d.Co) return;
  float dot = 0.f;
#pragma unroll
  for (int z = 0; z < WN_SPLITS; ++z) dot += part[((int64_t)blockIdx.y * WN_SPLITS + z) * max_co + co];
  const float inv = d.inv_norm[co], gi = d.g[co] * inv, k = inv * inv * dot;
  if (ty == 0 && blockIdx.z == 0) d.dg[co] += dot * inv;
  const int AB = d.A * d.B, per = (AB + WN_SPLITS - 1) / WN_SPLITS;
  const int e0 = blockIdx.z * per, e1 = min(AB, e0 + per);
  if (d.B == 1) {
    const int64_t Co = d.Co;
    int e = e0 + ty;
    for (; e + 24 < e1; e += 32) {
      float w[4], v[4], o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = (int64_t)(e + 8 * u) * Co + co;
        w[u] = d.dW[i]; v[u] = d.V[i]; o[u] = d.dV[i];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) d.dV[(int64_t)(e + 8 * u) * Co + co] = o[u] + gi * (w[u] - v[u] * k);
    }
    for (; e < e1; e += 8) {
      const int64_t i = (int64_t)e * Co + co;
      d.dV[i] += gi * (d.dW[i] - d.V[i] * k);
    }
    return;
  }
  for (int e = e0 + ty; e < e1; e += 8) {
    const int64_t i = wn_idx(e / d.B, co, e % d.B, d.Co, d.B);
    d.dV[i] += gi * (d.dW[i] - d.V[i] * k);
  }
}

// dst[t][n][k] (bf16, k < Kpad) = k < K ? src[taps[t]*st + n*sn + k*sk] * scale : 0, in 32x32 (n, k) tiles.  One of the
// source strides is 1 in every layout the step uses; when it is the n stride the tile is transposed through shared
// memory so that both the fp32 reads and the bf16 writes are coalesced.
// One wave of persistent CTAs over a FLAT list of 32x32 tiles of all tensors.  The kernel is latency-bound (a network has
// 0.3-10 M weights: 1-10 tiles per CTA), so everything that would sit in front of a tile's loads is hoisted into the CTA
// prologue -- descriptor table, a parallel scan of the per-tensor tile counts, the tap tables -- and the tile loop
// handles PACK_U tiles per iteration: all their global loads are issued before the first store.
// (v1: 592 CTAs per tensor, ~12,000 CTAs per network.  v2: flat list, but a serial prefix sum by thread 0, a dependent
// tap-table load and one tile per iteration: 9 / 22 / 30 us for D / C / G.)
constexpr int PACK_MAX_DESCS = 64;
constexpr int PACK_U = 2;
__global__ void __launch_bounds__(256) pack_multi_kernel(const tgan_pack_desc* __restrict__ descs, int n0, int n) {
  pdl_entry();
  __shared__ tgan_pack_desc ds[PACK_MAX_DESCS];
  __shared__ int first[PACK_MAX_DESCS + 1];      // first flat tile of every tensor
  __shared__ int wtot[2];
  __shared__ unsigned char tap_sm[PACK_MAX_DESCS][32];
  __shared__ float sm[PACK_U][32][33];
  for (int i = threadIdx.x; i < n * (int)(sizeof(tgan_pack_desc) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(ds)[i] = reinterpret_cast<const uint32_t*>(descs + n0)[i];
  __syncthreads();
  int incl = 0;
  if (threadIdx.x < PACK_MAX_DESCS) {      // two warps: inclusive scan of the tile counts
    const int i = threadIdx.x, lane = i & 31;
    int v = i < n ? ds[i].T * ((ds[i].Nr + 31) / 32) * ((ds[i].Kpad + 31) / 32) : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) wtot[i >> 5] = v;
    incl = v;
  }
  for (int idx = threadIdx.x; idx < n * 32; idx += blockDim.x) {
    const int i = idx >> 5, t = idx & 31;
    tap_sm[i][t] = (unsigned char)((ds[i].taps && t < ds[i].T) ? ds[i].taps[t] : t);
  }
  __syncthreads();
  if (threadIdx.x < PACK_MAX_DESCS) {
    first[threadIdx.x + 1] = incl + (threadIdx.x >= 32 ? wtot[0] : 0);
    if (threadIdx.x == 0) first[0] = 0;
  }
  __syncthreads();
  const int total = first[n];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int di = 0;
  // ncu (profiles/r2_ncu_pack_multi_summary.txt): the first versions of this loop were INSTRUCTION-bound -- 84 instructions
  // per element (64-bit index arithmetic on descriptor fields re-read from shared memory, per-element bounds checks).  Per
  // tile everything is now reduced to one source pointer + stride and one destination pointer + stride per thread; only
  // edge tiles (partial in n or k) take the checked path.
  for (int ft0 = blockIdx.x * PACK_U; ft0 < total; ft0 += gridDim.x * PACK_U) {
    float v[PACK_U][4];
    bf16* dptr[PACK_U];
    int64_t dstep[PACK_U];
    int mode[PACK_U];            // -1: no tile; bit 0: transposed; bit 1: edge tile (checked stores)
    int nrem[PACK_U], krem[PACK_U];
    // ---- phase 1: every global load of the PACK_U tiles ----
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      const int ft = ft0 + u;
      mode[u] = -1;
      if (ft >= total) continue;
      while (ft >= first[di + 1]) ++di;            // flat tiles are visited in increasing order
      const tgan_pack_desc& d = ds[di];
      const int K = d.K, Nr = d.Nr, Kpad = d.Kpad, son = d.scale_on;
      const int64_t sn = d.sn, sk = d.sk;
      const int tile = ft - first[di];
      const int nbk = (Kpad + 31) >> 5, nbn = (Nr + 31) >> 5;
      const int kb = tile % nbk, r2 = tile / nbk, nb = r2 % nbn, t = r2 / nbn;
      const int tap = t < 32 ? (int)tap_sm[di][t] : (d.taps ? d.taps[t] : t);
      const bool transpose = sk != 1 && sn == 1;
      const int smod = d.scale_mod;               // (scale index modulo: only the checked path handles it)
      const bool edge = (nb * 32 + 32 > Nr) || (kb * 32 + 32 > K) || smod > 0;
      mode[u] = (transpose ? 1 : 0) | (edge ? 2 : 0);
      // rows j = ty + 8 * jj of the tile; lanes tx along the source's contiguous axis
      const int n0 = transpose ? nb * 32 + tx : nb * 32 + ty;      // first n of this thread
      const int k0 = transpose ? kb * 32 + ty : kb * 32 + tx;      // first k of this thread
      const float* sp = d.src + (int64_t)tap * d.st + (int64_t)n0 * sn + (int64_t)k0 * sk;
      const int64_t sstep = 8 * (transpose ? sk : sn);
      const float* scp = son == 0 ? nullptr : d.scale + ((son == 1) == transpose ? (transpose ? n0 : k0) : 0);
      // scale index: son == 1 -> scale[n], son == 2 -> scale[k]; it is lane-constant when it follows the lane axis
      const bool sc_lane = son != 0 && ((son == 1) == transpose);   // transposed: lanes along n; else lanes along k
      float scl = 1.f;
      if (!edge) {
        if (sc_lane) scl = *scp;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) v[u][jj] = sp[jj * sstep];
        if (son != 0 && !sc_lane) {
          const float* sr = d.scale + (transpose ? k0 : n0);       // follows the row axis j
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[u][jj] *= sr[8 * jj];
        } else {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) v[u][jj] *= scl;
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int n_ = transpose ? n0 : n0 + 8 * jj, k = transpose ? k0 + 8 * jj : k0;
          float x = 0.f;
          if (k < K && n_ < Nr) {
            x = sp[jj * sstep];
            if (son == 1) x *= d.scale[smod > 0 ? n_ % smod : n_];
            else if (son == 2) x *= d.scale[smod > 0 ? k % smod : k];
          }
          v[u][jj] = x;
        }
      }
      // destination: rows nn = nb * 32 + ty + 8 * jj, lane k = kb * 32 + tx
      dptr[u] = reinterpret_cast<bf16*>(d.dst) + ((int64_t)t * Nr + nb * 32 + ty) * Kpad + kb * 32 + tx;
      dstep[u] = (int64_t)8 * Kpad;
      nrem[u] = Nr - (nb * 32 + ty);          // row jj is inside iff 8 * jj < nrem
      krem[u] = Kpad - (kb * 32 + tx);        // lane is inside iff krem > 0
    }
    // ---- phase 2: transposed tiles go through shared memory ----
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      if (mode[u] >= 0 && (mode[u] & 1)) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) sm[u][ty + 8 * jj][tx] = v[u][jj];
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      if (mode[u] < 0) continue;
      if (mode[u] & 1) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) v[u][jj] = sm[u][tx][ty + 8 * jj];
      }
      if (!(mode[u] & 2)) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) dptr[u][jj * dstep[u]] = __float2bfloat16_rn(v[u][jj]);
      } else if (krem[u] > 0) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (8 * jj < nrem[u]) dptr[u][jj * dstep[u]] = __float2bfloat16_rn(v[u][jj]);
      }
    }
    __syncthreads();
  }
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_sizeof_wn_desc(void) { return (int)sizeof(tgan_wn_desc); }
extern "C" int tgan_sizeof_pack_desc(void) { return (int)sizeof(tgan_pack_desc); }

extern "C" int tgan_weightnorm_fwd_multi(const tgan_wn_desc* descs_dev, int n, int max_co, float* ws, void* stream) {
  TGAN_CHECK_ARG(descs_dev && ws && n > 0 && max_co > 0, "weightnorm_fwd_multi: bad args");
  pdl_launch(wn_fwd_part_kernel, dim3(ceil_div(max_co, 32), n, WN_SPLITS), 256, 0, (cudaStream_t)((cudaStream_t)stream), descs_dev, ws, max_co);
  TGAN_LAUNCHED();
  pdl_launch(wn_fwd_final_kernel, dim3(ceil_div(max_co, 128), n), 128, 0, (cudaStream_t)((cudaStream_t)stream), descs_dev, ws, max_co);
  TGAN_LAUNCHED();
  return 0;
}
extern "C" int64_t tgan_weightnorm_bwd_multi_ws_floats(int n, int max_co) { return (int64_t)n * WN_SPLITS * max_co; }
extern "C" int tgan_weightnorm_bwd_multi(const tgan_wn_desc* descs_dev, int n, int max_co, float* ws, void* stream) {
  TGAN_CHECK_ARG(descs_dev && ws && n > 0 && max_co > 0, "weightnorm_bwd_multi: bad args");
  dim3 grid(ceil_div(max_co, 32), n, WN_SPLITS);
  pdl_launch(wn_bwd_dot_kernel, grid, 256, 0, (cudaStream_t)((cudaStream_t)stream), descs_dev, ws, max_co);
  TGAN_LAUNCHED();
  pdl_launch(wn_bwd_apply_kernel, grid, 256, 0, (cudaStream_t)((cudaStream_t)stream), descs_dev, ws, max_co);
  TGAN_LAUNCHED();
  return 0;
}
// the tile loop is grid-strided over RESIDENT CTAs: a grid larger than one wave would run its surplus CTAs after the
// first wave has finished all of its tiles
static int pack_grid() {
  static int grid = 0;
  if (!grid) {
    int per_sm = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pack_multi_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    grid = sms * per_sm;
  }
  return grid;
}
extern "C" int tgan_pack_weight_multi(const tgan_pack_desc* descs_dev, int n, void* stream) {
  TGAN_CHECK_ARG(descs_dev && n > 0, "pack_weight_multi: bad args");
  for (int n0 = 0; n0 < n; n0 += PACK_MAX_DESCS) {      // (a network has 10-40 packed operands: one launch)
    const int cnt = n - n0 < PACK_MAX_DESCS ? n - n0 : PACK_MAX_DESCS;
    pdl_launch(pack_multi_kernel, pack_grid(), 256, 0, (cudaStream_t)stream, descs_dev, n0, cnt);
    TGAN_LAUNCHED();
  }
  return 0;
}
