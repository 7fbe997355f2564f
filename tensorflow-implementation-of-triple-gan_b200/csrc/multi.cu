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
  if (co < d.Co)
    for (int e = e0 + ty; e < e1; e += 8) {
      const float v = d.V[wn_idx(e / d.B, co, e % d.B, d.Co, d.B)];
      s += v * v;
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
  if (co < d.Co)
    for (int e = e0 + ty; e < e1; e += 8) {
      const int64_t i = wn_idx(e / d.B, co, e % d.B, d.Co, d.B);
      s += d.dW[i] * d.V[i];
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
  for (int ft0 = blockIdx.x * PACK_U; ft0 < total; ft0 += gridDim.x * PACK_U) {
    float v[PACK_U][4];
    int dsel[PACK_U], kb_[PACK_U], nb_[PACK_U], t_[PACK_U];
    // ---- phase 1: every global load of the PACK_U tiles ----
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      const int ft = ft0 + u;
      dsel[u] = -1;
      if (ft >= total) continue;
      while (ft >= first[di + 1]) ++di;            // flat tiles are visited in increasing order
      const tgan_pack_desc& d = ds[di];
      const int tile = ft - first[di];
      const int nbk = (d.Kpad + 31) / 32, nbn = (d.Nr + 31) / 32;
      const int kb = tile % nbk, nb = (tile / nbk) % nbn, t = tile / (nbk * nbn);
      dsel[u] = di; kb_[u] = kb; nb_[u] = nb; t_[u] = t;
      const int tap = t < 32 ? (int)tap_sm[di][t] : (d.taps ? d.taps[t] : t);
      const int64_t base = (int64_t)tap * d.st;
      const bool transpose = d.sk != 1 && d.sn == 1;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = ty + 8 * jj;
        // transposed tiles: lanes along n (the contiguous source axis), rows j along k; else lanes along k, rows along n
        const int n_ = transpose ? nb * 32 + tx : nb * 32 + j;
        const int k = transpose ? kb * 32 + j : kb * 32 + tx;
        float x = 0.f;
        if (k < d.K && n_ < d.Nr) {
          x = d.src[base + (int64_t)n_ * d.sn + (int64_t)k * d.sk];
          if (d.scale_on == 1) x *= d.scale[n_];
          else if (d.scale_on == 2) x *= d.scale[k];
        }
        v[u][jj] = x;
      }
    }
    // ---- phase 2: transposed tiles go through shared memory ----
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      if (dsel[u] < 0) continue;
      const tgan_pack_desc& d = ds[dsel[u]];
      if (d.sk != 1 && d.sn == 1) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) sm[u][ty + 8 * jj][tx] = v[u][jj];
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PACK_U; ++u) {
      if (dsel[u] < 0) continue;
      const tgan_pack_desc& d = ds[dsel[u]];
      bf16* dst = reinterpret_cast<bf16*>(d.dst);
      const bool transpose = d.sk != 1 && d.sn == 1;
      const int k = kb_[u] * 32 + tx;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = ty + 8 * jj;
        const int nn = nb_[u] * 32 + j;
        if (nn < d.Nr && k < d.Kpad)
          dst[((int64_t)t_[u] * d.Nr + nn) * d.Kpad + k] = __float2bfloat16_rn(transpose ? sm[u][tx][j] : v[u][jj]);
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
