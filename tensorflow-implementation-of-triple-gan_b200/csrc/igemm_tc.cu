// igemm_tc.cu -- the tensor-core path of every dense contraction of the Triple-GAN step, hand-written
// for sm_100a: TMA (cp.async.bulk.tensor) operand staging, tcgen05.mma with fp32 accumulators in TMEM,
// tcgen05.ld epilogues, mbarrier producer/consumer pipelines, warp-specialised persistent CTAs.
//
//  * igemm_kernel  : implicit-GEMM convolution fprop / dgrad / transposed-conv parity classes / plain GEMM.
//                    A = activations [N,H,W,C] gathered tap by tap with SHIFTED-WINDOW 4-D TMA boxes
//                    (out-of-bounds zero fill == TF zero padding, traversal strides == conv stride),
//                    B = packed bf16 weights [T][Nout][Kpad]; both K-major, 128B-swizzled.
//                    Replaces tf.nn.conv2d / conv2d_transpose / matmul (nn.py:504,553; modle_base.py:40,102,
//                    149,161,250) and their input-gradients.
//  * wgrad_kernel  : filter gradients.  The GEMM K dimension is the PIXEL axis, so both operands (dz and the
//                    shifted x window) are MN-major UMMA operands straight out of the same NHWC TMA boxes;
//                    split-K over pixel tiles + deterministic second-stage reduction.
//
// Roofline: tensor pipe (bf16 in, fp32 accumulate).  Algorithmic FLOPs per launch = 2 * pixels * T * C * Nout.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace tgan {

// ------------------------------------------------------------------------------------------------
// host: tensor-map encoder via the runtime's driver entry point (no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
              (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)(rank > 2 ? gd[2] : 0),
              (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0], bx[1], rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// implicit-GEMM kernel
// ------------------------------------------------------------------------------------------------
constexpr int IG_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int A_STAGE_BYTES = 128 * 128; // 128 pixels x 64 bf16

struct IgParams {
  int N, th, tw, nb, tiles_y, tiles_x, m_tiles, n_tiles, BN, T, kchunks, klast, stages, sy, sx;
  int dy[25], dx[25];
  void* out;
  int odt, OH, OW, ldo, osy, osx, ooy, oox, vh, vw, Nout;
  const float* bias;
  float* colsum;
  int act;
  float alpha;
};

__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ IgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages, BN = p.BN;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128u;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)stages * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)stages * b_stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + stages;
  uint64_t* tfull = bars + 2 * stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 0) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int ksteps = p.T * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, ng = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.tw * p.sx, y0 = ty * p.th * p.sy, n0 = ng * p.nb;
        for (int t = 0; t < p.T; ++t) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], A_STAGE_BYTES + b_stage_bytes);
            tma_load_4d(sA + (size_t)stage * A_STAGE_BYTES, &tmA, &full[stage], kc * 64, x0 + p.dx[t], y0 + p.dy[t], n0);
            tma_load_3d(sB + (size_t)stage * b_stage_bytes, &tmB, &full[stage], kc * 64, nt * BN, t);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t accphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], accphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const int nk = ((ks % p.kchunks) == p.kchunks - 1) ? p.klast : 4;
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
          for (int k = 0; k < nk; ++k) {
            // K-major SWIZZLE_128B: 8-row groups 1024 B apart; a K step of 16 bf16 = +32 B inside the atom
            umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (ks | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);          // smem slot reusable once these MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; accphase ^= 1; }
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> (bias, activation, column sums) -> global ----
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int pix_per_img = p.th * p.tw;
    int acc = 0; uint32_t accphase = 0;
    const bool vec_ok = (p.odt == TGAN_BF16) ? ((p.ldo % 8 == 0) && (p.Nout % 8 == 0)) : ((p.ldo % 4 == 0) && (p.Nout % 4 == 0));
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, ng = mt / (p.tiles_x * p.tiles_y);
      const int nl = row / pix_per_img, rem = row % pix_per_img;
      const int oy = ty * p.th + rem / p.tw, ox = tx * p.tw + rem % p.tw, n = ng * p.nb + nl;
      const bool rvalid = (n < p.N) && (oy < p.vh) && (ox < p.vw);
      const int64_t poff = rvalid ? (((int64_t)n * p.OH + (oy * p.osy + p.ooy)) * p.OW + (ox * p.osx + p.oox)) * p.ldo : 0;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const int colbase = nt * BN + c0;
        if (colbase >= p.Nout) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(r[j]) * p.alpha;
          const int col = colbase + j;
          if (p.bias && col < p.Nout) x += p.bias[col];
          v[j] = act_fwd(x, p.act, 0.2f);
        }
        if (p.colsum) {
          float mine = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float s = warp_sum(rvalid ? v[j] : 0.f);
            if (lane == j) mine = s;
          }
          if (colbase + lane < p.Nout) atomicAdd(&p.colsum[colbase + lane], mine);
        }
        if (rvalid) {
          if (p.odt == TGAN_BF16) {
            bf16* o = reinterpret_cast<bf16*>(p.out) + poff + colbase;
            if (vec_ok) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (colbase + g * 8 < p.Nout) {
                  uint4 u;
                  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[g * 8 + 0], v[g * 8 + 1]);
                  __nv_bfloat162 h1 = __floats2bfloat162_rn(v[g * 8 + 2], v[g * 8 + 3]);
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[g * 8 + 4], v[g * 8 + 5]);
                  __nv_bfloat162 h3 = __floats2bfloat162_rn(v[g * 8 + 6], v[g * 8 + 7]);
                  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
                  u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
                  *reinterpret_cast<uint4*>(o + g * 8) = u;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (colbase + j < p.Nout) o[j] = __float2bfloat16_rn(v[j]);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out) + poff + colbase;
            if (vec_ok) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                if (colbase + g * 4 < p.Nout)
                  *reinterpret_cast<float4*>(o + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (colbase + j < p.Nout) o[j] = v[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel: D[co, (tap, ci)] += dz^T[co, pix] * xs_tap[pix, ci], pixel axis = K, MN-major operands
// ------------------------------------------------------------------------------------------------
constexpr int WG_KP = 32;                // pixels (GEMM-K) per pipeline stage
constexpr int WG_BOX_BYTES = WG_KP * 128;

struct WgParams {
  int N, bh, bw, bn, ptiles_y, ptiles_x, p_tiles;   // pixel tiling (K axis)
  int co_tiles, ci_tiles, BNc, tg, tap_groups, T, splits, stages, sy, sx;
  int dy[25], dx[25];
  float* ws;            // [splits][T][Cout][Cin]
  int Cout, Cin;
};

__global__ void __launch_bounds__(IG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = p.stages;
  const int nbB = p.BNc / 64;                                   // 64-channel boxes per tap operand
  const uint32_t stage_bytes = (uint32_t)(2 + p.tg * nbB) * WG_BOX_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + stages;
  uint64_t* done = bars + 2 * stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncols = p.tg * p.BNc;
  const uint32_t tmem_cols = ncols <= 32 ? 32 : ncols <= 64 ? 64 : ncols <= 128 ? 128 : ncols <= 256 ? 256 : 512;

  // CTA -> (co tile, ci tile, tap group, split)
  int b = blockIdx.x;
  const int split = b % p.splits; b /= p.splits;
  const int tgi = b % p.tap_groups; b /= p.tap_groups;
  const int cit = b % p.ci_tiles; const int cot = b / p.ci_tiles;
  const int t0 = tgi * p.tg, nt = min(p.tg, p.T - t0);
  const int per = (p.p_tiles + p.splits - 1) / p.splits;
  const int pt0 = split * per, pt1 = min(p.p_tiles, pt0 + per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 0) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        const int tx = pt % p.ptiles_x, ty = (pt / p.ptiles_x) % p.ptiles_y, ng = pt / (p.ptiles_x * p.ptiles_y);
        const int ox0 = tx * p.bw, oy0 = ty * p.bh, n0 = ng * p.bn;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], (uint32_t)(2 + nt * nbB) * WG_BOX_BYTES);   // exact: last group may be short
        uint8_t* s = smem + (size_t)stage * stage_bytes;
        tma_load_4d(s, &tmDz, &full[stage], cot * 128, ox0, oy0, n0);
        tma_load_4d(s + WG_BOX_BYTES, &tmDz, &full[stage], cot * 128 + 64, ox0, oy0, n0);
        for (int j = 0; j < nt; ++j)
          for (int bb = 0; bb < nbB; ++bb)
            tma_load_4d(s + (size_t)(2 + j * nbB + bb) * WG_BOX_BYTES, &tmX, &full[stage], cit * p.BNc + bb * 64,
                        ox0 * p.sx + p.dx[t0 + j], oy0 * p.sy + p.dy[t0 + j], n0);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BNc, 1, 1);
      int stage = 0; uint32_t phase = 0;
      bool first = true;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t s = smem_u32(smem + (size_t)stage * stage_bytes);
        for (int j = 0; j < nt; ++j) {
          const uint32_t b_addr = s + (uint32_t)(2 + j * nbB) * WG_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < WG_KP / 16; ++k) {
            // MN-major SWIZZLE_128B: LBO = distance between 64-channel boxes, SBO = 8 pixel rows = 1024 B;
            // a K step of 16 pixels = 2 row groups = +2048 B
            umma_bf16(tmem_base + (uint32_t)(j * p.BNc), umma_smem_desc(s + k * 2048, WG_BOX_BYTES, 1024),
                      umma_smem_desc(b_addr + k * 2048, WG_BOX_BYTES, 1024), idesc, (first && k == 0) ? 0u : 1u);
          }
        }
        first = false;
        umma_commit(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int co = cot * 128 + q * 32 + lane;
    if (pt1 > pt0) {
      mbar_wait(done, 0);
      tc_fence_after();
    }
    for (int j = 0; j < nt; ++j) {
      for (int c0 = 0; c0 < p.BNc; c0 += 32) {
        const int ci0 = cit * p.BNc + c0;
        if (ci0 >= p.Cin) break;
        uint32_t r[32];
        if (pt1 > pt0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.BNc + c0), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        if (co < p.Cout) {
          float* o = p.ws + (((int64_t)split * p.T + (t0 + j)) * p.Cout + co) * p.Cin + ci0;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (ci0 + i < p.Cin) o[i] = __uint_as_float(r[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dst[t*st + co*sco + ci*sci] = beta*dst + sum_splits ws[s][t][co][ci]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int T, int Cout, int Cin,
                                    float* __restrict__ dw, int64_t st, int64_t sco, int64_t sci, float beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t tot = (int64_t)T * Cout * Cin;
  if (i >= tot) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += ws[z * tot + i];
  const int ci = (int)(i % Cin);
  const int co = (int)((i / Cin) % Cout);
  const int t = (int)(i / ((int64_t)Cin * Cout));
  float* o = dw + t * st + co * sco + ci * sci;
  *o = (beta != 0.f ? beta * (*o) : 0.f) + s;
}

// dst[t][n][k] (bf16, k < Kpad) = k < K ? src[tap(t)*st + n*sn + k*sk] : 0
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int T, int Nr, int K,
                                   int Kpad, int64_t st, int64_t sn, int64_t sk, const int* __restrict__ taps,
                                   int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % Kpad);
  const int n = (int)((i / Kpad) % Nr);
  const int t = (int)(i / ((int64_t)Kpad * Nr));
  const int ts = taps ? taps[t] : t;
  dst[i] = __float2bfloat16_rn(k < K ? src[ts * st + n * sn + k * sk] : 0.f);
}

static void pick_tile(int gh, int gw, int& th, int& tw, int& nb, int total) {
  if (gh == 1) { tw = total; th = 1; nb = 1; return; }
  tw = 1; while (tw < gw && tw < total) tw <<= 1;
  th = 1; while (th < gh && th * tw < total) th <<= 1;
  nb = total / (tw * th);
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_igemm_bf16(const tgan_igemm_args* a, void* stream) {
  TGAN_CHECK_ARG(a && a->x && a->wp && a->out, "igemm: null pointer");
  TGAN_CHECK_ARG(a->T >= 1 && a->T <= 25, "igemm: T out of range");
  TGAN_CHECK_ARG(a->ldx % 8 == 0 && a->Kpad % 8 == 0, "igemm: ldx (%d) and Kpad (%d) must be multiples of 8", a->ldx, a->Kpad);
  TGAN_CHECK_ARG(((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->wp & 15) == 0, "igemm: operands must be 16B aligned");
  TGAN_CHECK_ARG(a->C >= 1 && a->C <= a->Kpad && a->Nout >= 1, "igemm: bad channel counts");
  IgParams p;
  memset(&p, 0, sizeof(p));
  const int sy = a->sy > 0 ? a->sy : 1, sx = a->sx > 0 ? a->sx : 1;
  pick_tile(a->gh, a->gw, p.th, p.tw, p.nb, 128);
  TGAN_CHECK_ARG(p.tw * sx <= 256 && p.th * sy <= 256, "igemm: strided box too large");
  p.N = a->N; p.sy = sy; p.sx = sx;
  p.tiles_y = ceil_div(a->gh, p.th); p.tiles_x = ceil_div(a->gw, p.tw);
  p.m_tiles = ceil_div(a->N, p.nb) * p.tiles_y * p.tiles_x;
  p.BN = a->Nout <= 32 ? 32 : a->Nout <= 64 ? 64 : a->Nout <= 128 ? 128 : 256;
  p.n_tiles = ceil_div(a->Nout, p.BN);
  p.T = a->T; p.kchunks = ceil_div(a->C, 64);
  p.klast = ceil_div(a->C - (p.kchunks - 1) * 64, 16);
  for (int t = 0; t < a->T; ++t) { p.dy[t] = a->dy[t]; p.dx[t] = a->dx[t]; }
  p.out = a->out; p.odt = a->odt; p.OH = a->OH; p.OW = a->OW; p.ldo = a->ldo;
  p.osy = a->osy > 0 ? a->osy : 1; p.osx = a->osx > 0 ? a->osx : 1; p.ooy = a->ooy; p.oox = a->oox;
  p.vh = a->vh > 0 ? a->vh : a->gh; p.vw = a->vw > 0 ? a->vw : a->gw; p.Nout = a->Nout;
  p.bias = a->bias; p.colsum = a->colsum; p.act = a->act; p.alpha = a->alpha == 0.f ? 1.f : a->alpha;
  const size_t stage_bytes = A_STAGE_BYTES + (size_t)p.BN * 128;
  int stages = (int)((232448 - 2048) / stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  const size_t smem_bytes = 1024 + stages * stage_bytes + 256;

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    uint64_t str[3] = {(uint64_t)a->ldx * 2, (uint64_t)a->W * a->ldx * 2, (uint64_t)a->H * a->W * a->ldx * 2};
    uint32_t box[4] = {64, (uint32_t)(p.tw * sx), (uint32_t)(p.th * sy), (uint32_t)p.nb};
    uint32_t es[4] = {1, (uint32_t)sx, (uint32_t)sy, 1};
    if (make_tmap_bf16(&tmA, a->x, 4, dims, str, box, es)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Kpad, (uint64_t)a->Nout, (uint64_t)a->T};
    uint64_t str[2] = {(uint64_t)a->Kpad * 2, (uint64_t)a->Nout * a->Kpad * 2};
    uint32_t box[3] = {64, (uint32_t)p.BN, 1};
    if (make_tmap_bf16(&tmB, a->wp, 3, dims, str, box, nullptr)) return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    TGAN_CHECK_ARG(e == cudaSuccess, "igemm: cannot set max dynamic smem: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < 148 ? total : 148;
  igemm_kernel<<<grid, IG_THREADS, smem_bytes, (cudaStream_t)stream>>>(tmA, tmB, p);
  TGAN_LAUNCHED();
  return 0;
}

static int wgrad_plan(const tgan_wgrad_args* a, WgParams& p) {
  memset(&p, 0, sizeof(p));
  const int sy = a->sy > 0 ? a->sy : 1, sx = a->sx > 0 ? a->sx : 1;
  p.N = a->N; p.sy = sy; p.sx = sx; p.T = a->T; p.Cout = a->Cout; p.Cin = a->Cin;
  pick_tile(a->gh, a->gw, p.bh, p.bw, p.bn, WG_KP);
  if (p.bw * sx > 256 || p.bh * sy > 256) { set_error("wgrad: strided box too large"); return 1; }
  p.ptiles_y = ceil_div(a->gh, p.bh); p.ptiles_x = ceil_div(a->gw, p.bw);
  p.p_tiles = ceil_div(a->N, p.bn) * p.ptiles_y * p.ptiles_x;
  p.co_tiles = ceil_div(a->Cout, 128);
  p.BNc = a->Cin <= 64 ? 64 : a->Cin <= 128 ? 128 : 256;
  p.ci_tiles = ceil_div(a->Cin, p.BNc);
  p.tg = 512 / p.BNc;
  if (p.tg > a->T) p.tg = a->T;
  p.tg = ceil_div(a->T, ceil_div(a->T, p.tg));     // balanced tap groups (T=9, max 4 -> 3+3+3)
  // smem: stages * (2 + tg*BNc/64) boxes of 4 KB
  while (p.tg > 1 && (size_t)2 * (2 + p.tg * (p.BNc / 64)) * WG_BOX_BYTES > 200 * 1024) --p.tg;
  p.tap_groups = ceil_div(a->T, p.tg);
  const int base = p.co_tiles * p.ci_tiles * p.tap_groups;
  int splits = (2 * 148 + base - 1) / base;
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  // keep >= 8 pixel tiles per split so the pipeline prologue is amortised
  while (splits > 1 && p.p_tiles / splits < 8) --splits;
  if (a->ws_bytes > 0) {   // clamp to the caller's workspace
    const int64_t per_split = (int64_t)a->T * a->Cout * a->Cin * 4;
    while (splits > 1 && (int64_t)splits * per_split > a->ws_bytes) --splits;
  }
  p.splits = splits;
  const size_t stage_bytes = (size_t)(2 + p.tg * (p.BNc / 64)) * WG_BOX_BYTES;
  int stages = (int)((232448 - 2048) / stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  for (int t = 0; t < a->T; ++t) { p.dy[t] = a->dy[t]; p.dx[t] = a->dx[t]; }
  return 0;
}

extern "C" int64_t tgan_wgrad_workspace_bytes(const tgan_wgrad_args* a) {
  WgParams p;
  if (!a || wgrad_plan(a, p)) return -1;
  return (int64_t)p.splits * a->T * a->Cout * a->Cin * 4;
}

extern "C" int tgan_wgrad_bf16(const tgan_wgrad_args* a, void* stream) {
  TGAN_CHECK_ARG(a && a->dz && a->x && a->dw && a->ws, "wgrad: null pointer");
  TGAN_CHECK_ARG(a->T >= 1 && a->T <= 25, "wgrad: T out of range");
  TGAN_CHECK_ARG(a->ldx % 8 == 0 && a->lddz % 8 == 0, "wgrad: pixel strides must be multiples of 8 elements");
  WgParams p;
  if (wgrad_plan(a, p)) return 1;
  const int64_t need = (int64_t)p.splits * a->T * a->Cout * a->Cin * 4;
  TGAN_CHECK_ARG(a->ws_bytes >= need, "wgrad: workspace too small (%lld < %lld)", (long long)a->ws_bytes, (long long)need);
  p.ws = a->ws;
  CUtensorMap tmDz, tmX;
  {
    uint64_t dims[4] = {(uint64_t)a->Cout, (uint64_t)a->gw, (uint64_t)a->gh, (uint64_t)a->N};
    uint64_t str[3] = {(uint64_t)a->lddz * 2, (uint64_t)a->gw * a->lddz * 2, (uint64_t)a->gh * a->gw * a->lddz * 2};
    uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if (make_tmap_bf16(&tmDz, a->dz, 4, dims, str, box, nullptr)) return 1;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    uint64_t str[3] = {(uint64_t)a->ldx * 2, (uint64_t)a->W * a->ldx * 2, (uint64_t)a->H * a->W * a->ldx * 2};
    uint32_t box[4] = {64, (uint32_t)(p.bw * p.sx), (uint32_t)(p.bh * p.sy), (uint32_t)p.bn};
    uint32_t es[4] = {1, (uint32_t)p.sx, (uint32_t)p.sy, 1};
    if (make_tmap_bf16(&tmX, a->x, 4, dims, str, box, es)) return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    TGAN_CHECK_ARG(e == cudaSuccess, "wgrad: cannot set max dynamic smem: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t stage_bytes = (size_t)(2 + p.tg * (p.BNc / 64)) * WG_BOX_BYTES;
  const size_t smem_bytes = 1024 + p.stages * stage_bytes + 256;
  const int grid = p.co_tiles * p.ci_tiles * p.tap_groups * p.splits;
  wgrad_kernel<<<grid, IG_THREADS, smem_bytes, (cudaStream_t)stream>>>(tmDz, tmX, p);
  TGAN_LAUNCHED();
  const int64_t tot = (int64_t)a->T * a->Cout * a->Cin;
  wgrad_reduce_kernel<<<ceil_div(tot, 256), 256, 0, (cudaStream_t)stream>>>(a->ws, p.splits, a->T, a->Cout, a->Cin,
                                                                          a->dw, a->dw_st, a->dw_sco, a->dw_sci, a->beta);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_pack_weight_bf16(const float* src, void* dst, int T, int Nrows, int K, int Kpad, int64_t st,
                                     int64_t sn, int64_t sk, const int* taps_dev, void* stream) {
  TGAN_CHECK_ARG(src && dst && T >= 1 && Nrows >= 1 && K >= 1 && Kpad >= K && Kpad % 8 == 0, "pack_weight: bad args");
  const int64_t total = (int64_t)T * Nrows * Kpad;
  pack_weight_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, T, Nrows, K, Kpad, st, sn,
                                                                            sk, taps_dev, total);
  TGAN_LAUNCHED();
  return 0;
}
