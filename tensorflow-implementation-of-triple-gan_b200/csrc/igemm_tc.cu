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
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

// Floor-measurement switches (skip loads / stores / the epilogue) exist only in builds made with
// -DTGAN_DEBUG_SWITCHES; in the release library the branches are compiled out and TGAN_IGEMM_DBG is never read.
#ifdef TGAN_DEBUG_SWITCHES
#define TGAN_DBG(bit) (p.dbg & (bit))
#else
#define TGAN_DBG(bit) 0
#endif

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
                   const uint32_t* box, const uint32_t* elem_strides, bool swizzle128) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return 1; }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
constexpr int IG_THREADS = 320;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue (see below)
constexpr int WG_THREADS = 192;          // wgrad: warp 0 producer, warp 1 issuer, warps 2-5 epilogue
constexpr int W_STAGE_BYTES = 128 * 128; // 128 output channels x 64 bf16 (UMMA A operand, M = 128)
constexpr int P_TILE_BYTES = 128 * 128;  // 128 pixels x 64 bf16
constexpr int IG_NPIX = 256;             // pixels per CTA tile = UMMA N (two 128-pixel TMA boxes)
constexpr int IG_STAGE_BYTES = W_STAGE_BYTES + 2 * P_TILE_BYTES;
constexpr int IG_STAGES = 4;
constexpr int IG_OUT_STAGE_BYTES = 128 * 256;   // epilogue staging: 128 pixels x 128 channels bf16 ([pixel][channel] rows)
// Row-halo mode (3x3 / stride-1 convolutions whose 256-pixel tile is a block of full-width rows of ONE image): for each
// of the three column offsets dx ONE activation box of (rows + 2) x width pixels is loaded, and the three row taps dy are
// row-shifted views of it (UMMA descriptor start address + dy * width * 128 B) -- 3 activation loads per K chunk instead
// of 9.  Shared memory: HALO_A_SLOTS activation slots + HALO_W_SLOTS weight slots (one per tap) replace the 4-stage ring.
constexpr int HALO_A_SLOT_BYTES = 40 * 1024;    // (8 + 2) rows x 32 pixels x 128 B is the largest box (32x32 images)
#ifndef TGAN_HALO_A_SLOTS
#define TGAN_HALO_A_SLOTS 2
#define TGAN_HALO_W_SLOTS 4
#endif
constexpr int HALO_A_SLOTS = TGAN_HALO_A_SLOTS;
constexpr int HALO_W_SLOTS = TGAN_HALO_W_SLOTS;
// Two activation slots keep the MMAs fed (three measured no faster); the 48 KB that frees inside the ring region hold a
// SECOND epilogue staging buffer in row-halo mode, so the shared-memory writes of box h+1 overlap the TMA store of box h.
constexpr int HALO_OSTAGE2_OFFSET = HALO_A_SLOTS * HALO_A_SLOT_BYTES + HALO_W_SLOTS * W_STAGE_BYTES;
static_assert(HALO_OSTAGE2_OFFSET + IG_OUT_STAGE_BYTES <= IG_STAGES * IG_STAGE_BYTES, "halo rings + second staging buffer fit");

// Orientation: D[co, pixel] = W[co, k] * X[pixel, k]^T.  The OUTPUT CHANNELS are the UMMA M dimension (TMEM lanes) and
// 256 PIXELS are the UMMA N dimension: measured on B200, one cta_group::1 tcgen05.mma (M=128, K=16, smem operands)
// retires every ~82 ns for any N <= 256, so only N = 256 instructions reach the tensor-pipe rate; with the pixels on N
// every layer (Cout = 32 ... 512) issues full-width MMAs, and per-channel epilogue work (bias, column sums for the
// batch-norm statistics) is per-THREAD state instead of cross-lane reductions.
// n / d for 0 <= n < 2^31 with a host-computed multiplier (a runtime integer division costs ~40 dependent instructions;
// the tile loops of all three warp roles decode tile -> (class, channel tile, pixel tile, image, row, column))
struct FastDiv {
  uint32_t mul, shr, d;
  __host__ void set(int denom) {
    d = (uint32_t)denom;
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    const uint32_t p = 31 + l;
    mul = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
    shr = p - 32;
  }
  __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr); }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * (int)d; }
};

struct IgParams {
  FastDiv d_tiles_x, d_tiles_y, d_tpi, d_ct, d_tpc;
  int N, th, tw, nb, ltw, lppi, tiles_y, tiles_x, m_tiles, pp_tiles, ct_tiles, T, kchunks, klast, sy, sx;
  int dy[25], dx[25];
  void* out;
  int odt, OH, OW, ldo, osy, osx, ooy, oox, vh, vw, Nout;
  const float* bias;
  long long* colsum;   // Q24 fixed point: integer atomics are order-independent, so the statistics are bit-reproducible
  int nseg, seg_end[4], segflat;   // segflat: plain GEMM (one row of pixels) -> segments are pixel ranges
  int act;
  float alpha;
  int tstore;   // 1: bf16 output staged in shared memory and written by TMA bulk-tensor stores
  // output-parity classes of a strided transposed convolution, all in ONE launch: class c owns taps
  // [ctap0[c], ctap0[c] + cT[c]) of dy/dx/the packed weights and writes at output offset (cooy[c], coox[c])
  int ncls, cT[4], ctap0[4], cooy[4], coox[4];
  int halo, halo_rows, halo_dy0;   // row-halo mode: box rows (2*th + 2), smallest dy
  int nstages;                     // ring depth of the tap-by-tap mode: 4, or 3 for launches with <= 2 stages per tile --
                                   // the fourth stage's 48 KB then hold the SECOND output staging buffer (as in row-halo mode)
  int nbox;                        // 128-pixel boxes per tile: 2 (UMMA N = 256), or 1 when that leaves SMs without a tile
  // ---- fused mean-only batch norm (nn.py:147-187 without its own pass over the activation) ----
  int bias_seg;                    // bias is [nseg][Nout]: b - mean of the tile's batch segment (the mean is known BEFORE the
                                   // launch: a linear function of border-aware input sums, tgan_mobn_mean_from_sums)
  long long* clsum;                // Q24 [nseg][9][Nout]: sums of the STORED (bf16-rounded) values by border class
                                   // (row first / interior / last) x (column first / interior / last) -> the next layer's mean
  int cls_lw, cls_lh;              // log2 of the image width / height the classes refer to (pixels are linear in the output)
  uint32_t* mask_out;              // [pixels / 32][Nout]: bit j = (stored value of pixel 32*i + j) > 0 (lrelu side)
  const uint32_t* mask_in;         // input gradient: multiply by 1 (bit set) or mask_alpha (lrelu') before the store
  float mask_alpha;
  int dbg;   // experiments only (TGAN_IGEMM_DBG): 1 skip activation loads, 2 skip weight loads, 4 skip stores,
             // 8 skip the epilogue body, 16 skip the smem-ring handshakes
};

// Slow path of the epilogue, deliberately OUT OF LINE and not unrolled: rare activations (tanh / sigmoid / softplus) and
// the direct global store used when the output cannot go through the TMA store (fp32 or a pixel stride that is not a
// multiple of 16 bytes: RGB images, 3-channel input gradients).  Keeping it compact matters more than its speed: fully
// unrolled it made the kernel ~300 KB of SASS and the hot epilogue stalled on instruction fetches.
__device__ __noinline__ void ig_slow_chunk(const IgParams& p, float* v, int pbase, int tx, int ty, int ng, int co,
                                           bool cvalid, float* csum, int ooy, int oox, int do_act, int do_store) {
  if (do_act && cvalid) {      // lanes beyond Nout (125 of 128 for the generator's RGB layer) skip the transcendental
#pragma unroll 4
    for (int j = 0; j < 32; ++j) v[j] = act_fwd(v[j], p.act, 0.2f);
  }
  if (!do_store) return;
  const int ppi = p.th * p.tw;
  const int pstep = p.osx * p.ldo;
  const int TW = p.tw < 32 ? p.tw : 32;
#pragma unroll 1
  for (int sgm = 0; sgm < 32 / TW; ++sgm) {
    const int pix = pbase + sgm * TW;
    const int nl = pix >> p.lppi, rem = pix & (ppi - 1);
    const int oy = ty * p.th + (rem >> p.ltw), ox0 = tx * p.tw + (rem & (p.tw - 1)), n = ng * p.nb + nl;
    const bool rowok = cvalid && (n < p.N) && (oy < p.vh);
    const int base = ((n * p.OH + (oy * p.osy + ooy)) * p.OW + (ox0 * p.osx + oox)) * p.ldo + co;
    const int lim = p.vw - ox0;          // pixel jc is inside the valid width iff jc < lim
    const int sg = (n >= p.seg_end[0]) + (n >= p.seg_end[1]) + (n >= p.seg_end[2]);   // batch segment of this image
#pragma unroll 4
    for (int jc = 0; jc < TW; ++jc) {
      const bool ok = rowok && (jc < lim);
      const float x = v[sgm * TW + jc];
      if (p.colsum && ok) {
        const int ox = ox0 + jc;
        const int sj = p.segflat ? (ox >= p.seg_end[0]) + (ox >= p.seg_end[1]) + (ox >= p.seg_end[2]) : sg;
        csum[sj] += x;
      }
      if (ok) {
        if (p.odt == TGAN_BF16) reinterpret_cast<bf16*>(p.out)[base + jc * pstep] = __float2bfloat16_rn(x);
        else reinterpret_cast<float*>(p.out)[base + jc * pstep] = x;
      }
    }
  }
}

__device__ __noinline__ void ig_sum_boundary(const IgParams& p, const float* v, int n, int ox0, int lim, float* cs) {
#pragma unroll 1
  for (int jc = 0; jc < n; ++jc) {
    const int ox = ox0 + jc;
    const float xs = (jc < lim) ? v[jc] : 0.f;
    cs[(ox >= p.seg_end[0]) + (ox >= p.seg_end[1]) + (ox >= p.seg_end[2])] += xs;
  }
}

// Per-segment channel sums of one chunk of 32 tile pixels (TMA-store path: the store itself needs no geometry, the
// statistics still must skip pixels outside the valid grid).  A run of TW pixels lies in one image row.
template <int TW>
__device__ __forceinline__ void ig_sum_chunk(const IgParams& p, const float (&v)[32], int pbase, int tx, int ty, int ng,
                                             float (&csum)[4]) {
  const int ppi = p.th * p.tw;
#pragma unroll
  for (int sgm = 0; sgm < 32 / TW; ++sgm) {
    const int pix = pbase + sgm * TW;
    const int nl = pix >> p.lppi, rem = pix & (ppi - 1);
    const int oy = ty * p.th + (rem >> p.ltw), ox0 = tx * p.tw + (rem & (p.tw - 1)), n = ng * p.nb + nl;
    const bool rowok = (n < p.N) && (oy < p.vh);
    const int lim = rowok ? p.vw - ox0 : 0;
    const int sga = (ox0 >= p.seg_end[0]) + (ox0 >= p.seg_end[1]) + (ox0 >= p.seg_end[2]);
    const int oxl = ox0 + TW - 1;
    const int sgb = (oxl >= p.seg_end[0]) + (oxl >= p.seg_end[1]) + (oxl >= p.seg_end[2]);
    if (p.segflat && sga != sgb) {      // a segment boundary falls inside this run of pixels: classify every pixel
      // (rare: at most three runs per launch.  Out of line on a copy, so that the hot path stays short and v[] in registers)
      float tmp[TW], cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int jc = 0; jc < TW; ++jc) tmp[jc] = v[sgm * TW + jc];
      ig_sum_boundary(p, tmp, TW, ox0, lim, cs);
#pragma unroll
      for (int k = 0; k < 4; ++k) csum[k] += cs[k];
      continue;
    }
    float s = 0.f;
    if (lim >= TW) {            // the whole run is inside the image (warp-uniform: lanes differ in the channel only)
#pragma unroll
      for (int jc = 0; jc < TW; ++jc) s += v[sgm * TW + jc];
    } else {
#pragma unroll
      for (int jc = 0; jc < TW; ++jc) s += (jc < lim) ? v[sgm * TW + jc] : 0.f;
    }
    const int sg = p.segflat ? sga : (n >= p.seg_end[0]) + (n >= p.seg_end[1]) + (n >= p.seg_end[2]);
    csum[0] += sg == 0 ? s : 0.f; csum[1] += sg == 1 ? s : 0.f;
    csum[2] += sg == 2 ? s : 0.f; csum[3] += sg == 3 ? s : 0.f;
  }
}

// Border-class sums of one chunk of 32 consecutive (linear) pixels starting at p0, image width W = 1 << LW in {16, 32}:
// a[rc * 3 + cc], rc / cc = 0 first, 1 interior, 2 last row / column.  p0 is warp-uniform (lanes differ in the channel),
// so the row-class branches do not diverge.  vr: the values as stored (bf16-rounded).
template <int LW>
__device__ __forceinline__ void ig_cls_chunk(const float (&vr)[32], int p0, int lh, int valid, float (&a)[9]) {
  constexpr int W = 1 << LW;
#pragma unroll
  for (int k = 0; k < 32 / W; ++k) {
    if (k * W >= valid) break;
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) tot += vr[k * W + j];
    const float fi = vr[k * W], la = vr[k * W + W - 1], mid = tot - fi - la;
    const int y = ((p0 >> LW) + k) & ((1 << lh) - 1);
    if (y == 0) { a[0] += fi; a[1] += mid; a[2] += la; }
    else if (y == (1 << lh) - 1) { a[6] += fi; a[7] += mid; a[8] += la; }
    else { a[3] += fi; a[4] += mid; a[5] += la; }
  }
}

// FUSED / RARE prune the epilogue at compile time.  ncu's source page of the all-in-one kernel (profiles/
// r2_ncu_igemm250_source_summary.txt): one pass of the epilogue body executes ~290 instructions scattered over 85 KB of SASS
// (every activation, the fused mean-only-BN stages, the direct-store path), and a third of the epilogue warps' stall
// samples are instruction fetches (stall_no_inst).  The launches of the step fall into three groups:
//   <0, false>  bf16 TMA-store output, activation none / leaky ReLU / ReLU, optional channel sums   (most launches)
//   <1, false>  + the forward stages of the fused mean-only BN (per-segment bias, sign masks out, border-class sums)
//   <2, false>  + its backward stage (sign masks in: du = dy * lrelu'(y), with the per-segment channel sums of du)
//   <3, true >  everything (tanh / sigmoid / softplus, fp32 or narrow outputs through the direct store)
template <int FUSED, bool RARE>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
             const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3,
             const __grid_constant__ IgParams p) {
  pdl_trigger();      // the wait sits after the prologue (barriers, TMEM allocation, descriptor prefetch)
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (pointer arithmetic on the __shared__ array keeps the address space, so the epilogue's
  // staging stores compile to STS instead of generic stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ostage = smem + (size_t)IG_STAGES * IG_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ostage + IG_OUT_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + IG_STAGES;
  uint64_t* tfull = bars + 2 * IG_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* fullA = tempty + 2;                 // row-halo mode rings
  uint64_t* emptyA = fullA + HALO_A_SLOTS;
  uint64_t* fullW = emptyA + HALO_A_SLOTS;
  uint64_t* emptyW = fullW + HALO_W_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(emptyW + HALO_W_SLOTS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < IG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < HALO_A_SLOTS; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < HALO_W_SLOTS; ++s) { mbar_init(&fullW[s], 1); mbar_init(&emptyW[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    if (p.halo) tma_prefetch_desc(&tmXh);
    if (p.tstore) tma_prefetch_desc(&tmO0);
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }   // 2 accumulators x 256 fp32 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_cls = p.pp_tiles * p.ct_tiles;
  const int total_tiles = tiles_per_cls * p.ncls;

  // Producer and MMA-issuer warps run their loops with ALL 32 lanes (warp-uniform control flow keeps addresses and
  // descriptors in uniform registers); only the TMA / tcgen05 instructions themselves are issued by one elected lane.
  if (warp == 0 && p.halo) {
    // ---- row-halo producer: per K chunk and column offset one activation box, then its three row taps' weights ----
    uint8_t* wring = smem + HALO_A_SLOTS * HALO_A_SLOT_BYTES;
    const uint32_t a_bytes = (uint32_t)p.halo_rows * p.tw * 128u;
    int sa = 0, sw = 0; uint32_t pa = 0, pw = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int ct, pp, ty, ng;
      p.d_ct.divmod(tile, pp, ct);
      p.d_tiles_y.divmod(2 * pp, ng, ty);          // tiles_x == 1, nb == 1: box 2*pp starts the 256-pixel tile
      const int y0 = ty * p.th + p.halo_dy0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        for (int c = 0; c < 3; ++c) {
          mbar_wait(&emptyA[sa], pa ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&fullA[sa], TGAN_DBG(1) ? 0u : a_bytes);
            if (!TGAN_DBG(1)) tma_load_4d(smem + (size_t)sa * HALO_A_SLOT_BYTES, &tmXh, &fullA[sa], kc * 64, p.dx[c], y0, ng);
          }
          __syncwarp();
          if (++sa == HALO_A_SLOTS) { sa = 0; pa ^= 1; }
          for (int r = 0; r < 3; ++r) {
            mbar_wait(&emptyW[sw], pw ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&fullW[sw], TGAN_DBG(2) ? 0u : (uint32_t)W_STAGE_BYTES);
              if (!TGAN_DBG(2)) tma_load_3d(wring + (size_t)sw * W_STAGE_BYTES, &tmW, &fullW[sw], kc * 64, ct * 128, r * 3 + c);
            }
            __syncwarp();
            if (++sw == HALO_W_SLOTS) { sw = 0; pw ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && p.halo) {
    // ---- row-halo MMA issuer ----
    const uint32_t idesc = umma_idesc_bf16(128, IG_NPIX, 0, 0);
    const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t a_base = smem_u32(smem) >> 4;
    const uint32_t w_base = a_base + ((HALO_A_SLOTS * HALO_A_SLOT_BYTES) >> 4);
    int sa = 0, sw = 0; uint32_t pa = 0, pw = 0; int acc = 0; uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], accphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * IG_NPIX);
      uint32_t accum = 0;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        const int nk = (kc == p.kchunks - 1) ? p.klast : 4;
        for (int c = 0; c < 3; ++c) {
          mbar_wait(&fullA[sa], pa);
          const uint32_t x_lo = a_base + (uint32_t)sa * (HALO_A_SLOT_BYTES >> 4);
          for (int r = 0; r < 3; ++r) {
            mbar_wait(&fullW[sw], pw);
            tc_fence_after();
            const uint32_t a_lo = w_base + (uint32_t)sw * (W_STAGE_BYTES >> 4);
            // the row tap dy = row-shifted view of the halo box: + (dy - dy0) * width pixel rows of 128 B
            const uint32_t b_lo = x_lo + (uint32_t)((p.dy[r * 3] - p.halo_dy0) * p.tw * 8);
            if (elect_one()) {
              for (int k = 0; k < nk; ++k) {
                umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, accum);
                accum = 1u;
              }
              umma_commit(&emptyW[sw]);
              if (r == 2) umma_commit(&emptyA[sa]);
              if (r == 2 && c == 2 && kc == p.kchunks - 1) umma_commit(&tfull[acc]);
            }
            accum = 1u;
            __syncwarp();
            if (++sw == HALO_W_SLOTS) { sw = 0; pw ^= 1; }
          }
          if (++sa == HALO_A_SLOTS) { sa = 0; pa ^= 1; }
        }
      }
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
  } else if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int cls, tin, ct, pp;
      p.d_tpc.divmod(tile, cls, tin);
      p.d_ct.divmod(tin, pp, ct);
      const int tap0 = p.ctap0[cls], tap1 = tap0 + p.cT[cls];
      int x0[2], y0[2], n0[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mt = p.nbox * pp + h;  // mt >= m_tiles -> image index >= N -> the box is zero-filled
        int t2, txx, tyy, ngg;
        p.d_tiles_x.divmod(mt, t2, txx);
        p.d_tiles_y.divmod(t2, ngg, tyy);
        x0[h] = txx * p.tw * p.sx;
        y0[h] = tyy * p.th * p.sy;
        n0[h] = ngg * p.nb;
      }
      for (int t = tap0; t < tap1; ++t) {
        const int ddx = p.dx[t], ddy = p.dy[t];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (TGAN_DBG(16)) continue;
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* s = smem + (size_t)stage * IG_STAGE_BYTES;
            mbar_arrive_expect_tx(&full[stage], (TGAN_DBG(2) ? 0u : (uint32_t)W_STAGE_BYTES) +
                                                    (TGAN_DBG(1) ? 0u : (uint32_t)p.nbox * P_TILE_BYTES));
            if (!TGAN_DBG(2)) tma_load_3d(s, &tmW, &full[stage], kc * 64, ct * 128, t);
            if (!TGAN_DBG(1)) {
              tma_load_4d(s + W_STAGE_BYTES, &tmX, &full[stage], kc * 64, x0[0] + ddx, y0[0] + ddy, n0[0]);
              if (p.nbox == 2)
                tma_load_4d(s + W_STAGE_BYTES + P_TILE_BYTES, &tmX, &full[stage], kc * 64, x0[1] + ddx, y0[1] + ddy, n0[1]);
            }
          }
          __syncwarp();
          if (++stage == p.nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, 128 * p.nbox, 0, 0);
    // K-major SWIZZLE_128B descriptor: LBO field 1 (unused), SBO = 1024 B (8 rows), version 1, layout 2
    const uint64_t desc_hi = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t s_base = smem_u32(smem) >> 4;
    int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ksteps = p.cT[p.d_tpc.div(tile)] * p.kchunks;
      mbar_wait(&tempty[acc], accphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * IG_NPIX);
      int kc = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        if (!TGAN_DBG(16)) mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_lo = s_base + (uint32_t)stage * (IG_STAGE_BYTES >> 4);
        const uint32_t b_lo = a_lo + (W_STAGE_BYTES >> 4);
        const bool last = (++kc == p.kchunks);
        if (last) kc = 0;
        if (elect_one()) {
          // a K step of 16 bf16 = +32 B inside the 128 B swizzle atom = +2 in the (>>4) start-address field
          if (!last || p.klast == 4) {
            umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 0), desc_hi | (uint64_t)(b_lo + 0), idesc, ks ? 1u : 0u);
            umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 2), desc_hi | (uint64_t)(b_lo + 2), idesc, 1u);
            umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 4), desc_hi | (uint64_t)(b_lo + 4), idesc, 1u);
            umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 6), desc_hi | (uint64_t)(b_lo + 6), idesc, 1u);
          } else {
            for (int k = 0; k < p.klast; ++k)
              umma_bf16(d_tmem, desc_hi | (uint64_t)(a_lo + 2 * k), desc_hi | (uint64_t)(b_lo + 2 * k), idesc, (ks | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                     // smem slot reusable once these MMAs retire
          if (ks == ksteps - 1) umma_commit(&tfull[acc]); // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
  } else {
    // ---- epilogue: thread = one output channel (TMEM lane), 32 consecutive pixels per tcgen05.ld.  EIGHT warps: a
    // warp runs alone on its scheduler, so the ~4 dependent instructions per element (convert, shared-memory store,
    // statistics) of a 128 x 256 tile cost ~6 us with four warps -- more than the tile's MMAs.  Warps w and w+4 share a
    // TMEM lane quadrant and split every 128-pixel box into two 64-pixel halves.
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int g = (warp - 2) >> 2;             // which 64 pixels of each 128-pixel box
    const bool plain = p.alpha == 1.f && p.bias == nullptr;
#ifdef TGAN_SINGLE_OSTAGE
    const bool two_stage = false;
#else
    const bool two_stage = p.halo != 0 || p.nstages == 3;
#endif
    // (offset arithmetic on the shared-memory base keeps the address space: the staging stores stay STS)
    const int OSTAGE_B_DELTA = p.halo ? HALO_OSTAGE2_OFFSET - IG_STAGES * IG_STAGE_BYTES : -IG_STAGE_BYTES;
    int obuf = 0;
    int acc = 0; uint32_t accphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int cls, tin, ct, pp;
      p.d_tpc.divmod(tile, cls, tin);
      p.d_ct.divmod(tin, pp, ct);
      const CUtensorMap* tmO = cls == 0 ? &tmO0 : cls == 1 ? &tmO1 : cls == 2 ? &tmO2 : &tmO3;
      const int co = ct * 128 + q * 32 + lane;
      const bool cvalid = co < p.Nout;
      float bias = (p.bias && cvalid && !p.bias_seg) ? p.bias[co] : 0.f;
      float csum[4] = {0.f, 0.f, 0.f, 0.f};
      float bsum[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // border-class sums of this tile
      int tile_seg = 0;
      mbar_wait(&tfull[acc], accphase);
      tc_fence_after();
      for (int h = 0; h < p.nbox; ++h) {
        const int mt = p.nbox * pp + h;
        if (mt >= p.m_tiles || TGAN_DBG(8)) break;
        int t2, tx, ty, ng;
        p.d_tiles_x.divmod(mt, t2, tx);
        p.d_tiles_y.divmod(t2, ng, ty);
        uint8_t* const ost = ostage + (obuf ? OSTAGE_B_DELTA : 0);
        if (!RARE || p.tstore) {           // the store that last used THIS staging buffer must have finished reading it
          if (warp == 2 && lane == 0) { if (two_stage) bulk_wait_read1(); else bulk_wait_read0(); }
          named_bar_sync(1, 256);
        }
        // linear index of the box's first pixel (host guarantees full-width tiles of one image, or one flat row, whenever
        // masks / class sums / per-segment biases are requested) and the batch segment the box belongs to
        const int box_p0 = ((ng * p.nb) * p.OH + ty * p.th) * p.OW + tx * p.tw;
        if ((FUSED & 1) && (p.bias_seg || p.clsum)) {
          const int key = p.segflat ? box_p0 : ng * p.nb;
          tile_seg = (key >= p.seg_end[0]) + (key >= p.seg_end[1]) + (key >= p.seg_end[2]);
          if (p.bias_seg) bias = cvalid ? p.bias[tile_seg * p.Nout + co] : 0.f;
        }
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int pbase = g * 64 + cc * 32;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * IG_NPIX + h * 128 + pbase), r);
          tmem_ld_wait();
          float v[32];
          if (plain) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha + bias;
          }
          // optional stages sit behind warp-uniform branches; every loop is fully unrolled so that v[] stays in registers
          // (a single dynamically indexed access would move the whole array to local memory)
          if (p.act == TGAN_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
          } else if (p.act == TGAN_ACT_RELU) {      // (one FMNMX per element: the generator's fc / transposed convs)
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (RARE && p.act == TGAN_ACT_TANH && p.tstore) {
            // bf16 output: 1 - 2/(e^2x + 1) with the fast exponential (abs. error ~1e-7, far below the bf16 rounding);
            // fp32 outputs take the exact tanhf of the out-of-line path
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 1.f - __fdividef(2.f, __expf(2.f * v[j]) + 1.f);
          } else if (RARE && p.act == TGAN_ACT_SIGMOID && p.tstore) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
          } else if (RARE && p.act != TGAN_ACT_NONE) {
            float tmp[32];            // a separate copy: taking v's address would move v to local memory on the hot path too
#pragma unroll
            for (int j = 0; j < 32; ++j) tmp[j] = v[j];
            ig_slow_chunk(p, tmp, pbase, tx, ty, ng, co, cvalid, csum, 0, 0, 1, 0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tmp[j];
          }
          if ((FUSED & 2) && p.mask_in && cvalid) {      // input gradient through the producer's leaky ReLU: du = dy * lrelu'(y)
            // bit 31 - j of the word = sign bit of the producer's stored value j (set = negative side = slope alpha)
            const uint32_t mw = p.mask_in[(size_t)((box_p0 + pbase) >> 5) * p.Nout + co];
            const float sl = p.mask_alpha;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (mw & (0x80000000u >> j)) v[j] *= sl;
          }
          if (!RARE || p.tstore) {
            // ---- staged path: [pixel][channel] rows in shared memory, one bulk-tensor store per 128-pixel box.  The
            // store needs no geometry (TMA clips at the valid extents); 32 lanes write 32 consecutive channels = 64 B.
            if ((FUSED & 1) && (p.mask_out || p.clsum)) {
              const int p0 = box_p0 + pbase;
              const int total_px = p.segflat ? p.vw : p.N * p.OH * p.OW;
              const int valid = total_px - p0;           // pixels of this chunk inside the tensor (flat GEMM tail)
              if (p.mask_out && cvalid && valid > 0) {
                // one funnel shift per element moves its SIGN bit into the word: bit 31 - j = (value j is negative)
                uint32_t mw = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) mw = __funnelshift_l(__float_as_uint(v[j]), mw, 1);
                p.mask_out[(size_t)(p0 >> 5) * p.Nout + co] = mw;
              }
              if (p.clsum && valid > 0) {
                // the values as stored: rounded in PAIRS (F2FP.PACK_AB on the ALU pipe) and unpacked with a shift / a mask;
                // element-wise F2F conversions go through the quarter-rate conversion unit (ncu: mio / short-scoreboard stalls
                // in the conv1_1 launch, profiles/r2_ncu_conv11_wgrad_summary.txt)
                float vr[32];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                  const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
                  vr[2 * j] = __uint_as_float(u << 16);
                  vr[2 * j + 1] = __uint_as_float(u & 0xffff0000u);
                }
                if (p.cls_lw == 5) ig_cls_chunk<5>(vr, p0, p.cls_lh, valid, bsum);
                else ig_cls_chunk<4>(vr, p0, p.cls_lh, valid, bsum);
              }
            }
            if (p.colsum) {
              switch (p.ltw) {
                case 2: ig_sum_chunk<4>(p, v, pbase, tx, ty, ng, csum); break;
                case 3: ig_sum_chunk<8>(p, v, pbase, tx, ty, ng, csum); break;
                case 4: ig_sum_chunk<16>(p, v, pbase, tx, ty, ng, csum); break;
                default: ig_sum_chunk<32>(p, v, pbase, tx, ty, ng, csum); break;
              }
            }
            bf16* srow = reinterpret_cast<bf16*>(ost) + (size_t)pbase * 128 + (q * 32 + lane);
#pragma unroll
            for (int j = 0; j < 32; ++j) srow[j * 128] = __float2bfloat16_rn(v[j]);
          } else if (RARE && !TGAN_DBG(4)) {
            float tmp[32], cs2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) tmp[j] = v[j];
            ig_slow_chunk(p, tmp, pbase, tx, ty, ng, co, cvalid, cs2, p.cooy[cls], p.coox[cls], 0, 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) csum[j] += cs2[j];
          }
        }
        if (!RARE || p.tstore) {
          fence_proxy_async();
          named_bar_sync(1, 256);
          if (warp == 2 && lane == 0 && !TGAN_DBG(4)) {
            tma_store_4d(tmO, ost, ct * 128, tx * p.tw, ty * p.th, ng * p.nb);
            bulk_commit();
          }
          if (two_stage) obuf ^= 1;
        }
      }
      if (p.colsum && cvalid) {
#pragma unroll
        for (int gg = 0; gg < 4; ++gg)
          if (gg < p.nseg && csum[gg] != 0.f)
            atomicAdd(reinterpret_cast<unsigned long long*>(&p.colsum[gg * p.Nout + co]),
                      (unsigned long long)__float2ll_rn(csum[gg] * 16777216.f));
      }
      if ((FUSED & 1) && p.clsum && cvalid) {        // a tile lies inside one batch segment (one image, or 256 pixels of one flat image)
#pragma unroll
        for (int k = 0; k < 9; ++k)
          if (bsum[k] != 0.f)
            atomicAdd(reinterpret_cast<unsigned long long*>(&p.clsum[((size_t)tile_seg * 9 + k) * p.Nout + co]),
                      (unsigned long long)__float2ll_rn(bsum[k] * 16777216.f));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; accphase ^= 1; }
    }
    if ((!RARE || p.tstore) && warp == 2 && lane == 0) bulk_wait_read0();   // staging buffer must outlive the last store's read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
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
  float* ws;            // [splits][T][Cin][Cout]
  int Cout, Cin;
  // row-halo mode (3x3 / stride 1, pixel tiles = kr full-width rows of one image): the tap group is one COLUMN offset dx,
  // its x operand is ONE box of kr + 2 rows and the three row taps dy are row-shifted views of it
  int halo, kp, xrows, dy0;     // pixels (GEMM K) per stage, pixel rows of the x box, smallest dy
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ WgParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET (pointer arithmetic on the __shared__ array keeps the address space, so the epilogue's
  // staging stores compile to STS instead of generic stores)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stages = p.stages;
  const int nbB = p.BNc / 64;                                   // 64-channel boxes per tap operand
  const uint32_t dz_box = (uint32_t)p.kp * 128u, x_box = (uint32_t)p.xrows * 128u;
  const uint32_t stage_bytes = 2u * dz_box + (uint32_t)((p.halo ? 1 : p.tg) * nbB) * x_box;
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
  const int t0 = p.halo ? tgi : tgi * p.tg, nt = p.halo ? 3 : min(p.tg, p.T - t0);   // halo: taps t0, t0 + 3, t0 + 6
  const int tstep = p.halo ? 3 : 1;
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
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      const int tx = pt % p.ptiles_x, ty = (pt / p.ptiles_x) % p.ptiles_y, ng = pt / (p.ptiles_x * p.ptiles_y);
      const int ox0 = tx * p.bw, oy0 = ty * p.bh, n0 = ng * p.bn;
      mbar_wait(&empty[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* s = smem + (size_t)stage * stage_bytes;
        if (p.halo) {
          mbar_arrive_expect_tx(&full[stage], 2u * dz_box + (uint32_t)nbB * x_box);
          tma_load_4d(s, &tmDz, &full[stage], cot * 128, ox0, oy0, n0);
          tma_load_4d(s + dz_box, &tmDz, &full[stage], cot * 128 + 64, ox0, oy0, n0);
          for (int bb = 0; bb < nbB; ++bb)
            tma_load_4d(s + 2 * dz_box + (size_t)bb * x_box, &tmX, &full[stage], cit * p.BNc + bb * 64, ox0 + p.dx[t0],
                        oy0 + p.dy0, n0);
        } else {
          mbar_arrive_expect_tx(&full[stage], (uint32_t)(2 + nt * nbB) * dz_box);   // exact: last group may be short
          tma_load_4d(s, &tmDz, &full[stage], cot * 128, ox0, oy0, n0);
          tma_load_4d(s + dz_box, &tmDz, &full[stage], cot * 128 + 64, ox0, oy0, n0);
          for (int j = 0; j < nt; ++j)
            for (int bb = 0; bb < nbB; ++bb)
              tma_load_4d(s + (size_t)(2 + j * nbB + bb) * dz_box, &tmX, &full[stage], cit * p.BNc + bb * 64,
                          ox0 * p.sx + p.dx[t0 + j], oy0 * p.sy + p.dy[t0 + j], n0);
        }
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const int span = max(1, 256 / p.BNc);
    const uint32_t idesc = umma_idesc_bf16(128, span * p.BNc, 1, 1);
    // MN-major SWIZZLE_128B: LBO = distance between 64-channel boxes, SBO = 8 pixel rows = 1024 B
    // (non-halo: the dz box and every x box hold kp pixel rows of 128 B; kp = 32, or gh*gw*bn for small odd grids)
    const uint64_t desc_hi = ((uint64_t)(dz_box >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t s_base = smem_u32(smem) >> 4;
    int stage = 0; uint32_t phase = 0;
    for (int pt = pt0; pt < pt1; ++pt) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t s = s_base + (uint32_t)stage * (stage_bytes >> 4);
      if (p.halo) {
        if (elect_one()) {
          // A (dz): MN-major, LBO = dz box; B (x): MN-major, LBO = x box; row tap r = + (dy_r - dy0) * width pixel rows
          const uint64_t hiA = ((uint64_t)(dz_box >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          const uint64_t hiB = ((uint64_t)(x_box >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
          const uint32_t idh = umma_idesc_bf16(128, p.BNc, 1, 1);
          const uint32_t xb = s + ((2u * dz_box) >> 4);
          for (int ks = 0; ks < p.kp / 16; ++ks)
            for (int r = 0; r < 3; ++r) {
              const uint32_t b_lo = xb + (uint32_t)((p.dy[t0 + 3 * r] - p.dy0) * p.bw * 8) + (uint32_t)ks * 128u;
              umma_bf16(tmem_base + (uint32_t)(r * p.BNc), hiA | (uint64_t)(s + (uint32_t)ks * 128u), hiB | (uint64_t)b_lo, idh,
                        (pt > pt0 || ks > 0) ? 1u : 0u);
            }
          umma_commit(&empty[stage]);
          if (pt == pt1 - 1) umma_commit(done);
        }
      } else if (elect_one()) {
        // consecutive taps are consecutive 64-channel boxes in smem and consecutive column blocks in TMEM, so one
        // MMA spans `span` taps: N = span*BNc <= 256 (an MMA costs the same ~82 ns for any N <= 256)
        for (int j = 0; j < nt; j += span) {
          const int sp = min(span, nt - j);
          const uint32_t b_lo = s + (uint32_t)(2 + j * nbB) * (dz_box >> 4);
          const uint32_t d = tmem_base + (uint32_t)(j * p.BNc);
          const uint32_t id = sp == span ? idesc : umma_idesc_bf16(128, sp * p.BNc, 1, 1);
          // a K step of 16 pixels = 2 row groups = +2048 B = +128 in the (>>4) start-address field
          for (int ks = 0; ks < p.kp / 16; ++ks)
            umma_bf16(d, desc_hi | (uint64_t)(s + (uint32_t)ks * 128u), desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 128u), id,
                      (pt > pt0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&empty[stage]);
        if (pt == pt1 - 1) umma_commit(done);
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
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
          // partial layout [split][t][ci][co]: for a fixed ci the 32 lanes write 32 consecutive co (one 128 B line)
          float* o = p.ws + (((int64_t)split * p.T + (t0 + j * tstep)) * p.Cin + ci0) * p.Cout + co;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (ci0 + i < p.Cin) o[(int64_t)i * p.Cout] = __uint_as_float(r[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dw[i] = beta*dw[i] + sum_splits ws[s][i]  (i over [T][Cin][Cout], same layout in and out -> fully coalesced)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int64_t tot, int64_t tot_store,
                                    float* __restrict__ dw, float beta) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot_store; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += ws[z * tot + i];
    dw[i] = (beta != 0.f ? beta * dw[i] : 0.f) + s;
  }
}

// 16-byte accesses, and the split range is cut into 4 groups summed by different threads (64 outputs x 4 groups per CTA,
// folded through shared memory in a fixed order): the single-chain version is bound by load latency -- 49 dependent
// rounds for the 128->128 layers -- not by bandwidth.  Deterministic (fixed association), not the same rounding as the
// scalar kernel.
__global__ void __launch_bounds__(256) wgrad_reduce_v4_kernel(const float4* __restrict__ ws, int splits, int64_t tot4,
                                                              int64_t tot_store4, float4* __restrict__ dw, float beta) {
  pdl_entry();
  __shared__ float4 part[4][64];
  const int tx = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + tx;
  const int per = (splits + 3) / 4;
  const int z0 = g * per, z1 = min(splits, z0 + per);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < tot_store4) {
    int z = z0;
    for (; z + 4 <= z1; z += 4) {
      const float4 a = ws[(int64_t)z * tot4 + i], b = ws[(int64_t)(z + 1) * tot4 + i], c = ws[(int64_t)(z + 2) * tot4 + i],
                   d = ws[(int64_t)(z + 3) * tot4 + i];
      s.x = ((s.x + a.x) + b.x) + c.x + d.x; s.y = ((s.y + a.y) + b.y) + c.y + d.y;
      s.z = ((s.z + a.z) + b.z) + c.z + d.z; s.w = ((s.w + a.w) + b.w) + c.w + d.w;
    }
    for (; z < z1; ++z) {
      const float4 a = ws[(int64_t)z * tot4 + i];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  part[g][tx] = s;
  __syncthreads();
  if (g == 0 && i < tot_store4) {
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float4 a = part[k][tx];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    if (beta != 0.f) {
      const float4 o = dw[i];
      s.x += beta * o.x; s.y += beta * o.y; s.z += beta * o.z; s.w += beta * o.w;
    }
    dw[i] = s;
  }
}

// dst[t][n][k] (bf16, k < Kpad) = k < K ? src[tap(t)*st + n*sn + k*sk] : 0
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int T, int Nr, int K,
                                   int Kpad, int64_t st, int64_t sn, int64_t sk, const int* __restrict__ taps,
                                   int64_t total) {
  pdl_entry();
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
  tw = 4; while (tw < gw && tw < total) tw <<= 1;   // >= 4: the epilogue is specialised for tile widths 4, 8, 16, >= 32
  th = 1; while (th < gh && th * tw < total) th <<= 1;
  nb = total / (tw * th);
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_igemm_bf16(const tgan_igemm_args* a, void* stream) {
  TGAN_CHECK_ARG(a && a->x && a->wp && a->out, "igemm: null pointer");
  TGAN_CHECK_ARG(a->T >= 1 && a->T <= 25, "igemm: T out of range");
  TGAN_CHECK_ARG(a->ncls >= 0 && a->ncls <= 4, "igemm: at most 4 output classes");
  TGAN_CHECK_ARG(a->ldx % 8 == 0 && a->Kpad % 8 == 0, "igemm: ldx (%d) and Kpad (%d) must be multiples of 8", a->ldx, a->Kpad);
  TGAN_CHECK_ARG(((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->wp & 15) == 0, "igemm: operands must be 16B aligned");
  TGAN_CHECK_ARG(a->C >= 1 && a->C <= a->Kpad && a->Nout >= 1, "igemm: bad channel counts");
  IgParams p;
  memset(&p, 0, sizeof(p));
  const int sy = a->sy > 0 ? a->sy : 1, sx = a->sx > 0 ? a->sx : 1;
  pick_tile(a->gh, a->gw, p.th, p.tw, p.nb, 128);
  TGAN_CHECK_ARG(p.tw * sx <= 256 && p.th * sy <= 256, "igemm: strided box too large");
  TGAN_CHECK_ARG(p.tw >= 4, "igemm: output grid narrower than 4 pixels is not supported");
  TGAN_CHECK_ARG((int64_t)a->N * a->OH * a->OW * a->ldo < (1ll << 31), "igemm: output larger than 2^31 elements");
  p.ltw = 0; while ((1 << p.ltw) < p.tw) ++p.ltw;
  p.lppi = 0; while ((1 << p.lppi) < p.tw * p.th) ++p.lppi;
  p.N = a->N; p.sy = sy; p.sx = sx;
  p.tiles_y = ceil_div(a->gh, p.th); p.tiles_x = ceil_div(a->gw, p.tw);
  p.m_tiles = ceil_div(a->N, p.nb) * p.tiles_y * p.tiles_x;
  p.nbox = 2;
  p.pp_tiles = ceil_div(p.m_tiles, 2);
  p.ct_tiles = ceil_div(a->Nout, 128);
  p.T = a->T; p.kchunks = ceil_div(a->C, 64);
  p.klast = ceil_div(a->C - (p.kchunks - 1) * 64, 16);
  for (int t = 0; t < a->T; ++t) { p.dy[t] = a->dy[t]; p.dx[t] = a->dx[t]; }
  p.out = a->out; p.odt = a->odt; p.OH = a->OH; p.OW = a->OW; p.ldo = a->ldo;
  p.osy = a->osy > 0 ? a->osy : 1; p.osx = a->osx > 0 ? a->osx : 1; p.ooy = a->ooy; p.oox = a->oox;
  p.vh = a->vh > 0 ? a->vh : a->gh; p.vw = a->vw > 0 ? a->vw : a->gw; p.Nout = a->Nout;
#ifdef TGAN_DEBUG_SWITCHES
  { const char* e = getenv("TGAN_IGEMM_DBG"); p.dbg = e ? atoi(e) : 0; }
#else
  p.dbg = 0;
#endif
  p.nseg = a->nseg > 1 ? a->nseg : 1;
  TGAN_CHECK_ARG(p.nseg <= 4, "igemm: at most 4 batch segments");
  for (int i = 0; i < 4; ++i) p.seg_end[i] = (a->nseg > 1 && i < a->nseg - 1) ? a->seg_end[i] : 0x7fffffff;
  p.segflat = (a->nseg > 1 && a->N == 1 && a->gh == 1) ? 1 : 0;
  p.bias = a->bias; p.colsum = reinterpret_cast<long long*>(a->colsum); p.act = a->act; p.alpha = a->alpha == 0.f ? 1.f : a->alpha;
  const size_t smem_bytes = 1024 + (size_t)IG_STAGES * IG_STAGE_BYTES + IG_OUT_STAGE_BYTES + 256;
  p.tstore = (a->odt == TGAN_BF16 && a->ldo % 8 == 0 && ((uintptr_t)a->out & 15) == 0) ? 1 : 0;
  { const char* e = getenv("TGAN_IGEMM_NO_TSTORE"); if (e && atoi(e)) p.tstore = 0; }

  p.bias_seg = a->bias_seg ? 1 : 0;
  p.clsum = reinterpret_cast<long long*>(a->clsum);
  p.mask_out = reinterpret_cast<uint32_t*>(a->mask_out);
  p.mask_in = reinterpret_cast<const uint32_t*>(a->mask_in);
  p.mask_alpha = a->mask_alpha;
  const bool fused = p.bias_seg || p.clsum || p.mask_out || p.mask_in;
  if (fused) {
    TGAN_CHECK_ARG(p.nb == 1 && (p.tiles_x == 1 || a->gh == 1) && a->ncls <= 1 && p.osy == 1 && p.osx == 1 && p.ooy == 0 &&
                       p.oox == 0 && p.vh == a->gh && p.vw == a->gw && a->OH == a->gh && a->OW == a->gw,
                   "igemm: fused mean-only-BN epilogue needs linear output pixels (full-width tiles of one image or a flat GEMM)");
    TGAN_CHECK_ARG(!p.bias_seg || a->bias, "igemm: bias_seg without bias");
    TGAN_CHECK_ARG((!p.clsum && !p.mask_out) || p.tstore, "igemm: class sums / masks need the bf16 TMA-store epilogue");
    if (p.clsum) {
      TGAN_CHECK_ARG((a->cls_w == 16 || a->cls_w == 32) && a->cls_h >= 2 && (a->cls_h & (a->cls_h - 1)) == 0,
                     "igemm: class sums need cls_w in {16, 32} and a power-of-two cls_h");
      p.cls_lw = a->cls_w == 32 ? 5 : 4;
      p.cls_lh = 0; while ((1 << p.cls_lh) < a->cls_h) ++p.cls_lh;
    }
    if (p.segflat)
      for (int i = 0; i + 1 < a->nseg; ++i)
        TGAN_CHECK_ARG(a->seg_end[i] % 256 == 0, "igemm: flat segments must end on multiples of 256 pixels for the fused epilogue");
  }

  p.ncls = a->ncls > 1 ? a->ncls : 1;
  if (a->ncls > 1) {
    int t0 = 0;
    for (int c = 0; c < p.ncls; ++c) {
      TGAN_CHECK_ARG(a->cls_T[c] >= 1, "igemm: empty output class");
      p.cT[c] = a->cls_T[c]; p.ctap0[c] = t0; p.cooy[c] = a->cls_ooy[c]; p.coox[c] = a->cls_oox[c];
      t0 += a->cls_T[c];
    }
    TGAN_CHECK_ARG(t0 == a->T, "igemm: class tap counts must add up to T");
    TGAN_CHECK_ARG(a->colsum == nullptr, "igemm: channel sums are not supported with output classes");
  } else {
    p.cT[0] = a->T; p.ctap0[0] = 0; p.cooy[0] = p.ooy; p.coox[0] = p.oox;
  }

  // small problems: with 256-pixel tiles fewer than 148 CTAs would have work, and an N = 128 MMA runs at the same rate
  if (2 * p.pp_tiles * p.ct_tiles * p.ncls <= 148 && p.m_tiles > 1) { p.nbox = 1; p.pp_tiles = p.m_tiles; }   // still one wave
  { const char* e = getenv("TGAN_IGEMM_NBOX"); if (e && atoi(e) == 2) { p.nbox = 2; p.pp_tiles = ceil_div(p.m_tiles, 2); } }
  // row-halo eligibility: a full 3x3 tap grid (row-major, unit steps), stride 1, one class, full-width tiles of one image
  // whose two 128-pixel boxes are vertically adjacent
  p.halo = 0;
  if (p.nbox == 2 && a->T == 9 && sy == 1 && sx == 1 && p.ncls == 1 && p.nb == 1 && p.tiles_x == 1 && p.tiles_y % 2 == 0 && p.tw >= 8 &&
      (2 * p.th + 2) * p.tw * 128 <= HALO_A_SLOT_BYTES) {
    bool grid = true;
    const int sr = a->dy[3] - a->dy[0], sc = a->dx[1] - a->dx[0];
    for (int t = 0; t < 9; ++t)
      grid = grid && a->dy[t] == a->dy[0] + (t / 3) * sr && a->dx[t] == a->dx[0] + (t % 3) * sc;
    if (grid && (sr == 1 || sr == -1) && (sc == 1 || sc == -1)) {
      p.halo = 1;
      p.halo_rows = 2 * p.th + 2;
      p.halo_dy0 = sr == 1 ? a->dy[0] : a->dy[6];
    }
  }
  { const char* e = getenv("TGAN_IGEMM_NO_HALO"); if (e && atoi(e)) p.halo = 0; }
  // short K loops (conv1_1 / D's first conv through im2col, small dense layers: <= 2 ring stages per tile) are bound by the
  // epilogue waiting for the previous box's TMA store to release the single staging buffer, not by load latency
  {
    int max_t = a->T;
    if (a->ncls > 1) { max_t = 0; for (int c = 0; c < a->ncls; ++c) max_t = a->cls_T[c] > max_t ? a->cls_T[c] : max_t; }
    p.nstages = (!p.halo && p.tstore && max_t * p.kchunks <= 2 && !getenv("TGAN_IGEMM_RING4")) ? 3 : IG_STAGES;
  }

  CUtensorMap tmX, tmW, tmXh, tmO[4];
  {
    uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->N};
    uint64_t str[3] = {(uint64_t)a->ldx * 2, (uint64_t)a->W * a->ldx * 2, (uint64_t)a->H * a->W * a->ldx * 2};
    uint32_t box[4] = {64, (uint32_t)(p.tw * sx), (uint32_t)(p.th * sy), (uint32_t)p.nb};
    uint32_t es[4] = {1, (uint32_t)sx, (uint32_t)sy, 1};
    if (make_tmap_bf16(&tmX, a->x, 4, dims, str, box, es)) return 1;
    tmXh = tmX;
    if (p.halo) {
      uint32_t hbox[4] = {64, (uint32_t)p.tw, (uint32_t)p.halo_rows, 1};
      if (make_tmap_bf16(&tmXh, a->x, 4, dims, str, hbox, nullptr)) return 1;
    }
  }
  {
    uint64_t dims[3] = {(uint64_t)a->Kpad, (uint64_t)a->Nout, (uint64_t)a->T};
    uint64_t str[2] = {(uint64_t)a->Kpad * 2, (uint64_t)a->Nout * a->Kpad * 2};
    uint32_t box[3] = {64, 128, 1};
    if (make_tmap_bf16(&tmW, a->wp, 3, dims, str, box, nullptr)) return 1;
  }
  for (int c = 0; c < 4; ++c) {
    if (!p.tstore || c >= p.ncls) { tmO[c] = tmX; continue; }
    // the valid output grid as a strided 4-D view [Nout, vw, vh, N] of out (parity-class placement = base offset +
    // pixel strides); staging rows are dense [pixel][128 channels], so the map is not swizzled
    const uint64_t ldo = (uint64_t)a->ldo;
    uint64_t dims[4] = {(uint64_t)a->Nout, (uint64_t)p.vw, (uint64_t)p.vh, (uint64_t)a->N};
    uint64_t str[3] = {(uint64_t)p.osx * ldo * 2, (uint64_t)p.osy * a->OW * ldo * 2, (uint64_t)a->OH * a->OW * ldo * 2};
    uint32_t box[4] = {128, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.nb};
    const char* base = (const char*)a->out + ((int64_t)p.cooy[c] * a->OW + p.coox[c]) * (int64_t)ldo * 2;
    if (make_tmap_bf16(&tmO[c], base, 4, dims, str, box, nullptr, false)) return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(igemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(igemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(igemm_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    TGAN_CHECK_ARG(e == cudaSuccess, "igemm: cannot set max dynamic smem: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  p.d_tiles_x.set(p.tiles_x); p.d_tiles_y.set(p.tiles_y); p.d_tpi.set(p.tiles_x * p.tiles_y);
  p.d_ct.set(p.ct_tiles); p.d_tpc.set(p.pp_tiles * p.ct_tiles);
  const int total = p.pp_tiles * p.ct_tiles * p.ncls;
  const int grid = total < 148 ? total : 148;
  // epilogue variant (see igemm_kernel): the pruned instantiations are exact subsets of the full one
  static const bool one_kernel = getenv("TGAN_IGEMM_ONE_KERNEL") != nullptr;
  const bool common = p.tstore && (p.act == TGAN_ACT_NONE || p.act == TGAN_ACT_LRELU || p.act == TGAN_ACT_RELU) && !one_kernel;
  const bool ffwd = p.bias_seg || p.clsum || p.mask_out, fbwd = p.mask_in != nullptr;
#define TGAN_IG_LAUNCH(F, R) pdl_launch(igemm_kernel<F, R>, grid, IG_THREADS, smem_bytes, (cudaStream_t)stream, tmX, tmW, tmXh, tmO[0], tmO[1], tmO[2], tmO[3], p)
  if (!common || (ffwd && fbwd)) TGAN_IG_LAUNCH(3, true);
  else if (ffwd) TGAN_IG_LAUNCH(1, false);
  else if (fbwd) TGAN_IG_LAUNCH(2, false);
  else TGAN_IG_LAUNCH(0, false);
#undef TGAN_IG_LAUNCH
  TGAN_LAUNCHED();
  return 0;
}

static size_t wg_stage_bytes(const WgParams& p) {
  return (size_t)2 * p.kp * 128 + (size_t)((p.halo ? 1 : p.tg) * (p.BNc / 64)) * p.xrows * 128;
}

static int wgrad_plan(const tgan_wgrad_args* a, WgParams& p) {
  memset(&p, 0, sizeof(p));
  const int sy = a->sy > 0 ? a->sy : 1, sx = a->sx > 0 ? a->sx : 1;
  p.N = a->N; p.sy = sy; p.sx = sx; p.T = a->T; p.Cout = a->Cout; p.Cin = a->Cin;
  // row-halo eligibility: full 3x3 tap grid in row-major order with unit steps, stride 1, width a power of two in
  // [16, 128] that divides 128, height a multiple of the 128 / width rows of a pixel tile
  p.halo = 0;
  if (a->T == 9 && sy == 1 && sx == 1 && a->gw >= 16 && a->gw <= 128 && (a->gw & (a->gw - 1)) == 0 && a->W == a->gw &&
      a->gh % (128 / a->gw) == 0 && a->Cin >= 64) {
    bool grid = true;
    const int sr = a->dy[3] - a->dy[0], sc = a->dx[1] - a->dx[0];
    for (int t = 0; t < 9; ++t)
      grid = grid && a->dy[t] == a->dy[0] + (t / 3) * sr && a->dx[t] == a->dx[0] + (t % 3) * sc;
    if (grid && (sr == 1 || sr == -1) && (sc == 1 || sc == -1)) { p.halo = 1; p.dy0 = sr == 1 ? a->dy[0] : a->dy[6]; }
  }
  { const char* e = getenv("TGAN_WGRAD_NO_HALO"); if (e && atoi(e)) p.halo = 0; }
  if (p.halo) {
    p.kp = 128; p.bw = a->gw; p.bh = 128 / a->gw; p.bn = 1;
    p.xrows = (p.bh + 2) * p.bw;
  } else {
    pick_tile(a->gh, a->gw, p.bh, p.bw, p.bn, WG_KP);
    p.kp = WG_KP; p.xrows = WG_KP;
    // small grids that are not powers of two (conv3: 6x6 VALID outputs): power-of-two pixel tiles would carry 44 % padding
    // through the tensor pipe.  Take the exact grid of bn images as one K tile when that is a multiple of 16 pixels.
    const bool pow2 = (a->gh & (a->gh - 1)) == 0 && (a->gw & (a->gw - 1)) == 0;
    if (!pow2 && a->gh <= 16 && a->gw <= 16 && !getenv("TGAN_WGRAD_NO_EXACT")) {
      for (int bn = 1; bn <= 8; ++bn) {
        const int kp = a->gh * a->gw * bn;
        if (kp % 16 == 0 && kp <= 256 && kp >= 64) {
          p.bh = a->gh; p.bw = a->gw; p.bn = bn; p.kp = kp; p.xrows = kp;
          break;
        }
      }
    }
  }
  if (p.bw * sx > 256 || p.bh * sy > 256) { set_error("wgrad: strided box too large"); return 1; }
  p.ptiles_y = ceil_div(a->gh, p.bh); p.ptiles_x = ceil_div(a->gw, p.bw);
  p.p_tiles = ceil_div(a->N, p.bn) * p.ptiles_y * p.ptiles_x;
  p.co_tiles = ceil_div(a->Cout, 128);
  if (p.halo) {
    p.BNc = a->Cin <= 64 ? 64 : 128;        // three row taps x BNc accumulator columns <= 512
    p.ci_tiles = ceil_div(a->Cin, p.BNc);
    p.tg = 3;
    p.tap_groups = 3;                        // one per column offset
  } else {
    p.BNc = a->Cin <= 64 ? 64 : a->Cin <= 128 ? 128 : 256;
    p.ci_tiles = ceil_div(a->Cin, p.BNc);
    p.tg = 512 / p.BNc;
    if (p.tg > a->T) p.tg = a->T;
    p.tg = ceil_div(a->T, ceil_div(a->T, p.tg));     // balanced tap groups (T=9, max 4 -> 3+3+3)
    // smem: stages * (2 + tg*BNc/64) boxes of kp * 128 B, at least two stages
    while (p.tg > 1 && (size_t)2 * (2 + p.tg * (p.BNc / 64)) * p.kp * 128 > 200 * 1024) --p.tg;
    p.tap_groups = ceil_div(a->T, p.tg);
  }
  const int base = p.co_tiles * p.ci_tiles * p.tap_groups;
  int splits = 148 / base;            // one wave of equal-work CTAs; fewer splits = fewer fp32 partials to fold
  if (splits > p.p_tiles) splits = p.p_tiles;
  if (splits < 1) splits = 1;
  // keep >= 8 pixel tiles (2 of the 4x larger row-halo tiles) per split so the pipeline prologue is amortised
  while (splits > 1 && p.p_tiles / splits < (p.halo ? 2 : 8)) --splits;
  if (a->ws_bytes > 0) {   // clamp to the caller's workspace
    const int64_t per_split = (int64_t)a->T * a->Cout * a->Cin * 4;
    while (splits > 1 && (int64_t)splits * per_split > a->ws_bytes) --splits;
  }
  p.splits = splits;
  const size_t stage_bytes = wg_stage_bytes(p);
  int stages = (int)((232448 - 2048) / stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  for (int t = 0; t < a->T; ++t) { p.dy[t] = a->dy[t]; p.dx[t] = a->dx[t]; }
  return 0;
}

extern "C" int tgan_sizeof_igemm_args(void) { return (int)sizeof(tgan_igemm_args); }
extern "C" int tgan_sizeof_wgrad_args(void) { return (int)sizeof(tgan_wgrad_args); }

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
    uint32_t box[4] = {64, (uint32_t)(p.bw * p.sx), (uint32_t)((p.halo ? p.bh + 2 : p.bh) * p.sy), (uint32_t)p.bn};
    uint32_t es[4] = {1, (uint32_t)p.sx, (uint32_t)p.sy, 1};
    if (make_tmap_bf16(&tmX, a->x, 4, dims, str, box, es)) return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    TGAN_CHECK_ARG(e == cudaSuccess, "wgrad: cannot set max dynamic smem: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t stage_bytes = wg_stage_bytes(p);
  const size_t smem_bytes = 1024 + p.stages * stage_bytes + 256;
  const int grid = p.co_tiles * p.ci_tiles * p.tap_groups * p.splits;
  pdl_launch(wgrad_kernel, grid, WG_THREADS, smem_bytes, (cudaStream_t)((cudaStream_t)stream), tmDz, tmX, p);
  TGAN_LAUNCHED();
  const int64_t tot = (int64_t)a->T * a->Cout * a->Cin;
  TGAN_CHECK_ARG(a->cin_store == 0 || (a->T == 1 && a->cin_store <= a->Cin), "wgrad: cin_store needs T == 1");
  const int64_t tot_store = a->cin_store > 0 ? (int64_t)a->cin_store * a->Cout : tot;
  if (tot % 4 == 0 && tot_store % 4 == 0 && ((uintptr_t)a->ws & 15) == 0 && ((uintptr_t)a->dw & 15) == 0) {
    const int rg = ceil_div(tot_store / 4, 64);
    pdl_launch(wgrad_reduce_v4_kernel, rg, 256, 0, (cudaStream_t)stream, (const float4*)a->ws, p.splits, tot / 4, tot_store / 4,
               (float4*)a->dw, a->beta);
  } else {
    int rg = ceil_div(tot_store, 256);
    if (rg > 148 * 8) rg = 148 * 8;
    pdl_launch(wgrad_reduce_kernel, rg, 256, 0, (cudaStream_t)stream, a->ws, p.splits, tot, tot_store, a->dw, a->beta);
  }
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_pack_weight_bf16(const float* src, void* dst, int T, int Nrows, int K, int Kpad, int64_t st,
                                     int64_t sn, int64_t sk, const int* taps_dev, void* stream) {
  TGAN_CHECK_ARG(src && dst && T >= 1 && Nrows >= 1 && K >= 1 && Kpad >= K && Kpad % 8 == 0, "pack_weight: bad args");
  const int64_t total = (int64_t)T * Nrows * Kpad;
  pdl_launch(pack_weight_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)((cudaStream_t)stream), src, (bf16*)dst, T, Nrows, K, Kpad, st, sn,
                                                                            sk, taps_dev, total);
  TGAN_LAUNCHED();
  return 0;
}
