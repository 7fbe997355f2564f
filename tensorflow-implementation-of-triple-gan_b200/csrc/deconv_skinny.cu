// deconv_skinny.cu -- the generator's last layer (Good_GAN_cifar10.py:56: tf.layers.conv2d_transpose 5x5 / stride 2 / SAME,
// 138 -> 3 channels, then tanh) as a kernel of its own.
//
// On the tcgen05 implicit GEMM (output channels on the 128 UMMA lanes) this layer uses 3 of 128 lanes: 44-58 us per call,
// 0.01 of the tensor peak ("What's missing" 9).  Here the PIXELS are the MMA rows: one CTA per image keeps the
// whole 16x16x138 input (+1 halo) and the bf16 filter in shared memory; a warp owns output rows, and for each of the two
// column parities of a row runs mma.sync.m16n8k16 (bf16 in, fp32 accumulate) over the taps that reach that parity
// (2 or 3 row taps x 2 or 3 column taps) -- 16 same-parity output pixels x 8 padded output channels per instruction.
// The legacy warp-level MMA is the right tool: the whole layer is 0.3 GFLOP and latency bound, not tensor-pipe bound.
//
//   y[n, oy, ox, co] = tanh(b[co] + sum_{r,c,ci} x[n, (oy + pt - r)/2, (ox + pl - c)/2, ci] * W[r, c, co, ci])   (taps with an
//   even numerator only; pt = pl = 1 for 5x5 / s2 SAME)
//
// Roofline: latency / L2 (74 KB input + 41 KB filter per CTA); algorithmic FLOPs 2 * N * 32*32 * 6.25 * 138 * 3.
#include <mma.h>

#include "common.cuh"

namespace tgan {

constexpr int DS_PIX = 152;      // padded channel stride of a staged pixel (304 B: ldmatrix rows fall in distinct banks)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// x: bf16 [N, 16, 16, ldx] (first Cin channels used); w: fp32 [5, 5, Cout, Cin]; y: fp32 [N, 32, 32, Cout]; Cout <= 8
__global__ void __launch_bounds__(256, 1) deconv5s2_skinny_kernel(const bf16* __restrict__ x, int ldx, int Cin,
                                                                  const float* __restrict__ w, const float* __restrict__ bias,
                                                                  int Cout, float* __restrict__ y, int act) {
  pdl_entry();
  extern __shared__ __align__(16) uint8_t ds_smem[];
  bf16* xs = reinterpret_cast<bf16*>(ds_smem);                        // [18][18][DS_PIX]
  bf16* ws = xs + 18 * 18 * DS_PIX;                                   // [25][144][8]
  const int n = blockIdx.x, tid = threadIdx.x;
  // ---- stage: zero everything (halo, channel padding), then the image (cp.async: every 16-byte piece in flight at once)
  // and the filter (eight independent loads per thread and round; one load per round made this phase latency bound)
  {
    uint4* z = reinterpret_cast<uint4*>(ds_smem);
    const int nz = (18 * 18 * DS_PIX + 25 * 144 * 8) * 2 / 16;
    for (int i = tid; i < nz; i += 256) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    const int vfull = Cin / 8, vpp = (Cin + 7) / 8;                   // whole / all 16-byte vectors per pixel with real channels
    const bf16* xi = x + (int64_t)n * 256 * ldx;
    for (int i = tid; i < 256 * vfull; i += 256) {
      const int px = i / vfull, v = i - px * vfull;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(xs + ((px / 16 + 1) * 18 + (px % 16 + 1)) * DS_PIX + v * 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(xi + (int64_t)px * ldx + v * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (vpp > vfull && tid < 256) {                                   // the last, partial vector of every pixel
      const int px = tid, v = vfull;
      uint4 u = *reinterpret_cast<const uint4*>(xi + (int64_t)px * ldx + v * 8);
      bf16* e = reinterpret_cast<bf16*>(&u);
      for (int j = 0; j < 8; ++j) if (v * 8 + j >= Cin) e[j] = __float2bfloat16_rn(0.f);
      *reinterpret_cast<uint4*>(xs + ((px / 16 + 1) * 18 + (px % 16 + 1)) * DS_PIX + v * 8) = u;
    }
    const int total = 25 * Cout * Cin;                                // w[t][co][ci] -> ws[t][ci][co]
    for (int base = 0; base < total; base += 256 * 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int i = base + k * 256 + tid; v[k] = i < total ? __ldg(w + i) : 0.f; }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = base + k * 256 + tid;
        if (i < total) {
          const int ci = i % Cin, co = (i / Cin) % Cout, t = i / (Cin * Cout);
          ws[(t * 144 + ci) * 8 + co] = __float2bfloat16_rn(v[k]);
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  const int ksteps = (Cin + 15) / 16;
  const int g = lane >> 2, q = lane & 3;
  float b0 = 0.f, b1 = 0.f;
  if (bias) { b0 = 2 * q < Cout ? bias[2 * q] : 0.f; b1 = 2 * q + 1 < Cout ? bias[2 * q + 1] : 0.f; }
  for (int oy = warp; oy < 32; oy += 8) {
    const int py = oy & 1;
    // both column parities of the row advance together: two independent accumulator chains hide the MMA latency
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int r = (py + 1) & 1; r < 5; r += 2) {
      const int iy = (oy + 1 - r) / 2;                                // exact: the numerator is even; -1 <= iy <= 16
      const bf16* xrow = xs + ((iy + 1) * 18 + (lane & 7) + 8 * ((lane >> 3) & 1)) * DS_PIX + 8 * (lane >> 4);
      const bf16* wrow = ws + (r * 5 * 144 + (lane & 15)) * 8;
      // px = 0 uses column taps c = 1, 3 (ix0 = 0, -1); px = 1 uses c = 0, 2, 4 (ix0 = 1, 0, -1)
      for (int ks = 0; ks < ksteps; ++ks) {
#pragma unroll
        for (int c = 0; c < 5; ++c) {                                 // consecutive MMAs alternate between the two chains
          const int px = (c & 1) ^ 1;                                 // parity of the output columns tap c reaches
          const int ix0 = (px + 1 - c) / 2;                           // input column of the first such output column
          uint32_t a[4], b[2];
          ldsm_x4(a, xrow + (ix0 + 1) * DS_PIX + ks * 16);
          ldsm_x2_trans(b, wrow + (c * 144 + ks * 16) * 8);
          mma_bf16_16816(acc[px], a, b);
        }
      }
    }
    // C fragment: rows g and g + 8 (output columns px + 2g, px + 2(g + 8)), channels 2q, 2q + 1
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ox = px + 2 * (g + 8 * h);
        float v0 = acc[px][2 * h] + b0, v1 = acc[px][2 * h + 1] + b1;
        if (act == TGAN_ACT_TANH) { v0 = tanhf(v0); v1 = tanhf(v1); }
        float* o = y + (((int64_t)n * 32 + oy) * 32 + ox) * Cout;
        if (2 * q < Cout) o[2 * q] = v0;
        if (2 * q + 1 < Cout) o[2 * q + 1] = v1;
      }
  }
}

}  // namespace tgan

using namespace tgan;

extern "C" int tgan_deconv5s2_skinny(const void* x, int N, int ldx, int Cin, const float* w, const float* bias, int Cout,
                                     float* y, int act, void* stream) {
  TGAN_CHECK_ARG(x && w && y && N > 0 && Cin >= 1 && Cin <= 144 && ldx >= Cin && ldx % 8 == 0 && Cout >= 1 && Cout <= 8 &&
                     ((uintptr_t)x & 15) == 0 && (act == TGAN_ACT_NONE || act == TGAN_ACT_TANH),
                 "deconv5s2_skinny: bf16 [N,16,16,ldx] input with Cin <= 144, ldx %% 8 == 0, Cout <= 8, act none / tanh");
  const size_t smem = (size_t)(18 * 18 * DS_PIX + 25 * 144 * 8) * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(deconv5s2_skinny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TGAN_CHECK_ARG(e == cudaSuccess, "deconv5s2_skinny: cannot set max dynamic smem: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  pdl_launch(deconv5s2_skinny_kernel, N, 256, smem, (cudaStream_t)stream, (const bf16*)x, ldx, Cin, w, bias, Cout, y, act);
  TGAN_LAUNCHED();
  return 0;
}
