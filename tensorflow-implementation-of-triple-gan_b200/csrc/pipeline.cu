// pipeline.cu -- the data formats either side of the training step:
//   * device-side batch formation from a uint8 dataset resident in HBM: gather by index + the reference's pixel map
//     (Input_Pipeline/cifar10Dataset.py:56-62, svhnDataset.py:59-65: x/255*2-1; mnistDataset.py:60-67: x/255) + one-hot
//     labels; latent / label draws of Train_goodGAN.py:232-237; the sample grid of utils.py:199-231 (merge);
//   * CRC-32C (Castagnoli) on the host for the TF tensor-bundle checkpoint files (Training/Saver.py:29-36 -> tf.train.Saver).
#include "common.cuh"

namespace tgan {
static inline int grid_for(int64_t n, int block = 256) {
  int64_t g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// 16 pixels bytes per thread: one 16-byte load, four 16-byte stores.  mode 0: v/255*2-1, mode 1: v/255 (IEEE division and
// an exact doubling, so the result is bit-identical to TF's float32 `image / 255 * 2 - 1`).
__global__ void gather_images_u8_kernel(const uint8_t* __restrict__ data, int64_t elems, const int64_t* __restrict__ idx,
                                        int64_t n_total, int64_t nvec, int vec_per_img, float* __restrict__ out, int mode) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / vec_per_img;
    const int v = (int)(i - img * vec_per_img);
    int64_t src = idx ? idx[img] : img;
    src = src < 0 ? 0 : (src >= n_total ? n_total - 1 : src);
    const uint4 raw = *reinterpret_cast<const uint4*>(data + src * elems + (int64_t)v * 16);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float4* o = reinterpret_cast<float4*>(out + img * elems + (int64_t)v * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float p = __fdiv_rn((float)((w[q] >> (8 * b)) & 0xffu), 255.f);
        f[b] = mode == 0 ? __fsub_rn(__fmul_rn(p, 2.f), 1.f) : p;
      }
      o[q] = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
}

// tf.one_hot(label, depth=K) of the gathered rows
__global__ void gather_onehot_kernel(const int32_t* __restrict__ labels, const int64_t* __restrict__ idx, int64_t n_total,
                                     int n, int K, float* __restrict__ out) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * K) return;
  const int r = i / K, k = i - r * K;
  int64_t src = idx ? idx[r] : r;
  src = src < 0 ? 0 : (src >= n_total ? n_total - 1 : src);
  out[i] = labels[src] == k ? 1.f : 0.f;
}

// z ~ U(-1, 1) [n, zdim] and y = one_hot(randint(0, K)) [n, K] from the step's Philox counter
// (np.random.uniform / np.random.randint of Train_goodGAN.py:232-237 -- a different generator, same distributions)
__global__ void draw_latent_kernel(float* __restrict__ z, int n, int zdim, float* __restrict__ y, int K, uint64_t seed,
                                   const uint64_t* __restrict__ counter, uint64_t stream_id) {
  pdl_entry();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t ctr = counter ? *counter : 0;
  const int64_t nz4 = ((int64_t)n * zdim + 3) / 4;
  if (i < nz4) {
    const uint4 r = Philox(seed)((uint64_t)i, stream_id + (ctr << 20));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t e = i * 4 + j;
      if (e < (int64_t)n * zdim) z[e] = __fmaf_rn((float)(w[j] >> 8), 2.f / 16777216.f, -1.f);   // [-1, 1)
    }
  } else if (i < nz4 + n) {
    const int r0 = (int)(i - nz4);
    const uint4 r = Philox(seed)((uint64_t)r0 + (1ull << 40), stream_id + (ctr << 20));
    const int cls = (int)(((uint64_t)r.x * (uint64_t)K) >> 32);                                      // randint(0, K)
    for (int k = 0; k < K; ++k) y[(int64_t)r0 * K + k] = k == cls ? 1.f : 0.f;
  }
}

// utils.py:199-231 `merge` after `inverse_transform` ((x+1)/2): n images [n,H,W,C] -> one [gh*H, gw*W, C] grid, image k
// at row k / gw, column k % gw; cells without an image stay 0 (np.zeros).
__global__ void image_grid_kernel(const float* __restrict__ x, int n, int H, int W, int C, int gh, int gw,
                                  float* __restrict__ grid, int inverse) {
  pdl_entry();
  const int64_t total = (int64_t)gh * H * gw * W * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t t = i / C;
  const int gx = (int)(t % ((int64_t)gw * W));
  const int gy = (int)(t / ((int64_t)gw * W));
  const int k = (gy / H) * gw + gx / W;
  float v = 0.f;
  if (k < n) {
    v = x[(((int64_t)k * H + gy % H) * W + gx % W) * C + c];
    if (inverse) v = __fmul_rn(__fadd_rn(v, 1.f), 0.5f);
  }
  grid[i] = v;
}

// ---- CRC-32C, slicing-by-8 (host) ----
static uint32_t g_crc_tab[8][256];
static bool g_crc_init = false;
static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1u)));
    g_crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xffu];
  g_crc_init = true;
}
}  // namespace tgan

using namespace tgan;

extern "C" int tgan_gather_images_u8(const void* data, int64_t n_total, int64_t elems, const int64_t* idx, int n,
                                     float* out, int mode, void* stream) {
  TGAN_CHECK_ARG(data && out && n > 0 && n_total > 0, "gather_images_u8: null / empty");
  TGAN_CHECK_ARG(elems > 0 && elems % 16 == 0 && aligned16(data) && aligned16(out), "gather_images_u8: image size must be a multiple of 16 bytes, 16-byte aligned buffers");
  TGAN_CHECK_ARG(mode == 0 || mode == 1, "gather_images_u8: mode 0 (x/255*2-1) or 1 (x/255)");
  const int vpi = (int)(elems / 16);
  const int64_t nvec = (int64_t)n * vpi;
  pdl_launch(gather_images_u8_kernel, grid_for(nvec), 256, 0, (cudaStream_t)stream, (const uint8_t*)data, elems, idx, n_total,
             nvec, vpi, out, mode);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_gather_onehot(const int32_t* labels, int64_t n_total, const int64_t* idx, int n, int K, float* out,
                                  void* stream) {
  TGAN_CHECK_ARG(labels && out && n > 0 && K > 0 && n_total > 0, "gather_onehot: null / empty");
  pdl_launch(gather_onehot_kernel, ceil_div((int64_t)n * K, 256), 256, 0, (cudaStream_t)stream, labels, idx, n_total, n, K, out);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_draw_latent(float* z, int n, int zdim, float* y, int K, uint64_t seed, const uint64_t* counter,
                                uint64_t stream_id, void* stream) {
  TGAN_CHECK_ARG(z && y && n > 0 && zdim > 0 && K > 0, "draw_latent: null / empty");
  const int64_t work = ((int64_t)n * zdim + 3) / 4 + n;
  pdl_launch(draw_latent_kernel, ceil_div(work, 256), 256, 0, (cudaStream_t)stream, z, n, zdim, y, K, seed, counter, stream_id);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" int tgan_image_grid(const float* x, int n, int H, int W, int C, int gh, int gw, float* grid, int inverse,
                               void* stream) {
  TGAN_CHECK_ARG(x && grid && n > 0 && H > 0 && W > 0 && C > 0 && gh > 0 && gw > 0, "image_grid: null / empty");
  const int64_t total = (int64_t)gh * H * gw * W * C;
  pdl_launch(image_grid_kernel, ceil_div(total, 256), 256, 0, (cudaStream_t)stream, x, n, H, W, C, gh, gw, grid, inverse);
  TGAN_LAUNCHED();
  return 0;
}

extern "C" uint64_t tgan_crc32c(uint64_t crc, const void* data, int64_t n) {
  if (!g_crc_init) crc_init();
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = ~(uint32_t)crc;
  while (n > 0 && ((uintptr_t)p & 7)) { c = g_crc_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8); --n; }
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_tab[7][w & 0xff] ^ g_crc_tab[6][(w >> 8) & 0xff] ^ g_crc_tab[5][(w >> 16) & 0xff] ^ g_crc_tab[4][(w >> 24) & 0xff] ^
        g_crc_tab[3][(w >> 32) & 0xff] ^ g_crc_tab[2][(w >> 40) & 0xff] ^ g_crc_tab[1][(w >> 48) & 0xff] ^ g_crc_tab[0][w >> 56];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) c = g_crc_tab[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
  return (uint64_t)(uint32_t)~c;
}
