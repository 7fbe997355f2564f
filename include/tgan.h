/* tgan.h -- C ABI of libtgan.so: the B200 (sm_100a) Triple-GAN training hot path.
 *
 * The reference (Wenyuan-Vincent-Li/Tensorflow-Implementation-of-Triple-GAN) has no FFI layer: every
 * numerical call goes Python -> TensorFlow op.  Each entry point below replaces the TensorFlow op(s)
 * at the cited reference call sites (file:line relative to the reference repository root); the Python
 * mirror of the reference operator surface (tgan/nn.py, tgan/model_base.py) binds them with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  Every function returns 0 on success, non-zero on
 *     error; tgan_last_error() returns a thread-local message.  No exception crosses the ABI.
 *   - All pointers are DEVICE pointers owned by the caller (incl. workspaces); no entry point
 *     allocates, frees or synchronises.  Every op is asynchronous on `stream` (a cudaStream_t passed
 *     as void*), and is CUDA-graph capturable.
 *   - Activations are NHWC, conv filters HWIO [kh,kw,Cin,Cout], transposed-conv filters
 *     [kh,kw,Cout,Cin], dense [in,out] -- the TF variable layouts.  `ld*` = row (pixel) stride in elements.
 *   - dtype codes: TGAN_F32 = 0, TGAN_BF16 = 1.  Parameters, statistics, gradients of parameters and
 *     loss scalars are always fp32.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef TGAN_H_
#define TGAN_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TGAN_F32 0
#define TGAN_BF16 1

/* activation codes for the fused epilogues */
#define TGAN_ACT_NONE 0
#define TGAN_ACT_RELU 1      /* tf.nn.relu          Good_GAN_cifar10.py:41,47,52 */
#define TGAN_ACT_LRELU 2     /* leakyReLu a=0.2     Good_GAN_cifar10.py:19-27; tf.nn.leaky_relu Good_GAN.py:99 */
#define TGAN_ACT_TANH 3      /* tf.nn.tanh          Good_GAN_cifar10.py:57 */
#define TGAN_ACT_SIGMOID 4   /* tf.nn.sigmoid       Good_GAN_cifar10.py:99, Good_GAN.py:32 */
#define TGAN_ACT_SOFTPLUS 5  /* tf.nn.softplus      Good_GAN.py:22,27 */

const char* tgan_last_error(void);
int tgan_version(void);
/* number of kernels libtgan has launched in this process since the last reset (bench.py gpu_launches) */
int64_t tgan_launch_count(void);
void tgan_launch_count_reset(void);

/* ---------------------------------------------------------------- dense contractions (fp32 SIMT) --
 * C[M,N] = alpha * op(A) * op(B) + beta * C, row-major fp32.  Replaces tf.matmul / tf.layers.dense
 * (nn.py:553, modle_base.py:40,67) and, with tgan_im2col / tgan_col2im, tf.nn.conv2d and
 * tf.nn.conv2d_transpose and their gradients (nn.py:504; modle_base.py:102,149,161,250) in the
 * fp32 parity mode and for the skinny layers (Cout in {1,3,10}).  splits > 1 = deterministic split-K:
 * `ws` must hold splits*M*N floats. */
int tgan_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
               const float* B, int ldb, float beta, float* C, int ldc, int splits, float* ws,
               void* stream);

/* col[(n,ho,wo),(r,s,c)] = x[n, ho*sh + r - pt, wo*sw + s - pl, c] (0 outside).  x has dtype `xdt`,
 * pixel stride ldx; col is fp32 [N*Ho*Wo, kh*kw*C].  TF SAME/VALID geometry is the caller's pt/pl. */
int tgan_im2col(const void* x, int xdt, int N, int H, int W, int C, int ldx, int kh, int kw, int sh,
                int sw, int pt, int pl, int Ho, int Wo, float* col, void* stream);
/* bf16 im2col for the tensor-core path of convolutions with few input channels (conv1_1: 3 channels, D's first
 * conv: 13): col is bf16 [N*Ho*Wo, ldc], ldc % 8 == 0, columns >= kh*kw*C are zero.  The convolution then runs as one
 * plain tcgen05 GEMM with K = kh*kw*C instead of kh*kw narrow taps. */
int tgan_im2col_bf16(const void* x, int xdt, int N, int H, int W, int C, int ldx, int kh, int kw, int sh, int sw, int pt,
                     int pl, int Ho, int Wo, void* col, int ldc, void* stream);
/* adjoint of tgan_im2col (gather form, deterministic): x[n,h,w,c] = sum of the col entries that read it.
 * Writes dtype `xdt`; only channels [0,Cx) of the C channels in col are produced (label-concat slice). */
int tgan_col2im(const float* col, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int pt,
                int pl, int Ho, int Wo, void* x, int xdt, int Cx, int ldx, void* stream);

/* ---------------------------------------------------------------- implicit-GEMM tcgen05 path (bf16) --
 * One tcgen05/TMEM/TMA kernel family.  A = activations gathered by shifted-window TMA loads (zero fill =
 * TF padding), B = packed bf16 weights [T][Nout][Kpad], fp32 accumulation in TMEM.
 *   out[n, oy*osy + ooy, ox*osx + oox, co] = sum_t sum_ci x[n, oy*sy + dy[t], ox*sx + dx[t], ci] * Wp[t][co][ci]
 * for (oy,ox) on the [gh, gw] output grid.  Covers conv fprop / dgrad (stride 1), the four output-parity
 * sub-convolutions of the 5x5/s2 transposed conv (modle_base.py:250) and plain GEMMs (gh=1, T=1).
 * Replaces the same reference call sites as tgan_sgemm in the bf16 tensor-core mode. */
typedef struct {
  const void* x;       /* bf16 [N, H, W, ldx]                                  */
  int N, H, W, C, ldx; /* C = contraction channels actually used (<= ldx)      */
  const void* wp;      /* bf16 packed weights [T][Nout][Kpad], Kpad % 8 == 0   */
  int T, Nout, Kpad;
  int dy[25], dx[25];  /* tap offsets                                          */
  int gh, gw;          /* output grid per image (tile = th x tw pixels)        */
  int sy, sx;          /* input traversal stride: x[n, oy*sy + dy, ox*sx + dx]  (conv stride; 0 = 1) */
  void* out;           /* dtype odt, [N, OH, OW, ldo]                          */
  int odt, OH, OW, ldo;
  int osy, osx, ooy, oox; /* output placement (stride / offset)                */
  int vh, vw;          /* number of valid grid rows/cols actually stored       */
  const float* bias;   /* optional fp32 [Nout] added before the store          */
  void* colsum;        /* optional int64 [nseg][Nout], Q24 fixed point (value * 2^24): += per-channel sums of the stored
                          values, per batch segment.  Integer atomics: the result does not depend on the CTA order */
  int nseg;            /* 0/1 = whole batch; up to 4 segments of consecutive images (several network calls grouped
                          in one launch keep their own mean-only-BN statistics) */
  int seg_end[4];      /* exclusive image index where segment i ends (plain GEMM, N == 1 && gh == 1: pixel index) */
  int act;             /* TGAN_ACT_* applied after bias (before store)         */
  float alpha;
  int ncls;            /* 0/1 = one tap list.  2..4 = output classes run in ONE launch (the output-parity sub-convolutions
                          of a stride-2 transposed conv / stride-2 input gradient): class c owns the next cls_T[c] entries
                          of dy/dx/wp (sum = T) and writes at output offset (cls_ooy[c], cls_oox[c]) instead of (ooy, oox);
                          list the classes with the most taps first */
  int cls_T[4], cls_ooy[4], cls_oox[4];
  /* ---- mean-only batch norm fused into the contraction (nn.py:147-187, :504-518 without a pass of its own).  All four
   * need an output whose pixels are LINEAR in memory per launch tile (full-width row blocks of one image, or a flat
   * GEMM) and the TMA-store epilogue (bf16 output, ldo % 8 == 0); the entry point rejects other geometries. ---- */
  int bias_seg;        /* 1: bias is fp32 [nseg][Nout] -- segment s adds bias[s][:] (= b - mean_s, the mean known before
                          the launch from tgan_mobn_mean_from_sums) */
  void* clsum;         /* optional int64 Q24 [nseg][9][Nout]: += sums of the STORED values by border class, index
                          3 * (row first / interior / last) + (column first / interior / last) of a cls_h x cls_w image */
  int cls_h, cls_w;    /* powers of two; cls_w in {16, 32} */
  void* mask_out;      /* optional uint32 [ceil(pixels / 32)][Nout]: bit 31 - j of word (i, c) = SIGN bit of the stored
                          value of pixel 32 i + j, channel c (set = the negative side of the leaky ReLU) */
  const void* mask_in; /* optional, same layout: the stored value is multiplied by mask_alpha where the bit is set --
                          the input gradient passes through the PRODUCER's leaky ReLU inside this launch's epilogue */
  float mask_alpha;
} tgan_igemm_args;
int tgan_igemm_bf16(const tgan_igemm_args* a, void* stream);

/* The batch mean of a convolution output is a linear function of border-aware sums of its INPUT:
 *   mean_s[co] = sum_{t, ci} W[t][co][ci] * S_s[t][ci] / count_s,
 *   S_s[t = (r, c)][ci] = sum of class sums over the row classes tap row r reads x the column classes tap column c reads
 * (a 3x3 / stride 1 / SAME tap at row offset -1 never reads the last input row, +1 never the first; kh = kw = 1: all).
 * clsum: int64 Q24 [nseg][9][Cin] written by the producer (tgan_igemm_bf16.clsum, tgan_class_sums, the pooling kernel);
 * wp: the packed bf16 fprop weights the contraction itself reads (so the mean matches its arithmetic), element
 * (t, co, ci) at wp[t * w_tap_stride + co * w_co_stride + ci]: [T][Cout][Kpad] -> (Cout*Kpad, Kpad); the im2col form
 * [Cout][T*Cin padded] -> (Cin, Kpad).  T = 9 (3x3 SAME, taps row-major) or 1.  count[s] = output pixels of segment s.
 * Writes shift[s][co] = b[co] - mean_s[co] (the per-segment bias of the fused epilogue) and, in call order,
 * pop_mean = pop_mean*decay + mean_s*(1-decay) (nn.py:181). */
int tgan_mobn_mean_from_sums(const void* clsum, int nseg, const int64_t* count, const void* wp, int T, int Cout, int Cin,
                             int64_t w_tap_stride, int64_t w_co_stride, const float* b, float* pop_mean, float decay,
                             float* shift, void* stream);
/* border-class sums (same layout as tgan_igemm_bf16.clsum) of [N, H, W, C]: a small-channel fp32 / bf16 tensor (C <= 16,
 * values rounded to bf16 as the tensor-core path reads them: the classifier's 3-channel input) or a contiguous bf16
 * activation with C % 8 == 0 (the pooled + dropped tensor in front of conv2_1) */
int tgan_class_sums(const void* x, int xdt, int N, int H, int W, int C, int ld, int nseg, const int* seg_end_images,
                    void* clsum, void* stream);
/* int64 Q24 per-segment channel sums [nseg][C] (tgan_igemm_bf16.colsum) -> fp32 colsums[4][C] (unused segments 0) and
 * grad_acc[c] += sum over segments (the bias gradient of a mean-only-BN layer); grad_acc may be NULL */
int tgan_seg_sums_finalize(const void* q24, int nseg, int C, float* colsums, float* grad_acc, void* stream);

/* wgrad on tensor cores: dW[t][ci][co] (+)= sum_{n,oy,ox} dz[n,oy,ox,co] * x[n, oy + dy[t], ox + dx[t], ci]
 * (pixel dimension is the GEMM K; both operands are MN-major UMMA operands loaded by TMA).
 * Writes fp32 partials with split-K over pixels followed by a deterministic reduce. */
typedef struct {
  const void* dz;      /* bf16 [N, gh, gw, lddz]                               */
  int N, gh, gw, Cout, lddz;
  const void* x;       /* bf16 [N, H, W, ldx]                                  */
  int H, W, Cin, ldx;
  int sy, sx;          /* x traversal stride (0 = 1): x[n, oy*sy + dy[t], ox*sx + dx[t], ci] */
  int T; int dy[25], dx[25];
  float* dw;           /* fp32 [T][Cin][Cout] (== HWIO for a conv; == [kh,kw,Cout,Cin] of a transposed conv when
                          the operand roles are exchanged); dw = beta*dw + sum */
  float beta;
  float* ws; int64_t ws_bytes;
  int cin_store;       /* T == 1 only: store the first cin_store rows of dw (0 = all Cin); the x operand may carry zero
                          padding columns that dw has no room for */
} tgan_wgrad_args;
int tgan_wgrad_bf16(const tgan_wgrad_args* a, void* stream);
int64_t tgan_wgrad_workspace_bytes(const tgan_wgrad_args* a);
/* sizeof() of the two argument structs, for binding-side layout checks */
int tgan_sizeof_igemm_args(void);
int tgan_sizeof_wgrad_args(void);

/* The generator's last layer (Good_GAN_cifar10.py:56): 5x5 / stride 2 / SAME transposed convolution with <= 8 output
 * channels on a 16x16 input, bias and optional tanh fused: y[N,32,32,Cout] fp32 = act(conv2d_transpose(x, w) + bias).
 * x: bf16 [N,16,16,ldx] (first Cin <= 144 channels), w: fp32 [5,5,Cout,Cin] (the TF filter as it is).  Pixels are the MMA
 * rows (mma.sync m16n8k16), one CTA per image: replaces a 128-lane tcgen05 launch that used 3 lanes. */
int tgan_deconv5s2_skinny(const void* x, int N, int ldx, int Cin, const float* w, const float* bias, int Cout, float* y,
                          int act, void* stream);

/* weight preparation for the tcgen05 path: fp32 weights (any TF layout, addressed by strides) ->
 * bf16 K-major [T][Nrows][Kpad]:  dst[t][n][k] = k < K ? src[taps[t]*st + n*sn + k*sk] : 0
 * (taps_dev = NULL -> identity).  fprop HWIO: Nrows=Cout,K=Cin,sn=1,sk=Cout; dgrad HWIO: Nrows=Cin,K=Cout,
 * sn=Cout,sk=1; transposed-conv [T][Cout][Cin]: Nrows=Cout,K=Cin,sn=Cin,sk=1. */
int tgan_pack_weight_bf16(const float* src, void* dst, int T, int Nrows, int K, int Kpad, int64_t st, int64_t sn,
                          int64_t sk, const int* taps_dev, void* stream);

/* ---------------------------------------------------------------- multi-tensor weight preparation --
 * One launch for every layer of a network (descriptor tables live in DEVICE memory; the pointers inside are device
 * pointers, normally views into the flat per-network parameter / gradient buffers).
 *   weightnorm_fwd_multi : inv_norm[co], scale[co] = g[co]*inv_norm[co] for each tensor (same maths as
 *                          tgan_weightnorm_fwd, W is not materialised: tgan_pack_weight_multi folds `scale`)
 *   weightnorm_bwd_multi : dg[co] += <dW,V>*inv ; dV += g*inv*(dW - V*inv^2*<dW,V>)   (accumulating)
 *   pack_weight_multi    : tgan_pack_weight_bf16 for each entry, times scale[n] (scale_on = 1) or scale[k] (2)
 * max_co = largest Co in the table (sizes the grid). */
typedef struct {
  const float* V; const float* g; float* inv_norm; float* scale;
  const float* dW; float* dV; float* dg;      /* backward only */
  int A, Co, B, eps_mode;
} tgan_wn_desc;
typedef struct {
  const float* src; void* dst; const float* scale; const int* taps;
  int64_t st, sn, sk;
  int T, Nr, K, Kpad, scale_on, scale_mod;   /* scale_mod > 0: the scale index is taken modulo scale_mod (k = (tap, co) flattened) */
} tgan_pack_desc;
int tgan_weightnorm_fwd_multi(const tgan_wn_desc* descs_dev, int n, int max_co, float* ws, void* stream);
int tgan_weightnorm_bwd_multi(const tgan_wn_desc* descs_dev, int n, int max_co, float* ws, void* stream);
int64_t tgan_weightnorm_bwd_multi_ws_floats(int n, int max_co);   /* size of ws (fwd and bwd) */
int tgan_pack_weight_multi(const tgan_pack_desc* descs_dev, int n, void* stream);
int tgan_sizeof_wn_desc(void);
int tgan_sizeof_pack_desc(void);

/* ---------------------------------------------------------------- weight normalisation ------------
 * V viewed as [A, Co, B]; per output channel co: nrm = sqrt(sum_{a,b} V^2);
 *   W = V * g[co] / nrm          (eps_mode 0: no epsilon, nn.py:554)
 *   W = V * g[co] * rsqrt(max(nrm^2, 1e-12))  (eps_mode 1: tf.nn.l2_normalize, nn.py:502; modle_base.py:66,101,148)
 * HWIO / [in,out]: A = kh*kw*Cin, B = 1.  Transposed-conv [kh,kw,Cout,Cin] (axes [0,1,3]): A = kh*kw, B = Cin. */
int tgan_weightnorm_fwd(const float* V, const float* g, float* W, float* inv_norm, float* scale, int A, int Co,
                        int B, int eps_mode, float* ws, void* stream);
/* also writes scale[co] = g[co]*inv_norm[co]; W may be NULL (the tcgen05 path folds `scale` into
 * tgan_pack_weight_bf16 instead).  ws: (4*TGAN_STATS_MAX_PARTS + 1)*Co floats.
 * dg[co] (+)= <dW, V>*inv_norm ; dV (+)= (g*inv_norm) * (dW - V*inv_norm * <dW,V>*inv_norm)    */
int tgan_weightnorm_bwd(const float* V, const float* g, const float* inv_norm, const float* dW, float* dV,
                        float* dg, int A, int Co, int B, float beta, float* ws, void* stream);

/* ---------------------------------------------------------------- per-channel statistics ----------
 * x [rows, C] (dtype xdt, contiguous): sum[c] = beta*sum[c] + sum_r x, sumsq[c] likewise (sumsq may be NULL).
 * One launch, deterministic (the last CTA folds the partials in a fixed order); ws: 4*TGAN_STATS_MAX_PARTS*C floats (fp64 partials). */
#define TGAN_STATS_MAX_PARTS 256
#define TGAN_ACT_BWD_SEG_PARTS 444   /* tgan_act_bwd_seg: ws must hold 4*TGAN_ACT_BWD_SEG_PARTS*C floats */
int tgan_channel_stats(const void* x, int xdt, int64_t rows, int C, float* sum, float* sumsq, float beta, float* ws,
                       void* stream);

/* mean-only batch norm (nn.py:147-187) fused with the nonlinearity (nn.py:517), one pass:
 *   training: mean = sum[c]/rows (sum from tgan_channel_stats or from the tgan_igemm_bf16 epilogue);
 *             y = act(x - mean + b);  pop_mean = pop_mean*decay + mean*(1-decay)        (nn.py:172-183)
 *   testing : y = act(x - pop_mean + b)                                                 (nn.py:170-171) */
int tgan_mobn_apply(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, const float* sum, const float* b,
                    float* pop_mean, float decay, int train, int act, float alpha, void* stream);

/* Segment-aware variants for GROUPED batches: several calls of one network (e.g. C(x_l), C(x_u), C(G(z)) of
 * Good_GAN_cifar10.py:220-254) are concatenated along the batch axis and run as one pass; every call keeps its own
 * batch mean.  nseg <= 4 segments of consecutive rows; r0,r1,r2 = exclusive end rows of segments 0..2 (the last
 * segment ends at `rows`; unused values ignored).  bf16 tensors, C % 8 == 0 (16-byte vector accesses).
 *   mobn_apply_seg: y = act(x - sums[seg]/rows_seg + b) (train) | act(x - pop_mean + b) (test); pop_mean is updated once
 *                   per segment in call order (nn.py:176-183).  sums: [nseg][C], fp32 (sums_q24 = 0) or the int64 Q24
 *                   fixed-point sums a tgan_igemm_bf16 epilogue accumulated (sums_q24 = 1).
 *   act_bwd_seg:    du = dy * act'(y); colsums[seg][c] = sum_{rows in seg} du; grad_acc[c] += sum over all rows (db).
 *   sub_channel_mean_seg: dz = du - colsums[seg]/rows_seg (the mean-subtraction's gradient, nn.py:179). */
int tgan_mobn_apply_seg(const void* x, void* y, int64_t rows, int C, int nseg, int64_t r0, int64_t r1, int64_t r2,
                        const void* sums, int sums_q24, const float* b, float* pop_mean, float decay, int train, int act,
                        float alpha, void* stream);
int tgan_act_bwd_seg(const void* dy, int dydt, const void* y, int ydt, void* du, int dudt, int64_t rows, int C,
                     int nseg, int64_t r0, int64_t r1, int64_t r2, int act, float alpha, float* colsums,
                     float* grad_acc, float* ws, void* stream);
int tgan_sub_channel_mean_seg(const void* du, void* dz, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                              int64_t r2, const float* colsums, void* stream);

/* The same layer on a SMALL fp32 tensor (the classifier's logits [rows, 10], Good_GAN_cifar10.py:166 through
 * nn.py:147-187 / :517): every segment of the grouped batch in ONE single-CTA launch instead of two launches per segment.
 * C <= 32, rows * C <= 2^20, nseg <= 4, r0..r2 as above.
 *   mobn_small_fwd: per-segment column means in a fixed order; y = act(z - mean_seg + b) (train) | act(z - pop_mean + b);
 *                   pop_mean <- decay*pop_mean + (1-decay)*mean_seg once per segment, in segment order.
 *   mobn_small_bwd: du = dy * act'(y); grad_acc[c] += sum_rows du (may be NULL);
 *                   dz = du - mean_seg(du) if subtract_mean else du (dz may be NULL: bias gradient only). */
int tgan_mobn_small_fwd(const float* z, float* y, int64_t rows, int C, int nseg, int64_t r0, int64_t r1, int64_t r2,
                        const float* b, float* pop_mean, float decay, int train, int act, float alpha, void* stream);
int tgan_mobn_small_bwd(const float* dy, const float* y, float* dz, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                        int64_t r2, int act, float alpha, int subtract_mean, float* grad_acc, void* stream);

/* Training-mode tf.contrib.layers.batch_norm (modle_base.py:229-237) of a GROUPED batch: rows [0, r0), [r0, r1), ... are
 * the calls of one network (segments as in tgan_mobn_apply_seg, nseg <= 4); every call keeps its own batch statistics and
 * updates the moving statistics once, in call order -- what the separate calls of the reference graph compute
 * (Good_GAN.py:220-299 normalises after every convolution of the classifier).  Three launches per direction for all
 * segments (partial sums -> fold / finalize -> apply), fp64 partials folded in a fixed order.
 *   bn_fwd_seg: y = (x - mean_seg) * rstd_seg * gamma + beta; mean / rstd: [nseg][C] outputs kept for the backward.
 *   bn_bwd_seg: dx = gamma * rstd_seg * (dy - s1_seg/n_seg - xhat * s2_seg/n_seg), s1 = sum dy, s2 = sum dy * xhat;
 *               dbeta = beta_acc * dbeta + sum_seg s1, dgamma likewise with s2 (either may be NULL).
 * ws: 1036 * C floats, 16-byte aligned; rows * C < 2^31. */
int tgan_bn_fwd_seg(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, int nseg, int64_t r0, int64_t r1,
                    int64_t r2, const float* gamma, const float* beta, float eps, float decay, int unbiased_moving_var,
                    float* moving_mean, float* moving_var, float* mean, float* rstd, float* ws, void* stream);
int tgan_bn_bwd_seg(const void* dy, int dydt, const void* x, int xdt, void* dx, int dxdt, int64_t rows, int C, int nseg,
                    int64_t r0, int64_t r1, int64_t r2, const float* mean, const float* rstd, const float* gamma,
                    float* dgamma, float* dbeta, float beta_acc, float* ws, void* stream);

/* tf.contrib.layers.batch_norm training statistics (modle_base.py:229-237):
 *   mean, rstd = rsqrt(var_biased + eps); scale = gamma*rstd; shift = beta - mean*scale;
 *   moving_mean/variance <- decay-EMA; unbiased_moving_var = 1: the variance fed to the EMA is the unbiased estimate
 *   (contrib's fused path), 0: the biased tf.nn.moments variance (nn.batch_norm_impl, nn.py:207-214). */
int tgan_bn_finalize(const float* sum, const float* sumsq, int64_t rows, int C, const float* gamma,
                     const float* beta, float eps, float decay, int unbiased_moving_var, float* moving_mean,
                     float* moving_var, float* mean, float* rstd, float* scale, float* shift, void* stream);
int tgan_bn_eval_affine(const float* gamma, const float* beta, const float* moving_mean,
                        const float* moving_var, float eps, int C, float* scale, float* shift, void* stream);

/* y = act(x*scale[c] + shift[c]) ; scale/shift may be NULL (1 / 0).  x [rows,C] dtype xdt -> y dtype ydt.
 * The one fused epilogue: bias + nonlinearity (nn.py:508,517), mean-only-BN apply + lrelu (nn.py:183,517),
 * BN apply (modle_base.py:230), tanh/sigmoid heads. */
int tgan_affine_act(const void* x, int xdt, void* y, int ydt, int64_t rows, int C, const float* scale,
                    const float* shift, int act, float alpha, void* stream);
/* du = dy * act'(.) computed from the activation OUTPUT y; also per-channel partial sums of du
 * (-> bias gradient / mean-only-BN backward).  colsum[c] = sum_r du and grad_acc[c] += sum_r du (either may be
 * NULL; deterministic, ws as channel_stats). */
int tgan_act_bwd(const void* dy, int dydt, const void* y, int ydt, void* du, int dudt, int64_t rows, int C,
                 int act, float alpha, float* colsum, float* grad_acc, float* ws, void* stream);
/* same with y in a channel-padded buffer (row stride ldy >= C), e.g. a conv output written straight into its
 * label-concatenated successor (modle_base.py:239-244) */
int tgan_act_bwd_ld(const void* dy, int dydt, const void* y, int ydt, int ldy, void* du, int dudt, int64_t rows, int C,
                    int act, float alpha, float* colsum, float* grad_acc, float* ws, void* stream);
/* mean-only BN backward: dz = du - colsum[c]/rows */
int tgan_sub_channel_mean(const void* du, int dudt, void* dz, int dzdt, int64_t rows, int C,
                          const float* colsum, void* stream);
/* BN backward: s1[c] = sum dy, s2[c] = sum dy*xhat, then dx = scale*(dy - s1/M - xhat*s2/M);
 * dgamma (+)= s2, dbeta (+)= s1. */
int tgan_bn_bwd(const void* dy, int dydt, const void* x, int xdt, void* dx, int dxdt, int64_t rows, int C,
                const float* mean, const float* rstd, const float* gamma, float* dgamma, float* dbeta,
                float beta_acc, float* ws, void* stream);

/* ---------------------------------------------------------------- pointwise layers ----------------
 * Gaussian noise (modle_base.py:193-202; Good_GAN_cifar10.py:29-31): y = x + std * n,
 * n = noise[i] if noise != NULL (parity mode, fp32 standard normal) else Philox4x32-10(seed,stream_id,*counter). */
int tgan_add_noise(const void* x, int xdt, void* y, int ydt, int64_t n, float std, const float* noise,
                   uint64_t seed, uint64_t stream_id, const uint64_t* counter, void* stream);
/* inverted dropout (tf.layers.dropout, modle_base.py:190; Good_GAN_cifar10.py:124,143):
 * y = x * keep / (1-rate).  gen != 0: draw keep from Philox and WRITE mask; gen == 0: READ mask (uint8 0/1). */
int tgan_dropout(const void* x, int xdt, void* y, int ydt, uint8_t* mask, int64_t n, float rate, int gen,
                 uint64_t seed, uint64_t stream_id, const uint64_t* counter, void* stream);
/* 2x2/s2 max pool on even extents (tf.nn.max_pool SAME, Good_GAN_cifar10.py:123,142; max_pooling2d(2,2)
 * Good_GAN.py:223): writes the winner index (first max, row-major window order) for the backward. */
int tgan_maxpool2_fwd(const void* x, int dt, void* y, uint8_t* idx, int N, int H, int W, int C, void* stream);
int tgan_maxpool2_bwd(const void* dy, int dt, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                      void* stream);
/* 2x2/s2 max pool fused with the inverted dropout that follows it (Good_GAN_cifar10.py:123-124, 142-143): bf16 NHWC,
 * C % 8 == 0.  y = keep ? max * 1/(1-rate) : 0;  code[N,H/2,W/2,C] = winner (bits 0-1, first maximum) | keep << 2.
 * mask != NULL: keep flags supplied (parity mode); else Philox(seed, stream_id, *counter), the same draws as tgan_dropout
 * for that stream; rate == 0: pool only.  bwd: dx[winner] = keep ? dy/(1-rate) : 0, zeros elsewhere. */
int tgan_maxpool2_dropout_fwd(const void* x, void* y, uint8_t* code, int N, int H, int W, int C, float rate,
                              const uint8_t* mask, uint64_t seed, uint64_t stream_id, const uint64_t* counter,
                              void* stream);
int tgan_maxpool2_dropout_bwd(const void* dy, const uint8_t* code, void* dx, int N, int H, int W, int C, float rate,
                              void* stream);
/* mean-only BN apply + nonlinearity + 2x2 max pool + dropout in one pass (conv1_3 / conv2_3 of the classifier,
 * Good_GAN_cifar10.py:118-124, 137-143); the full-resolution activation is never materialised.  z: bf16 [N,H,W,C],
 * segments in IMAGES (n0,n1,n2 = exclusive end images of segments 0..2), sums fp32 [nseg][C] as tgan_mobn_apply_seg.
 * y / code as tgan_maxpool2_dropout_fwd.  bwd: du[N,H,W,C] = at the window winner keep * dy/(1-rate) * act'(y_winner),
 * zero elsewhere; colsums[seg][c] = per-segment sums of du, grad_acc[c] += total (the bias gradient).
 * ws: 4*TGAN_ACT_BWD_SEG_PARTS*C floats. */
int tgan_mobn_pool_dropout_fwd(const void* z, void* y, uint8_t* code, int N, int H, int W, int C, int nseg, int64_t n0,
                               int64_t n1, int64_t n2, const void* sums, int sums_q24, const float* b, float* pop_mean,
                               float decay, int train, int act, float alpha, float rate, const uint8_t* mask, uint64_t seed,
                               uint64_t stream_id, const uint64_t* counter, void* stream);
int tgan_mobn_pool_dropout_bwd(const void* dy, const void* y, const uint8_t* code, void* du, int N, int H, int W, int C,
                               int nseg, int64_t n0, int64_t n1, int64_t n2, int act, float alpha, float rate,
                               float* colsums, float* grad_acc, float* ws, void* stream);
/* global pooling over H*W: mode 0 = max (max_pooling2d(6,1) 'avg_pool_0', Good_GAN_cifar10.py:163),
 * mode 1 = mean (average_pooling2d(8,1) :94; reduce_mean([1,2]) Good_GAN.py:157,243,295). */
int tgan_global_pool_fwd(const void* x, int xdt, void* y, int ydt, uint8_t* idx, int N, int HW, int C,
                         int mode, void* stream);
int tgan_global_pool_bwd(const void* dy, int dydt, const uint8_t* idx, void* dx, int dxdt, int N, int HW,
                         int C, int mode, void* stream);
/* _conv_cond_concat / tf.concat([h, y], 1) (modle_base.py:239-244; Good_GAN_cifar10.py:38,97):
 * out[r, 0:C] = x[r, 0:C]; out[r, C:C+K] = lab[r / rows_per_sample, 0:K]; out[r, C+K:ldo] = 0. */
int tgan_concat_label(const void* x, int xdt, int64_t rows, int C, int ldx, const float* lab, int K,
                      int rows_per_sample, void* out, int odt, int ldo, void* stream);
/* dropout / per-channel affine (batch-norm apply) written straight into a label-concatenated tensor: y[r, 0:C] =
 * dropout(x)[r] resp. x*scale + shift, y[r, C:C+K] = lab[r / rows_per_sample], y[r, C+K:ldy] = 0 -- the
 * `dropout -> _conv_cond_concat` (Good_GAN_cifar10.py:73-75, 83-85) and `batch_norm -> _conv_cond_concat` (:43-51) pairs
 * as one pass.  Dropout arguments and Philox stream as tgan_dropout (mask is [rows, C], dense). */
int tgan_dropout_concat(const void* x, int xdt, void* y, int ydt, int ldy, uint8_t* mask, int64_t rows, int C, float rate,
                        int gen, uint64_t seed, uint64_t stream_id, const uint64_t* counter, const float* lab, int K,
                        int rows_per_sample, void* stream);
int tgan_affine_concat(const void* x, int xdt, void* y, int ydt, int ldy, int64_t rows, int C, const float* scale,
                       const float* shift, const float* lab, int K, int rows_per_sample, void* stream);
/* label planes only: out[r, C + j] = j < K ? lab[r / rows_per_sample, j] : 0 for j in [0, ldo - C) -- used when the
 * producing GEMM epilogue already wrote channels [0, C) of the concatenated tensor in place */
int tgan_fill_label(const float* lab, int K, int rows_per_sample, void* out, int odt, int64_t rows, int C, int ldo,
                    void* stream);
/* strided channel slice / cast: dst[r, 0:C] = src[r, 0:C] */
int tgan_copy_channels(const void* src, int sdt, int lds, void* dst, int ddt, int ldd, int64_t rows, int C,
                       void* stream);
/* dst = concat(src[0], ..., src[n-1]) (n <= 8), each source contiguous with counts[i] elements of dtype dts[i]
 * (fp32 / bf16), converted to ddt: the grouped batch of several calls of one network (Good_GAN_cifar10.py:228-240 /
 * :264-270 run as one pass) formed by ONE launch */
int tgan_gather_rows(const void* const* srcs, const int* dts, const int64_t* counts, int n, void* dst, int ddt,
                     void* stream);
/* y[i] = sum_p x[p][i] over `parts` contiguous bf16 slices of n elements (n % 8 == 0): folds the split-K partial outputs
 * of a contraction whose tap groups ran as output classes of one tgan_igemm_bf16 launch */
int tgan_sum_slices_bf16(const void* x, void* y, int64_t n, int parts, void* stream);
/* y (+)= x elementwise (gradient accumulation), fp32 or bf16 */
int tgan_accumulate(void* y, const void* x, int dt, int64_t n, void* stream);
int tgan_fill_f32(float* p, float v, int64_t n, void* stream);

/* ---------------------------------------------------------------- pseudo-labels -------------------
 * tf.argmax(axis=1) -> int64 (lowest index wins ties; NaN never wins over a number) and
 * tf.one_hot(depth=K) fp32 (Good_GAN_cifar10.py:232,237,259,270; Good_GAN.py:447,451,458,468). */
int tgan_argmax_onehot(const float* logits, int N, int K, int64_t* idx, float* onehot, void* stream);

/* ---------------------------------------------------------------- losses (train_base.py:113-154) --
 * d_loss = CE(1,dr) + .5 CE(0,df) + .5 CE(0,du)  and its dlogits (train_base.py:123-126). */
int tgan_loss_d(const float* dr, int nr, const float* df, int nf, const float* du, int nu, float* loss,
                float* g_dr, float* g_df, float* g_du, void* stream);
/* g_loss = .5 CE(1,df) (train_base.py:128) */
int tgan_loss_g(const float* df, int nf, float* loss, float* g_df, void* stream);
/* c_loss (train_base.py:130-152) with its four dlogits; c_rep / g_rep may be NULL (non-cifar10).
 * lambdas = device float[2] {lambda_1, lambda_2}. */
int tgan_loss_c(const float* c_real, const float* y_l_c, int n_real, const float* c_unl, const float* c_rep,
                const float* d_unl_logits, int n_unl, const float* c_fake, const float* y_g, int n_fake,
                int K, const float* lambdas, float* loss, float* g_real, float* g_unl, float* g_rep,
                float* g_fake, void* stream);

/* ---------------------------------------------------------------- optimiser -----------------------
 * tf.train.AdamOptimizer kernel form (train_base.py:91-97; Train_goodGAN.py:85-94) over one flat buffer:
 *   a = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= a*m/(sqrt(v)+eps)
 * state = device float[3] {lr, beta1_power, beta2_power}; grad is multiplied by grad_scale first
 * (1/world_size after the NCCL sum).  If ema != NULL also applies tf.train.ExponentialMovingAverage
 * (Train_goodGAN.py:101-103): ema -= (ema - theta)*(1-ema_decay). */
int tgan_adam(float* theta, float* m, float* v, const float* grad, int64_t n, const float* state,
              float beta1, float beta2, float eps, float grad_scale, float* ema, float ema_decay,
              void* stream);
/* ---------------------------------------------------------------- data-parallel update over peer memory ----
 * NEW (the reference is single-GPU): reduce-scatter -> Adam on the owned shard -> all-gather of the parameters as one
 * kernel over NVLink peer memory (csrc/dp_fused.cu).  grad_ptrs / theta_ptrs / flag_ptrs: HOST arrays of `world` device
 * addresses (uint64), entry p = rank p's buffer (symmetric allocations, identical layout on every rank).
 *   tgan_dp_barrier : every rank arrives, nobody leaves before all did; slot 0 / 1 = two independent barriers; flags:
 *                     int [2][8] per rank, zero-initialised; epoch: LOCAL int [2] zero-initialised (device counters)
 *   tgan_dp_adam    : rank r sums elements [r*per, (r+1)*per) of all gradient buffers in rank order, divides by world,
 *                     applies tgan_adam's update with its local m / v slice and writes the new parameters into EVERY
 *                     rank's theta buffer.  Call between tgan_dp_barrier(slot 0) and tgan_dp_barrier(slot 1).
 *                     mc_grad / mc_theta (optional, both or neither): NVSwitch MULTICAST addresses of the same buffers;
 *                     then the sum is one multimem.ld_reduce (added inside the switch) and the broadcast one multimem.st.
 *   tgan_ema        : shadow -= (shadow - theta) * (1 - decay) over the gathered parameters (Train_goodGAN.py:101-103) */
int tgan_dp_barrier(const uint64_t* flag_ptrs, int rank, int world, int slot, int* epoch, void* stream);
int tgan_dp_adam(const uint64_t* grad_ptrs, const uint64_t* theta_ptrs, const float* mc_grad, float* mc_theta, float* m,
                 float* v, int64_t n, int rank, int world, const float* state, float beta1, float beta2, float eps,
                 void* stream);
int tgan_ema(float* ema, const float* theta, int64_t n, float decay, void* stream);
/* beta1_power *= beta1, beta2_power *= beta2 (TF's _finish) */
int tgan_adam_advance(float* state, float beta1, float beta2, void* stream);
/* Philox counter bump once per step: *counter += inc */
int tgan_counter_advance(uint64_t* counter, uint64_t inc, void* stream);

/* ---------------------------------------------------------------- data formats either side of the step --
 * Device-side batch formation from a uint8 dataset resident in HBM (replaces the tf.data / TFRecord host path of
 * Input_Pipeline/cifar10Dataset.py:42-85, svhnDataset.py:41-88, mnistDataset.py:42-90 for the step's inputs):
 * out[i] = map(data[idx[i]]) with map = x/255*2-1 (mode 0; cifar10Dataset.py:60, svhnDataset.py:63) or x/255 (mode 1;
 * mnistDataset.py:65), bit-identical to the float32 TF expression.  data: [n_total, elems] uint8, elems % 16 == 0;
 * idx: device int64[n] (NULL = the first n images); indices are clamped to [0, n_total). */
int tgan_gather_images_u8(const void* data, int64_t n_total, int64_t elems, const int64_t* idx, int n,
                          float* out, int mode, void* stream);
/* tf.one_hot(label[idx[i]], depth=K) (cifar10Dataset.py:62); labels: device int32[n_total] */
int tgan_gather_onehot(const int32_t* labels, int64_t n_total, const int64_t* idx, int n, int K, float* out,
                       void* stream);
/* z ~ U(-1,1) [n, zdim] and y = one_hot(randint(0,K)) [n, K] (Train_goodGAN.py:232-237) from Philox(seed, stream_id,
 * *counter); counter may be NULL (= 0). */
int tgan_draw_latent(float* z, int n, int zdim, float* y, int K, uint64_t seed, const uint64_t* counter,
                     uint64_t stream_id, void* stream);
/* utils.py:199-231: merge(inverse_transform(images), [gh, gw]) -> grid [gh*H, gw*W, C]; inverse = 1 applies (x+1)/2 */
int tgan_image_grid(const float* x, int n, int H, int W, int C, int gh, int gw, float* grid, int inverse,
                    void* stream);
/* CRC-32C (Castagnoli, reflected 0x82f63b78) of a HOST buffer, continuing from `crc` (0 to start): the checksum of the
 * TF tensor-bundle files written by tf.train.Saver (Training/Saver.py:29-36). */
uint64_t tgan_crc32c(uint64_t crc, const void* data, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* TGAN_H_ */
