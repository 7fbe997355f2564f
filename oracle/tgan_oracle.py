"""Torch-CPU restatement of the reference Triple-GAN training step (TEST INFRASTRUCTURE).

PARITY UNPINNED (see oracle/__init__.py): TensorFlow cannot run here and the
reference holds no golden vectors; this file restates the reference graph code
with TensorFlow's published op semantics and is pinned against
`oracle/tf_semantics_np.py` + finite differences in tests/test_oracle.py.

Engine: torch CPU autograd, float64 for golden values, float32 for the
fp32-vs-fp64 noise floor and for the timed CPU baseline (bench.py cpu_baseline).

All `file:line` citations are relative to /root/reference.
Layouts: activations NHWC, conv filters HWIO [kh,kw,Cin,Cout], transposed-conv
filters [kh,kw,Cout,Cin], dense [in,out] -- exactly as the TF variables.
"""
import math
import zlib

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# RNG: every stochastic op draws by *tag*, so the oracle and the CUDA path can be fed the
# identical noise / dropout tensors independent of call order (TF's Philox streams cannot
# be reproduced without TF -- SURVEY.md §7 "Non-reproducible randomness").
# --------------------------------------------------------------------------------------


class TagRNG:
    """Deterministic per-tag standard-normal / keep-mask source (float32 draws)."""

    def __init__(self, seed=0):
        self.seed = int(seed)

    def _gen(self, tag):
        g = torch.Generator(device='cpu')
        g.manual_seed((zlib.crc32(tag.encode()) * 2654435761 + self.seed) % (2 ** 63))
        return g

    def normal(self, tag, shape):
        return torch.randn(tuple(shape), generator=self._gen('n:' + tag), dtype=torch.float32)

    def keep_mask(self, tag, shape, rate):
        u = torch.rand(tuple(shape), generator=self._gen('d:' + tag), dtype=torch.float32)
        return (u >= rate).to(torch.uint8)


# --------------------------------------------------------------------------------------
# Optional bf16 rounding points.  The reference computes in fp32; the tensor-core mode of the CUDA path
# stores activations and feeds the MMAs in bf16 (fp32 accumulation).  Rounding flips discrete routing
# decisions (max-pool winners, lrelu sides), so its gradients cannot be compared point-wise with an
# un-rounded run.  `with quantized():` restates the SAME rounding points in the oracle (straight-through
# for autograd), which isolates the kernels' arithmetic from the precision choice:
#   * operands of every tensor-core contraction (Cout >= 16 and Cout % 8 == 0) and its result,
#   * the output of every epilogue / normalisation with >= 16 channels, of noise, dropout and concat.
# --------------------------------------------------------------------------------------
_QUANT = False


class quantized:
    def __enter__(self):
        global _QUANT
        self._old, _QUANT = _QUANT, True

    def __exit__(self, *a):
        global _QUANT
        _QUANT = self._old


class _RoundBoth(torch.autograd.Function):
    """bf16 rounding of an ACTIVATION: the value in forward and, in backward, the gradient that flows back through the
    same tensor -- the tensor-core path stores activation gradients (dy, du = dy*act', dz = du - mean) in bf16 exactly
    where it stores the activations."""

    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def q(t):
    """round an activation to bf16 (and its gradient on the way back) when the bf16 restatement is active"""
    if not _QUANT:
        return t
    return _RoundBoth.apply(t)


def qw(t):
    """round a weight operand to bf16; its gradient stays fp32 (filter gradients accumulate in fp32 / TMEM)"""
    if not _QUANT:
        return t
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _tc(cout):
    return _QUANT and cout >= 16 and cout % 8 == 0


def qc(t):
    """epilogue outputs keep fp32 when they have < 16 channels (logits, RGB)"""
    return q(t) if t.shape[-1] >= 16 else t


# --------------------------------------------------------------------------------------
# TF op semantics
# --------------------------------------------------------------------------------------


def same_pad(n, k, s):
    """TF SAME geometry: (out, pad_before, pad_after); extra pad goes after."""
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return out, tot // 2, tot - tot // 2


def conv2d_tf(x, w, stride=1, padding='SAME', round_out=True):
    """tf.nn.conv2d on NHWC x, HWIO w (nn.py:504, modle_base.py:102,161).  round_out=False (bf16 restatement only): the
    result stays in the fp32 accumulator because the consuming epilogue is fused into the contraction."""
    kh, kw = w.shape[0], w.shape[1]
    tc = _tc(w.shape[3])
    if tc:
        x, w = q(x), qw(w)
    xc = x.permute(0, 3, 1, 2)
    if padding.upper() == 'SAME':
        _, pt, pb = same_pad(x.shape[1], kh, stride)
        _, pl, pr = same_pad(x.shape[2], kw, stride)
        xc = F.pad(xc, (pl, pr, pt, pb))
    y = F.conv2d(xc, w.permute(3, 2, 0, 1), stride=stride)
    return q(y.permute(0, 2, 3, 1)) if (tc and round_out) else y.permute(0, 2, 3, 1)


def conv2d_transpose_tf(x, w, stride=2, padding='SAME'):
    """tf.nn.conv2d_transpose / tf.layers.conv2d_transpose (modle_base.py:149,250).
    w is [kh,kw,Cout,Cin].  == input-gradient of the SAME conv: full scatter of size
    (n-1)s+k then crop [before : before + s*n] with before = (k-s)//2."""
    kh, kw = w.shape[0], w.shape[1]
    tc = _tc(w.shape[2])
    if tc:
        x, w = q(x), qw(w)
    y = F.conv_transpose2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=stride)
    if padding.upper() == 'SAME':
        Ho, Wo = x.shape[1] * stride, x.shape[2] * stride
        _, pt, _ = same_pad(Ho, kh, stride)
        _, pl, _ = same_pad(Wo, kw, stride)
        # the full scatter (n-1)s+k is shorter than pt+s*n when k < s: the missing tail is zeros
        y = F.pad(y, (0, max(0, pl + Wo - y.shape[3]), 0, max(0, pt + Ho - y.shape[2])))
        y = y[:, :, pt:pt + Ho, pl:pl + Wo]
    return q(y.permute(0, 2, 3, 1)) if tc else y.permute(0, 2, 3, 1)


def matmul_tf(x, w):
    """tf.matmul / tf.layers.dense contraction (nn.py:553; modle_base.py:40,67)"""
    if _tc(w.shape[1]):
        return q(q(x) @ qw(w))
    return x @ w


def l2_normalize(v, axes, eps=1e-12):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps))."""
    ss = (v * v).sum(dim=axes, keepdim=True)
    return v * torch.rsqrt(torch.clamp(ss, min=eps))


def lrelu_cifar(x, alpha=0.2):
    """Good_GAN_cifar10.py:26-27: relu(x) - alpha*relu(-x)."""
    return F.relu(x) - alpha * F.relu(-x)


def leaky_relu_tf(x, alpha=0.2):
    """tf.nn.leaky_relu default alpha=0.2 (Good_GAN.py:99...)."""
    return torch.maximum(x, alpha * x)


def dropout_tf(x, keep_mask, rate):
    """tf.layers.dropout(training=True): x * mask / (1-rate)  (modle_base.py:190)."""
    return q(x * keep_mask.to(x.dtype) * (1.0 / (1.0 - rate)))


def max_pool_tf(x, k, s):
    """tf.nn.max_pool / tf.layers.max_pooling2d on even extents (no padding needed)."""
    return F.max_pool2d(x.permute(0, 3, 1, 2), k, s).permute(0, 2, 3, 1)


def cond_concat(x, yb):
    """modle_base.py:239-244."""
    return q(torch.cat([x, yb.expand(x.shape[0], x.shape[1], x.shape[2], yb.shape[3])], dim=3))


def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def softmax_ce(logits, labels):
    """tf.nn.softmax_cross_entropy_with_logits_v2 (per row)."""
    return -(labels * F.log_softmax(logits, dim=1)).sum(dim=1)


def argmax_onehot(logits, depth=10):
    """tf.argmax(axis=1) + tf.one_hot (Good_GAN_cifar10.py:232,237,259,270).

    tf.argmax reduces with Eigen's ArgMaxTupleReducer (TF 1.12-1.15): the accumulator starts at
    (index 0, NumTraits<float>::lowest() = -FLT_MAX) and an element replaces it only if it is strictly GREATER.
    Consequences restated here: the lowest index wins ties; a NaN never wins (NaN > x is false) and never blocks
    a later number; a row of only NaN / -inf / -FLT_MAX yields index 0.  torch.argmax would return the NaN's
    index, so NaN and -inf are mapped to -FLT_MAX first (torch.argmax on CPU returns the first maximal index;
    checked against the loop restatement oracle/tf_semantics_np.py in tests/test_oracle.py)."""
    lo = float(torch.finfo(torch.float32).min)
    x = torch.nan_to_num(logits.detach().double(), nan=lo, neginf=lo).clamp_(min=lo)
    idx = torch.argmax(x, dim=1)
    return idx, F.one_hot(idx, depth).to(logits.dtype)


# --------------------------------------------------------------------------------------
# Layers (each mutates the non-trainable running statistics in `S` like the TF assigns)
# --------------------------------------------------------------------------------------


def mean_only_bn(x, pop_mean_key, b, S, train, conv, decay=0.9):
    """nn.py:147-187."""
    if train:
        axes = (0, 1, 2) if conv else (0,)
        m = x.mean(dim=axes)
        S[pop_mean_key] = (S[pop_mean_key] * decay + m.detach() * (1 - decay))
        return x - m + b
    return x - S[pop_mean_key] + b


def conv2d_WN(P, S, scope, x, pad, train, nonlin=lrelu_cifar, stride=1, fused=False):
    """nn.conv2d_WN with use_weight_normalization + use_mean_only_batch_normalization,
    init=False branch (nn.py:501-518).  fused (bf16 restatement): `- mean + b` and the nonlinearity run in the
    contraction's epilogue on the fp32 accumulator, so the convolution output itself is never rounded."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    W = g.view(1, 1, 1, -1) * l2_normalize(V, (0, 1, 2))
    x = conv2d_tf(x, W, stride, pad, round_out=not fused)
    x = mean_only_bn(x, scope + '/meanOnlyBatchNormalization/pop_mean', b, S, train, True)
    return qc(nonlin(x) if nonlin is not None else x)


def dense_WN(P, S, scope, x, train, nonlin=None):
    """nn.dense_WN, WN + mean-only BN (nn.py:552-570): matmul first, then g/sqrt(sum V^2), no eps."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if _QUANT:      # the CUDA path folds g/||V|| into the weight before the (bf16) contraction
        x = matmul_tf(x, V * (g / torch.sqrt((V * V).sum(dim=0))).view(1, -1))
    else:
        x = x @ V
        x = (g / torch.sqrt((V * V).sum(dim=0))).view(1, -1) * x
    x = mean_only_bn(x, scope + '/meanOnlyBatchNormalization/pop_mean', b.view(1, -1), S, train, False)
    return qc(nonlin(x) if nonlin is not None else x)


def NiN_WN(P, S, scope, x, train, nonlin):
    """nn.NiN_WN (nn.py:577-589); variable scope is doubled name/name."""
    s = x.shape
    y = dense_WN(P, S, scope + '/' + scope.split('/')[-1], x.reshape(-1, s[-1]), train, nonlin)
    return y.reshape(s[0], s[1], s[2], -1)


def bn_contrib(P, S, scope, x, train, eps=1e-5, decay=0.9):
    """tf.contrib.layers.batch_norm(decay .9, eps 1e-5, scale=True, updates_collections=None)
    (modle_base.py:229-237)."""
    gamma, beta = P[scope + '/gamma'], P[scope + '/beta']
    axes = tuple(range(x.dim() - 1))
    if train:
        mu = x.mean(dim=axes)
        var = ((x - mu) ** 2).mean(dim=axes)
        n = x.numel() // x.shape[-1]
        S[scope + '/moving_mean'] = S[scope + '/moving_mean'] * decay + mu.detach() * (1 - decay)
        S[scope + '/moving_variance'] = (S[scope + '/moving_variance'] * decay
                                         + var.detach() * (n / max(n - 1, 1)) * (1 - decay))
    else:
        mu, var = S[scope + '/moving_mean'], S[scope + '/moving_variance']
    return qc(gamma * (x - mu) * torch.rsqrt(var + eps) + beta)


def linear_fc(P, scope, x):
    """NN_Base._linear_fc -> tf.layers.dense, scope doubled (modle_base.py:27-48)."""
    n = scope + '/' + scope.split('/')[-1]
    return matmul_tf(x, P[n + '/kernel']) + P[n + '/bias']


def conv2d_layer(P, scope, x, stride):
    """NN_Base._conv2d -> tf.layers.conv2d padding='same' + bias (modle_base.py:157-168)."""
    n = scope + '/' + scope.split('/')[-1]
    return conv2d_tf(x, P[n + '/kernel'], stride, 'SAME') + P[n + '/bias']


def deconv2d_layer(P, scope, x, stride=2):
    """NN_Base._deconv2d -> tf.layers.conv2d_transpose padding='same' + bias (modle_base.py:246-259)."""
    n = scope + '/' + scope.split('/')[-1]
    return conv2d_transpose_tf(x, P[n + '/kernel'], stride, 'SAME') + P[n + '/bias']


def WN_dense(P, scope, x):
    """NN_Base._WN_dense init=False (modle_base.py:50-73)."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if _QUANT:      # g folded into the weight before the contraction (same function of (V, g))
        return matmul_tf(x, g.view(1, -1) * l2_normalize(V, (0,))) + b.view(1, -1)
    return g.view(1, -1) * (x @ l2_normalize(V, (0,))) + b.view(1, -1)


def WN_conv2d(P, scope, x, stride):
    """NN_Base._WN_conv2d init=False (modle_base.py:75-108): conv(x, l2norm(V))*g + b."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if _QUANT:
        return conv2d_tf(x, g.view(1, 1, 1, -1) * l2_normalize(V, (0, 1, 2)), stride, 'SAME') + b.view(1, 1, 1, -1)
    return g.view(1, 1, 1, -1) * conv2d_tf(x, l2_normalize(V, (0, 1, 2)), stride, 'SAME') + b.view(1, 1, 1, -1)


def WN_deconv2d(P, scope, x, stride=2):
    """NN_Base._WN_deconv2d init=False (modle_base.py:130-155): normalise over axes [0,1,3]."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if _QUANT:
        return conv2d_transpose_tf(x, g.view(1, 1, -1, 1) * l2_normalize(V, (0, 1, 3)), stride, 'SAME') \
            + b.view(1, 1, 1, -1)
    y = conv2d_transpose_tf(x, l2_normalize(V, (0, 1, 3)), stride, 'SAME')
    return g.view(1, 1, 1, -1) * y + b.view(1, 1, 1, -1)


def batch_norm_impl(P, S, scope, x, train, conv=True, decay=0.9):
    """nn.batch_norm_impl (nn.py:192-217): tf.nn.batch_normalization with eps 1e-3, own scale / beta; the
    population statistics follow the BIASED batch variance of tf.nn.moments (:207-214)."""
    n = scope + '/BatchNormalization'
    scale, beta = P[n + '/scale'], P[n + '/beta']
    if train:
        axes = (0, 1, 2) if conv else (0,)
        mu = x.mean(dim=axes)
        var = ((x - mu) ** 2).mean(dim=axes)
        S[n + '/pop_mean'] = S[n + '/pop_mean'] * decay + mu.detach() * (1 - decay)
        S[n + '/pop_var'] = S[n + '/pop_var'] * decay + var.detach() * (1 - decay)
    else:
        mu, var = S[n + '/pop_mean'], S[n + '/pop_var']
    return qc((x - mu) * torch.rsqrt(var + 0.001) * scale + beta)


def _salimans_init(z, axes, init_scale, eps, nonlin):
    """the data-dependent branch shared by nn.dense / conv2d / deconv2d (nn.py:229-238, 264-274, 306-316):
    x_init = init_scale / sqrt(var + eps) * (x_init - mean), then the nonlinearity."""
    m = z.mean(dim=axes)
    v = ((z - m) ** 2).mean(dim=axes)
    y = (init_scale / torch.sqrt(v + eps)) * (z - m)
    return qc(nonlin(y) if nonlin is not None else y)


def salimans_dense(P, scope, x, nonlin=None, init=False, init_scale=1.0):
    """nn.dense (nn.py:220-252).  init=True: x @ l2_normalize(V,[0]), data-normalised (eps 1e-10); else
    (x @ V) * g / sqrt(sum V^2) + b (no epsilon)."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if init:
        return _salimans_init(matmul_tf(x, l2_normalize(V, (0,))), (0,), init_scale, 1e-10, nonlin)
    sc = g / torch.sqrt((V * V).sum(dim=0))
    y = (matmul_tf(x, V * sc.view(1, -1)) if _QUANT else sc.view(1, -1) * (x @ V)) + b.view(1, -1)
    return qc(nonlin(y) if nonlin is not None else y)


def salimans_conv2d(P, scope, x, stride=1, pad='SAME', nonlin=None, init=False, init_scale=1.0):
    """nn.conv2d (nn.py:255-289)."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if init:
        return _salimans_init(conv2d_tf(x, l2_normalize(V, (0, 1, 2)), stride, pad), (0, 1, 2), init_scale, 1e-8, nonlin)
    y = conv2d_tf(x, g.view(1, 1, 1, -1) * l2_normalize(V, (0, 1, 2)), stride, pad) + b
    return qc(nonlin(y) if nonlin is not None else y)


def salimans_deconv2d(P, scope, x, stride=1, nonlin=None, init=False, init_scale=1.0):
    """nn.deconv2d, pad='SAME' (nn.py:292-332); V is [kh,kw,Cout,Cin], normalised over axes [0,1,3]."""
    V, g, b = P[scope + '/V'], P[scope + '/g'], P[scope + '/b']
    if init:
        return _salimans_init(conv2d_transpose_tf(x, l2_normalize(V, (0, 1, 3)), stride, 'SAME'), (0, 1, 2), init_scale,
                              1e-8, nonlin)
    y = conv2d_transpose_tf(x, g.view(1, 1, -1, 1) * l2_normalize(V, (0, 1, 3)), stride, 'SAME') + b
    return qc(nonlin(y) if nonlin is not None else y)


def salimans_nin(P, scope, x, **kw):
    """nn.nin (nn.py:335-340): reshape -> dense -> reshape."""
    s = x.shape
    return salimans_dense(P, scope, x.reshape(-1, s[-1]), **kw).reshape(s[0], s[1], s[2], -1)


# --------------------------------------------------------------------------------------
# Configs (batch-size constants of Training/Train_goodGAN.py:484-530, 560-607, 641-685)
# --------------------------------------------------------------------------------------


class OracleConfig:
    def __init__(self, data_name, scale=1):
        self.DATA_NAME = data_name
        self.NUM_CLASSES = 10
        self.Z_DIM = 100
        self.BETA1 = 0.5
        self.BATCH_SIZE_G = 100 // scale
        self.BATCH_SIZE_L_D = 20 // scale
        self.BATCH_SIZE_U_D = 80 // scale
        if data_name == 'mnist':
            self.IMAGE_DIM = [28, 28, 1]
            self.BATCH_SIZE_L_C = 100 // scale
            self.BATCH_SIZE_U_C = 100 // scale
            self.LEARNING_RATE, self.CLA_LEARNINIG_RATE, self.FAKE_G_LAMBDA = 1e-3, 3e-4, 0.1
        else:
            self.IMAGE_DIM = [32, 32, 3]
            self.BATCH_SIZE_L_C = 50 // scale
            self.BATCH_SIZE_U_C = 50 // scale
            if data_name == 'cifar10':
                self.LEARNING_RATE, self.CLA_LEARNINIG_RATE, self.FAKE_G_LAMBDA = 3e-4, 3e-3, 0.3
            else:
                self.LEARNING_RATE, self.CLA_LEARNINIG_RATE, self.FAKE_G_LAMBDA = 3e-4, 3e-4, 0.03
        self.BATCH_SIZE = self.BATCH_SIZE_G


# --------------------------------------------------------------------------------------
# Parameter inventory + synthetic init (SURVEY.md §8a variable inventory, §8d init)
# --------------------------------------------------------------------------------------


def _he(rng, shape, fan_in):
    return (rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)


def _tfdefault(rng, shape):
    # tf.random_normal_initializer(0.02) / tf.truncated_normal_initializer(0.02): the positional
    # argument is the MEAN; stddev stays 1.0 (modle_base.py:28,159,248).  Synthetic init keeps the
    # mean but uses a He-scaled spread so a random-init net stays in a sane numeric range.
    fan_in = int(np.prod(shape[:-1]))
    return (0.02 + rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)


def _wn(P, S, rng, scope, vshape, cout, pop_mean=False):
    P[scope + '/V'] = (rng.standard_normal(vshape) * 0.05).astype(np.float32)
    P[scope + '/g'] = np.ones(cout, np.float32)
    P[scope + '/b'] = np.zeros(cout, np.float32)
    if pop_mean:
        S[scope + '/meanOnlyBatchNormalization/pop_mean'] = np.zeros(cout, np.float32)


def _bn(P, S, scope, c):
    P[scope + '/gamma'] = np.ones(c, np.float32)
    P[scope + '/beta'] = np.zeros(c, np.float32)
    S[scope + '/moving_mean'] = np.zeros(c, np.float32)
    S[scope + '/moving_variance'] = np.ones(c, np.float32)


def _dense(P, rng, scope, cin, cout, init):
    n = scope + '/' + scope.split('/')[-1]
    P[n + '/kernel'] = init(rng, (cin, cout))
    P[n + '/bias'] = np.zeros(cout, np.float32)


def _conv(P, rng, scope, k, cin, cout, init, transpose=False):
    n = scope + '/' + scope.split('/')[-1]
    shape = (k, k, cout, cin) if transpose else (k, k, cin, cout)
    P[n + '/kernel'] = init(rng, shape)
    P[n + '/bias'] = np.zeros(cout, np.float32)


def init_params(data_name, seed=1234):
    """Returns (P, S): trainable numpy float32 params and non-trainable running stats,
    keyed by the TF variable names.  Order of insertion == creation order in forward_pass."""
    rng = np.random.default_rng(seed)
    P, S = {}, {}
    he = lambda r, s: _he(r, s, int(np.prod(s[:-1])))
    if data_name == 'cifar10':   # Model/Good_GAN_cifar10.py
        _dense(P, rng, 'good_generator/gg_h0_lin', 110, 8192, he)
        _bn(P, S, 'good_generator/gg_bn0', 8192)
        _conv(P, rng, 'good_generator/gg_dconv0', 5, 522, 256, he, True)
        _bn(P, S, 'good_generator/gg_bn1', 256)
        _conv(P, rng, 'good_generator/gg_dconv1', 5, 266, 128, he, True)
        _bn(P, S, 'good_generator/gg_bn2', 128)
        _conv(P, rng, 'good_generator/gg_dconv2', 5, 138, 3, he, True)
        for n, ci, co in [('conv1_1', 3, 128), ('conv1_2', 128, 128), ('conv1_3', 128, 128),
                          ('conv2_1', 128, 256), ('conv2_2', 256, 256), ('conv2_3', 256, 256),
                          ('conv3', 256, 512)]:
            _wn(P, S, rng, 'classifier/' + n, (3, 3, ci, co), co, True)
        _wn(P, S, rng, 'classifier/NiN1/NiN1', (512, 256), 256, True)
        _wn(P, S, rng, 'classifier/NiN2/NiN2', (256, 128), 128, True)
        _wn(P, S, rng, 'classifier/output_dense', (128, 10), 10, True)
        for n, ci, co in [('conv2d_00', 13, 32), ('conv2d_01', 42, 32), ('conv2d_10', 42, 64),
                          ('conv2d_11', 74, 64), ('conv2d_20', 74, 128), ('conv2d_21', 138, 128)]:
            _conv(P, rng, 'discriminator/' + n, 3, ci, co, he)
        _dense(P, rng, 'discriminator/lin', 138, 1, he)
    elif data_name == 'svhn':    # Model/Good_GAN.py svhn branches
        _dense(P, rng, 'good_generator/gg_h0_lin', 110, 8192, _tfdefault)
        _bn(P, S, 'good_generator/gg_bn0', 512)
        _conv(P, rng, 'good_generator/gg_dconv0', 5, 522, 256, _tfdefault, True)
        _bn(P, S, 'good_generator/gg_bn1', 256)
        _conv(P, rng, 'good_generator/gg_dconv1', 5, 266, 128, _tfdefault, True)
        _bn(P, S, 'good_generator/gg_bn2', 128)
        _wn(P, S, rng, 'good_generator/gg_wndconv0', (5, 5, 3, 138), 3)
        cs = [('c_h0_conv0', 3, 128, 'c_h0_bn0'), ('c_h0_conv1', 128, 128, 'c_h0_bn1'),
              ('c_h0_conv2', 128, 128, 'c_h0_bn2'), ('c_h1_conv0', 128, 256, 'c_h1_bn0'),
              ('c_h1_conv1', 256, 256, 'c_h1_bn1'), ('c_h1_conv2', 256, 256, 'c_h1_bn2'),
              ('c_h2_conv0', 256, 512, 'c_h2_bn0')]
        for n, ci, co, bn in cs:
            _conv(P, rng, 'classifier/' + n, 3, ci, co, _tfdefault)
            _bn(P, S, 'classifier/' + bn, co)
        _wn(P, S, rng, 'classifier/c_h2_nin0', (512, 256), 256)
        _bn(P, S, 'classifier/c_h2_bn1', 256)
        _wn(P, S, rng, 'classifier/c_h2_nin1', (256, 128), 128)
        _bn(P, S, 'classifier/c_h2_bn2', 128)
        _dense(P, rng, 'classifier/c_h2_lin', 128, 10, _tfdefault)
        _bn(P, S, 'classifier/c_h3_bn0', 10)
        for n, ci, co in [('d_h0_wnconv0', 13, 32), ('d_h0_wnconv1', 42, 32), ('d_h1_wnconv0', 42, 64),
                          ('d_h1_wnconv1', 74, 64), ('d_h2_wnconv0', 74, 128), ('d_h2_wnconv1', 148, 128)]:
            _wn(P, S, rng, 'discriminator/' + n, (3, 3, ci, co), co)
        _wn(P, S, rng, 'discriminator/d_h3_wndense', (138, 1), 1)
    elif data_name == 'mnist':   # Model/Good_GAN.py mnist branches
        _dense(P, rng, 'good_generator/gg_h0_lin', 110, 500, _tfdefault)
        _bn(P, S, 'good_generator/gg_bn0', 500)
        _dense(P, rng, 'good_generator/gg_h1_lin', 510, 500, _tfdefault)
        _bn(P, S, 'good_generator/gg_bn1', 500)
        _wn(P, S, rng, 'good_generator/gg_h2_lin', (510, 784), 784)
        for n, ci, co, bn in [('c_h0_conv0', 1, 32, 'c_h0_bn0'), ('c_h1_conv0', 32, 64, 'c_h1_bn0'),
                              ('c_h1_conv1', 64, 64, 'c_h1_bn1'), ('c_h2_conv0', 64, 128, 'c_h2_bn0'),
                              ('c_h2_conv1', 128, 128, 'c_h2_bn1')]:
            _conv(P, rng, 'classifier/' + n, 3, ci, co, _tfdefault)
            _bn(P, S, 'classifier/' + bn, co)
        _dense(P, rng, 'classifier/c_h2_lin', 128, 10, _tfdefault)
        _bn(P, S, 'classifier/c_h3_bn0', 10)
        for n, ci, co in [('d_h0_wndense0', 794, 1000), ('d_h1_wndense0', 1010, 500),
                          ('d_h2_wndense0', 510, 250), ('d_h3_wndense0', 260, 250),
                          ('d_h4_wndense0', 260, 250), ('d_h5_wndense0', 260, 1)]:
            _wn(P, S, rng, 'discriminator/' + n, (ci, co), co)
    else:
        raise ValueError("The specified dataset is not yet implemented!")
    return P, S


def make_zca(seed=1234):
    """Synthetic stand-in for DataSet/cifar_10/cifar10_zca_{mean,mat}.npy (absent from the repo,
    Good_GAN_cifar10.py:289-290): mean 0, seeded random orthogonal 3072x3072."""
    rng = np.random.default_rng(seed + 7)
    q, r = np.linalg.qr(rng.standard_normal((3072, 3072)))
    q = q * np.sign(np.diag(r))
    return np.zeros(3072, np.float32), q.astype(np.float32)


def make_batch(cfg, seed=1234):
    """Synthetic per-rank step inputs, numpy float32."""
    rng = np.random.default_rng(seed)
    lo = 0.0 if cfg.DATA_NAME == 'mnist' else -1.0
    img = lambda n: rng.uniform(lo, 1.0, [n] + cfg.IMAGE_DIM).astype(np.float32)

    def onehot(n):
        y = np.zeros((n, cfg.NUM_CLASSES), np.float32)
        y[np.arange(n), rng.integers(0, cfg.NUM_CLASSES, n)] = 1
        return y
    return dict(z_g=rng.uniform(-1, 1, (cfg.BATCH_SIZE_G, cfg.Z_DIM)).astype(np.float32),
                y_g=onehot(cfg.BATCH_SIZE_G),
                x_l_c=img(cfg.BATCH_SIZE_L_C), y_l_c=onehot(cfg.BATCH_SIZE_L_C),
                x_l_d=img(cfg.BATCH_SIZE_L_D), y_l_d=onehot(cfg.BATCH_SIZE_L_D),
                x_u_d=img(cfg.BATCH_SIZE_U_D), x_u_c=img(cfg.BATCH_SIZE_U_C))


# --------------------------------------------------------------------------------------
# Model builders
# --------------------------------------------------------------------------------------


class OracleModel:
    """Restates Model/Good_GAN_cifar10.py:33-278 and the mnist/svhn branches of Model/Good_GAN.py.
    P: dict name->torch tensor (trainable), S: dict name->torch tensor (running stats)."""

    def __init__(self, cfg, P, S, zca=None):
        self.cfg, self.P, self.S = cfg, P, S
        self.zca = zca           # (mean, mat) torch tensors, cifar10 only
        self.acts = None         # optional dict: captures named intermediate activations

    def _cap(self, name, x):
        if self.acts is not None:
            self.acts[name] = x
        return x

    # ---- generator ----
    def good_generator(self, z, y, rng, tag):
        P, S, name = self.P, self.S, self.cfg.DATA_NAME
        if name == 'mnist':      # Good_GAN.py:19-33
            h = qc(F.softplus(linear_fc(P, 'good_generator/gg_h0_lin', q(torch.cat([z, y], 1)))))
            h = bn_contrib(P, S, 'good_generator/gg_bn0', h, True)
            h = qc(F.softplus(linear_fc(P, 'good_generator/gg_h1_lin', q(torch.cat([h, y], 1)))))
            h = bn_contrib(P, S, 'good_generator/gg_bn1', h, True)
            return qc(torch.sigmoid(WN_dense(P, 'good_generator/gg_h2_lin', q(torch.cat([h, y], 1)))))
        yb = y.view(y.shape[0], 1, 1, -1)
        h = linear_fc(P, 'good_generator/gg_h0_lin', q(torch.cat([z, y], 1)))
        if name == 'cifar10':    # Good_GAN_cifar10.py:40-43  fc -> relu -> BN over [N,8192] -> reshape
            h = bn_contrib(P, S, 'good_generator/gg_bn0', qc(F.relu(h)), True).view(-1, 4, 4, 512)
        else:                    # Good_GAN.py:40-43  fc -> reshape -> relu -> BN over [N,4,4,512]
            h = bn_contrib(P, S, 'good_generator/gg_bn0', qc(F.relu(h.view(-1, 4, 4, 512))), True)
        h = self._cap(tag + '/h0', cond_concat(h, yb))
        h = qc(F.relu(deconv2d_layer(P, 'good_generator/gg_dconv0', h)))
        h = cond_concat(bn_contrib(P, S, 'good_generator/gg_bn1', h, True), yb)
        h = qc(F.relu(deconv2d_layer(P, 'good_generator/gg_dconv1', h)))
        h = cond_concat(bn_contrib(P, S, 'good_generator/gg_bn2', h, True), yb)
        if name == 'cifar10':
            return torch.tanh(deconv2d_layer(P, 'good_generator/gg_dconv2', h))
        return torch.tanh(WN_deconv2d(P, 'good_generator/gg_wndconv0', h))

    # ---- discriminator ----
    def discriminator(self, image, y, rng, tag):
        P, name = self.P, self.cfg.DATA_NAME
        if name == 'mnist':      # Good_GAN.py:93-124
            h = image.reshape(-1, 784)
            h = q(h + 0.2 * rng.normal(tag + '/noise0', h.shape).to(h.dtype))
            h = q(torch.cat([h, y], 1))
            for i, _ in enumerate([1000, 500, 250, 250, 250]):
                h = qc(leaky_relu_tf(WN_dense(P, 'discriminator/d_h%d_wndense0' % i, h)))
                h = q(h + 0.2 * rng.normal(tag + '/noise%d' % (i + 1), h.shape).to(h.dtype))
                h = q(torch.cat([h, y], 1))
            h = WN_dense(P, 'discriminator/d_h5_wndense0', h)
            return torch.sigmoid(h), h
        yb = y.view(y.shape[0], 1, 1, -1)
        if name == 'cifar10':    # Good_GAN_cifar10.py:60-99 (plain tf.layers.conv2d + cifar lrelu)
            names = ['conv2d_00', 'conv2d_01', 'conv2d_10', 'conv2d_11', 'conv2d_20', 'conv2d_21']
            conv = lambda n, x, s: qc(lrelu_cifar(conv2d_layer(P, 'discriminator/' + n, x, s)))
        else:                    # Good_GAN.py:126-165 (WN convs + tf.nn.leaky_relu)
            names = ['d_h0_wnconv0', 'd_h0_wnconv1', 'd_h1_wnconv0', 'd_h1_wnconv1', 'd_h2_wnconv0',
                     'd_h2_wnconv1']
            conv = lambda n, x, s: qc(leaky_relu_tf(WN_conv2d(P, 'discriminator/' + n, x, s)))
        h = dropout_tf(image, rng.keep_mask(tag + '/drop0', image.shape, 0.2), 0.2)
        h = conv(names[0], cond_concat(h, yb), 1)
        h = conv(names[1], cond_concat(h, yb), 2)
        h = dropout_tf(h, rng.keep_mask(tag + '/drop1', h.shape, 0.2), 0.2)
        h = conv(names[2], cond_concat(h, yb), 1)
        h = conv(names[3], cond_concat(h, yb), 2)
        h = dropout_tf(h, rng.keep_mask(tag + '/drop2', h.shape, 0.2), 0.2)
        h = conv(names[4], cond_concat(h, yb), 1)
        h = cond_concat(h, yb)
        if name != 'cifar10':
            h = cond_concat(h, yb)            # the double concat of Good_GAN.py:151-153 (148 channels)
        h = self._cap(tag + '/h5', conv(names[5], h, 1))
        h = q(torch.cat([q(h.mean(dim=(1, 2))), y], 1))     # avg-pool 8x8 == mean over H,W
        if name == 'cifar10':
            h = linear_fc(P, 'discriminator/lin', h)
        else:
            h = WN_dense(P, 'discriminator/d_h3_wndense', h)
        return torch.sigmoid(h), h

    # ---- classifier ----
    def classifier(self, inp, train, rng, tag):
        P, S, name = self.P, self.S, self.cfg.DATA_NAME
        if name == 'cifar10':    # Good_GAN_cifar10.py:101-174
            x = inp.reshape(-1, 32, 32, 3)
            x = q(x + 0.15 * rng.normal(tag + '/noise', x.shape).to(x.dtype))
            # training-mode layers whose mean-only BN + leaky ReLU the tensor-core path fuses into the GEMM epilogue
            # (tgan/ops.py conv2d_mobn): no max pool behind them, 16 / 32 pixel wide
            fz = lambda n: train and n in ('conv1_1', 'conv1_2', 'conv2_1', 'conv2_2')
            for n in ['conv1_1', 'conv1_2', 'conv1_3']:
                x = self._cap(tag + '/' + n, conv2d_WN(P, S, 'classifier/' + n, x, 'SAME', train, fused=fz(n)))
            x = max_pool_tf(x, 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop1', x.shape, 0.5), 0.5)
            for n in ['conv2_1', 'conv2_2', 'conv2_3']:
                x = self._cap(tag + '/' + n, conv2d_WN(P, S, 'classifier/' + n, x, 'SAME', train, fused=fz(n)))
            x = max_pool_tf(x, 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop2', x.shape, 0.5), 0.5)
            x = self._cap(tag + '/conv3', conv2d_WN(P, S, 'classifier/conv3', x, 'VALID', train))
            x = NiN_WN(P, S, 'classifier/NiN1', x, train, lrelu_cifar)
            x = NiN_WN(P, S, 'classifier/NiN2', x, train, lrelu_cifar)
            x = max_pool_tf(x, 6, 1).reshape(x.shape[0], -1)       # 'avg_pool_0' is a MAX pool (:163)
            inter = x
            return dense_WN(P, S, 'classifier/output_dense', x, train, None), inter
        # Good_GAN.py:216-247 (mnist) / :249-299 (svhn): conv(bias) -> leaky_relu -> BN
        blk = lambda c, b, x: bn_contrib(P, S, 'classifier/' + b,
                                         qc(leaky_relu_tf(conv2d_layer(P, 'classifier/' + c, x, 1))), train)
        if name == 'mnist':
            x = inp.reshape(-1, 28, 28, 1)
            x = q(x + 0.3 * rng.normal(tag + '/noise', x.shape).to(x.dtype))
            x = max_pool_tf(blk('c_h0_conv0', 'c_h0_bn0', x), 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop1', x.shape, 0.5), 0.5)
            x = blk('c_h1_conv1', 'c_h1_bn1', blk('c_h1_conv0', 'c_h1_bn0', x))
            x = max_pool_tf(x, 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop2', x.shape, 0.5), 0.5)
            x = blk('c_h2_conv1', 'c_h2_bn1', blk('c_h2_conv0', 'c_h2_bn0', x))
        else:
            x = inp
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop0', x.shape, 0.2), 0.2)
            for i in range(3):
                x = blk('c_h0_conv%d' % i, 'c_h0_bn%d' % i, x)
            x = max_pool_tf(x, 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop1', x.shape, 0.5), 0.5)
            for i in range(3):
                x = blk('c_h1_conv%d' % i, 'c_h1_bn%d' % i, x)
            x = max_pool_tf(x, 2, 2)
            if train:
                x = dropout_tf(x, rng.keep_mask(tag + '/drop2', x.shape, 0.5), 0.5)
            x = blk('c_h2_conv0', 'c_h2_bn0', x)
            for i in range(2):   # NN_Base._nin (modle_base.py:204-209)
                s = x.shape
                h = WN_dense(P, 'classifier/c_h2_nin%d' % i, x.reshape(-1, s[-1])).reshape(s[0], s[1], s[2], -1)
                x = bn_contrib(P, S, 'classifier/c_h2_bn%d' % (i + 1), qc(leaky_relu_tf(h)), train)
        fm = q(x.mean(dim=(1, 2)))
        h = linear_fc(P, 'classifier/c_h2_lin', fm)
        return bn_contrib(P, S, 'classifier/c_h3_bn0', h, train), fm

    def zca_apply(self, x):
        """cifar10_ZCA.apply (Good_GAN_cifar10.py:294-299)."""
        mean, mat = self.zca
        if _QUANT:      # CUDA path: x @ mat on the tensor cores, -(mean @ mat) rides in the epilogue
            return qc(matmul_tf(x.reshape(x.shape[0], -1), mat) - mean @ mat).reshape(x.shape)
        return ((x.reshape(x.shape[0], -1) - mean) @ mat).reshape(x.shape)

    def forward_pass(self, z_g, y_g, x_l_c, y_l_c, x_l_d, y_l_d, x_u_d, x_u_c, train, rng, tag='F'):
        """Good_GAN_cifar10.forward_pass (:204-278) / Good_GAN.forward_pass (Good_GAN.py:428-472):
        the un-pruned graph, one value per returned tensor, fresh noise per pass."""
        cif = self.cfg.DATA_NAME == 'cifar10'
        pre = self.zca_apply if cif else (lambda t: t)
        G = self.good_generator(z_g, y_g, rng, tag + '/G')
        C_real, _ = self.classifier(pre(x_l_c), train, rng, tag + '/C_real')
        C_unl, _ = self.classifier(pre(x_u_c), train, rng, tag + '/C_unl')
        _, unl_oh = argmax_onehot(C_unl)
        C = [C_real, C_unl]
        if cif:
            C_rep, _ = self.classifier(pre(x_u_c), train, rng, tag + '/C_unl_rep')
        C_unl_d, _ = self.classifier(pre(x_u_d), train, rng, tag + '/C_unl_d')
        _, unl_d_oh = argmax_onehot(C_unl_d)
        C_fake, _ = self.classifier(pre(G), train, rng, tag + '/C_fake')
        C += [C_unl_d, C_fake] + ([C_rep] if cif else [])
        X_P, Y_P = torch.cat([x_l_d, x_u_d], 0), torch.cat([y_l_d, unl_d_oh], 0)
        D_real, D_real_l = self.discriminator(X_P, Y_P, rng, tag + '/D_real')
        D_fake, D_fake_l = self.discriminator(G, y_g, rng, tag + '/D_fake')
        D_unl, D_unl_l = self.discriminator(x_u_c, unl_oh, rng, tag + '/D_unl')
        return [G, [D_real, D_real_l, D_fake, D_fake_l, D_unl, D_unl_l], C]


# --------------------------------------------------------------------------------------
# Losses (Training/train_base.py:43-57, 75-84, 113-154)
# --------------------------------------------------------------------------------------


def entropy(logits):
    p = F.softmax(logits, dim=1)
    return (-(p * logits).sum(dim=1, keepdim=True) + torch.logsumexp(logits, dim=1, keepdim=True)).mean()


def balance_entropy(logits):
    q = F.softmax(logits, dim=1).mean(dim=0)
    return -(1.0 / logits.shape[1] * torch.log(q + 1e-12)).sum()


def d_loss_fn(D_real_l, D_fake_l, D_unl_l):
    return (sigmoid_ce(D_real_l, torch.ones_like(D_real_l)).mean()
            + 0.5 * sigmoid_ce(D_fake_l, torch.zeros_like(D_fake_l)).mean()
            + 0.5 * sigmoid_ce(D_unl_l, torch.zeros_like(D_unl_l)).mean())


def g_loss_fn(D_fake_l):
    return 1 / 2 * sigmoid_ce(D_fake_l, torch.ones_like(D_fake_l)).mean()


def c_loss_fn(C_real, C_unl, C_fake, D_unl_l, y_l_c, y_g, lambda_1, C_unl_rep=None, lambda_2=0.0):
    c_real = softmax_ce(C_real, y_l_c).mean()
    c_fake = softmax_ce(C_fake, y_g).mean()
    p = F.softmax(C_unl, dim=1).max(dim=1).values
    s = sigmoid_ce(D_unl_l, torch.ones_like(D_unl_l)).mean(dim=1)
    c_unl = (p * s).mean()
    c_real = c_real + 1e-6 * entropy(C_unl) + 1e-3 * balance_entropy(C_unl)
    loss = 0.01 * 0.5 * c_unl + c_real + lambda_1 * c_fake
    if C_unl_rep is not None:
        loss = loss + lambda_2 * ((C_unl - C_unl_rep) ** 2).mean()
    return loss


# --------------------------------------------------------------------------------------
# Optimiser (tf.train.AdamOptimizer kernel form + tf.train.ExponentialMovingAverage)
# --------------------------------------------------------------------------------------


class TFAdam:
    def __init__(self, names, P, beta1, beta2=0.999, eps=1e-8):
        self.names, self.b1, self.b2, self.eps = list(names), beta1, beta2, eps
        self.m = {n: torch.zeros_like(P[n]) for n in self.names}
        self.v = {n: torch.zeros_like(P[n]) for n in self.names}
        self.b1p, self.b2p = beta1, beta2        # beta*_power accumulators start at beta (TF)

    def apply(self, P, grads, lr):
        a = lr * math.sqrt(1 - self.b2p) / (1 - self.b1p)
        with torch.no_grad():
            for n in self.names:
                g = grads[n]
                self.m[n] += (g - self.m[n]) * (1 - self.b1)
                self.v[n] += (g * g - self.v[n]) * (1 - self.b2)
                P[n] -= a * self.m[n] / (torch.sqrt(self.v[n]) + self.eps)
        self.b1p *= self.b1
        self.b2p *= self.b2


class OracleTrainer:
    """The D -> G -> C step of Training/Train_goodGAN.py:78-103, 249-276 with the per-`sess.run`
    pruning of SURVEY.md §3.3."""

    def __init__(self, data_name, P_np, S_np, zca_np=None, dtype=torch.float64, scale=1):
        self.cfg = OracleConfig(data_name, scale)
        self.dtype = dtype
        self.P = {k: torch.tensor(v, dtype=dtype).requires_grad_(True) for k, v in P_np.items()}
        self.S = {k: torch.tensor(v, dtype=dtype) for k, v in S_np.items()}
        zca = None
        if data_name == 'cifar10':
            zca = tuple(torch.tensor(a, dtype=dtype) for a in zca_np)
        self.model = OracleModel(self.cfg, self.P, self.S, zca)
        self.g_vars = [k for k in self.P if 'good_generator' in k]
        self.d_vars = [k for k in self.P if 'discriminator' in k]
        self.c_vars = [k for k in self.P if 'classifier' in k]
        self.opt_d = TFAdam(self.d_vars, self.P, self.cfg.BETA1)
        self.opt_g = TFAdam(self.g_vars, self.P, self.cfg.BETA1)
        self.opt_c = TFAdam(self.c_vars, self.P, 0.5)
        self.ema = {k: self.P[k].detach().clone() for k in self.c_vars}
        self.last_grads = {}
        self.last_aux = {}

    def _pseudo(self, key, logits):
        idx, oh = argmax_onehot(logits)
        forced = getattr(self, '_labels', {}).get(key)
        if forced is not None:
            idx = torch.as_tensor(np.asarray(forced), dtype=torch.int64)
            oh = F.one_hot(idx, logits.shape[1]).to(logits.dtype)
        return idx, oh

    def _grads(self, loss, names):
        gs = torch.autograd.grad(loss, [self.P[n] for n in names], allow_unused=True)
        return {n: (g if g is not None else torch.zeros_like(self.P[n])) for n, g in zip(names, gs)}

    def step(self, batch, rng, lambda_1, lambda_2=0.0, lr=None, cla_lr=None, train=True, update=True, phases='DGC',
             labels=None):
        """labels (optional): {'idx_unl_d', 'idx_unl', 'idx_unl_c'} int arrays that REPLACE the three pseudo-label
        argmaxes (Good_GAN_cifar10.py:232,237 in phase D, :232 again in phase C).  Used by the reduced-precision
        parity tests: the run under test supplies its own labels so gradients are compared on identical discrete
        routing; the labels themselves are checked separately wherever the top-2 margin exceeds the noise."""
        cfg, m = self.cfg, self.model
        self._labels = labels or {}
        lr = cfg.LEARNING_RATE if lr is None else lr
        cla_lr = cfg.CLA_LEARNINIG_RATE if cla_lr is None else cla_lr
        b = {k: torch.as_tensor(v).to(self.dtype) for k, v in batch.items()}
        cif = cfg.DATA_NAME == 'cifar10'
        pre = m.zca_apply if cif else (lambda t: t)
        if phases == 'C':      # PRE_TRAIN iteration (Train_goodGAN.py:182-224): only sess.run([c_solver, c_loss])
            self._labels = labels or {}
            return self._phase_c(b, rng, lambda_1, lambda_2, cla_lr, train, update)
        d_loss, gd = self.phase_d(b, rng, lr, train, update)
        g_loss, gg = self.phase_g(b, rng, lr, update)
        _, _, c_loss_v = self._phase_c(b, rng, lambda_1, lambda_2, cla_lr, train, update)
        self.last_grads = {'D': gd, 'G': gg, 'C': self.last_grads['C']}
        return d_loss, g_loss, c_loss_v

    def phase_d(self, b, rng, lr, train=True, update=True):
        """sess.run([d_solver, d_loss]) (Train_goodGAN.py:267) -> (d_loss, {name: grad}); b: dict of tensors"""
        m = self.model
        pre = m.zca_apply if self.cfg.DATA_NAME == 'cifar10' else (lambda t: t)
        with torch.no_grad():
            c_unl_d, _ = m.classifier(pre(b['x_u_d']), train, rng, 'D/C_unl_d')
            c_unl, _ = m.classifier(pre(b['x_u_c']), train, rng, 'D/C_unl')
            idx_d, oh_d = self._pseudo('idx_unl_d', c_unl_d)
            idx_u, oh_u = self._pseudo('idx_unl', c_unl)
            G = m.good_generator(b['z_g'], b['y_g'], rng, 'D/G')
        X_P, Y_P = torch.cat([b['x_l_d'], b['x_u_d']], 0), torch.cat([b['y_l_d'], oh_d], 0)
        _, dr = m.discriminator(X_P, Y_P, rng, 'D/D_real')
        _, df = m.discriminator(G, b['y_g'], rng, 'D/D_fake')
        _, du = m.discriminator(b['x_u_c'], oh_u, rng, 'D/D_unl')
        d_loss = d_loss_fn(dr, df, du)
        gd = self._grads(d_loss, self.d_vars)
        self.last_aux['D'] = dict(idx_unl_d=idx_d, idx_unl=idx_u, G=G, c_unl_d=c_unl_d, c_unl=c_unl, logits=(dr.detach(), df.detach(), du.detach()))
        if update:
            self.opt_d.apply(self.P, gd, lr)
        return float(d_loss.detach()), gd

    def phase_g(self, b, rng, lr, update=True):
        """sess.run([g_solver, g_loss]) (Train_goodGAN.py:270)"""
        m = self.model
        G = m.good_generator(b['z_g'], b['y_g'], rng, 'G/G')
        _, df = m.discriminator(G, b['y_g'], rng, 'G/D_fake')
        g_loss = g_loss_fn(df)
        gg = self._grads(g_loss, self.g_vars)
        if update:
            self.opt_g.apply(self.P, gg, lr)
        return float(g_loss.detach()), gg

    def apply_c(self, gc, cla_lr):
        """c_solver: Adam on c_vars, then ema.apply(c_vars) (Train_goodGAN.py:101-103)"""
        self.opt_c.apply(self.P, gc, cla_lr)
        with torch.no_grad():
            for k in self.c_vars:
                self.ema[k] -= (self.ema[k] - self.P[k]) * (1 - 0.9999)

    def tensors(self, batch):
        return {k: torch.as_tensor(v).to(self.dtype) for k, v in batch.items()}

    def _phase_c(self, b, rng, lambda_1, lambda_2, cla_lr, train, update):
        cfg, m = self.cfg, self.model
        cif = cfg.DATA_NAME == 'cifar10'
        pre = m.zca_apply if cif else (lambda t: t)
        # ---- phase C (:275) ----
        c_real, _ = m.classifier(pre(b['x_l_c']), train, rng, 'C/C_real')
        c_unl, _ = m.classifier(pre(b['x_u_c']), train, rng, 'C/C_unl')
        c_rep = m.classifier(pre(b['x_u_c']), train, rng, 'C/C_unl_rep')[0] if cif else None
        with torch.no_grad():
            G = m.good_generator(b['z_g'], b['y_g'], rng, 'C/G')
            idx_c, oh_u = self._pseudo('idx_unl_c', c_unl)
            _, du = m.discriminator(b['x_u_c'], oh_u, rng, 'C/D_unl')
        c_fake, _ = m.classifier(pre(G), train, rng, 'C/C_fake')
        c_loss = c_loss_fn(c_real, c_unl, c_fake, du, b['y_l_c'], b['y_g'], lambda_1, c_rep, lambda_2)
        gc = self._grads(c_loss, self.c_vars)
        self.last_aux['C'] = dict(logits=(c_real.detach(), c_unl.detach(), c_fake.detach(),
                                          None if c_rep is None else c_rep.detach()), d_unl=du, idx_unl_c=idx_c)
        if update:
            self.apply_c(gc, cla_lr)
        self.last_grads = dict(self.last_grads, C=gc)
        return None, None, float(c_loss.detach())

    def evaluate(self, x, rng):
        """validation forward of Train_goodGAN.py:296-351: classifier logits with train=False"""
        m = self.model
        pre = m.zca_apply if self.cfg.DATA_NAME == 'cifar10' else (lambda t: t)
        with torch.no_grad():
            return m.classifier(pre(torch.as_tensor(x).to(self.dtype)), False, rng, 'V/C_real')[0]
