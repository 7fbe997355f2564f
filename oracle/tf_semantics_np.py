"""Definition-level numpy restatement of the TensorFlow-1.x kernel semantics the
Triple-GAN hot path relies on.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function here is a direct, loop-level transcription of the *published* TF
op definition (TensorFlow is a third-party dependency of the reference, version
unpinned, API bracket TF 1.12-1.15; it is not vendored under /root/reference and
is not installable in this image).  They are deliberately slow and obvious: they
exist to pin `oracle/tgan_oracle.py` (the vectorised torch restatement) against an
independent second restatement on small shapes, because the reference ships no
golden vectors ("parity unpinned").

Call sites in the reference that fix which semantics matter:
  tf.nn.conv2d 'SAME'/'VALID'      Model/nn.py:504, Model/modle_base.py:102,161
  tf.nn.conv2d_transpose 'SAME'    Model/modle_base.py:149,250
  tf.nn.max_pool 'SAME' 2x2/s2     Model/Good_GAN_cifar10.py:123,142
  tf.layers.max_pooling2d 'valid'  Model/Good_GAN_cifar10.py:163, Model/Good_GAN.py:223
  tf.argmax / tf.one_hot           Model/Good_GAN_cifar10.py:232,237,259,270
"""
import numpy as np


def same_padding(n, k, s):
    """TF 'SAME': out = ceil(n/s); pad_total = max((out-1)*s + k - n, 0);
    before = pad_total // 2 (the extra pixel goes at the END)."""
    out = -(-n // s)
    pad_total = max((out - 1) * s + k - n, 0)
    before = pad_total // 2
    return out, before, pad_total - before


def conv2d(x, w, stride, padding):
    """x [N,H,W,Cin], w [kh,kw,Cin,Cout] (HWIO), cross-correlation."""
    N, H, W, Cin = x.shape
    kh, kw, _, Cout = w.shape
    if padding == 'SAME':
        Ho, pt, _ = same_padding(H, kh, stride)
        Wo, pl, _ = same_padding(W, kw, stride)
    else:
        Ho, pt = (H - kh) // stride + 1, 0
        Wo, pl = (W - kw) // stride + 1, 0
    y = np.zeros((N, Ho, Wo, Cout), dtype=np.float64)
    for n in range(N):
        for i in range(Ho):
            for j in range(Wo):
                for r in range(kh):
                    for s in range(kw):
                        hi, wi = i * stride + r - pt, j * stride + s - pl
                        if 0 <= hi < H and 0 <= wi < W:
                            y[n, i, j, :] += x[n, hi, wi, :] @ w[r, s]
    return y


def conv2d_transpose(x, w, stride, padding='SAME'):
    """tf.nn.conv2d_transpose == input-gradient of conv2d.
    x [N,h,w,Cin], w [kh,kw,Cout,Cin]; 'SAME' output is [N,h*s,w*s,Cout].
    For every forward-conv tap (i*s + r - pt) the input pixel i scatters into the output."""
    N, h, wd, Cin = x.shape
    kh, kw, Cout, _ = w.shape
    if padding == 'SAME':
        Ho, Wo = h * stride, wd * stride
        _, pt, _ = same_padding(Ho, kh, stride)
        _, pl, _ = same_padding(Wo, kw, stride)
    else:
        Ho, Wo = h * stride + kh - 1, wd * stride + kw - 1
        pt = pl = 0
    y = np.zeros((N, Ho, Wo, Cout), dtype=np.float64)
    for n in range(N):
        for i in range(h):
            for j in range(wd):
                for r in range(kh):
                    for s in range(kw):
                        ho, wo = i * stride + r - pt, j * stride + s - pl
                        if 0 <= ho < Ho and 0 <= wo < Wo:
                            y[n, ho, wo, :] += w[r, s] @ x[n, i, j, :]
    return y


def max_pool(x, k, s, padding):
    N, H, W, C = x.shape
    if padding == 'SAME':
        Ho, pt, _ = same_padding(H, k, s)
        Wo, pl, _ = same_padding(W, k, s)
    else:
        Ho, pt = (H - k) // s + 1, 0
        Wo, pl = (W - k) // s + 1, 0
    y = np.full((N, Ho, Wo, C), -np.inf)
    for i in range(Ho):
        for j in range(Wo):
            for r in range(k):
                for q in range(k):
                    hi, wi = i * s + r - pt, j * s + q - pl
                    if 0 <= hi < H and 0 <= wi < W:
                        y[:, i, j, :] = np.maximum(y[:, i, j, :], x[:, hi, wi, :])
    return y


def argmax_onehot(logits, depth):
    """tf.argmax(axis=1) -> int64 + tf.one_hot(depth), as Eigen's ArgMaxTupleReducer computes it: accumulator
    (0, -FLT_MAX), replaced only by a strictly greater element -> lowest index wins ties, NaN never wins."""
    idx = np.zeros(logits.shape[0], dtype=np.int64)
    for n in range(logits.shape[0]):
        best, m = 0, np.float32(-3.4028234663852886e38)
        for k in range(logits.shape[1]):
            if logits[n, k] > m:
                best, m = k, logits[n, k]
        idx[n] = best
    oh = np.zeros((logits.shape[0], depth), dtype=np.float32)
    oh[np.arange(logits.shape[0]), idx] = 1.0
    return idx, oh


def sigmoid_ce(x, z):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|))."""
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
