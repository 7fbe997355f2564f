"""-m gpu: the tcgen05/TMEM/TMA kernels (tgan_igemm_bf16, tgan_wgrad_bf16) through the public ops, against
the float64 oracle evaluated on the SAME bf16-rounded operands.

Tolerance (relative to max-abs): bf16 operands are exact in both arms, accumulation is fp32 in TMEM, so the
only differences are the fp32 summation order and the final bf16 rounding of activations
(2^-9 = 2e-3 per element) -> 6e-3 for bf16 outputs; fp32 outputs (filter gradients) 2e-3 because the
upstream gradient is itself rounded to bf16 before the MMA.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402
from util_gpu import param, relerr, run_bwd, setup, tnp   # noqa: E402


def bf(a):
    return torch.tensor(np.asarray(a, np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


def T(a, rg=False):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=rg)


@pytest.fixture(autouse=True)
def _ctx():
    setup('bf16')
    yield


CASES = [
    dict(N=3, H=32, W=32, Cin=72, Cout=64, k=3, s=1, pad='SAME'),     # 2 K-chunks, short last chunk
    dict(N=2, H=16, W=16, Cin=128, Cout=256, k=3, s=1, pad='SAME'),   # BN=256
    dict(N=3, H=8, W=8, Cin=64, Cout=128, k=3, s=1, pad='SAME'),      # 2 images / tile, odd N
    dict(N=5, H=8, W=8, Cin=64, Cout=512, k=3, s=1, pad='VALID'),     # conv3: 8x8 -> 6x6, 2 N-tiles
    dict(N=2, H=32, W=32, Cin=48, Cout=32, k=3, s=2, pad='SAME'),     # D conv2d_01: TF SAME pads (0,1)
    dict(N=2, H=16, W=16, Cin=80, Cout=64, k=3, s=2, pad='SAME'),     # D conv2d_11
    dict(N=2, H=32, W=32, Cin=3, Cout=128, k=3, s=1, pad='SAME'),     # conv1_1: K = 3 channels
    dict(N=4, H=28, W=28, Cin=32, Cout=64, k=3, s=1, pad='SAME'),     # MNIST extent (not a power of two)
    dict(N=300, H=1, W=1, Cin=110, Cout=136, k=1, s=1, pad='SAME'),   # dense (tf.matmul)
    dict(N=2, H=16, W=16, Cin=256, Cout=256, k=3, s=1, pad='SAME'),   # conv2_2: Cin 256 -> wgrad N tile 256
    dict(N=3, H=8, W=8, Cin=256, Cout=512, k=3, s=1, pad='VALID'),    # conv3 at its real channel counts
    dict(N=16, H=6, W=6, Cin=512, Cout=256, k=1, s=1, pad='SAME'),    # NiN1
    dict(N=16, H=6, W=6, Cin=256, Cout=128, k=1, s=1, pad='SAME'),    # NiN2
    dict(N=16, H=1, W=1, Cin=110, Cout=8192, k=1, s=1, pad='SAME'),   # G fc 110 -> 8192 (32 N-tiles)
    # the MNIST networks (Good_GAN.py:93-124, :31-57): widths that are not multiples of 8 on either side
    dict(N=300, H=1, W=1, Cin=794, Cout=1000, k=1, s=1, pad='SAME'),
    dict(N=300, H=1, W=1, Cin=1010, Cout=500, k=1, s=1, pad='SAME'),
    dict(N=100, H=1, W=1, Cin=510, Cout=250, k=1, s=1, pad='SAME'),
    dict(N=120, H=1, W=1, Cin=260, Cout=250, k=1, s=1, pad='SAME'),
]


# the heavy classifier layers at the sizes the benchmark runs them (Train_goodGAN.py:566-572): one call of batch
# 100 (G) and the grouped phase-C batch of 250 (50 + 50 + 50 + 100): multi-wave persistent tile schedule, TMEM
# double-buffer phase flips, 49-way split-K filter gradients
FULL = [
    dict(N=100, H=32, W=32, Cin=128, Cout=128, k=3, s=1, pad='SAME'),   # conv1_2 / conv1_3
    dict(N=250, H=32, W=32, Cin=128, Cout=128, k=3, s=1, pad='SAME'),
    dict(N=100, H=16, W=16, Cin=256, Cout=256, k=3, s=1, pad='SAME'),   # conv2_2 / conv2_3
    dict(N=250, H=16, W=16, Cin=256, Cout=256, k=3, s=1, pad='SAME'),
    dict(N=250, H=8, W=8, Cin=256, Cout=512, k=3, s=1, pad='VALID'),    # conv3
    dict(N=250, H=16, W=16, Cin=128, Cout=256, k=3, s=1, pad='SAME'),   # conv2_1
]


@pytest.mark.parametrize('c', CASES + FULL)
def test_tc_conv_fwd_bwd(c):
    from tgan import core, ops, tc
    rng = np.random.default_rng(1)
    dense = c['H'] == 1
    xs = (c['N'], c['Cin']) if dense else (c['N'], c['H'], c['W'], c['Cin'])
    x = bf(rng.standard_normal(xs))
    w = bf(rng.standard_normal((c['k'], c['k'], c['Cin'], c['Cout'])) * 0.1)
    xt, wt = T(x, True), T(w, True)
    yt = O.conv2d_tf(xt.view(c['N'], c['H'], c['W'], c['Cin']), wt, c['s'], c['pad'])
    gy = bf(rng.standard_normal(tuple(yt.shape)))
    yt.backward(T(gy))
    p = param(w.reshape(c['Cin'], c['Cout']) if dense else w)
    with core.recording():
        xv = ops.Var(torch.tensor(x, dtype=torch.float32).cuda().to(torch.bfloat16), xs, requires_grad=True)
        out = ops.conv2d(xv, ops.PlainWeight(p), c['k'], c['k'], c['s'], c['pad'])
        assert out.data.dtype == torch.bfloat16          # proves the tcgen05 route was taken
        fwd = tnp(out.data).reshape(tuple(yt.shape))
        run_bwd(out, gy.reshape(out.shape))
    assert relerr(fwd, yt.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad).reshape(xs), xt.grad.numpy()) < 6e-3
    assert relerr(tnp(p.grad).reshape(w.shape), wt.grad.numpy()) < 2e-3


@pytest.mark.parametrize('c', [dict(N=3, h=4, w=4, Cin=24, Cout=16, k=5, s=2), dict(N=2, h=8, w=8, Cin=272, Cout=128, k=5, s=2),
                               dict(N=5, h=4, w=4, Cin=528, Cout=256, k=5, s=2)])
def test_tc_deconv_fwd_bwd(c):
    from tgan import core, ops
    rng = np.random.default_rng(2)
    x = bf(rng.standard_normal((c['N'], c['h'], c['w'], c['Cin'])))
    w = bf(rng.standard_normal((c['k'], c['k'], c['Cout'], c['Cin'])) * 0.05)
    xt, wt = T(x, True), T(w, True)
    yt = O.conv2d_transpose_tf(xt, wt, c['s'])
    gy = bf(rng.standard_normal(tuple(yt.shape)))
    yt.backward(T(gy))
    p = param(w)
    with core.recording():
        xv = ops.Var(torch.tensor(x, dtype=torch.float32).cuda().to(torch.bfloat16), x.shape, requires_grad=True)
        out = ops.conv2d_transpose(xv, ops.PlainWeight(p), c['k'], c['k'], c['s'])
        assert out.data.dtype == torch.bfloat16
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 6e-3
    assert relerr(tnp(p.grad), wt.grad.numpy()) < 2e-3


def test_tc_weightnorm_rgb_deconv():
    """the Good_GAN generator's last layer (Good_GAN.py:57: `_WN_deconv2d`, 138 -> 3 channels, norm over axes [0, 1, 3]) on
    the tensor-core path: the packed operands fold g / ||V|| per OUTPUT channel -- in the input-gradient operand the GEMM K
    axis is (row, column, channel) flattened, so the scale index is taken modulo the channel count (tgan_pack_desc.scale_mod)"""
    from tgan import core, ops
    rng = np.random.default_rng(8)
    N, Cin, Cout = 4, 138, 3
    x = bf(rng.standard_normal((N, 16, 16, Cin)))
    V = rng.standard_normal((5, 5, Cout, Cin)) * 0.05
    g = rng.uniform(0.5, 1.5, Cout)
    xt, Vt, gt = T(x, True), T(V, True), T(g, True)
    W = gt.view(1, 1, -1, 1) * O.l2_normalize(Vt, (0, 1, 3))
    Wq = W + (torch.tensor(bf(W.detach().numpy())) - W.detach())      # the MMA reads the bf16-rounded effective filter
    yt = O.conv2d_transpose_tf(xt, Wq, 2)
    gy = bf(rng.standard_normal(tuple(yt.shape)))
    yt.backward(T(gy))
    pV, pg = param(V), param(g)
    core.ctx.store.bump()
    with core.recording():
        xv = ops.Var(torch.tensor(x, dtype=torch.float32).cuda().to(torch.bfloat16), x.shape, requires_grad=True)
        out = ops.conv2d_transpose(xv, ops.WNWeight(pV, pg, 25, Cout, Cin, 1), 5, 5, 2)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 6e-3
    assert relerr(tnp(pV.grad), Vt.grad.numpy()) < 4e-3
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 4e-3


def test_tc_padded_concat_input():
    """label-concatenated activations (13/42/74/138/522 channels) are stored with a zero-padded pixel stride."""
    from tgan import core, ops
    rng = np.random.default_rng(3)
    x = bf(rng.standard_normal((2, 16, 16, 32)))
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 2)]
    w = bf(rng.standard_normal((3, 3, 42, 64)) * 0.1)
    xt, wt = T(x, True), T(w, True)
    yt = O.conv2d_tf(O.cond_concat(xt, T(y).view(2, 1, 1, 10)), wt, 1, 'SAME')
    gy = bf(rng.standard_normal(tuple(yt.shape)))
    yt.backward(T(gy))
    p = param(w)
    with core.recording():
        xv = ops.Var(torch.tensor(x, dtype=torch.float32).cuda().to(torch.bfloat16), x.shape, requires_grad=True)
        h = ops.concat_label(xv, ops.Var(torch.tensor(y).cuda(), y.shape))
        assert h.ld == 48 and h.C == 42
        out = ops.conv2d(h, ops.PlainWeight(p), 3, 3, 1, 'SAME')
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 6e-3
    assert relerr(tnp(p.grad), wt.grad.numpy()) < 2e-3


def test_tc_full_size_linearity():
    """BASELINE-size layer (conv1_2, batch 100): size-independent property instead of the oracle --
    conv(a*x1 + x2) == a*conv(x1) + conv(x2) within bf16 rounding, and a zero input gives exact zeros."""
    from tgan import ops
    rng = np.random.default_rng(4)
    w = param(bf(rng.standard_normal((3, 3, 128, 128)) * 0.03))
    mk = lambda a: ops.Var(torch.tensor(a, dtype=torch.float32).cuda().to(torch.bfloat16), a.shape)
    x1 = bf(rng.standard_normal((100, 32, 32, 128)))
    x2 = bf(rng.standard_normal((100, 32, 32, 128)))
    f = lambda a: tnp(ops.conv2d(mk(a), ops.PlainWeight(w), 3, 3, 1, 'SAME').data).astype(np.float64)
    y1, y2, y12 = f(x1), f(x2), f(bf(2.0 * x1 + x2))
    assert relerr(y12, 2.0 * y1 + y2) < 2e-2
    assert np.abs(f(np.zeros_like(x1))).max() == 0.0
