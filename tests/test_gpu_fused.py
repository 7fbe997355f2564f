"""-m gpu: op-level parity of the fused / grouped / re-routed tensor-core paths against the float64 oracle on the SAME
bf16-rounded operands (tolerances as in test_gpu_tc.py: 6e-3 relative-to-max for bf16 outputs, 2e-3 for fp32 filter
gradients):

  * grouped batches: several calls of one layer in one pass, per-call mean-only-BN statistics (segments), pop_mean
    updated per call in call order, per-call mean subtraction in the backward
  * max-pool + dropout as one kernel (forward and backward)
  * deferred contraction: conv + bias + leaky-ReLU + label concat in one GEMM epilogue (+ label-plane fill), input
    gradient restricted to the non-label channels
  * the generator's 3-channel transposed convolution (parity classes forward, im2col GEMMs backward)
  * row-halo implicit GEMM / wgrad against the per-tap kernels (TGAN_IGEMM_NO_HALO / TGAN_WGRAD_NO_HALO)
"""
import os

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


def dev(a, dtype=torch.bfloat16):
    return torch.tensor(np.asarray(a, np.float32)).cuda().to(dtype)


@pytest.fixture(autouse=True)
def _ctx():
    setup('bf16')
    yield


@pytest.mark.parametrize('geom', [dict(H=16, C=64, Cout=128, segs=[3, 2, 2, 4]), dict(H=32, C=32, Cout=64, segs=[2, 3]),
                                  dict(H=6, C=128, Cout=64, segs=[5, 4, 7], k=1)])
def test_grouped_conv_mobn_segments(geom):
    """conv2d_WN-style layer on a grouped batch == the separate calls of the reference graph."""
    from tgan import core, ops
    rng = np.random.default_rng(11)
    segs, H, C, Cout, k = geom['segs'], geom['H'], geom['C'], geom['Cout'], geom.get('k', 3)
    N = sum(segs)
    x = bf(rng.standard_normal((N, H, H, C)))
    w = bf(rng.standard_normal((k, k, C, Cout)) * 0.1)
    b = rng.standard_normal(Cout) * 0.1
    pm = rng.standard_normal(Cout) * 0.1
    gy = bf(rng.standard_normal((N, H, H, Cout)))
    # oracle: one call per segment, in order (pop_mean is updated sequentially)
    xt, wt, bt = T(x, True), T(w, True), T(b, True)
    S = {'pm': T(pm)}
    ys, o = [], 0
    for n in segs:
        z = O.conv2d_tf(xt[o:o + n], wt, 1, 'SAME')
        # the CUDA path: batch mean from the fp32 accumulators in the GEMM epilogue (as the fp32 reference would), applied to
        # the bf16-stored z
        m = z.mean(dim=(0, 1, 2))
        zr = T(bf(z.detach().numpy())) + (z - z.detach())
        S['pm'] = S['pm'] * 0.9 + m.detach() * 0.1                   # nn.py:176-183, once per call, in call order
        ys.append(O.lrelu_cifar(zr - m + bt))
        o += n
    yt = torch.cat(ys, 0)
    yt.backward(T(gy))
    pw, pb, ppm = param(w), param(b), param(pm, False)
    with core.recording():
        parts, o = [], 0
        for n in segs:
            parts.append(ops.Var(dev(x[o:o + n]), (n, H, H, C)))
            o += n
        xg = ops.group_batch(parts)
        assert xg.aux['segs'] == segs
        xg.requires_grad = True
        z = ops.conv2d(xg, ops.PlainWeight(pw), k, k, 1, 'SAME', colsum=True)
        out = ops.mobn_act(z, pb, ppm, True, 'lrelu', 0.2)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 1e-2
    assert relerr(tnp(ppm.data), S['pm'].numpy()) < 2e-3
    # db: the two arms round z to bf16 after different fp32 summation orders, so a handful of elements with |u| below one
    # bf16 ulp sit on the other side of the leaky-ReLU kink; each flips one 0.8*dy term of a channel's bias gradient
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 2.5e-2
    assert relerr(tnp(pw.grad), wt.grad.numpy()) < 6e-3
    assert relerr(tnp(xg.grad), xt.grad.numpy()) < 1e-2


@pytest.mark.parametrize('shape,rate', [((6, 32, 32, 128), 0.5), ((5, 16, 16, 256), 0.5), ((3, 8, 8, 64), 0.2)])
def test_maxpool_dropout_fused(shape, rate):
    from tgan import core, ops
    rng = np.random.default_rng(12)
    x = bf(rng.standard_normal(shape))
    oshape = (shape[0], shape[1] // 2, shape[2] // 2, shape[3])
    src = O.TagRNG(5)
    core.ctx.rng = core.InjectedSource(src)
    xt = T(x, True)
    yt = O.dropout_tf(O.max_pool_tf(xt, 2, 2), src.keep_mask('t/drop', oshape, rate), rate)
    gy = bf(rng.standard_normal(oshape))
    yt.backward(T(gy))
    with core.recording():
        xv = ops.Var(dev(x), shape, requires_grad=True)
        pooled = ops.max_pool2(xv)
        assert pooled._data is None, 'the pool must wait for the dropout that follows it'
        out = ops.dropout(pooled, rate, 't/drop', True)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, bf(yt.detach().numpy())) < 1e-6          # exact up to the bf16 store
    assert relerr(tnp(xv.grad), bf(xt.grad.numpy())) < 1e-6
    # a pool that is NOT followed by a dropout materialises as a plain pool
    with core.no_grad():
        p2 = ops.max_pool2(ops.Var(dev(x), shape))
        assert relerr(tnp(p2.data), O.max_pool_tf(T(x), 2, 2).numpy()) == 0.0


def test_deferred_conv_bias_lrelu_concat():
    """one D block: conv (+bias) -> leaky ReLU -> label concat -> conv, all on the tensor-core path"""
    from tgan import core, ops
    rng = np.random.default_rng(13)
    N, H, C, C1, C2, K = 5, 16, 42, 64, 64, 10
    x = bf(rng.standard_normal((N, H, H, C)))
    y = np.eye(K, dtype=np.float32)[rng.integers(0, K, N)]
    w1, b1 = bf(rng.standard_normal((3, 3, C, C1)) * 0.1), rng.standard_normal(C1) * 0.1
    w2, b2 = bf(rng.standard_normal((3, 3, C1 + K, C2)) * 0.1), rng.standard_normal(C2) * 0.1
    xt, w1t, b1t, w2t, b2t = T(x, True), T(w1, True), T(b1, True), T(w2, True), T(b2, True)
    h = O.lrelu_cifar(O.conv2d_tf(xt, w1t, 1, 'SAME') + b1t)
    h = T(bf(h.detach().numpy())) + (h - h.detach())              # bf16 activation storage
    yt = O.lrelu_cifar(O.conv2d_tf(O.cond_concat(h, T(y).view(N, 1, 1, K)), w2t, 1, 'SAME') + b2t)
    gy = bf(rng.standard_normal(tuple(yt.shape)))
    yt.backward(T(gy))
    pw1, pb1, pw2, pb2 = param(w1), param(b1), param(w2), param(b2)
    with core.recording():
        xv = ops.Var(dev(x), x.shape, requires_grad=True)
        z1 = ops.lazy_bias(ops.conv2d(xv, ops.PlainWeight(pw1), 3, 3, 1, 'SAME'), pb1)
        a1 = ops.activation(z1, 'lrelu', 0.2)
        assert a1._data is None, 'bias and activation stay pending until the consumer is known'
        c1 = ops.concat_label(a1, ops.Var(torch.tensor(y).cuda(), y.shape))
        assert c1.ld == 80 and c1.C == C1 + K and a1.ld == 80, 'the GEMM wrote straight into the padded concat buffer'
        z2 = ops.lazy_bias(ops.conv2d(c1, ops.PlainWeight(pw2), 3, 3, 1, 'SAME'), pb2)
        out = ops.activation(z2, 'lrelu', 0.2)
        fwd = tnp(out.data)
        lab = tnp(c1.data)[..., C1:C1 + K]
        run_bwd(out, gy)
    assert np.array_equal(lab, np.broadcast_to(y[:, None, None, :], lab.shape))
    assert relerr(fwd, yt.detach().numpy()) < 1e-2
    for p, t in ((pw1, w1t), (pb1, b1t), (pw2, w2t), (pb2, b2t)):
        assert relerr(tnp(p.grad), t.grad.numpy()) < 1e-2
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 1e-2


@pytest.mark.parametrize('c', [dict(N=4, h=16, w=16, Cin=138, Cout=3), dict(N=3, h=8, w=8, Cin=64, Cout=3)])
def test_skinny_deconv_tanh(c):
    """gg_dconv2 (Good_GAN_cifar10.py:56): 5x5/s2 transposed conv to 3 channels + bias + tanh"""
    from tgan import core, ops
    rng = np.random.default_rng(14)
    x = bf(rng.standard_normal((c['N'], c['h'], c['w'], c['Cin'])))
    w = bf(rng.standard_normal((5, 5, c['Cout'], c['Cin'])) * 0.05)
    b = rng.standard_normal(c['Cout']) * 0.1
    xt, wt, bt = T(x, True), T(w, True), T(b, True)
    yt = torch.tanh(O.conv2d_transpose_tf(xt, wt, 2) + bt)
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    pw, pb = param(w), param(b)
    with core.recording():
        xv = ops.Var(dev(x), x.shape, requires_grad=True)
        out = ops.activation(ops.lazy_bias(ops.conv2d_transpose(xv, ops.PlainWeight(pw), 5, 5, 2), pb), 'tanh')
        assert out.data.dtype == torch.float32 and out.shape[-1] == c['Cout']
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 1e-2
    assert relerr(tnp(pw.grad), wt.grad.numpy()) < 6e-3
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 6e-3


@pytest.mark.parametrize('c', [dict(N=7, H=32, Cin=128, Cout=128), dict(N=5, H=16, Cin=256, Cout=256),
                               dict(N=9, H=16, Cin=72, Cout=192), dict(N=3, H=32, Cin=64, Cout=32)])
def test_row_halo_equals_per_tap(c):
    """the row-halo kernels (shifted UMMA descriptors into one (rows+2)-row box) against the per-tap kernels on the same
    operands: identical products, only the fp32 accumulation order differs"""
    from tgan import core, ops
    rng = np.random.default_rng(15)
    x = bf(rng.standard_normal((c['N'], c['H'], c['H'], c['Cin'])))
    w = bf(rng.standard_normal((3, 3, c['Cin'], c['Cout'])) * 0.1)
    gy = bf(rng.standard_normal((c['N'], c['H'], c['H'], c['Cout'])))
    res = []
    for off in ('0', '1'):
        os.environ['TGAN_IGEMM_NO_HALO'] = off
        os.environ['TGAN_WGRAD_NO_HALO'] = off
        try:
            setup('bf16')
            p = param(w)
            with core.recording():
                xv = ops.Var(dev(x), x.shape, requires_grad=True)
                out = ops.conv2d(xv, ops.PlainWeight(p), 3, 3, 1, 'SAME')
                fwd = tnp(out.data)
                run_bwd(out, gy)
            res.append((fwd, tnp(xv.grad), tnp(p.grad)))
        finally:
            os.environ.pop('TGAN_IGEMM_NO_HALO')
            os.environ.pop('TGAN_WGRAD_NO_HALO')
    for a, b_ in zip(*res):
        assert relerr(a, b_) < 4e-3
    yt = O.conv2d_tf(T(x), T(w), 1, 'SAME')
    assert relerr(res[0][0], yt.numpy()) < 6e-3


@pytest.mark.parametrize('geom', [dict(H=16, C=64, Cout=128, segs=[3, 2, 4], rate=0.5), dict(H=32, C=32, Cout=64, segs=[4], rate=0.5),
                                  dict(H=8, C=64, Cout=256, segs=[2, 3], rate=0.0)])
def test_mobn_pool_dropout_one_pass(geom):
    """conv -> mean-only BN + leaky ReLU -> 2x2 max pool -> dropout (Good_GAN_cifar10.py:118-124): the last three run as ONE
    kernel that never writes the full-resolution activation; backward through the pooled output only."""
    from tgan import core, ops
    rng = np.random.default_rng(21)
    segs, H, C, Cout, rate = geom['segs'], geom['H'], geom['C'], geom['Cout'], geom['rate']
    N = sum(segs)
    x = bf(rng.standard_normal((N, H, H, C)))
    w = bf(rng.standard_normal((3, 3, C, Cout)) * 0.1)
    b = rng.standard_normal(Cout) * 0.1
    pm = rng.standard_normal(Cout) * 0.1
    oshape = (N, H // 2, H // 2, Cout)
    gy = bf(rng.standard_normal(oshape))
    src = O.TagRNG(9)
    core.ctx.rng = core.InjectedSource(src)
    xt, wt, bt = T(x, True), T(w, True), T(b, True)
    S = {'pm': T(pm)}
    ys, o = [], 0
    for n in segs:
        z = O.conv2d_tf(xt[o:o + n], wt, 1, 'SAME')
        m = z.mean(dim=(0, 1, 2))
        zr = T(bf(z.detach().numpy())) + (z - z.detach())
        S['pm'] = S['pm'] * 0.9 + m.detach() * 0.1
        y = O.lrelu_cifar(zr - m + bt)
        ys.append(T(bf(y.detach().numpy())) + (y - y.detach()))          # y is rounded to bf16 before the pool
        o += n
    pooled = O.max_pool_tf(torch.cat(ys, 0), 2, 2)
    yt = O.dropout_tf(pooled, src.keep_mask('t/drop', oshape, rate), rate) if rate > 0 else pooled
    yt.backward(T(gy))
    pw, pb, ppm = param(w), param(b), param(pm, False)
    with core.recording():
        parts, o = [], 0
        for n in segs:
            parts.append(ops.Var(dev(x[o:o + n]), (n, H, H, C)))
            o += n
        xg = ops.group_batch(parts)
        xg.requires_grad = True
        z = ops.conv2d(xg, ops.PlainWeight(pw), 3, 3, 1, 'SAME', colsum=True)
        act = ops.mobn_act(z, pb, ppm, True, 'lrelu', 0.2)
        assert act._data is None, 'the mean-only-BN apply waits for its consumer'
        pl = ops.max_pool2(act)
        out = ops.dropout(pl, rate, 't/drop', True) if rate > 0 else pl
        fwd = tnp(out.data)
        assert act._data is None, 'the full-resolution activation must not have been materialised'
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 1e-2
    assert relerr(tnp(ppm.data), S['pm'].numpy()) < 2e-3
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 2.5e-2
    assert relerr(tnp(pw.grad), wt.grad.numpy()) < 1e-2
    assert relerr(tnp(xg.grad), xt.grad.numpy()) < 1e-2


def test_dropout_and_batchnorm_write_into_the_concat():
    """`dropout -> _conv_cond_concat` (D) and `batch_norm -> _conv_cond_concat` (G) are one launch each: the producer writes
    the label-concatenated tensor itself"""
    from tgan import core, ops
    rng = np.random.default_rng(31)
    N, H, C, K = 6, 8, 64, 10
    x = bf(rng.standard_normal((N, H, H, C)))
    y = np.eye(K, dtype=np.float32)[rng.integers(0, K, N)]
    yv = lambda: ops.Var(torch.tensor(y).cuda(), y.shape)
    src = O.TagRNG(3)
    core.ctx.rng = core.InjectedSource(src)
    # dropout -> concat
    xt = T(x, True)
    ref = O.cond_concat(O.dropout_tf(xt, src.keep_mask('d', x.shape, 0.2), 0.2), T(y).view(N, 1, 1, K))
    gy = rng.standard_normal(tuple(ref.shape))
    ref.backward(T(gy))
    with core.recording():
        xv = ops.Var(dev(x), x.shape, requires_grad=True)
        d = ops.dropout(xv, 0.2, 'd', True)
        assert d._data is None
        c = ops.concat_label(d, yv())
        assert c.ld == 80 and d.ld == 80 and c.data.data_ptr() == d.data.data_ptr()
        got = tnp(c.data)[..., :C + K]
        assert np.abs(tnp(c.data)[..., C + K:]).max() == 0.0
        run_bwd(c, gy)
    assert relerr(got, bf(ref.detach().numpy())) < 1e-6
    assert relerr(tnp(xv.grad), bf(xt.grad.numpy())) < 8e-3      # dy and dy*1.25 are each rounded to bf16
    # batch norm -> concat
    gam, bet = rng.standard_normal(C) * 0.1 + 1, rng.standard_normal(C) * 0.1
    xt, gt, bt = T(x, True), T(gam, True), T(bet, True)
    P = {'s/gamma': gt, 's/beta': bt}
    S = {'s/moving_mean': T(np.zeros(C)), 's/moving_variance': T(np.ones(C))}
    ref = O.cond_concat(O.bn_contrib(P, S, 's', xt, True), T(y).view(N, 1, 1, K))
    ref.backward(T(gy))
    pg, pb = param(gam), param(bet)
    mm, mv = param(np.zeros(C), False), param(np.ones(C), False)
    with core.recording():
        xv = ops.Var(dev(x), x.shape, requires_grad=True)
        h = ops.batch_norm(xv, pg, pb, mm, mv, True)
        assert h._data is None
        c = ops.concat_label(h, yv())
        assert c.ld == 80 and c.data.data_ptr() == h.data.data_ptr()
        got = tnp(c.data)[..., :C + K]
        run_bwd(c, gy)
    assert relerr(got, ref.detach().numpy()) < 6e-3
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 1e-2
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 1e-2 and relerr(tnp(pb.grad), bt.grad.numpy()) < 1e-2


def _wn_layer_params(rng, name, cin, cout):
    return {name + '/V': rng.standard_normal((3, 3, cin, cout)) * 0.05, name + '/g': rng.uniform(0.7, 1.3, cout),
            name + '/b': rng.standard_normal(cout) * 0.1}, {name + '/meanOnlyBatchNormalization/pop_mean': rng.standard_normal(cout) * 0.1}


# (depth, bound on every parameter gradient relative to max-abs): bf16 rounding noise grows with the number of layers the
# gradient crosses -- the two-kernel path shows the same or larger errors (measured: 5 layers 4e-2 ... 1.6e-1 either way)
DEPTH_TOL = {1: 2e-2, 2: 6e-2, 5: 2.5e-1}


@pytest.mark.parametrize('depth', [1, 2, 5])
@pytest.mark.parametrize('segs', [[2, 1, 3], [4]])
@pytest.mark.parametrize('fusion', [True, False])
def test_mobn_fused_into_gemm_epilogue_chain(segs, fusion, depth):
    """The CIFAR-10 classifier's head chain conv1_1 -> conv1_2 -> conv1_3 -> max pool -> dropout -> conv2_1 -> conv2_2
    (nn.conv2d_WN, Good_GAN_cifar10.py:106-131) on a grouped batch, training mode: with mean-only batch norm FUSED
    into the GEMM epilogues (batch mean from border-class input sums, lrelu mask, input-gradient epilogue applying
    lrelu') and with the two-kernel path (TGAN_NO_MOBN_FUSION) -- both against the float64 oracle with the same bf16
    rounding points, forward value, pop_mean of every layer and every parameter gradient."""
    from tgan import core, nn, ops
    old = os.environ.pop('TGAN_NO_MOBN_FUSION', None)
    if not fusion:
        os.environ['TGAN_NO_MOBN_FUSION'] = '1'
    try:
        rng = np.random.default_rng(21)
        N = sum(segs)
        chans = [('conv1_1', 3, 128), ('conv1_2', 128, 128), ('conv1_3', 128, 128), ('conv2_1', 128, 256), ('conv2_2', 256, 256)]
        Pn, Sn = {}, {}
        for n, ci, co in chans:
            p_, s_ = _wn_layer_params(rng, 'classifier/' + n, ci, co)
            Pn.update(p_)
            Sn.update(s_)
        chans = chans[:depth]
        x = bf(rng.standard_normal((N, 32, 32, 3)))
        keep = (rng.uniform(size=(N, 16, 16, 128)) >= 0.5).astype(np.uint8)
        R = bf(rng.standard_normal((N, 16, 16, 256) if depth == 5 else (N, 32, 32, 128)))
        fused = lambda n: fusion and n in ('conv1_1', 'conv1_2', 'conv2_1', 'conv2_2')

        def oracle_net(P, S, xt):
            outs, o = [], 0
            for ns in segs:                      # one call per segment, in call order (pop_mean chain)
                h = xt[o:o + ns]
                for n in ('conv1_1', 'conv1_2', 'conv1_3')[:depth]:
                    h = O.conv2d_WN(P, S, 'classifier/' + n, h, 'SAME', True, fused=fused(n))
                if depth == 5:
                    h = O.dropout_tf(O.max_pool_tf(h, 2, 2), torch.tensor(keep[o:o + ns]), 0.5)
                    for n in ('conv2_1', 'conv2_2'):
                        h = O.conv2d_WN(P, S, 'classifier/' + n, h, 'SAME', True, fused=fused(n))
                outs.append(h)
                o += ns
            return torch.cat(outs, 0)
        P = {k: T(v, True) for k, v in Pn.items()}
        S = {k: T(v) for k, v in Sn.items()}
        with O.quantized():
            yt = oracle_net(P, S, T(x))
            (yt * T(R)).sum().backward()
        # CUDA
        store = core.ctx.store
        prm = {}
        for k, v in list(Pn.items()) + list(Sn.items()):
            parts = k.split('/')
            import contextlib
            with contextlib.ExitStack() as st:
                for s_ in parts[:-1]:
                    st.enter_context(core.variable_scope(s_))
                p = core.get_variable(parts[-1], list(np.shape(v)), core.constant_initializer(0.0), trainable=k in Pn)
            p.data = torch.from_numpy(np.asarray(v, np.float32).copy()).cuda()
            p.grad = torch.zeros_like(p.data)
            p.requires_grad = k in Pn
            prm[k] = p
        store.bump()
        store.reuse[0] = True

        class FixedMask:
            injected = True

            def keep_mask(self, tag, shape, rate):
                return torch.from_numpy(keep).cuda()
        core.ctx.rng = FixedMask()
        lrelu = nn.leaky_relu
        kw = dict(init=False, use_weight_normalization=True, use_mean_only_batch_normalization=True, deterministic=False)
        ops.arena_reset()
        with core.recording(), core.variable_scope('classifier', reuse=True):
            xv = ops.Var(dev(x), x.shape)
            if len(segs) > 1:
                xv.aux = {'segs': list(segs)}
            h = xv
            for n, _, co in chans[:3]:
                h = nn.conv2d_WN(h, num_filters=co, name=n, nonlinearity=lrelu, **kw)
            if depth == 5:
                h = ops.dropout(ops.max_pool2(h), 0.5, 'T/drop1')
                for n, _, co in chans[3:]:
                    h = nn.conv2d_WN(h, num_filters=co, name=n, nonlinearity=lrelu, **kw)
            fwd = tnp(h.data)
            if fusion:
                assert h.aux.get('mask') is not None      # proves the fused route ran
            run_bwd(h, R)
        assert relerr(fwd, yt.detach().numpy()) < 2e-2
        for k in Sn:
            if k.split('/')[1] in [c[0] for c in chans]:
                assert relerr(tnp(prm[k].data), S[k].numpy()) < 5e-3, k
        used = [k for k in Pn if k.split('/')[1] in [c[0] for c in chans]]
        gscale = max(float(P[k].grad.abs().max()) for k in used)
        for k in used:
            ref = P[k].grad.numpy()
            den = max(np.abs(ref).max(), 1e-2 * gscale)
            e = np.abs(tnp(prm[k].grad) - ref).max() / den
            print(fusion, segs, depth, k, '%.3e' % e)
            assert e < DEPTH_TOL[depth], (k, e)
    finally:
        os.environ.pop('TGAN_NO_MOBN_FUSION', None)
        if old is not None:
            os.environ['TGAN_NO_MOBN_FUSION'] = old
