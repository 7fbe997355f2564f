"""-m gpu: per-op forward/backward parity of the CUDA path (through the C ABI) against the oracle.

Tolerances (relative to max-abs): fp32 CUDA-core path 2e-5 vs the float64 oracle
(fp32 accumulation over K <= ~2k terms); integer / index results bit-exact.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tf_semantics_np as tfnp          # noqa: E402
from oracle import tgan_oracle as O                 # noqa: E402
from util_gpu import param, relerr, run_bwd, setup, tnp, var   # noqa: E402

TOL = 2e-5


def T(a, rg=False):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=rg)


@pytest.fixture(autouse=True)
def _ctx():
    setup('fp32')
    yield


@pytest.mark.parametrize('ta,tb,M,N,K,splits', [(0, 0, 70, 33, 50, 1), (0, 1, 129, 65, 17, 1), (1, 0, 45, 130, 300, 1),
                                                (1, 0, 64, 64, 5000, 7), (1, 1, 10, 3, 4100, 5)])
def test_sgemm(ta, tb, M, N, K, splits):
    from tgan import _lib
    rng = np.random.default_rng(0)
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    a, b, c = (torch.from_numpy(t).cuda() for t in (A, B, C0.copy()))
    ws = torch.empty(max(1, splits * M * N), device='cuda')
    _lib.call('tgan_sgemm', ta, tb, M, N, K, 0.5, a.data_ptr(), A.shape[1], b.data_ptr(), B.shape[1], 2.0,
              c.data_ptr(), N, splits, ws.data_ptr(), 0)
    torch.cuda.synchronize()
    ref = 0.5 * ((A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)) + 2.0 * C0
    assert relerr(tnp(c), ref) < TOL


CONV_CASES = [
    dict(N=3, H=8, W=8, Cin=5, Cout=7, k=3, s=1, pad='SAME'),
    dict(N=2, H=8, W=8, Cin=6, Cout=4, k=3, s=2, pad='SAME'),      # TF SAME pads (0,1) here
    dict(N=2, H=8, W=8, Cin=4, Cout=9, k=3, s=1, pad='VALID'),
    dict(N=2, H=7, W=9, Cin=3, Cout=5, k=5, s=2, pad='SAME'),
    dict(N=4, H=6, W=6, Cin=8, Cout=8, k=1, s=1, pad='SAME'),
]


@pytest.mark.parametrize('c', CONV_CASES)
def test_conv2d_fwd_bwd(c):
    from tgan import core, ops
    rng = np.random.default_rng(1)
    x = rng.standard_normal((c['N'], c['H'], c['W'], c['Cin']))
    w = rng.standard_normal((c['k'], c['k'], c['Cin'], c['Cout'])) * 0.3
    xt, wt = T(x, True), T(w, True)
    yt = O.conv2d_tf(xt, wt, c['s'], c['pad'])
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    # second, independent restatement (loop-level numpy) pins the oracle's TF geometry
    assert relerr(yt.detach().numpy(), tfnp.conv2d(x, w, c['s'], c['pad'])) < 1e-12
    p = param(w)
    with core.recording():
        xv = var(x, True)
        out = ops.conv2d(xv, ops.PlainWeight(p), c['k'], c['k'], c['s'], c['pad'])
        assert tuple(out.shape) == tuple(yt.shape)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < TOL
    assert relerr(tnp(p.grad), wt.grad.numpy()) < TOL


@pytest.mark.parametrize('c', [dict(N=2, h=4, w=4, Cin=6, Cout=5, k=5, s=2), dict(N=3, h=3, w=5, Cin=4, Cout=3, k=5, s=2),
                               dict(N=2, h=4, w=4, Cin=3, Cout=4, k=3, s=2)])
def test_conv2d_transpose_fwd_bwd(c):
    from tgan import core, ops
    rng = np.random.default_rng(2)
    x = rng.standard_normal((c['N'], c['h'], c['w'], c['Cin']))
    w = rng.standard_normal((c['k'], c['k'], c['Cout'], c['Cin'])) * 0.3
    xt, wt = T(x, True), T(w, True)
    yt = O.conv2d_transpose_tf(xt, wt, c['s'])
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    assert relerr(yt.detach().numpy(), tfnp.conv2d_transpose(x, w, c['s'])) < 1e-12
    p = param(w)
    with core.recording():
        xv = var(x, True)
        out = ops.conv2d_transpose(xv, ops.PlainWeight(p), c['k'], c['k'], c['s'])
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < TOL
    assert relerr(tnp(p.grad), wt.grad.numpy()) < TOL


def test_weightnorm_conv_and_dense():
    from tgan import core, ops
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 6, 6, 5))
    V = rng.standard_normal((3, 3, 5, 8)) * 0.05
    g = rng.uniform(0.5, 1.5, 8)
    xt, Vt, gt = T(x, True), T(V, True), T(g, True)
    yt = O.conv2d_tf(xt, gt.view(1, 1, 1, -1) * O.l2_normalize(Vt, (0, 1, 2)), 1, 'SAME')
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    pV, pg = param(V), param(g)
    core.ctx.store.bump()
    with core.recording():
        xv = var(x, True)
        out = ops.conv2d(xv, ops.WNWeight(pV, pg, 45, 8, 1, 1), 3, 3, 1, 'SAME')
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(pV.grad), Vt.grad.numpy()) < 5e-5
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 5e-5
    # transposed-conv layout [kh,kw,Cout,Cin], norm over axes [0,1,3] (modle_base.py:148)
    Vd = rng.standard_normal((5, 5, 3, 6)) * 0.05
    gd = rng.uniform(0.5, 1.5, 3)
    xd = rng.standard_normal((2, 4, 4, 6))
    xt, Vt, gt = T(xd, True), T(Vd, True), T(gd, True)
    yt = gt.view(1, 1, 1, -1) * O.conv2d_transpose_tf(xt, O.l2_normalize(Vt, (0, 1, 3)), 2)
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    pV, pg = param(Vd), param(gd)
    core.ctx.store.bump()
    with core.recording():
        xv = var(xd, True)
        out = ops.conv2d_transpose(xv, ops.WNWeight(pV, pg, 25, 3, 6, 1), 5, 5, 2)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(pV.grad), Vt.grad.numpy()) < 5e-5
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 5e-5
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < TOL


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('shape', [(4, 6, 6, 8), (7, 10), (3, 5, 5, 3), (16, 16, 16, 256)])
def test_mobn_act(train, shape):
    from tgan import core, ops
    rng = np.random.default_rng(4)
    z = rng.standard_normal(shape) + 0.5
    b = rng.standard_normal(shape[-1])
    pm = rng.standard_normal(shape[-1])
    zt, bt = T(z, True), T(b, True)
    S = {'pm': T(pm)}
    yt = O.lrelu_cifar(O.mean_only_bn(zt, 'pm', bt, S, train, len(shape) == 4))
    gy = rng.standard_normal(shape)
    yt.backward(T(gy))
    pb, ppm = param(b), param(pm, False)
    with core.recording():
        zv = var(z, True)
        out = ops.mobn_act(zv, pb, ppm, train, 'lrelu', 0.2)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(zv.grad), zt.grad.numpy()) < TOL
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < TOL
    assert relerr(tnp(ppm.data), S['pm'].numpy()) < TOL


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('segs', [[7], [5, 3], [4, 9, 2], [50, 50, 50, 100]])
@pytest.mark.parametrize('act', ['none', 'lrelu'])
def test_mobn_small_segments(train, segs, act):
    """the classifier's logits layer on a grouped batch (tgan_mobn_small_fwd / _bwd: every segment in one launch) against
    the oracle applied call by call: outputs, dz, db and the pop_mean chain (one update per call, in call order)"""
    from tgan import core, ops
    rng = np.random.default_rng(14)
    n, C = sum(segs), 10
    z = rng.standard_normal((n, C)) + 0.5
    b, pm = rng.standard_normal(C), rng.standard_normal(C)
    gy = rng.standard_normal((n, C))
    zt, bt = T(z, True), T(b, True)
    S = {'pm': T(pm)}
    ys, r0 = [], 0
    for k in segs:
        y = O.mean_only_bn(zt[r0:r0 + k], 'pm', bt, S, train, False)
        ys.append(O.lrelu_cifar(y) if act == 'lrelu' else y)
        r0 += k
    yt = torch.cat(ys, 0)
    yt.backward(T(gy))
    pb, ppm = param(b), param(pm, False)
    with core.recording():
        zv = var(z, True)
        zv.aux = {'segs': list(segs)}
        out = ops.mobn_act(zv, pb, ppm, train, act, 0.2)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(zv.grad), zt.grad.numpy()) < TOL
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < TOL
    assert relerr(tnp(ppm.data), S['pm'].numpy()) < TOL


@pytest.mark.parametrize('shape', [(4, 6, 6, 8), (16, 12), (5, 4, 4, 3), (16, 8, 8, 128), (16, 32, 32, 128)])
def test_batch_norm_train(shape):
    from tgan import core, ops
    rng = np.random.default_rng(5)
    x = rng.standard_normal(shape) * 2 + 1
    C = shape[-1]
    gm, bt_ = rng.uniform(0.5, 1.5, C), rng.standard_normal(C)
    xt, gt, bt = T(x, True), T(gm, True), T(bt_, True)
    P = {'s/gamma': gt, 's/beta': bt}
    S = {'s/moving_mean': T(np.zeros(C)), 's/moving_variance': T(np.ones(C))}
    yt = O.bn_contrib(P, S, 's', xt, True)
    gy = rng.standard_normal(shape)
    yt.backward(T(gy))
    pg, pb, pmm, pmv = param(gm), param(bt_), param(np.zeros(C), False), param(np.ones(C), False)
    with core.recording():
        xv = var(x, True)
        out = ops.batch_norm(xv, pg, pb, pmm, pmv, True)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 5e-5
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 2e-4
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 5e-5
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 5e-5
    assert relerr(tnp(pmm.data), S['s/moving_mean'].numpy()) < 5e-5
    assert relerr(tnp(pmv.data), S['s/moving_variance'].numpy()) < 5e-5


@pytest.mark.parametrize('shape,segs', [((9, 8, 8, 16), [3, 2, 4]), ((20, 12), [5, 15]), ((9, 4, 4, 3), [4, 5]),
                                        ((12, 16, 16, 64), [2, 3, 3, 4])])
@pytest.mark.parametrize('seg_kernels', [True, False])
def test_batch_norm_segments(shape, segs, seg_kernels, monkeypatch):
    """training-mode contrib batch norm of a grouped batch (tgan_bn_fwd_seg / _bwd_seg: every call of the group in three
    launches per direction; seg_kernels=False: the per-call launches) against the oracle applied call by call: outputs,
    dx, dgamma, dbeta, and the moving statistics after one update per call in call order"""
    from tgan import core, ops
    if not seg_kernels:
        monkeypatch.setenv('TGAN_NO_BN_SEG', '1')
    rng = np.random.default_rng(15)
    x = rng.standard_normal(shape) * 2 + 1
    C = shape[-1]
    gm, bt_ = rng.uniform(0.5, 1.5, C), rng.standard_normal(C)
    xt, gt, bt = T(x, True), T(gm, True), T(bt_, True)
    P = {'s/gamma': gt, 's/beta': bt}
    S = {'s/moving_mean': T(np.zeros(C)), 's/moving_variance': T(np.ones(C))}
    ys, r0 = [], 0
    for k in segs:
        ys.append(O.bn_contrib(P, S, 's', xt[r0:r0 + k], True))
        r0 += k
    yt = torch.cat(ys, 0)
    gy = rng.standard_normal(shape)
    yt.backward(T(gy))
    pg, pb, pmm, pmv = param(gm), param(bt_), param(np.zeros(C), False), param(np.ones(C), False)
    with core.recording():
        xv = var(x, True)
        xv.aux = {'segs': list(segs)}
        out = ops.batch_norm(xv, pg, pb, pmm, pmv, True)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 5e-5
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 2e-4
    assert relerr(tnp(pg.grad), gt.grad.numpy()) < 5e-5
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 5e-5
    assert relerr(tnp(pmm.data), S['s/moving_mean'].numpy()) < 5e-5
    assert relerr(tnp(pmv.data), S['s/moving_variance'].numpy()) < 5e-5


@pytest.mark.parametrize('shape', [(5, 4, 4, 12), (16, 8, 8, 128)])
@pytest.mark.parametrize('act', ['none', 'relu', 'lrelu', 'tanh', 'sigmoid', 'softplus'])
def test_bias_act(act, shape):
    from tgan import core, ops
    import torch.nn.functional as F
    rng = np.random.default_rng(6)
    z, b = rng.standard_normal(shape) * 2, rng.standard_normal(shape[-1])
    f = dict(none=lambda t: t, relu=F.relu, lrelu=O.lrelu_cifar, tanh=torch.tanh, sigmoid=torch.sigmoid,
             softplus=F.softplus)[act]
    zt, bt = T(z, True), T(b, True)
    yt = f(zt + bt)
    gy = rng.standard_normal(z.shape)
    yt.backward(T(gy))
    pb = param(b)
    with core.recording():
        zv = var(z, True)
        out = ops.activation(ops.lazy_bias(zv, pb), act)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < TOL
    assert relerr(tnp(zv.grad), zt.grad.numpy()) < 5e-5
    assert relerr(tnp(pb.grad), bt.grad.numpy()) < 5e-5


def test_pools_concat_dropout_noise():
    from tgan import core, ops
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 8, 8, 6))
    xt = T(x, True)
    yt = O.max_pool_tf(xt, 2, 2)
    assert relerr(yt.detach().numpy(), tfnp.max_pool(x, 2, 2, 'SAME')) == 0
    gy = rng.standard_normal(tuple(yt.shape))
    yt.backward(T(gy))
    with core.recording():
        xv = var(x, True)
        out = ops.max_pool2(xv)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    assert relerr(fwd, yt.detach().numpy()) < 1e-6
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < 1e-6
    for mode in ('max', 'mean'):
        x6 = rng.standard_normal((4, 6, 6, 10))
        xt = T(x6, True)
        yt = O.max_pool_tf(xt, 6, 1).reshape(4, 10) if mode == 'max' else xt.mean(dim=(1, 2))
        gy = rng.standard_normal((4, 10))
        yt.backward(T(gy))
        with core.recording():
            xv = var(x6, True)
            out = ops.global_pool(xv, mode)
            fwd = tnp(out.data)
            run_bwd(out, gy)
        assert relerr(fwd, yt.detach().numpy()) < 1e-6
        assert relerr(tnp(xv.grad), xt.grad.numpy()) < 1e-6
    # label concat (modle_base.py:239-244) fwd + slice backward
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 3)]
    ref = O.cond_concat(T(x), T(y).view(3, 1, 1, 10)).numpy()
    with core.recording():
        xv = var(x, True)
        out = ops.concat_label(xv, var(y))
        assert relerr(out.numpy(), ref) < 1e-6
        gy = rng.standard_normal(ref.shape)
        run_bwd(out, gy)
    assert relerr(tnp(xv.grad), gy[..., :6].astype(np.float32)) < 1e-6
    # injected dropout / noise reproduce the oracle's draws exactly
    r = O.TagRNG(11)
    ctx = setup('fp32', injected=r)
    xv = var(x)
    d = ops.dropout(xv, 0.2, 't/drop', True)
    ref = O.dropout_tf(T(x), r.keep_mask('t/drop', x.shape, 0.2), 0.2).numpy()
    assert relerr(tnp(d.data), ref) < 1e-6
    nz = ops.add_noise(xv, 0.15, 't/noise')
    ref = (T(x) + 0.15 * r.normal('t/noise', x.shape).double()).numpy()
    assert relerr(tnp(nz.data), ref) < 1e-6


def test_philox_statistics():
    from tgan import ops
    setup('fp32')
    x = var(np.ones((64, 32, 32, 16), np.float32))
    d = tnp(ops.dropout(x, 0.2, 'a', True).data)
    keep = (d != 0).mean()
    assert abs(keep - 0.8) < 2e-3 and np.allclose(d[d != 0], 1.25)
    d2 = tnp(ops.dropout(x, 0.2, 'b', True).data)
    assert (d != d2).mean() > 0.2          # different tags -> different streams
    n = tnp(ops.add_noise(var(np.zeros((64, 32, 32, 16), np.float32)), 0.5, 'n').data)
    assert abs(n.mean()) < 2e-3 and abs(n.std() - 0.5) < 2e-3
    assert abs(((n / 0.5) ** 4).mean() - 3.0) < 0.05       # Gaussian kurtosis


def test_argmax_onehot_bit_exact():
    from tgan import ops
    rng = np.random.default_rng(8)
    lg = rng.standard_normal((257, 10)).astype(np.float32)
    lg[3, 2] = lg[3, 7] = lg[3].max() + 1        # tie -> lowest index
    lg[4, :] = 0.25                               # all equal -> 0
    lg[5, 9] = np.inf
    lg[6, 0] = -np.inf
    # NaN rows: Eigen's ArgMaxTupleReducer starts at (0, -FLT_MAX) and takes strictly greater
    # elements only, so a NaN never wins and never blocks a later number; all-NaN / all -inf rows give index 0
    lg[7, 0] = np.nan
    lg[8, 5] = np.nan
    lg[9, :] = np.nan
    lg[10, :] = -np.inf
    lg[11, :4] = np.nan
    lg[11, 4:] = -np.inf
    lg[12, :] = np.nan
    lg[12, 7] = -1e30
    idx_ref, oh_ref = tfnp.argmax_onehot(lg, 10)
    assert idx_ref[7] != 0 and idx_ref[9] == 0 and idx_ref[10] == 0 and idx_ref[11] == 0 and idx_ref[12] == 7
    idx, oh = ops.argmax_onehot(var(lg), 10)
    assert idx.data.dtype == torch.int64
    assert np.array_equal(idx.data.cpu().numpy(), idx_ref)
    assert np.array_equal(oh.data.cpu().numpy(), oh_ref)
    assert np.array_equal(idx_ref, O.argmax_onehot(torch.from_numpy(lg))[0].numpy())


@pytest.mark.parametrize('rep', [True, False])
def test_losses(rep):
    from tgan import ops
    rng = np.random.default_rng(9)
    dr, df, du = (rng.standard_normal((n, 1)) * 2 for n in (10, 12, 6))
    t = [T(a, True) for a in (dr, df, du)]
    ld = O.d_loss_fn(*t)
    ld.backward()
    vs = [var(a, True) for a in (dr, df, du)]
    L = ops.loss_d(*vs)
    L.seed()
    assert abs(L.item() - float(ld)) < 1e-5 * max(1, abs(float(ld)))
    for v, tt in zip(vs, t):
        assert relerr(tnp(v.grad), tt.grad.numpy()) < 1e-5
    tf_ = T(df, True)
    lg = O.g_loss_fn(tf_)
    lg.backward()
    v = var(df, True)
    L = ops.loss_g(v)
    L.seed()
    assert abs(L.item() - float(lg)) < 1e-5 and relerr(tnp(v.grad), tf_.grad.numpy()) < 1e-5
    K = 10
    c_real, c_unl, c_fake, c_rep = (rng.standard_normal((n, K)) * 3 for n in (5, 6, 12, 6))
    y_l = np.eye(K)[rng.integers(0, K, 5)]
    y_g = np.eye(K)[rng.integers(0, K, 12)]
    tt = [T(a, True) for a in (c_real, c_unl, c_fake, c_rep)]
    lc = O.c_loss_fn(tt[0], tt[1], tt[2], T(du), T(y_l), T(y_g), 0.3, tt[3] if rep else None, 0.5)
    lc.backward()
    vv = [var(a, True) for a in (c_real, c_unl, c_fake, c_rep)]
    lam = torch.tensor([0.3, 0.5], device='cuda')
    L = ops.loss_c(vv[0], var(y_l), vv[1], vv[3] if rep else None, var(du), vv[2], var(y_g), lam)
    L.seed()
    assert abs(L.item() - float(lc)) < 2e-5 * max(1, abs(float(lc)))
    for i in range(4 if rep else 3):
        assert relerr(tnp(vv[i].grad), tt[i].grad.numpy()) < 2e-5, i


def test_adam_ema_multi_step():
    from tgan import _lib
    from tgan.train_base import AdamOptimizer
    rng = np.random.default_rng(10)
    n = 1003
    th = rng.standard_normal(n)
    P = {'w': torch.tensor(th.copy())}
    opt = O.TFAdam(['w'], P, 0.5)
    ema_ref = P['w'].clone()
    fb = dict(theta=torch.zeros(1004, device='cuda'), n=1004)
    fb['theta'][:n] = torch.from_numpy(th.astype(np.float32)).cuda()
    fb['m'], fb['v'], fb['grad'] = (torch.zeros(1004, device='cuda') for _ in range(3))
    ema = fb['theta'].clone()
    mine = AdamOptimizer(3e-3, 0.5)
    for step in range(5):
        g = rng.standard_normal(n)
        opt.apply(P, {'w': torch.tensor(g)}, 3e-3)
        ema_ref -= (ema_ref - P['w']) * (1 - 0.9999)
        fb['grad'][:n] = torch.from_numpy((g * 4).astype(np.float32)).cuda()
        mine.apply_flat(fb, 0.25, ema, 0.9999)       # grad_scale = 1/world folds the DP average
    assert relerr(tnp(fb['theta'][:n]), P['w'].numpy()) < 1e-5
    assert relerr(tnp(ema[:n]), ema_ref.numpy()) < 1e-6


@pytest.mark.parametrize('rows,C,dt', [(16384, 128, 'f32'), (16384, 128, 'bf16'), (4096, 512, 'f32'), (1024, 256, 'bf16'),
                                       (102400, 128, 'bf16'), (1600, 8192, 'f32'), (1000, 10, 'f32'), (77, 3, 'f32')])
def test_channel_stats_large(rows, C, dt):
    """single-launch column reduction (last-CTA fold) at real activation sizes, incl. accumulate-in-place"""
    from tgan import _lib
    rng = np.random.default_rng(12)
    x = (rng.standard_normal((rows, C)) + 0.3).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    if dt == 'bf16':
        xt = xt.to(torch.bfloat16)
        x = xt.float().cpu().numpy()
    s = torch.full((C,), 5.0, device='cuda')
    ss = torch.zeros(C, device='cuda')
    ws = torch.empty(4 * 256 * C + 2 * C, device="cuda")
    for beta in (0.0, 1.0):
        _lib.call('tgan_channel_stats', xt.data_ptr(), 0 if dt == 'f32' else 1, rows, C, s.data_ptr(), ss.data_ptr(), beta,
                  ws.data_ptr(), 0)
    ref = x.astype(np.float64).sum(0)
    assert relerr(tnp(s), 2 * ref) < 1e-5
    assert relerr(tnp(ss), 2 * (x.astype(np.float64) ** 2).sum(0)) < 1e-5
