"""-m gpu: the parts of the Model/nn.py operator surface that no Triple-GAN builder calls but the north_star names
(SURVEY.md 8a rows a5 / a6): nn.batch_norm_impl (nn.py:192-217) and the Salimans & Kingma weight-normalised layers
nn.dense / nn.conv2d / nn.deconv2d / nn.nin (nn.py:220-340), each in its init=False form (forward, input gradient,
V / g / b gradients) and its init=True data-dependent form (the data-normalised output the reference returns),
through tgan.nn with the reference's own call signatures, against the float64 oracle.

Tolerances (relative to max-abs): fp32 mode 5e-5 forward / 2e-4 gradients; bf16 mode against the oracle with the
same bf16 rounding points: 1e-2 forward, 2e-2 gradients.
"""
import contextlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402
from util_gpu import relerr, run_bwd, setup, tnp, var   # noqa: E402

TOL = {'fp32': (5e-5, 2e-4), 'bf16': (1e-2, 2e-2)}


def T(a, rg=False):
    return torch.tensor(np.asarray(a, np.float64), requires_grad=rg)


def _store_set(name, a, trainable=True):
    """create the variable `name` (TF-style full name) in the current store with value a"""
    from tgan import core
    a = np.asarray(a, np.float32)
    parts = name.split('/')
    with contextlib.ExitStack() as st:
        for s in parts[:-1]:
            st.enter_context(core.variable_scope(s))
        p = core.get_variable(parts[-1], list(a.shape), core.constant_initializer(0.0), trainable=trainable)
    p.data = torch.from_numpy(a.copy()).cuda()
    p.grad = torch.zeros_like(p.data)
    p.requires_grad = trainable
    return p


def _q(math):
    return O.quantized() if math == 'bf16' else contextlib.nullcontext()


def _xv(x, math, rg=True):
    v = var(x, rg, dtype=torch.bfloat16 if (math == 'bf16' and x.shape[-1] >= 16) else None)
    return v


def _bfr(a, math):
    """inputs are rounded to bf16 first in bf16 mode so both arms see identical operands"""
    if math == 'bf16':
        return torch.tensor(np.asarray(a, np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)
    return np.asarray(a, np.float64)


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('shape', [(6, 5, 5, 16), (12, 24)])
def test_batch_norm_impl(shape, train, math):
    from tgan import core, nn
    setup(math)
    rng = np.random.default_rng(11)
    C = shape[-1]
    x = _bfr(rng.standard_normal(shape) * 1.5 + 0.3, math)
    sc, be = rng.uniform(0.5, 1.5, C), rng.standard_normal(C)
    pm, pv = rng.standard_normal(C) * 0.2, rng.uniform(0.5, 2.0, C)
    n = 'L/BatchNormalization'
    P = {n + '/scale': T(sc, True), n + '/beta': T(be, True)}
    S = {n + '/pop_mean': T(pm), n + '/pop_var': T(pv)}
    xt = T(x, True)
    with _q(math):
        yt = O.batch_norm_impl(P, S, 'L', xt, train, conv=len(shape) == 4)
    gy = _bfr(rng.standard_normal(shape), math)
    yt.backward(T(gy))
    ps, pb = _store_set(n + '/scale', sc), _store_set(n + '/beta', be)
    ppm, ppv = _store_set(n + '/pop_mean', pm, False), _store_set(n + '/pop_var', pv, False)
    with core.recording(), core.variable_scope('L', reuse=True):
        xv = _xv(x, math)
        out = nn.batch_norm_impl(xv, is_conv_out=len(shape) == 4, deterministic=not train)
        fwd = tnp(out.data)
        run_bwd(out, gy)
    tf_, tg = TOL[math]
    assert relerr(fwd, yt.detach().numpy()) < tf_
    assert relerr(tnp(xv.grad), xt.grad.numpy()) < tg
    if train:      # (deterministic=True is the test-time graph: nothing differentiates scale / beta through it)
        assert relerr(tnp(ps.grad), P[n + '/scale'].grad.numpy()) < tg
        assert relerr(tnp(pb.grad), P[n + '/beta'].grad.numpy()) < tg
    # population statistics: decay-0.9 EMA of the batch mean and of the BIASED batch variance (nn.py:207-214)
    assert relerr(tnp(ppm.data), S[n + '/pop_mean'].numpy()) < 1e-4
    assert relerr(tnp(ppv.data), S[n + '/pop_var'].numpy()) < 1e-4


def _wn_case(kind, math, init, rng):
    """returns (oracle fn, tgan fn, x, V shape, Cout)"""
    from tgan import nn
    if kind == 'dense':
        x = rng.standard_normal((24, 40))
        vs, co = (40, 32), 32
        o = lambda P, xt: O.salimans_dense(P, 'dense_0', xt, nonlin=O.leaky_relu_tf, init=init, init_scale=0.8)
        g = lambda xv: nn.dense(xv, 32, nonlinearity=nn.leaky_relu, init_scale=0.8, counters={}, init=init)
        scope = 'dense_0'
    elif kind == 'conv2d':
        x = rng.standard_normal((4, 8, 8, 24))
        vs, co = (3, 3, 24, 32), 32
        o = lambda P, xt: O.salimans_conv2d(P, 'conv2d_0', xt, 2, 'SAME', nonlin=torch.relu, init=init, init_scale=1.3)
        g = lambda xv: nn.conv2d(xv, 32, filter_size=[3, 3], stride=[2, 2], pad='SAME', nonlinearity=nn.relu,
                                 init_scale=1.3, counters={}, init=init)
        scope = 'conv2d_0'
    elif kind == 'deconv2d':
        x = rng.standard_normal((3, 4, 4, 32))
        vs, co = (5, 5, 16, 32), 16
        o = lambda P, xt: O.salimans_deconv2d(P, 'deconv2d_0', xt, 2, nonlin=torch.tanh, init=init)
        g = lambda xv: nn.deconv2d(xv, 16, filter_size=[5, 5], stride=[2, 2], pad='SAME', nonlinearity=nn.tanh,
                                   counters={}, init=init)
        scope = 'deconv2d_0'
    else:
        x = rng.standard_normal((3, 6, 6, 48))
        vs, co = (48, 16), 16
        o = lambda P, xt: O.salimans_nin(P, 'dense_0', xt, nonlin=None, init=init)
        g = lambda xv: nn.nin(xv, 16, nonlinearity=None, counters={}, init=init)
        scope = 'dense_0'
    return o, g, x, vs, co, scope


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
@pytest.mark.parametrize('init', [False, True])
@pytest.mark.parametrize('kind', ['dense', 'conv2d', 'deconv2d', 'nin'])
def test_salimans_layers(kind, init, math):
    from tgan import core
    setup(math)
    rng = np.random.default_rng(12)
    o, g, x, vs, co, scope = _wn_case(kind, math, init, rng)
    x = _bfr(x, math)
    V = rng.standard_normal(vs) * 0.05
    gg, b = rng.uniform(0.5, 1.5, co), rng.standard_normal(co) * 0.1
    P = {scope + '/V': T(V, True), scope + '/g': T(gg, True), scope + '/b': T(b, True)}
    xt = T(x, True)
    with _q(math):
        yt = o(P, xt)
    gy = _bfr(rng.standard_normal(tuple(yt.shape)), math)
    if not init:
        yt.backward(T(gy))
    pV, pg, pb = (_store_set(scope + '/' + k, a) for k, a in (('V', V), ('g', gg), ('b', b)))
    core.ctx.store.bump()
    core.ctx.store.reuse[0] = True            # tf.variable_scope(..., reuse=True) at the root: share V / g / b
    tf_, tg = TOL[math]
    if init:
        # the data-dependent branch (nn.py:224-238, 259-274, 301-316) normalises with V.initialized_value(): a forward
        # quantity only -- no gradient reaches V, g or b through it in the reference either
        with core.no_grad():
            out = g(_xv(x, math, rg=False))
            fwd = tnp(out.data).reshape(tuple(yt.shape))
        assert relerr(fwd, yt.detach().numpy()) < tf_, (kind, init)
        # per-channel zero mean / init_scale standard deviation before the nonlinearity is what the branch is for
        if kind == 'nin':
            flat = fwd.reshape(-1, fwd.shape[-1])
            assert np.abs(flat.mean(0)).max() < 2e-2 and np.abs(flat.std(0) - 1.0).max() < 2e-2
        return
    with core.recording():
        xv = _xv(x, math)
        out = g(xv)
        fwd = tnp(out.data).reshape(tuple(yt.shape))
        run_bwd(out, gy.reshape(out.shape))
    assert relerr(fwd, yt.detach().numpy()) < tf_, (kind, init)
    assert relerr(tnp(xv.grad).reshape(x.shape), xt.grad.numpy()) < tg
    assert relerr(tnp(pV.grad), P[scope + '/V'].grad.numpy()) < tg
    assert relerr(tnp(pg.grad), P[scope + '/g'].grad.numpy()) < tg
    assert relerr(tnp(pb.grad), P[scope + '/b'].grad.numpy()) < tg
