"""-m gpu: whole-network forward + backward parity (generator, discriminator, classifier of all three
model families) with FIXED inputs and labels, so the comparison is not confounded by pseudo-label flips:
per-network logits and every parameter gradient against the float64 oracle.

Stated tolerances, relative to max-abs (gradients: relative to max(own max, 1e-2 (fp32) / 1e-1 (bf16) * the
network's largest gradient tensor)):
  fp32 CUDA-core mode : logits 1e-4; gradients max(1e-3, 5x the float32-vs-float64 floor of the oracle itself)
  bf16 tcgen05 mode   : logits 2e-2, gradients 1.5e-1 -- or, for the deep classifier at batch 16, 2x (rms) /
                        3x (max) the error the oracle itself shows when the same bf16 rounding points are
                        inserted into it (oracle.quantized), i.e. the precision's own noise floor.
"""
import contextlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402
from util_gpu import tnp                            # noqa: E402

TOL = {'fp32': (1e-4, 1e-3), 'bf16': (2e-2, 1.5e-1)}
B = 16


def _setup(data_name, math):
    import tgan
    from tgan import core
    P, S = O.init_params(data_name, seed=5)
    zca = O.make_zca(3) if data_name == 'cifar10' else None
    orc = O.OracleTrainer(data_name, P, S, zca, dtype=torch.float64, scale=10)
    o32 = O.OracleTrainer(data_name, P, S, zca, dtype=torch.float32, scale=10)
    tgan.init('cuda:0', math=math)
    tr = tgan.make_trainer(data_name, scale=10, init=(P, S), zca=zca)
    rng = O.TagRNG(7)
    core.ctx.rng = core.InjectedSource(rng)
    core.ctx.store = tr.store
    return orc, o32, tr, rng


def _ctx(math):
    return O.quantized() if math == 'bf16' else contextlib.nullcontext()


def _oracle(o, names, fn, math):
    """fn(model, to_tensor) -> (logits, loss); returns (logits ndarray, {name: grad ndarray})"""
    t = lambda a: torch.tensor(np.asarray(a), dtype=o.dtype)
    with _ctx(math):
        logits, loss = fn(o.model, t)
        gs = torch.autograd.grad(loss, [o.P[n] for n in names], allow_unused=True)
    return logits.detach().double().numpy(), \
        {n: (g if g is not None else torch.zeros_like(o.P[n])).double().numpy() for n, g in zip(names, gs)}


def _compare(orc, o32, tr, group, names, fn, logits_v, math, what):
    """CUDA result vs the float64 oracle; the bound is the stated tolerance or a multiple of the error the
    ORACLE ITSELF shows at the same precision (float32 run for fp32 mode, bf16 rounding points for bf16 mode).
    bf16: a 10-layer net decorrelates after ~3 layers (each 1-ulp rounding flip perturbs the next layer's
    roundings and max-pool routes), so two valid bf16 executions differ by the precision's noise floor --
    measured here, not assumed."""
    tl, tg = TOL[math]
    lt, ref = _oracle(orc, names, fn, 'fp32')
    scale = max(np.abs(r).max() for r in ref.values())
    guard = (1e-2 if math == 'fp32' else 1e-1) * scale      # parameters whose exact gradient is ~0
    den = {n: max(np.abs(r).max(), guard) for n, r in ref.items()}
    rden = {n: max(np.sqrt((r ** 2).mean()), guard) for n, r in ref.items()}
    if math == 'fp32':
        l2, r2 = _oracle(o32, names, fn, 'fp32')
    else:
        l2, r2 = _oracle(orc, names, fn, 'bf16')
    lden = max(np.abs(lt).max(), 1e-9)
    lfloor = np.abs(l2 - lt).max() / lden
    floor = max(np.abs(r2[n] - ref[n]).max() / den[n] for n in names)
    rfloor = max(np.sqrt(((r2[n] - ref[n]) ** 2).mean()) / rden[n] for n in names)
    el = np.abs(tnp(logits_v.data).reshape(lt.shape) - lt).max() / lden
    fb = tr.store.flat[group]
    worst, rworst, wn = 0.0, 0.0, None
    for p, o in zip(fb['params'], fb['offsets']):
        g = tnp(fb['grad'][o:o + p.size]).reshape(p.shape)
        e = np.abs(g - ref[p.name]).max() / den[p.name]
        rworst = max(rworst, np.sqrt(((g - ref[p.name]) ** 2).mean()) / rden[p.name])
        if e > worst:
            worst, wn = e, p.name
    print('%s %s: logits err %.2e (oracle floor %.2e) | grad max-err %.2e (floor %.2e) rms-err %.2e (floor %.2e) worst %s'
          % (what, math, el, lfloor, worst, floor, rworst, rfloor, wn))
    assert el < max(tl, 3 * lfloor), (what, el, lfloor)
    if math == 'fp32':
        # parameters whose exact gradient is identically zero (a bias in front of a batch-mean subtraction) carry
        # pure cancellation noise whose size depends on the summation order: bound it at 2e-3 of the net's scale
        zero = {n for n, r in ref.items() if np.abs(r).max() < 1e-6 * scale}
        worst_nz = 0.0
        for p, o in zip(fb['params'], fb['offsets']):
            g = tnp(fb['grad'][o:o + p.size]).reshape(p.shape)
            if p.name in zero:
                assert np.abs(g).max() < 2e-3 * scale, (what, p.name, np.abs(g).max(), scale)
            else:
                worst_nz = max(worst_nz, np.abs(g - ref[p.name]).max() / den[p.name])
        assert worst_nz < max(tg, 5 * floor), (what, wn, worst_nz, floor)
    else:
        assert rworst < max(tg, 2 * rfloor), (what, rworst, rfloor)
        assert worst < max(tg, 3 * floor), (what, wn, worst, floor)


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
@pytest.mark.parametrize('data_name', ['cifar10', 'svhn', 'mnist'])
def test_classifier_fwd_bwd(data_name, math):
    from tgan import core, ops
    orc, o32, tr, rng = _setup(data_name, math)
    nrng = np.random.default_rng(1)
    lo = 0.0 if data_name == 'mnist' else -1.0
    x = nrng.uniform(lo, 1, [B] + orc.cfg.IMAGE_DIM).astype(np.float32)
    R = nrng.standard_normal((B, 10))

    def fn(m, t):
        pre = m.zca_apply if data_name == 'cifar10' else (lambda a: a)
        lt, _ = m.classifier(pre(t(x)), True, rng, 'T/C')
        return lt, (lt * t(R)).sum()
    tr._begin('classifier', tr.c_vars)
    with core.recording():
        lv, _ = tr.model.classifier(tr._pre()(ops.constant(x)), True, reuse=True, tag='T/C')
        lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
        core.ctx.tape.backward()
    _compare(orc, o32, tr, 'classifier', orc.c_vars, fn, lv, math, data_name + ' C')


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
@pytest.mark.parametrize('data_name', ['cifar10', 'svhn', 'mnist'])
def test_generator_discriminator_fwd_bwd(data_name, math):
    """D(x, y) -> gradients of D's variables; D(G(z, y), y) -> gradients of G's variables (exercises every
    dgrad path of D, the transposed-conv backward and BN backward of G)."""
    from tgan import core, ops
    orc, o32, tr, rng = _setup(data_name, math)
    nrng = np.random.default_rng(2)
    lo = 0.0 if data_name == 'mnist' else -1.0
    x = nrng.uniform(lo, 1, [B] + orc.cfg.IMAGE_DIM).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[nrng.integers(0, 10, B)]
    z = nrng.uniform(-1, 1, (B, 100)).astype(np.float32)
    R = nrng.standard_normal((B, 1))

    def fn_d(m, t):
        _, lt = m.discriminator(t(x), t(y), rng, 'T/D')
        return lt, (lt * t(R)).sum()
    tr._begin('discriminator', tr.d_vars)
    with core.recording():
        _, lv = tr.model.discriminator(ops.constant(x), ops.constant(y), reuse=True, tag='T/D')
        lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
        core.ctx.tape.backward()
    _compare(orc, o32, tr, 'discriminator', orc.d_vars, fn_d, lv, math, data_name + ' D')

    def fn_g(m, t):
        _, lt = m.discriminator(m.good_generator(t(z), t(y), rng, 'T/G'), t(y), rng, 'T/DG')
        return lt, (lt * t(R)).sum()
    tr._begin('good_generator', tr.g_vars)
    with core.recording():
        Gv = tr.model.good_generator(ops.constant(z), ops.constant(y), reuse=True, tag='T/G')
        _, lv = tr.model.discriminator(Gv, ops.constant(y), reuse=True, tag='T/DG')
        lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
        core.ctx.tape.backward()
    t64 = lambda a: torch.tensor(a, dtype=torch.float64)
    Gt = orc.model.good_generator(t64(z), t64(y), rng, 'T/G').detach().numpy()
    with _ctx(math):
        Gq = orc.model.good_generator(t64(z), t64(y), rng, 'T/G').detach().numpy()
    eg, gfloor = np.abs(Gv.numpy().reshape(Gt.shape) - Gt).max(), np.abs(Gq - Gt).max()
    print('%s %s G image abs err %.2e (oracle floor at this precision %.2e)' % (data_name, math, eg, gfloor))
    assert eg < (1e-4 if math == 'fp32' else max(3e-2, 2 * gfloor))
    _compare(orc, o32, tr, 'good_generator', orc.g_vars, fn_g, lv, math, data_name + ' G<-D')
