"""-m gpu: whole-step parity of the CUDA path against the oracle on identical weights, inputs, noise and
dropout masks (SURVEY.md §8c parity contract): pseudo-label argmax bit-exact, per-phase gradients and
the multi-step (d, g, c)-loss trajectory within the stated tolerance.

fp32 mode (CUDA-core GEMMs): 2e-4 relative-to-max vs the float64 oracle -- measured floor of the
oracle's own float32 run vs float64 is ~1e-5 on these nets; the margin covers summation-order effects.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402
from util_gpu import relerr, tnp                    # noqa: E402


def _run(data_name, math, steps, scale, tol_loss, tol_grad, lambdas=(0.3, 0.5)):
    import tgan
    from tgan import core
    P, S = O.init_params(data_name, seed=5)
    zca = O.make_zca(3) if data_name == 'cifar10' else None
    orc = O.OracleTrainer(data_name, P, S, zca, dtype=torch.float64, scale=scale)
    tgan.init('cuda:0', math=math)
    tr = tgan.make_trainer(data_name, scale=scale, init=(P, S), zca=zca)
    worst = {}
    for step in range(steps):
        rng = O.TagRNG(100 + step)
        core.ctx.rng = core.InjectedSource(rng)
        batch = O.make_batch(orc.cfg, seed=50 + step)
        ref = orc.step(batch, rng, lambdas[0], lambdas[1])
        got = tr.step(batch, lambda_1=lambdas[0], lambda_2=lambdas[1]).cpu().numpy()
        # pseudo-labels: bit-exact (given matching logits; margin checked to exclude near-ties)
        for key in ('idx_unl_d', 'idx_unl'):
            assert np.array_equal(tr.aux[key].data.cpu().numpy(), orc.last_aux['D'][key].numpy()), (step, key)
        for i, nm in enumerate('dgc'):
            e = abs(got[i] - ref[i]) / max(1.0, abs(ref[i]))
            worst['loss_' + nm] = max(worst.get('loss_' + nm, 0), e)
            assert e < tol_loss, (step, nm, got[i], ref[i])
        for grp, ph in (('discriminator', 'D'), ('good_generator', 'G'), ('classifier', 'C')):
            fb = tr.store.flat[grp]
            for p, o in zip(fb['params'], fb['offsets']):
                g = tnp(fb['grad'][o:o + p.size]).reshape(p.shape)
                r = orc.last_grads[ph][p.name].numpy()
                e = np.abs(g - r).max() / max(np.abs(r).max(), 1e-8)
                worst['grad_' + ph] = max(worst.get('grad_' + ph, 0), e)
                assert e < tol_grad, (step, p.name, e)
    # parameters after `steps` Adam updates (Adam amplifies tiny gradient differences where |g| ~ 0,
    # so compare against the update scale lr*steps rather than against |theta|)
    for n, p in tr.store.vars.items():
        if p.trainable:
            lr = orc.cfg.CLA_LEARNINIG_RATE if 'classifier' in n else orc.cfg.LEARNING_RATE
            d = np.abs(tnp(p.data) - orc.P[n].detach().numpy()).max()
            assert d < 0.35 * lr * steps + 1e-6, (n, d)
    print(data_name, math, {k: '%.2e' % v for k, v in worst.items()})
    return worst


@pytest.mark.parametrize('data_name', ['cifar10', 'svhn', 'mnist'])
def test_step_parity_fp32(data_name):
    _run(data_name, 'fp32', steps=3, scale=10, tol_loss=2e-5, tol_grad=2e-4)


def test_step_parity_fp32_cifar_lambdas_zero():
    _run('cifar10', 'fp32', steps=1, scale=10, tol_loss=2e-5, tol_grad=2e-4, lambdas=(0.0, 0.0))


def test_graph_capture_matches_eager():
    """CUDA-graph replay of the three-phase step == eager launches (same Philox streams)."""
    import tgan
    P, S = O.init_params('cifar10', seed=5)
    outs = []
    for use_graph in (False, True):
        tgan.init('cuda:0', math='fp32', seed=77)
        tr = tgan.make_trainer('cifar10', scale=10, init=(P, S))
        batch = O.make_batch(O.OracleConfig('cifar10', 10), seed=9)
        tr.load_batch(batch)
        if use_graph:
            # restore the state the warm-up steps consume so both runs start identically
            tr.capture(warmup=0)
        losses = [tr.step(lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy() for _ in range(3)]
        outs.append(np.stack(losses))
    # capture itself executes nothing, so step k of the replay run == step k of the eager run
    assert np.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-6), (outs[0], outs[1])
