"""CPU: host-side logic of the product package -- graph construction in shape-only mode, TF variable naming and
scope semantics, configs, schedules, and loud failure without a GPU."""
import numpy as np
import pytest

import tgan
from oracle import tgan_oracle as O
from tgan import core, nn, ops


@pytest.mark.parametrize('name,counts', [('cifar10', (327467, 5129201, 3121812)), ('svhn', (339436, 5113844, 3124254)),
                                         ('mnist', (1561262, 714408, 279326))])
def test_variable_inventory_matches_tf_names(name, counts):
    tr = tgan.make_trainer(name, build_only=True)
    P, S = O.init_params(name)
    mine = {n: p.shape for n, p in tr.store.vars.items()}
    ref = {k: tuple(v.shape) for k, v in list(P.items()) + list(S.items())}
    assert mine == ref
    assert set(n for n, p in tr.store.vars.items() if not p.trainable) == set(S)
    got = tuple(sum(p.size for p in v) for v in (tr.d_vars, tr.g_vars, tr.c_vars))
    assert got == counts        # D 327,467 / G 5,129,201 / C 3,121,812 for the CIFAR file
    if name == 'cifar10':
        assert 'classifier/NiN1/NiN1/V' in mine and 'good_generator/gg_h0_lin/gg_h0_lin/kernel' in mine
        assert mine['good_generator/gg_dconv0/gg_dconv0/kernel'] == (5, 5, 256, 522)
        assert mine['discriminator/conv2d_21/conv2d_21/kernel'] == (3, 3, 138, 128)


def test_build_mode_shapes_of_forward_pass():
    tr = tgan.Train(tgan.make_config('cifar10', ZCA=(np.zeros(3072), np.eye(3072))))
    ph, G, D, C, model = tr._build_train_graph(tgan.Good_GAN_cifar10)
    assert G.shape == (100, 32, 32, 3)
    assert [d.shape for d in D[1::2]] == [(100, 1), (100, 1), (50, 1)]
    assert [c.shape for c in C] == [(50, 10), (50, 10), (80, 10), (100, 10), (50, 10)]
    tr2 = tgan.Train(tgan.make_config('mnist'))
    _, G, D, C, _ = tr2._build_train_graph(tgan.Good_GAN)
    assert G.shape == (100, 784) and len(C) == 4 and C[0].shape == (100, 10)


def test_variable_scope_reuse_semantics():
    store = core.VariableStore()
    core.ctx.store = store
    with core.building(), core.no_grad():
        x = ops.Var(None, (4, 8, 8, 3))
        with core.variable_scope('net'):
            nn.conv2d_WN(x, 16, name='c1', use_weight_normalization=True, use_mean_only_batch_normalization=True)
        assert set(store.vars) == {'net/c1/V', 'net/c1/b', 'net/c1/meanOnlyBatchNormalization/pop_mean', 'net/c1/g'}
        with pytest.raises(ValueError):            # TF: "Variable net/c1/V already exists"
            with core.variable_scope('net'):
                nn.conv2d_WN(x, 16, name='c1', use_weight_normalization=True)
        with core.variable_scope('net', reuse=True):
            y = nn.conv2d_WN(x, 16, name='c1', use_weight_normalization=True, use_mean_only_batch_normalization=True)
            assert y.shape == (4, 8, 8, 16)
            with pytest.raises(ValueError):        # TF: "Variable net/c2/V does not exist"
                nn.conv2d_WN(x, 16, name='c2', use_weight_normalization=True)
            with pytest.raises(ValueError):        # shape mismatch on reuse
                nn.conv2d_WN(x, 32, name='c1', use_weight_normalization=True)


def test_nn_surface_shapes_and_scopes():
    store = core.VariableStore()
    core.ctx.store = store
    with core.building(), core.no_grad():
        x = ops.Var(None, (2, 8, 8, 6))
        assert nn.conv2d_WN(x, 12, pad='VALID', name='a', use_weight_normalization=True).shape == (2, 6, 6, 12)
        assert nn.conv2d_WN(x, 12, stride=[2, 2], name='b').shape == (2, 4, 4, 12)
        assert nn.NiN_WN(x, 5, name='nin', use_weight_normalization=True,
                         use_mean_only_batch_normalization=True).shape == (2, 8, 8, 5)
        assert 'nin/nin/V' in store.vars and store.vars['nin/nin/V'].shape == (6, 5)
        assert nn.dense_WN(ops.Var(None, (7, 6)), 3, name='d', use_weight_normalization=True).shape == (7, 3)
        c = {}
        assert nn.conv2d(x, 4, counters=c).shape == (2, 8, 8, 4) and 'conv2d_0/V' in store.vars
        assert nn.deconv2d(x, 4, filter_size=[5, 5], stride=[2, 2], counters=c).shape == (2, 16, 16, 4)
        assert store.vars['deconv2d_0/V'].shape == (5, 5, 4, 6)
        assert nn.dense(ops.Var(None, (3, 6)), 9, counters=c).shape == (3, 9)
        assert nn.nin(x, 7, counters=c).shape == (2, 8, 8, 7) and 'dense_1/V' in store.vars
        assert nn._linear_fc(ops.Var(None, (3, 6)), 4, 'fc').shape == (3, 4) and 'fc/fc/kernel' in store.vars
        assert nn._deconv2d(x, 3, name='dc').shape == (2, 16, 16, 3)
        assert nn.batch_norm_contrib(x, 'bn', train=True).shape == (2, 8, 8, 6)
        assert {'bn/beta', 'bn/gamma', 'bn/moving_mean', 'bn/moving_variance'} <= set(store.vars)
        h = ops.concat_label(x, ops.Var(None, (2, 10)))
        assert h.shape == (2, 8, 8, 16)


def test_config_presets_and_schedule():
    c = tgan.make_config('cifar10')
    assert (c.BATCH_SIZE_G, c.BATCH_SIZE_L_C, c.BATCH_SIZE_U_C, c.BATCH_SIZE_L_D, c.BATCH_SIZE_U_D) == (100, 50, 50, 20, 80)
    assert c.IMAGE_DIM == [32, 32, 3] and c.CLA_LEARNINIG_RATE == 3e-3 and c.FAKE_G_LAMBDA == 0.3 and c.BETA1 == 0.5
    m = tgan.make_config('mnist')
    assert m.BATCH_SIZE_L_C == 100 and m.IMAGE_DIM == [28, 28, 1] and m.LEARNING_RATE == 1e-3
    with pytest.raises(ValueError):
        tgan.make_config('prostate')
    tr = tgan.Train(c)
    assert tr.schedule(1) == (0.0, 0, 3e-4, 3e-3)              # Train_goodGAN.py:165-177
    assert tr.schedule(68)[1] == 0.5 and tr.schedule(201)[0] == 0.3
    l1, l2, lr, clr = tr.schedule(301)
    assert abs(lr - 3e-4 * 0.995 ** 2) < 1e-12 and abs(clr - 3e-3 * 0.99 ** 2) < 1e-12


def test_unknown_dataset_raises_like_the_reference():
    class Cfg(tgan.Config):
        DATA_NAME, NUM_CLASSES, MINIBATCH_DIS = 'prostate', 10, False
    core.ctx.store = core.VariableStore()
    with core.building(), pytest.raises(ValueError):
        tgan.Good_GAN(Cfg()).good_generator(ops.Var(None, (2, 100)), ops.Var(None, (2, 10)))


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        tgan.init('cuda:0')


def test_train_epoch_uses_schedule_and_pretrain_phase():
    """Train.train_epoch (Train_goodGAN.py:160-276): lambda / lr schedule per epoch, classifier-only iterations while
    PRE_TRAIN and epoch <= 30 -- checked on the host with the step stubbed out (no GPU)."""
    import torch
    import tgan
    tr = tgan.make_trainer('cifar10', build_only=True)
    calls = []

    def fake_step(batch, **kw):
        calls.append(kw)
        return torch.tensor([1.0, 2.0, 3.0])
    tr.step = fake_step
    tr.config.PRE_TRAIN = True
    out = tr.train_epoch([{}, {}], epoch=5)
    assert out == [1.0, 2.0, 3.0] and len(calls) == 2 and all(c['phases'] == 'C' for c in calls)
    assert calls[0]['lambda_1'] == 0.0 and calls[0]['lambda_2'] == 0
    calls.clear()
    tr.train_epoch([{}], epoch=31)
    assert calls[0]['phases'] == 'DGC'
    calls.clear()
    tr.config.PRE_TRAIN = False
    tr.train_epoch([{}], epoch=301)
    c = tr.config
    assert calls[0]['lambda_1'] == c.FAKE_G_LAMBDA and calls[0]['lambda_2'] == 0.5
    assert abs(calls[0]['lr'] - c.LEARNING_RATE * 0.995 ** 2) < 1e-12 and abs(calls[0]['cla_lr'] - c.CLA_LEARNINIG_RATE * 0.99 ** 2) < 1e-12


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = the float32 oracle port, timed on the host cores): one JSON
    line with the GPU arm's metric / unit / config plus impl, cpu_baseline and a zero-copy e2e block; other ranks print
    nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK='0', WORLD_SIZE='1')
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j['impl'] == 'reference' and j['metric'] == 'CIFAR-10 Triple-GAN train images/sec' and j['unit'] == 'images/s'
    assert j['higher_is_better'] is True and j['value'] > 0 and j['steps'] == 1
    # the same config block as the GPU arm: the FULL batch tuple (100 images per step), not a reduced sample
    assert j['config'] == {'workload': j['config']['workload'], 'global_batch': 100, 'parallelism': 'dp1'}
    assert j['config']['workload'].startswith('CIFAR-10 32x32x3 Triple-GAN')
    assert abs(j['value'] - 100.0 / (j['ms_per_step'] * 1e-3)) < 1e-6 * j['value']
    assert j['cpu_baseline']['kind'] == 'port' and 1 <= j['cpu_baseline']['cores'] <= os.cpu_count()
    assert '100 images/step' in j['cpu_baseline']['sample']
    assert j['cpu_baseline']['value'] == j['value']
    assert j['e2e'] == {'value': j['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    env['RANK'] = '1'
    env['WORLD_SIZE'] = '2'
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0 and out.stdout.strip() == ''
    # BASELINE.json configs[0]: the MNIST step on the reference's CPU path
    env = dict(os.environ, RANK='0', WORLD_SIZE='1')
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--workload', 'mnist',
                          '--steps', '1', '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads(out.stdout.strip().splitlines()[-1])
    assert j['metric'] == 'MNIST Triple-GAN train images/sec' and j['config']['workload'].startswith('MNIST 28x28x1')
    assert j['value'] > 0 and j['config']['global_batch'] == 100


def test_pre_train_rule_of_the_dataset_mains():
    """Train_goodGAN.py:537-538, 614-615: PRE_TRAIN (classifier-only iterations for the first 30 epochs) switches on
    when NUM_LABEL < 1000 in the svhn and cifar10 mains; the mnist main has the rule commented out (:692-693)."""
    from tgan.config import make_config
    assert make_config('svhn').NUM_LABEL == 500 and make_config('svhn').PRE_TRAIN is True
    assert make_config('cifar10').NUM_LABEL == 4000 and make_config('cifar10').PRE_TRAIN is False
    assert make_config('cifar10', NUM_LABEL=500).PRE_TRAIN is True
    assert make_config('mnist').NUM_LABEL == 100 and make_config('mnist').PRE_TRAIN is False
    assert make_config('svhn', PRE_TRAIN=False).PRE_TRAIN is False
