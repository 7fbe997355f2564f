"""-m gpu: the data formats either side of the step -- TF-format checkpoints through
Training/Saver.py's surface, the device-side input pipeline, the sample grid.  Integer / byte work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import tgan_oracle as O

pytestmark = pytest.mark.gpu


def _trainer(math, pseed, scale=10):
    import tgan
    tgan.init('cuda:0', math=math, seed=77)
    P, S = O.init_params('cifar10', seed=pseed)
    tr = tgan.make_trainer('cifar10', scale=scale, init=(P, S), zca=O.make_zca(3))
    tr.load_batch(O.make_batch(O.OracleConfig('cifar10', scale), seed=9))
    return tr


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
def test_checkpoint_resume_is_bitwise(tmp_path, math):
    """save after two steps, restore into a differently initialised trainer, run the third step: identical bits to the
    uninterrupted run -- i.e. the checkpoint holds ALL the state of the step (variables, pop_mean / BN moving statistics,
    Adam slots and beta powers, EMA shadows) and the restore path puts every tensor back in its place."""
    import tgan
    from tgan import core
    a = _trainer(math, 5)
    for _ in range(2):
        a.step(lambda_1=0.3, lambda_2=0.5)
    sv = tgan.Saver(str(tmp_path))
    sv.set_save_path(comments='resume test')
    sv.save(a, 'model_0002.ckpt')
    files = sorted(os.listdir(sv.save_dir))
    assert files == ['Comments.txt', 'checkpoint', 'model_0002.ckpt.data-00000-of-00001', 'model_0002.ckpt.index']
    want_l = a.step(lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy()
    want = tgan.checkpoint.state_dict(a)

    b = _trainer(math, 6)                           # different weights, fresh optimiser state
    start = tgan.Saver(str(tmp_path)).restore(b)
    assert start == 2
    core.ctx.rng.counter().fill_(2)                 # the Philox step counter is not a TF variable (nor is TF's RNG state)
    got_l = b.step(lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy()
    got = tgan.checkpoint.state_dict(b)
    assert np.array_equal(want_l, got_l), (want_l, got_l)
    assert sorted(want) == sorted(got)
    for k in want:
        assert np.array_equal(want[k], got[k]), k


def test_checkpoint_names_and_partial_restore(tmp_path):
    """the saved names are the ones tf.train.Saver() would write for the reference graph (SURVEY.md §8a inventory);
    a file with model variables only (e.g. exported from elsewhere) restores, a missing model variable fails loudly."""
    import tgan
    from tgan import checkpoint as ck
    a = _trainer('fp32', 5)
    sd = ck.state_dict(a)
    for k in ('classifier/conv1_1/V', 'classifier/conv1_1/g', 'classifier/conv1_1/b',
              'classifier/conv1_1/meanOnlyBatchNormalization/pop_mean', 'classifier/NiN1/NiN1/V',
              'classifier/output_dense/V', 'discriminator/conv2d_00/conv2d_00/kernel', 'discriminator/lin/lin/bias',
              'good_generator/gg_h0_lin/gg_h0_lin/kernel', 'good_generator/gg_bn0/moving_variance',
              'good_generator/gg_dconv2/gg_dconv2/kernel', 'classifier/conv1_1/V/Adam_optimizer',
              'classifier/conv1_1/V/Adam_optimizer_1', 'classifier/conv1_1/V/ExponentialMovingAverage',
              'Train/beta1_power', 'Train/beta2_power_2'):
        assert k in sd, k
    assert sd['good_generator/gg_dconv0/gg_dconv0/kernel'].shape == (5, 5, 256, 522)
    assert sd['Train/beta1_power'].shape == () and sd['Train/beta1_power'] == np.float32(0.5)
    assert not any('ExponentialMovingAverage' in k for k in sd if not k.startswith('classifier/'))
    model_only = {k: v for k, v in sd.items() if k in a.store.vars}
    ck.write_bundle(str(tmp_path / 'w.ckpt'), model_only)
    b = _trainer('fp32', 6)
    r = ck.BundleReader(str(tmp_path / 'w.ckpt'))
    ck.load_state_dict(b, r.get, r.has)
    sb = ck.state_dict(b)
    for k in model_only:
        assert np.array_equal(sb[k], model_only[k]), k
    assert np.array_equal(sb['classifier/conv1_1/V/ExponentialMovingAverage'], model_only['classifier/conv1_1/V'])
    del model_only['classifier/conv2_2/g']
    ck.write_bundle(str(tmp_path / 'x.ckpt'), model_only)
    r = ck.BundleReader(str(tmp_path / 'x.ckpt'))
    with pytest.raises(KeyError, match='classifier/conv2_2/g'):
        ck.load_state_dict(b, r.get, r.has)


@pytest.mark.parametrize('name,shape', [('cifar10', (32, 32, 3)), ('svhn', (32, 32, 3)), ('mnist', (28, 28, 1))])
def test_gather_is_bit_exact(name, shape):
    """pixel map of cifar10Dataset.py:56-62 / svhnDataset.py:59-65 / mnistDataset.py:60-67 in float32, every byte value"""
    import tgan
    from tgan import pipeline
    tgan.init('cuda:0', math='fp32')
    rng = np.random.default_rng(1)
    M = 300
    img = rng.integers(0, 256, (M,) + shape, dtype=np.uint8)
    img[0].reshape(-1)[:256] = np.arange(256, dtype=np.uint8)
    lab = rng.integers(0, 10, M)
    ds = pipeline.DeviceDataset(img, lab, 10, name)
    idx = torch.from_numpy(np.concatenate([[0, M - 1, 5, 5], rng.integers(0, M, 93)])).cuda()
    x = torch.empty((97,) + shape, device='cuda')
    y = torch.empty((97, 10), device='cuda')
    ds.gather(idx, x, y)
    f = img[idx.cpu().numpy()].astype(np.float32)
    want = f / np.float32(255) * np.float32(2) - np.float32(1) if name != 'mnist' else f / np.float32(255.0)
    assert want.dtype == np.float32
    assert np.array_equal(x.cpu().numpy(), want)
    oh = np.zeros((97, 10), np.float32)
    oh[np.arange(97), lab[idx.cpu().numpy()]] = 1
    assert np.array_equal(y.cpu().numpy(), oh)
    ds.gather(None, x)                                            # head of the dataset
    f = img[:97].astype(np.float32)
    assert np.array_equal(x.cpu().numpy(), f / np.float32(255) * np.float32(2) - np.float32(1) if name != 'mnist' else f / np.float32(255))
    bad = torch.tensor([-3, M + 7] + [1] * 95, device='cuda')      # out-of-range indices are clamped, never read outside
    ds.gather(bad, x)
    got = x.cpu().numpy()
    m = (lambda a: a.astype(np.float32) / np.float32(255) * np.float32(2) - np.float32(1)) if name != 'mnist' else (lambda a: a.astype(np.float32) / np.float32(255))
    assert np.array_equal(got[0], m(img[0])) and np.array_equal(got[1], m(img[M - 1]))
    with pytest.raises(ValueError):
        pipeline.DeviceDataset(img.astype(np.float32), lab, 10, name)
    with pytest.raises(ValueError):
        pipeline.DeviceDataset(img, lab, 10, 'imagenet')


def _indexed_dataset(M, seed):
    """images whose first 4 bytes spell their own index"""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (M, 32, 32, 3), dtype=np.uint8)
    flat = img.reshape(M, -1)
    flat[:, :4] = np.arange(M, dtype='<u4').view(np.uint8).reshape(M, 4)
    return img, rng.integers(0, 10, M)


def _ids(x):
    b = np.rint((x.reshape(x.shape[0], -1)[:, :4].cpu().numpy().astype(np.float64) + 1) / 2 * 255).astype(np.int64)
    return b[:, 0] + 256 * b[:, 1] + 65536 * b[:, 2]


def test_input_stream_epoch_semantics():
    """one epoch visits every unlabelled image at most once, x_u is split [:U_D] / [U_D:U_D+U_C] (Train_goodGAN.py:255-256),
    the labelled stream repeats, z ~ U(-1,1), y_g one-hot; same seed -> same stream"""
    import tgan
    from tgan import pipeline
    tgan.init('cuda:0', math='fp32')
    cfg = tgan.make_config('cifar10', 1)
    Mu, Ml = 1000, 130
    iu, lu = _indexed_dataset(Mu, 1)
    il, ll = _indexed_dataset(Ml, 2)
    unl = pipeline.DeviceDataset(iu, lu, 10, 'cifar10')
    lab = pipeline.DeviceDataset(il, ll, 10, 'cifar10')

    def run(seed):
        inp = pipeline.TripleGANInput(cfg, lab, unl, seed=seed)
        shapes = dict(z_g=(100, 100), y_g=(100, 10), x_l_c=(50, 32, 32, 3), y_l_c=(50, 10), x_l_d=(20, 32, 32, 3),
                      y_l_d=(20, 10), x_u_d=(80, 32, 32, 3), x_u_c=(50, 32, 32, 3))
        bufs = {k: torch.zeros(v, device='cuda') for k, v in shapes.items()}
        seen_u, seen_l, zs, ys = [], [], [], []
        n = 0
        for _ in inp.epoch(bufs):
            n += 1
            seen_u.append(np.concatenate([_ids(bufs['x_u_d']), _ids(bufs['x_u_c'])]))
            lc, ld = _ids(bufs['x_l_c']), _ids(bufs['x_l_d'])
            assert np.array_equal(bufs['y_l_c'].cpu().numpy().argmax(1), ll[lc])
            assert np.array_equal(bufs['y_l_d'].cpu().numpy().argmax(1), ll[ld])
            seen_l.append(np.concatenate([lc, ld]))
            zs.append(bufs['z_g'].cpu().numpy().copy())
            ys.append(bufs['y_g'].cpu().numpy().copy())
        # the reference's epoch: int(TRAIN_SIZE / BATCH_SIZE) iterations (Train_goodGAN.py:230), TRAIN_SIZE = the
        # number of unlabelled images; each iteration takes 130 of them, so the repeating stream wraps
        assert n == inp.steps_per_epoch() == Mu // 100
        return np.concatenate(seen_u), np.concatenate(seen_l), np.stack(zs), np.stack(ys)
    su, sl, z, y = run(11)
    first = su[:7 * 130]                                                       # one pass over the shuffled stream
    assert len(np.unique(first)) == len(first) and su.max() < Mu               # no repeats before the wrap-around
    assert len(su) == 10 * 130 and len(np.unique(su)) > 900                    # ... then it reshuffles and repeats
    lc = sl.reshape(10, 70)[:, :50].reshape(-1)
    ld = sl.reshape(10, 70)[:, 50:].reshape(-1)
    assert sl.max() < Ml and len(np.unique(lc[:100])) == 100 and len(np.unique(ld[:120])) == 120    # separate shuffled streams
    assert not np.array_equal(lc[:20], ld[:20]) and len(np.unique(sl)) >= Ml - 3
    assert z.min() >= -1 and z.max() < 1 and abs(z.mean()) < 0.02 and abs(z.std() - 3 ** -0.5) < 0.02
    assert not np.array_equal(z[0], z[1])
    assert np.array_equal(y.sum(-1), np.ones(y.shape[:2])) and set(np.unique(y)) == {0.0, 1.0}
    h = y.reshape(-1, 10).sum(0)
    assert h.min() > 55 and h.max() < 150                                      # 1000 draws over 10 classes
    su2, sl2, z2, y2 = run(11)
    assert np.array_equal(su, su2) and np.array_equal(sl, sl2) and np.array_equal(z, z2) and np.array_equal(y, y2)
    su3, _, z3, _ = run(12)
    assert not np.array_equal(su, su3) and not np.array_equal(z, z3)


@pytest.mark.parametrize('C', [3, 1])
def test_image_grid_matches_merge(C):
    """utils.py:199-231 merge(inverse_transform(images), size), restated in numpy"""
    import tgan
    from tgan import pipeline
    tgan.init('cuda:0', math='fp32')
    rng = np.random.default_rng(C)
    H = W = 28 if C == 1 else 32
    x = rng.uniform(-1, 1, (64, H, W, C)).astype(np.float32)
    g = pipeline.image_grid(torch.from_numpy(x).cuda(), pipeline.image_manifold_size(64)).cpu().numpy()
    inv = (x + np.float32(1.)) / np.float32(2.)
    want = np.zeros((8 * H, 8 * W, C), np.float32)
    for k in range(64):
        i, j = k % 8, k // 8
        want[j * H:(j + 1) * H, i * W:(i + 1) * W] = inv[k]
    assert np.array_equal(g, want[..., 0] if C == 1 else want)
    g2 = pipeline.image_grid(torch.from_numpy(x[:6]).cuda(), (2, 4), inverse=False).cpu().numpy()     # 2 empty cells stay 0
    assert g2.shape[:2] == (2 * H, 4 * W) and not g2[H:, 2 * W:].any() and np.array_equal(g2[:H, :W].reshape(H, W, C), x[0])
    with pytest.raises(AssertionError):
        pipeline.image_manifold_size(50)


def test_epoch_from_device_pipeline_and_sample():
    """Train.train_epoch fed by the device pipeline through the captured graph; the per-epoch sample grid"""
    import tgan
    from tgan import pipeline
    tr = _trainer('bf16', 5, scale=1)
    iu, lu = _indexed_dataset(400, 1)
    il, ll = _indexed_dataset(200, 2)
    inp = pipeline.TripleGANInput(tr.config, pipeline.DeviceDataset(il, ll, 10, 'cifar10'),
                                  pipeline.DeviceDataset(iu, lu, 10, 'cifar10'), seed=3)
    tr.capture()
    d, g, c = tr.train_epoch(inp.epoch(tr.inputs), epoch=1)
    assert all(np.isfinite(v) for v in (d, g, c)) and inp.steps_per_epoch() == 4
    ids = _ids(tr.inputs['x_u_c'])
    assert ids.max() < 400 and len(np.unique(ids)) == 50                  # the graph's static buffers hold the last batch
    rng = np.random.default_rng(0)
    y = np.zeros((100, 10), np.float32)
    y[np.arange(100), np.arange(100) % 10] = 1
    grid = tr.sample(rng.uniform(-1, 1, (100, 100)).astype(np.float32), y)
    assert tuple(grid.shape) == (256, 256, 3)
    gn = grid.cpu().numpy()
    assert np.isfinite(gn).all() and gn.min() >= 0 and gn.max() <= 1


@pytest.mark.parametrize('graph', [False, True])
def test_host_feed_matches_direct_steps(graph):
    """Train.host_feed(): the double-buffered host feed (upload of batch k+1 overlapped with step k, losses read one step
    late) gives bit-identical losses and parameters to loading each batch and stepping directly."""
    cfg = O.OracleConfig('cifar10', 10)
    batches = [O.make_batch(cfg, seed=20 + i) for i in range(4)]
    pinned = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items()} for b in batches]
    tr = _trainer('bf16', 5)
    if graph:
        tr.capture()
    direct = [tr.step(b, lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy() for b in batches]
    theta = {g: tr.store.flat[g]['theta'].cpu().numpy().copy() for g in tr.store.GROUPS}
    tr2 = _trainer('bf16', 5)
    if graph:
        tr2.capture()
    feed = tr2.host_feed()
    feed.prime(pinned[0])
    got = []
    for i in range(4):
        prev = feed.step(pinned[i + 1] if i + 1 < 4 else None, lambda_1=0.3, lambda_2=0.5)
        assert (prev is None) == (i == 0)
        if prev is not None:
            got.append(prev.numpy().copy())
    got.append(feed.drain().numpy().copy())
    assert np.array_equal(np.stack(direct), np.stack(got)), (direct, got)
    for g in theta:
        assert np.array_equal(theta[g], tr2.store.flat[g]['theta'].cpu().numpy()), g
