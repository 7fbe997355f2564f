"""Synthetic step inputs of the benchmark shapes: the datasets and the
cifar10_zca_{mean,mat}.npy files of the reference are not in its repository, so the benchmark draws
z ~ U(-1,1), images ~ U(-1,1) (cifar10 / svhn, cifar10Dataset.py:60) or U(0,1) (mnist, mnistDataset.py:65),
random one-hot labels (Train_goodGAN.py:234-239) and a seeded random orthogonal ZCA matrix with zero mean."""
import numpy as np


def make_zca(seed=1234):
    rng = np.random.default_rng(seed + 7)
    q, r = np.linalg.qr(rng.standard_normal((3072, 3072)))
    return np.zeros(3072, np.float32), (q * np.sign(np.diag(r))).astype(np.float32)


def make_batch(cfg, seed=1234):
    rng = np.random.default_rng(seed)
    lo = 0.0 if cfg.DATA_NAME == 'mnist' else -1.0
    img = lambda n: rng.uniform(lo, 1.0, [n] + list(cfg.IMAGE_DIM)).astype(np.float32)

    def onehot(n):
        y = np.zeros((n, cfg.NUM_CLASSES), np.float32)
        y[np.arange(n), rng.integers(0, cfg.NUM_CLASSES, n)] = 1
        return y
    return dict(z_g=rng.uniform(-1, 1, (cfg.BATCH_SIZE_G, cfg.Z_DIM)).astype(np.float32), y_g=onehot(cfg.BATCH_SIZE_G),
                x_l_c=img(cfg.BATCH_SIZE_L_C), y_l_c=onehot(cfg.BATCH_SIZE_L_C), x_l_d=img(cfg.BATCH_SIZE_L_D),
                y_l_d=onehot(cfg.BATCH_SIZE_L_D), x_u_d=img(cfg.BATCH_SIZE_U_D), x_u_c=img(cfg.BATCH_SIZE_U_C))
