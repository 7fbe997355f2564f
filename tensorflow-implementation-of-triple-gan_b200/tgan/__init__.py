"""tgan -- B200-native (sm_100a) Triple-GAN training hot path behind the reference's Python operator
surface (Model/nn.py, Model/modle_base.py, Model/Good_GAN*.py, Training/Train_goodGAN.py step).

    import tgan
    tgan.init('cuda:0', math='bf16')                  # fails loudly without a GPU / libtgan.so
    tr = tgan.make_trainer('cifar10')                 # builds the graph, initialises variables
    losses = tr.step(batch)                           # device tensor [d_loss, g_loss, c_loss]
"""
from . import checkpoint, config, core, ddp, nn, ops, pipeline, synthetic, tfrecord  # noqa: F401
from .checkpoint import Saver  # noqa: F401
from .config import Config, make_config  # noqa: F401
from .core import InjectedSource, PhiloxSource, building, ctx, init, no_grad, recording  # noqa: F401
from .good_gan import Good_GAN  # noqa: F401
from .good_gan_cifar10 import Good_GAN_cifar10, cifar10_ZCA  # noqa: F401
from .model_base import NN_Base  # noqa: F401
from .train import Train  # noqa: F401
from .train_base import AdamOptimizer, Train_base  # noqa: F401


def model_for(data_name):
    """Train_goodGAN.py:477-479, :555, :631: cifar10 -> Good_GAN_cifar10, svhn / mnist -> Good_GAN."""
    return Good_GAN_cifar10 if data_name == 'cifar10' else Good_GAN


def make_trainer(data_name, scale=1, init=None, zca=None, seed=1234, build_only=False, **cfg_over):
    """Config preset -> Train -> graph -> (optionally) initialised variables."""
    cfg = make_config(data_name, scale, **cfg_over)
    if data_name == 'cifar10':
        if zca is None:
            import numpy as np
            zca = (np.zeros(3072, np.float32), np.eye(3072, dtype=np.float32))
        cfg.ZCA = zca
    tr = Train(cfg, seed=seed)
    tr._build_train_graph(model_for(data_name))
    if not build_only:
        tr.initialize(init)
    return tr
