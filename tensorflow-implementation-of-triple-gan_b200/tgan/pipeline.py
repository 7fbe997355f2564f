"""Device-side input pipeline and sample grid.

The reference feeds the step from tf.data on the host: TFRecord -> decode_raw uint8 -> float32 pixel map -> one-hot label
-> shuffle -> repeat -> batch (Input_Pipeline/cifar10Dataset.py:42-128, svhnDataset.py:41-131, mnistDataset.py:42-133),
draws z / y_g with numpy every iteration (Train_goodGAN.py:232-237), slices the unlabelled batch into the D and C parts
(:255-256) and copies ~2.5 MB of float32 to the device per step.  Here the datasets (CIFAR-10: 150 MB of uint8) stay
resident in HBM and a step's eight input tensors are formed by a handful of kernels directly into the trainer's static
input buffers (the ones the captured CUDA graph reads): no host work, no H2D traffic inside the training loop.

  DeviceDataset      uint8 images + int32 labels in HBM; gather(idx) applies the reference's pixel map and one-hot
  TripleGANInput     the NNIO tuple (x_l_c, y_l_c, x_l_d, y_l_d, x_u) + z_g / y_g: epoch permutations on the device,
                     `next_into(trainer.inputs)` fills the 8 buffers; labelled stream repeats forever (repeat=-1),
                     the unlabelled one defines the epoch
  image_grid         utils.py:199-231 save_images' `merge(inverse_transform(x), image_manifold_size(n))` on the device

TFRecord files themselves are not parsed (the reference's record files are not in its repository); a dataset is handed
over as numpy / torch uint8 arrays.
"""
import numpy as np
import torch

from . import _lib, ops
from .core import ctx

PIXEL_MODE = {'cifar10': 0, 'svhn': 0, 'mnist': 1}      # 0: x/255*2-1, 1: x/255


class DeviceDataset:
    def __init__(self, images, labels, num_classes, data_name='cifar10', device=None):
        dev = device or ctx.device
        im = torch.as_tensor(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
        if im.dtype != torch.uint8:
            raise ValueError('DeviceDataset: images must be uint8 (decode_raw(..., tf.uint8), cifar10Dataset.py:52)')
        if data_name not in PIXEL_MODE:
            raise ValueError('unknown dataset %r' % (data_name,))
        self.n = int(im.shape[0])
        self.image_shape = tuple(int(s) for s in im.shape[1:])
        self.elems = int(np.prod(self.image_shape))
        if self.elems % 16:
            raise ValueError('DeviceDataset: image size must be a multiple of 16 bytes')
        self.images = im.contiguous().to(dev)
        lb = torch.as_tensor(np.asarray(labels)) if not torch.is_tensor(labels) else labels
        if lb.numel() != self.n:
            raise ValueError('DeviceDataset: %d labels for %d images' % (lb.numel(), self.n))
        self.labels = lb.reshape(-1).to(torch.int32).contiguous().to(dev)
        self.K = int(num_classes)
        self.mode = PIXEL_MODE[data_name]

    def gather(self, idx, out_x, out_y=None):
        """out_x[i] = pixel_map(images[idx[i]]), out_y[i] = one_hot(labels[idx[i]]); idx: device int64 (None = head)."""
        n = int(out_x.shape[0])
        if idx is not None and (idx.dtype != torch.int64 or idx.numel() < n or not idx.is_contiguous()):
            raise ValueError('gather: idx must be a contiguous int64 tensor with at least %d entries' % n)
        if out_x.dtype != torch.float32 or out_x.numel() != n * self.elems or not out_x.is_contiguous():
            raise ValueError('gather: out_x must be contiguous float32 [n, %d]' % self.elems)
        ip = None if idx is None else idx.data_ptr()
        _lib.call('tgan_gather_images_u8', self.images.data_ptr(), self.n, self.elems, ip, n, out_x.data_ptr(), self.mode,
                  ops._st())
        if out_y is not None:
            _lib.call('tgan_gather_onehot', self.labels.data_ptr(), self.n, ip, n, self.K, out_y.data_ptr(), ops._st())
        return out_x, out_y


class TripleGANInput:
    """The step's inputs, formed on the device.  One instance per rank; `seed` should differ per rank (ddp.rank_seed)."""

    def __init__(self, config, labelled, unlabelled, seed=1234, train_size=None):
        self.c, self.lab, self.unl = config, labelled, unlabelled
        self.seed = int(seed)
        self.gen = torch.Generator(device=labelled.images.device)
        self.gen.manual_seed(self.seed)
        self.counter = torch.zeros(1, dtype=torch.int64, device=labelled.images.device)   # Philox step counter for z / y_g
        # TRAIN_SIZE = number of unlabelled images (Train_goodGAN.py:523, 599, 677)
        self.train_size = int(unlabelled.n if train_size is None else train_size)
        # three independently shuffled, endlessly repeating streams (the reference builds separate tf.data pipelines
        # for x_l_c, x_l_d and x_u, each shuffle -> repeat(-1) -> batch)
        self._streams = {}
        self._step = 0
        self.start_epoch()

    def _perm(self, n):
        return torch.randperm(n, device=self.lab.images.device, generator=self.gen)

    def start_epoch(self):
        """init_op_train (Train_goodGAN.py:180): re-initialise the iterators -> every stream restarts with a fresh shuffle."""
        self._streams = {k: [self._perm(n), 0] for k, n in (('l_c', self.lab.n), ('l_d', self.lab.n), ('u', self.unl.n))}
        self._step = 0

    def steps_per_epoch(self):
        """int(TRAIN_SIZE / BATCH_SIZE) iterations (Train_goodGAN.py:230); the streams repeat, so the 130 unlabelled
        images a CIFAR-10 iteration consumes wrap around within the epoch."""
        return int(self.train_size / self.c.BATCH_SIZE)

    def _take(self, key, n, total):
        if n > total:
            raise ValueError('dataset smaller than one batch')
        st = self._streams[key]
        if st[1] + n > total:                            # repeat(-1): reshuffle and carry on
            st[0], st[1] = self._perm(total), 0
        idx = st[0][st[1]:st[1] + n]
        st[1] += n
        return idx

    def next_into(self, inputs):
        """Fill the eight step inputs (train.INPUT_NAMES) in place.  Raises StopIteration at the end of the epoch
        after int(TRAIN_SIZE / BATCH_SIZE) iterations (the range of Train_goodGAN.py:230)."""
        c = self.c
        nu = c.BATCH_SIZE_U_D + c.BATCH_SIZE_U_C
        if self._step >= self.steps_per_epoch():
            raise StopIteration
        self._step += 1
        iu = self._take('u', nu, self.unl.n)
        self.lab.gather(self._take('l_c', c.BATCH_SIZE_L_C, self.lab.n), inputs['x_l_c'], inputs['y_l_c'])
        self.lab.gather(self._take('l_d', c.BATCH_SIZE_L_D, self.lab.n), inputs['x_l_d'], inputs['y_l_d'])
        self.unl.gather(iu[:c.BATCH_SIZE_U_D], inputs['x_u_d'])                         # x_u[:U_D]        (:255)
        self.unl.gather(iu[c.BATCH_SIZE_U_D:], inputs['x_u_c'])                         # x_u[U_D:U_D+U_C] (:256)
        _lib.call('tgan_draw_latent', inputs['z_g'].data_ptr(), c.BATCH_SIZE_G, c.Z_DIM, inputs['y_g'].data_ptr(),
                  c.NUM_CLASSES, self.seed, self.counter.data_ptr(), 0x7a5eed, ops._st())
        _lib.call('tgan_counter_advance', self.counter.data_ptr(), 1, ops._st())
        return inputs

    def epoch(self, inputs):
        """Iterator for Train.train_epoch: yields None after filling `inputs` in place (Train.step(None) then runs on
        the static buffers): int(TRAIN_SIZE / BATCH_SIZE) iterations."""
        self.start_epoch()
        while True:
            try:
                self.next_into(inputs)
            except StopIteration:
                return
            yield None


def image_manifold_size(num_images):
    """utils.py:192-196"""
    h = int(np.floor(np.sqrt(num_images)))
    w = int(np.ceil(np.sqrt(num_images)))
    assert h * w == num_images
    return h, w


def image_grid(images, size=None, inverse=True):
    """merge(inverse_transform(images), size) of utils.py:199-231 on the device: images float32 [n,H,W,C] device tensor
    -> [gh*H, gw*W, C] (C squeezed when 1, like np.squeeze in imsave)."""
    if images.dim() != 4 or images.dtype != torch.float32:
        raise ValueError('image_grid: float32 [n, H, W, C] expected')
    n, H, W, C = (int(s) for s in images.shape)
    if C not in (1, 3, 4):
        raise ValueError('in merge(images,size) images parameter must have dimensions: HxW or HxWx3 or HxWx4')
    gh, gw = size if size is not None else image_manifold_size(n)
    grid = torch.empty((gh * H, gw * W, C), dtype=torch.float32, device=images.device)
    x = images.contiguous()
    _lib.call('tgan_image_grid', x.data_ptr(), n, H, W, C, int(gh), int(gw), grid.data_ptr(), 1 if inverse else 0, ops._st())
    return grid[..., 0] if C == 1 else grid
