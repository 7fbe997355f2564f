"""Good_GAN_cifar10: the CIFAR-10 Triple-GAN builders of the reference's Model/Good_GAN_cifar10.py
(class Good_GAN_cifar10 :14-285, cifar10_ZCA :287-299) on the sm_100a path.  Same method names and
argument order; the only additions are the trailing `tag=` keyword naming the RNG streams of a pass.
"""
import os

import numpy as np
import torch

from . import model_base, nn, ops
from .core import Param, ctx, variable_scope, variance_scaling_initializer

he_init = variance_scaling_initializer()


class Good_GAN_cifar10(model_base.NN_Base):
    def __init__(self, config):
        super(Good_GAN_cifar10, self).__init__(config.BATCH_NORM_DECAY, config.BATCH_NORM_EPSILON)
        self.config = config
        self._zca = None

    def leakyReLu(self, x, alpha=0.2, name=None):
        return self._leakyReLu_impl(x, alpha)

    leakyReLu.tgan_act = ('lrelu', 0.2)      # fusable into the producing layer's epilogue (nn._fused)

    def _leakyReLu_impl(self, x, alpha):
        # tf.nn.relu(x) - alpha * tf.nn.relu(-x)  (Good_GAN_cifar10.py:26-27)
        return ops.activation(x, 'lrelu', alpha)

    def gaussian_noise_layer(self, input_layer, std, tag='noise'):
        return ops.add_noise(input_layer, std, tag)

    def good_generator(self, z, y, init=False, reuse=False, tag='G'):
        with variable_scope('good_generator', reuse=reuse):
            yb = ops.reshape(y, [y.shape[0], 1, 1, self.config.NUM_CLASSES])
            z = ops.concat_label(z, y)

            h0 = self._linear_fc(z, 4 * 4 * 512, 'gg_h0_lin', kernel_initializer=he_init)
            h0 = nn.relu(h0, 'gg_rl0')  # [4,4]
            h0 = self._batch_norm_contrib(h0, 'gg_bn0', train=True)
            h0 = ops.reshape(h0, [-1, 4, 4, 512])
            h0 = self._conv_cond_concat(h0, yb)

            h0 = self._deconv2d(h0, 256, k_w=5, k_h=5, d_w=2, d_h=2, name='gg_dconv0', kernel_initializer=he_init)
            h0 = nn.relu(h0, 'gg_rl1')
            h0 = self._batch_norm_contrib(h0, 'gg_bn1', train=True)
            h0 = self._conv_cond_concat(h0, yb)

            h1 = self._deconv2d(h0, 128, k_w=5, k_h=5, d_w=2, d_h=2, name='gg_dconv1', kernel_initializer=he_init)
            h1 = nn.relu(h1, 'gg_rl2')  # [16,16]
            h1 = self._batch_norm_contrib(h1, 'gg_bn2', train=True)
            h1 = self._conv_cond_concat(h1, yb)

            h2 = self._deconv2d(h1, 3, k_w=5, k_h=5, d_w=2, d_h=2, name='gg_dconv2', kernel_initializer=he_init)
            h2 = nn.tanh(h2)
        return h2

    def discriminator(self, image, y, init=False, reuse=False, getter=None, tag='D'):
        with variable_scope('discriminator', reuse=reuse):
            image = self._drop_out(image, 0.2, True, tag=tag + '/drop0')
            yb = ops.reshape(y, [image.shape[0], 1, 1, self.config.NUM_CLASSES])
            image = self._conv_cond_concat(image, yb)

            h0 = self._conv2d(image, 32, k_h=3, k_w=3, d_h=1, d_w=1, kernel_initializer=he_init, name="conv2d_00")
            h0 = self.leakyReLu(h0)
            h0 = self._conv_cond_concat(h0, yb)

            h0 = self._conv2d(h0, 32, k_h=3, k_w=3, d_h=2, d_w=2, kernel_initializer=he_init, name="conv2d_01")
            h0 = self.leakyReLu(h0)
            h0 = self._drop_out(h0, 0.2, True, tag=tag + '/drop1')  # [16, 16]

            h1 = self._conv_cond_concat(h0, yb)

            h1 = self._conv2d(h1, 64, k_h=3, k_w=3, d_h=1, d_w=1, kernel_initializer=he_init, name="conv2d_10")
            h1 = self.leakyReLu(h1)
            h1 = self._conv_cond_concat(h1, yb)

            h1 = self._conv2d(h1, 64, k_h=3, k_w=3, d_h=2, d_w=2, kernel_initializer=he_init, name="conv2d_11")
            h1 = self.leakyReLu(h1)
            h1 = self._drop_out(h1, 0.2, True, tag=tag + '/drop2')  # [8, 8]

            h2 = self._conv_cond_concat(h1, yb)
            h2 = self._conv2d(h2, 128, k_h=3, k_w=3, d_h=1, d_w=1, kernel_initializer=he_init, name="conv2d_20")
            h2 = self.leakyReLu(h2)
            h2 = self._conv_cond_concat(h2, yb)

            h2 = self._conv2d(h2, 128, k_h=3, k_w=3, d_h=1, d_w=1, kernel_initializer=he_init, name="conv2d_21")
            h2 = self.leakyReLu(h2)

            h3 = ops.global_pool(h2, 'mean')      # average_pooling2d(8, 1) + squeeze (:94-96)
            h3 = ops.concat_label(h3, y)
            h3 = self._linear_fc(h3, 1, 'lin', kernel_initializer=he_init)
            h3 = ops.force(h3)      # logits are consumed as is: apply the pending bias now
        return _LazySigmoid(h3), h3

    def classifier(self, inp, is_training, init=False, reuse=False, getter=None, tag='C'):
        det = not is_training
        kw = dict(init=init, use_weight_normalization=True, use_batch_normalization=False,
                  use_mean_only_batch_normalization=True, deterministic=det)
        with variable_scope('classifier', reuse=reuse):
            x = ops.reshape(inp, [-1, 32, 32, 3])
            x = self._add_noise(x, stddev=0.15, tag=tag + '/noise')

            x = nn.conv2d_WN(x, num_filters=128, name='conv1_1', nonlinearity=self.leakyReLu, **kw)
            x = nn.conv2d_WN(x, num_filters=128, name='conv1_2', nonlinearity=self.leakyReLu, **kw)
            x = nn.conv2d_WN(x, num_filters=128, name='conv1_3', nonlinearity=self.leakyReLu, **kw)

            x = ops.max_pool2(x)                                                      # max_pool_1
            x = ops.dropout(x, 0.5, tag + '/drop1', training=is_training)             # dropout_1

            x = nn.conv2d_WN(x, num_filters=256, name='conv2_1', nonlinearity=self.leakyReLu, **kw)
            x = nn.conv2d_WN(x, num_filters=256, name='conv2_2', nonlinearity=self.leakyReLu, **kw)
            x = nn.conv2d_WN(x, num_filters=256, name='conv2_3', nonlinearity=self.leakyReLu, **kw)

            x = ops.max_pool2(x)                                                      # max_pool_2
            x = ops.dropout(x, 0.5, tag + '/drop2', training=is_training)             # dropout_2

            x = nn.conv2d_WN(x, num_filters=512, name='conv3', nonlinearity=self.leakyReLu, pad='VALID', **kw)
            x = nn.NiN_WN(x, num_units=256, nonlinearity=self.leakyReLu, name='NiN1', **kw)
            x = nn.NiN_WN(x, num_units=128, nonlinearity=self.leakyReLu, name='NiN2', **kw)

            x = ops.global_pool(x, 'max')      # max_pooling2d(6, 1) named 'avg_pool_0' + squeeze (:163-165)
            intermediate_layer = x

            logits = nn.dense_WN(x, num_units=10, nonlinearity=None, name='output_dense', **kw)
        return logits, intermediate_layer

    def good_sampler(self, z, y):
        return self.good_generator(z, y, init=False, reuse=True, tag='sampler')

    def _whitener(self):
        if self._zca is None:
            self._zca = cifar10_ZCA(self.config)
        return self._zca

    def forward_pass(self, z_g, y_g, x_l_c, y_l_c, x_l_d, y_l_d, x_u_d, x_u_c, train, tag='F'):
        """Good_GAN_cifar10.forward_pass (:204-278): the whole (un-pruned) graph, evaluated eagerly."""
        store = ctx.store
        if not store.has('good_generator/gg_h0_lin/gg_h0_lin/kernel'):
            self.good_generator(z_g, y_g, init=True, reuse=False, tag=tag + '/init')
        G = self.good_generator(z_g, y_g, init=False, reuse=True, tag=tag + '/G')

        cif = self.config.DATA_NAME == "cifar10"
        pre = self._whitener().apply if cif else (lambda t: t)
        x_u_c_zca, x_l_c_zca, x_u_d_zca, G_zca = pre(x_u_c), pre(x_l_c), pre(x_u_d), pre(G)
        if not store.has('classifier/conv1_1/V'):
            self.classifier(x_u_c_zca, train, init=True, reuse=False, tag=tag + '/init')
        C_real_logits, _ = self.classifier(x_l_c_zca, train, init=False, reuse=True, tag=tag + '/C_real')
        C_unl_logits, _ = self.classifier(x_u_c_zca, train, init=False, reuse=True, tag=tag + '/C_unl')
        _, C_unl_onehot = ops.argmax_onehot(C_unl_logits, self.config.NUM_CLASSES)
        if cif:
            C_unl_logits_rep, _ = self.classifier(x_u_c_zca, train, init=False, reuse=True, tag=tag + '/C_unl_rep')
        C_unl_d_logits, _ = self.classifier(x_u_d_zca, train, init=False, reuse=True, tag=tag + '/C_unl_d')
        _, C_unl_d_onehot = ops.argmax_onehot(C_unl_d_logits, self.config.NUM_CLASSES)
        C_fake_logits, _ = self.classifier(G_zca, train, init=False, reuse=True, tag=tag + '/C_fake')

        X_P = concat_batch([x_l_d, x_u_d])
        Y_P = concat_batch([y_l_d, C_unl_d_onehot])
        if not store.has('discriminator/conv2d_00/conv2d_00/kernel'):
            self.discriminator(X_P, Y_P, init=True, reuse=False, tag=tag + '/init')
        D_real, D_real_logits = self.discriminator(X_P, Y_P, init=False, reuse=True, tag=tag + '/D_real')
        D_fake, D_fake_logits = self.discriminator(G, y_g, init=False, reuse=True, tag=tag + '/D_fake')
        D_unl, D_unl_logits = self.discriminator(x_u_c, C_unl_onehot, init=False, reuse=True, tag=tag + '/D_unl')

        C = [C_real_logits, C_unl_logits, C_unl_d_logits, C_fake_logits]
        if cif:
            C.append(C_unl_logits_rep)
        return [G, [D_real, D_real_logits, D_fake, D_fake_logits, D_unl, D_unl_logits], C]

    def forward_pass_CGAN(self, z, image, y):
        G = self.good_generator(z, y, reuse=False)
        D_real, D_real_logits = self.discriminator(image, y, reuse=False)
        D_fake, D_fake_logits = self.discriminator(G, y, reuse=True)
        return G, [D_real, D_real_logits, D_fake, D_fake_logits]


class _LazySigmoid:
    """`tf.nn.sigmoid(h3)` returned next to the logits (:99).  No loss reads it, so it is only
    evaluated on demand."""

    def __init__(self, logits):
        self._logits, self._v = logits, None

    def value(self):
        if self._v is None:
            self._v = nn.sigmoid(self._logits)
        return self._v

    @property
    def data(self):
        return self.value().data

    def numpy(self):
        return self.value().numpy()


def concat_batch(vs):
    """tf.concat(axis=0) of same-shape-suffix tensors (X_P / Y_P, :258-259); no gradient needed
    (inputs / pseudo-labels)."""
    shape = (sum(v.shape[0] for v in vs),) + tuple(vs[0].shape[1:])
    if ctx.building:
        return ops.Var(None, shape)
    out = torch.empty(shape, dtype=vs[0].data.dtype, device=ctx.device)
    o = 0
    for v in vs:          # device-to-device copies through the copy engine (cudaMemcpyAsync), graph capturable
        out[o:o + v.shape[0]].copy_(v.data)
        o += v.shape[0]
    return ops.Var(out, shape)


class cifar10_ZCA():
    """(flat(x) - mean) @ mat  (Good_GAN_cifar10.py:287-299), evaluated as x @ mat + (-(mean @ mat)) so the
    mean subtraction rides in the GEMM epilogue.  DATA_DIR/cifar10_zca_{mean,mat}.npy are used when
    present; otherwise `config.ZCA = (mean, mat)` supplies them (synthetic benchmark)."""

    def __init__(self, config):
        d = getattr(config, 'DATA_DIR', None)
        if d and os.path.exists(os.path.join(d, "cifar10_zca_mean.npy")):
            m = np.load(os.path.join(d, "cifar10_zca_mean.npy"))
            mat = np.load(os.path.join(d, "cifar10_zca_mat.npy"))
        else:
            m, mat = config.ZCA
        m, mat = np.asarray(m, np.float64).reshape(-1), np.asarray(mat, np.float64)
        self.mat = Param('_zca/mat', mat.shape, False, mat.astype(np.float32))
        self.bias = Param('_zca/bias', (mat.shape[1],), False, (-(m @ mat)).astype(np.float32))
        if not ctx.building:
            self._upload()

    def _upload(self):
        for p in (self.mat, self.bias):
            if p.data is None:
                p.data = torch.from_numpy(p.init_value).to(ctx.device)

    def apply(self, image):
        if not ctx.building:
            self._upload()
        s = image.shape
        flat = ops.reshape(image, [s[0], s[1] * s[2] * s[3]])
        z = ops.conv2d(flat, ops.PlainWeight(self.mat), 1, 1)
        return ops.reshape(ops.bias_act(z, self.bias, 'none'), [-1, s[1], s[2], s[3]])
