"""NN_Base: the layer-method surface of the reference's Model/modle_base.py (class NN_Base,
modle_base.py:17-259), backed by the sm_100a kernels.  Same method names, argument order, keyword names
and defaults; dead methods of the reference that call undefined helpers (`_conv_batch_*`,
modle_base.py:211-227) and the unreachable keras `Dense` (:298-345) are not reproduced.
"""
import numpy as np

from . import nn, ops
from .core import (constant_initializer, get_variable, ones_initializer, random_normal_initializer,
                   truncated_normal_initializer, variable_scope, zeros_initializer)


class NN_Base(object):
    def __init__(self, batch_norm_decay=0.9, batch_norm_epsilon=1e-5):
        self._batch_norm_decay = batch_norm_decay
        self._batch_norm_epsilon = batch_norm_epsilon

    def forward_pass(self, x):
        raise NotImplementedError('forward_pass() is implemented in Model sub classes')

    # tf.layers.dense under a doubled scope (modle_base.py:27-48).  NB the reference default
    # `tf.random_normal_initializer(0.02)` sets the MEAN to 0.02 (stddev stays 1.0).
    def _linear_fc(self, input_, output_size, scope=None, bias_start=0.0, use_bias=True,
                   kernel_initializer=random_normal_initializer(0.02)):
        with variable_scope(scope):
            with variable_scope(scope):
                k = get_variable('kernel', [int(input_.shape[-1]), output_size], kernel_initializer)
                z = ops.conv2d(input_, ops.PlainWeight(k), 1, 1)
                if not use_bias:
                    return z
                return ops.lazy_bias(z, get_variable('bias', [output_size], constant_initializer(bias_start)))

    def _WN_dense(self, input_, output_size, scope, init_scale=1.0, init=False):
        """Weight normalization dense layer (modle_base.py:50-73): x @ l2norm(V) * g + b."""
        with variable_scope(scope):
            cin = int(input_.shape[1])
            V = get_variable('V', [cin, output_size], random_normal_initializer(0, 0.05))
            g = get_variable('g', [output_size], constant_initializer(1.))
            b = get_variable('b', [output_size], constant_initializer(0.))
            if init:
                raise NotImplementedError('_WN_dense(init=True): every reference call site passes init=False')
            z = ops.conv2d(input_, ops.WNWeight(V, g, cin, output_size, 1, 1), 1, 1)
            return ops.lazy_bias(z, b)

    def _WN_conv2d(self, input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2, padding='SAME', init_scale=1.0,
                   init=False, name="conv2d"):
        """Weight normalization conv2d layer (modle_base.py:75-108): conv(x, l2norm(V)) * g + b."""
        output_dim = int(output_dim)
        with variable_scope(name):
            cin = int(input_.shape[-1])
            V = get_variable('V', [k_h, k_w, cin, output_dim], random_normal_initializer(0, 0.05))
            g = get_variable('g', [output_dim], constant_initializer(1.))
            b = get_variable('b', [output_dim], constant_initializer(0.))
            if init:
                raise NotImplementedError('_WN_conv2d(init=True): every reference call site passes init=False')
            assert d_h == d_w
            z = ops.conv2d(input_, ops.WNWeight(V, g, k_h * k_w * cin, output_dim, 1, 1), k_h, k_w, d_h, padding)
            return ops.lazy_bias(z, b)

    def _WN_deconv2d(self, input_, output_dim, k_h=3, k_w=3, d_h=2, d_w=2, padding='SAME', init_scale=1.0,
                     init=False, name="deconv2d"):
        """modle_base.py:130-155: conv2d_transpose(x, l2norm(V, [0,1,3])) * g + b; V is [kh,kw,Cout,Cin]."""
        num_filters = int(output_dim)
        if padding != 'SAME':
            raise NotImplementedError('_WN_deconv2d: only SAME padding is used by the Triple-GAN models')
        with variable_scope(name):
            cin = int(input_.shape[-1])
            V = get_variable('V', [k_h, k_w, num_filters, cin], random_normal_initializer(0, 0.05))
            g = get_variable('g', [num_filters], constant_initializer(1.))
            b = get_variable('b', [num_filters], constant_initializer(0.))
            if init:
                raise NotImplementedError('_WN_deconv2d(init=True): every reference call site passes init=False')
            assert d_h == d_w
            z = ops.conv2d_transpose(input_, ops.WNWeight(V, g, k_h * k_w, num_filters, cin, 1), k_h, k_w, d_h)
            return ops.lazy_bias(z, b)

    # tf.layers.conv2d, always padding='same' (modle_base.py:157-168).  NB the reference default
    # `tf.truncated_normal_initializer(0.02)` sets the MEAN to 0.02.
    def _conv2d(self, input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2,
                kernel_initializer=truncated_normal_initializer(0.02), name="conv2d"):
        with variable_scope(name):
            with variable_scope(name):
                k = get_variable('kernel', [k_h, k_w, int(input_.shape[-1]), output_dim], kernel_initializer)
                assert d_h == d_w
                z = ops.conv2d(input_, ops.PlainWeight(k), k_h, k_w, d_h, 'SAME')
                return ops.lazy_bias(z, get_variable('bias', [output_dim], zeros_initializer()))

    def _relu(self, x):
        return nn.relu(x)

    def _leaky_relu(self, x, alpha):
        return ops.activation(x, 'lrelu', alpha)

    def _softplus(self, x):
        return nn.softplus(x)

    def _fully_connected(self, x, out_dim, use_bias=True, name=None):
        with variable_scope(name or 'dense'):
            k = get_variable('kernel', [int(x.shape[-1]), out_dim], random_normal_initializer(0., 0.05))
            z = ops.conv2d(x, ops.PlainWeight(k), 1, 1)
            return ops.lazy_bias(z, get_variable('bias', [out_dim], zeros_initializer())) if use_bias else z

    def _drop_out(self, x, rate=0.5, train=False, tag='dropout'):
        return ops.dropout(x, rate, tag, bool(train))

    def _add_noise(self, inputs, mean=0.0, stddev=0.001, tag='noise'):
        assert mean == 0.0
        return ops.add_noise(inputs, stddev, tag)

    def _nin(self, input, num_units, name):
        """ a network in network layer (1x1 CONV) (modle_base.py:204-209) """
        s = list(input.shape)
        x = ops.reshape(input, [int(np.prod(s[:-1])), s[-1]])
        x = self._WN_dense(x, num_units, name)
        return ops.reshape_pending(x, s[:-1] + [num_units])      # keeps the pending bias fusable through the reshape

    def _batch_norm_contrib(self, x, name, train=False):
        return nn.batch_norm_contrib(x, name, train, self._batch_norm_decay, self._batch_norm_epsilon)

    def _conv_cond_concat(self, x, y):
        """Concatenate conditioning vector on feature map axis (modle_base.py:239-244)."""
        return ops.concat_label(x, y)

    def _deconv2d(self, input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, name="deconv2d", use_bias=True,
                  kernel_initializer=random_normal_initializer(0.02)):
        with variable_scope(name):
            with variable_scope(name):
                k = get_variable('kernel', [k_h, k_w, output_shape, int(input_.shape[-1])], kernel_initializer)
                assert d_h == d_w
                z = ops.conv2d_transpose(input_, ops.PlainWeight(k), k_h, k_w, d_h)
                if not use_bias:
                    return z
                return ops.lazy_bias(z, get_variable('bias', [output_shape], zeros_initializer()))
