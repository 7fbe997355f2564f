"""Operator surface of the reference's Model/nn.py, backed by the sm_100a kernels.

Same function names, positional order, keyword names and defaults as the reference
(signatures at nn.py:147, 192, 220, 255, 292, 335, 404, 427, 441, 456, 469-471, 525-528, 577-578), so
model code written against `Model/nn.py` runs unchanged on `Var`s.  Differences forced by leaving TF:
`deterministic` / `train` are Python booleans (eager), `nonlinearity` is a callable on Var (callables
that carry a `tgan_act = (kind, alpha)` attribute are fused into the layer's epilogue kernel), and the
stochastic layers take an optional `tag=` naming their RNG stream.
"""
import numpy as np

from . import ops
from .core import (Param, constant_initializer, ctx, get_variable, ones_initializer, random_normal_initializer,
                   truncated_normal_initializer, variable_scope, zeros_initializer)


def int_shape(x):
    return list(map(int, x.shape))


def get_name(layer_name, counters):
    ''' utlity for keeping track of layer names (nn.py:138-144) '''
    if layer_name not in counters:
        counters[layer_name] = 0
    name = layer_name + '_' + str(counters[layer_name])
    counters[layer_name] += 1
    return name


# ---- nonlinearities (callables carrying their fused-epilogue code) ----

def _mk_act(kind, alpha=0.2):
    def f(x, *a, **k):
        return ops.activation(x, kind, alpha)
    f.tgan_act = (kind, alpha)
    f.__name__ = kind
    return f


relu = _mk_act('relu')
leaky_relu = _mk_act('lrelu', 0.2)      # tf.nn.leaky_relu default alpha 0.2
tanh = _mk_act('tanh')
sigmoid = _mk_act('sigmoid')
softplus = _mk_act('softplus')


def _fused(nonlinearity):
    """-> (kind, alpha) if the callable is one of ours, else None."""
    if nonlinearity is None:
        return ('none', 0.0)
    return getattr(nonlinearity, 'tgan_act', None)      # bound methods forward attribute reads to their function


def _tmp_param(value):
    """A constant fp32 vector that is not a variable (used by the init=True data-dependent branches)."""
    import torch
    p = Param('_tmp', np.shape(value), False, np.asarray(value, np.float32))
    if not ctx.building:
        p.data = torch.from_numpy(p.init_value).to(ctx.device)
    return p


def _finish(z, b, nonlinearity):
    fa = _fused(nonlinearity)
    if fa is not None:
        return ops.bias_act(z, b, fa[0], fa[1])
    return nonlinearity(ops.bias_act(z, b, 'none'))


def get_var_maybe_avg(param, ema):
    ''' utility for retrieving polyak averaged params (nn.py:98-110) '''
    return param if ema is None else ema.average(param)


# ---- mean-only batch norm / batch norm (nn.py:147-217) ----

def mean_only_batch_norm_impl(x, pop_mean, b, is_conv_out=True, deterministic=False, decay=0.9,
                              name='meanOnlyBatchNormalization'):
    '''input comes in which is t=(g*V/||V||)*x ; deterministic separates training and testing phases'''
    return ops.mobn_act(x, b, pop_mean, not deterministic, 'none', 0.0, decay)


def batch_norm_impl(x, is_conv_out=True, deterministic=False, decay=0.9, name='BatchNormalization'):
    with variable_scope(name):
        C = x.shape[-1]
        scale = get_variable('scale', [C], ones_initializer(), trainable=True)
        beta = get_variable('beta', [C], zeros_initializer(), trainable=True)
        pop_mean = get_variable('pop_mean', [C], zeros_initializer(), trainable=False)
        pop_var = get_variable('pop_var', [C], ones_initializer(), trainable=False)
        return ops.batch_norm(x, scale, beta, pop_mean, pop_var, not deterministic, eps=0.001, decay=decay,
                              unbiased=False)


# ---- Salimans & Kingma weight-normalised layers (nn.py:220-340) ----

def _data_init(z, init_scale, eps, nonlinearity):
    """x_init = scale_init * (x - m_init), scale_init = init_scale / sqrt(var + eps)."""
    C = z.shape[-1]
    y = ops.batch_norm(z, _tmp_param(np.full(C, init_scale)), _tmp_param(np.zeros(C)), None, None, True, eps=eps)
    return nonlinearity(y) if nonlinearity is not None else y


def dense(x, num_units, nonlinearity=None, init_scale=1., counters={}, init=False, ema=None, train_scale=True,
          init_w=random_normal_initializer(0, 0.05), **kwargs):
    ''' fully connected layer '''
    name = get_name('dense', counters)
    with variable_scope(name):
        V = get_variable('V', [int(x.shape[1]), num_units], init_w, trainable=True)
        g = get_variable('g', [num_units], ones_initializer(), trainable=train_scale)
        b = get_variable('b', [num_units], zeros_initializer(), trainable=True)
        if init:
            z = ops.conv2d(x, ops.WNWeight(V, _tmp_param(np.ones(num_units)), V.shape[0], num_units, 1, 1), 1, 1)
            return _data_init(z, init_scale, 1e-10, nonlinearity)
        V, g, b = (get_var_maybe_avg(p, ema) for p in (V, g, b))
        z = ops.conv2d(x, ops.WNWeight(V, g, V.shape[0], num_units, 1, 0), 1, 1)
        return _finish(z, b, nonlinearity)


def conv2d(x, num_filters, filter_size=[3, 3], stride=[1, 1], pad='SAME', nonlinearity=None, init_scale=1.,
           counters={}, init=False, ema=None, **kwargs):
    ''' convolutional layer '''
    name = get_name('conv2d', counters)
    kh, kw = filter_size
    assert stride[0] == stride[1]
    with variable_scope(name):
        cin = int(x.shape[-1])
        V = get_variable('V', [kh, kw, cin, num_filters], random_normal_initializer(0, 0.05), trainable=True)
        g = get_variable('g', [num_filters], ones_initializer(), trainable=True)
        b = get_variable('b', [num_filters], zeros_initializer(), trainable=True)
        A = kh * kw * cin
        if init:
            z = ops.conv2d(x, ops.WNWeight(V, _tmp_param(np.ones(num_filters)), A, num_filters, 1, 1), kh, kw,
                           stride[0], pad)
            return _data_init(z, init_scale, 1e-8, nonlinearity)
        V, g, b = (get_var_maybe_avg(p, ema) for p in (V, g, b))
        z = ops.conv2d(x, ops.WNWeight(V, g, A, num_filters, 1, 1), kh, kw, stride[0], pad)
        return _finish(z, b, nonlinearity)


def deconv2d(x, num_filters, filter_size=[3, 3], stride=[1, 1], pad='SAME', nonlinearity=None, init_scale=1.,
             counters={}, init=False, ema=None, **kwargs):
    ''' transposed convolutional layer '''
    name = get_name('deconv2d', counters)
    kh, kw = filter_size
    assert stride[0] == stride[1]
    if pad != 'SAME':
        raise NotImplementedError('deconv2d: only SAME padding is used by the Triple-GAN models')
    with variable_scope(name):
        cin = int(x.shape[-1])
        V = get_variable('V', [kh, kw, num_filters, cin], random_normal_initializer(0, 0.05), trainable=True)
        g = get_variable('g', [num_filters], ones_initializer(), trainable=True)
        b = get_variable('b', [num_filters], zeros_initializer(), trainable=True)
        if init:
            z = ops.conv2d_transpose(x, ops.WNWeight(V, _tmp_param(np.ones(num_filters)), kh * kw, num_filters, cin,
                                                     1), kh, kw, stride[0])
            return _data_init(z, init_scale, 1e-8, nonlinearity)
        V, g, b = (get_var_maybe_avg(p, ema) for p in (V, g, b))
        z = ops.conv2d_transpose(x, ops.WNWeight(V, g, kh * kw, num_filters, cin, 1), kh, kw, stride[0])
        return _finish(z, b, nonlinearity)


def nin(x, num_units, **kwargs):
    """ a network in network layer (1x1 CONV) """
    s = int_shape(x)
    x = ops.reshape(x, [int(np.prod(s[:-1])), s[-1]])
    x = dense(x, num_units, **kwargs)
    return ops.reshape(x, s[:-1] + [num_units])


# ---- thin tf.layers wrappers (nn.py:404-464) ----

def _linear_fc(input_, output_size, scope=None, stddev=0.02, bias_start=0.0, use_bias=True):
    with variable_scope(scope):
        with variable_scope(scope):
            k = get_variable('kernel', [int(input_.shape[-1]), output_size], random_normal_initializer(stddev=stddev))
            z = ops.conv2d(input_, ops.PlainWeight(k), 1, 1)
            if not use_bias:
                return z
            return ops.lazy_bias(z, get_variable('bias', [output_size], constant_initializer(bias_start)))


def _conv2d(input_, output_dim, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="conv2d"):
    with variable_scope(name):
        with variable_scope(name):
            k = get_variable('kernel', [k_h, k_w, int(input_.shape[-1]), output_dim],
                             truncated_normal_initializer(stddev=stddev))
            z = ops.conv2d(input_, ops.PlainWeight(k), k_h, k_w, d_h, 'SAME')
            return ops.lazy_bias(z, get_variable('bias', [output_dim], zeros_initializer()))


def _deconv2d(input_, output_shape, k_h=5, k_w=5, d_h=2, d_w=2, stddev=0.02, name="deconv2d", use_bias=True):
    with variable_scope(name):
        with variable_scope(name):
            k = get_variable('kernel', [k_h, k_w, output_shape, int(input_.shape[-1])],
                             random_normal_initializer(stddev=stddev))
            z = ops.conv2d_transpose(input_, ops.PlainWeight(k), k_h, k_w, d_h)
            if not use_bias:
                return z
            return ops.lazy_bias(z, get_variable('bias', [output_shape], zeros_initializer()))


def batch_norm_contrib(x, name, train=False, decay=0.9, epsilon=1e-5):
    with variable_scope(name):
        C = x.shape[-1]
        beta = get_variable('beta', [C], zeros_initializer())
        gamma = get_variable('gamma', [C], ones_initializer())
        mm = get_variable('moving_mean', [C], zeros_initializer(), trainable=False)
        mv = get_variable('moving_variance', [C], ones_initializer(), trainable=False)
        return ops.batch_norm(x, gamma, beta, mm, mv, train, eps=epsilon, decay=decay)


# ---- zoli333-style WN / mean-only-BN layers used by the CIFAR-10 classifier (nn.py:469-589) ----

def conv2d_WN(x, num_filters, filter_size=[3, 3], pad='SAME', stride=[1, 1], nonlinearity=None, init_scale=1.,
              init=False, use_weight_normalization=False, use_batch_normalization=False,
              use_mean_only_batch_normalization=False, deterministic=False, name=''):
    '''deterministic : used for batch normalizations (separates the training and testing phases)'''
    kh, kw = filter_size
    assert stride[0] == stride[1]
    with variable_scope(name):
        cin = int(x.shape[-1])
        V = get_variable('V', [kh, kw, cin, num_filters], random_normal_initializer(0, 0.05), trainable=True)
        b = pop_mean = g = None
        if use_batch_normalization is False:
            b = get_variable('b', [num_filters], constant_initializer(0.), trainable=True)
        if use_mean_only_batch_normalization:
            pop_mean = get_variable('meanOnlyBatchNormalization/pop_mean', [num_filters], zeros_initializer(),
                                    trainable=False)
        A = kh * kw * cin
        if use_weight_normalization:
            g = get_variable('g', [num_filters], constant_initializer(1.), trainable=True)
            if init:
                # data-dependent init output; the g/b assigns of nn.py:497-499 are built but never run
                z = ops.conv2d(x, ops.WNWeight(V, _tmp_param(np.ones(num_filters)), A, num_filters, 1, 1), kh, kw,
                               stride[0], pad)
                return _data_init(z, init_scale, 1e-08, nonlinearity)
            if use_mean_only_batch_normalization:
                fa = _fused(nonlinearity)
                if fa is not None:      # conv + mean-only BN + nonlinearity as one unit (ops.conv2d_mobn)
                    return ops.conv2d_mobn(x, ops.WNWeight(V, g, A, num_filters, 1, 1), kh, kw, stride[0], pad, b, pop_mean,
                                           not deterministic, fa[0], fa[1])
                z = ops.conv2d(x, ops.WNWeight(V, g, A, num_filters, 1, 1), kh, kw, stride[0], pad, colsum=True)
                return nonlinearity(ops.mobn_act(z, b, pop_mean, not deterministic))
            z = ops.conv2d(x, ops.WNWeight(V, g, A, num_filters, 1, 1), kh, kw, stride[0], pad)
            return _finish(z, b, nonlinearity)
        z = ops.conv2d(x, ops.PlainWeight(V), kh, kw, stride[0], pad)
        if use_batch_normalization:
            y = batch_norm_impl(z, is_conv_out=True, deterministic=deterministic)
            return nonlinearity(y) if nonlinearity is not None else y
        return _finish(z, b, nonlinearity)


def dense_WN(x, num_units, nonlinearity=None, init_scale=1., init=False, use_weight_normalization=False,
             use_batch_normalization=False, use_mean_only_batch_normalization=False, deterministic=False, name=''):
    with variable_scope(name):
        cin = int(x.shape[1])
        V = get_variable('V', [cin, num_units], random_normal_initializer(0, 0.05), trainable=True)
        b = pop_mean = None
        if use_batch_normalization is False:
            b = get_variable('b', [num_units], constant_initializer(0.), trainable=True)
        if use_mean_only_batch_normalization:
            pop_mean = get_variable('meanOnlyBatchNormalization/pop_mean', [num_units], zeros_initializer(),
                                    trainable=False)
        if use_weight_normalization:
            g = get_variable('g', [num_units], constant_initializer(1.), trainable=True)
            if init:
                z = ops.conv2d(x, ops.WNWeight(V, _tmp_param(np.ones(num_units)), cin, num_units, 1, 1), 1, 1)
                return _data_init(z, init_scale, 1e-10, nonlinearity)
            # (x @ V) * g/sqrt(sum V^2): the same function of (V, g) as x @ (g V/||V||), no epsilon (nn.py:553-555)
            z = ops.conv2d(x, ops.WNWeight(V, g, cin, num_units, 1, 0), 1, 1,
                           colsum=bool(use_mean_only_batch_normalization))
            if use_mean_only_batch_normalization:
                fa = _fused(nonlinearity)
                if fa is not None:
                    return ops.mobn_act(z, b, pop_mean, not deterministic, fa[0], fa[1])
                return nonlinearity(ops.mobn_act(z, b, pop_mean, not deterministic))
            return _finish(z, b, nonlinearity)
        z = ops.conv2d(x, ops.PlainWeight(V), 1, 1)
        if use_batch_normalization:
            y = batch_norm_impl(z, is_conv_out=False, deterministic=deterministic)
            return nonlinearity(y) if nonlinearity is not None else y
        return _finish(z, b, nonlinearity)


def NiN_WN(x, num_units, nonlinearity=None, init=False, use_weight_normalization=False,
           use_batch_normalization=False, use_mean_only_batch_normalization=False, deterministic=False, name=''):
    """ a network in network layer (1x1 CONV) """
    with variable_scope(name):
        s = int_shape(x)
        x = ops.reshape(x, [int(np.prod(s[:-1])), s[-1]])
        x = dense_WN(x, num_units=num_units, nonlinearity=nonlinearity, init=init,
                     use_weight_normalization=use_weight_normalization,
                     use_batch_normalization=use_batch_normalization,
                     use_mean_only_batch_normalization=use_mean_only_batch_normalization,
                     deterministic=deterministic, name=name)
        return ops.reshape(x, s[:-1] + [num_units])
