"""Execution context, variables, and the reverse-mode tape of the B200 Triple-GAN path.

This replaces what the reference gets from the TF-1.x runtime:
  * tf.get_variable / tf.variable_scope(reuse=...)      -> VariableStore / variable_scope
    (names are the TF variable names: they are the parameter keys)
  * tf.gradients (optimizer.minimize, train_base.py:64-68) -> Tape (closures pushed by every op)
  * graph construction without execution                -> `building()` mode (shape-only, no kernels)
PyTorch is used for device memory and streams only.
"""
import contextlib
import math

import numpy as np
import torch

F32, BF16 = 0, 1


def dt_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('unsupported dtype %s' % t.dtype)


class Context:
    """Process-wide execution state."""

    def __init__(self):
        self.device = None
        self.math = 'fp32'         # 'fp32' (CUDA-core parity mode) | 'bf16' (tcgen05 tensor-core mode)
        self.building = False      # shape-only graph construction (no GPU needed)
        self.tape = None
        self.store = None
        self.rng = None
        self._ws = None
        self._ws_side = None
        self.side = None            # second stream: independent small-kernel chains run beside the big GEMMs
        self.on_side = False
        self.ws_floats = 64 * 1024 * 1024
        self._arena = None          # zero-initialised fp32 scratch for epilogue-accumulated statistics
        self.arena_floats = 256 * 1024
        self.arena_off = 0

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.math == 'bf16' else torch.float32

    def arena(self):
        if self._arena is None:
            self._arena = torch.zeros(self.arena_floats, dtype=torch.float32, device=self.device)
            self.arena_off = 0
        return self._arena

    def ws(self):
        """scratch workspace of the CURRENT stream (kernels on the side stream must not share partial sums / ticket
        counters with kernels on the main stream)"""
        if self.on_side:
            if self._ws_side is None:
                self._ws_side = torch.empty(self.ws_floats // 4, dtype=torch.float32, device=self.device)
            return self._ws_side
        if self._ws is None:
            self._ws = torch.empty(self.ws_floats, dtype=torch.float32, device=self.device)
        return self._ws

    def side_stream(self):
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.device)
        return self.side


ctx = Context()


def init(device='cuda:0', math='fp32', seed=1234):
    """Bind the path to a CUDA device.  Fails loudly without one (there is no CPU fallback)."""
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError('tgan.init: no CUDA device -- the Triple-GAN B200 path has no CPU fallback')
    _lib.load()
    if math not in ('fp32', 'bf16'):
        raise ValueError('math must be fp32 or bf16')
    ctx.device = torch.device(device)
    torch.cuda.set_device(ctx.device)
    ctx.math = math
    ctx._ws = None
    ctx._ws_side = None
    ctx.side = None
    ctx.on_side = False
    ctx._arena = None
    ctx.rng = PhiloxSource(seed)
    return ctx


@contextlib.contextmanager
def building():
    """Shape-only execution: variables are created, no kernel is launched (TF graph construction)."""
    old = ctx.building
    ctx.building = True
    try:
        yield
    finally:
        ctx.building = old


# ----------------------------------------------------------------------------------------------
# RNG sources
# ----------------------------------------------------------------------------------------------


class PhiloxSource:
    """Perf mode: noise / dropout masks are generated in-kernel (Philox4x32-10).  Every stochastic
    call site (tag) owns a stream id; a device-side counter advances once per step so CUDA-graph
    replays draw fresh numbers."""
    injected = False

    def __init__(self, seed):
        self.seed = int(seed)
        self.streams = {}
        self._counter = None

    def stream_id(self, tag):
        if tag not in self.streams:
            self.streams[tag] = len(self.streams) + 1
        return self.streams[tag]

    def counter(self):
        if self._counter is None:
            self._counter = torch.zeros(1, dtype=torch.int64, device=ctx.device)
        return self._counter


class InjectedSource:
    """Parity mode: wraps a provider with .normal(tag, shape) / .keep_mask(tag, shape, rate) returning
    CPU tensors (the oracle's TagRNG) and uploads them, so both sides see identical draws."""
    injected = True

    def __init__(self, provider):
        self.p = provider

    def normal(self, tag, shape):
        return self.p.normal(tag, shape).to(ctx.device, torch.float32).contiguous()

    def keep_mask(self, tag, shape, rate):
        return self.p.keep_mask(tag, shape, rate).to(ctx.device, torch.uint8).contiguous()


# ----------------------------------------------------------------------------------------------
# Var + Tape
# ----------------------------------------------------------------------------------------------


class Var:
    """An activation.  `data` is a contiguous device tensor [..., ld]; the logical shape is `shape`
    (last dim C <= ld; ld > C only for label-concatenated tensors padded for 16-byte TMA strides)."""
    __slots__ = ('_data', 'shape', 'ld', 'grad', 'requires_grad', '_lazy', 'tag', 'aux')

    def __init__(self, data, shape=None, ld=None, requires_grad=False):
        self._data = data
        self.shape = tuple(shape if shape is not None else data.shape)
        self.ld = int(ld if ld is not None else self.shape[-1])
        self.grad = None
        self.requires_grad = requires_grad
        self._lazy = None          # pending fused epilogue: (z Var, bias Param)
        self.tag = None
        self.aux = None            # e.g. {'colsum': per-channel sums accumulated by the producing GEMM epilogue}

    @property
    def data(self):
        if self._data is None and self._lazy is not None:
            from . import ops
            ops._materialize(self)
        return self._data

    @property
    def C(self):
        return self.shape[-1]

    @property
    def rows(self):
        return int(np.prod(self.shape[:-1]))

    def numpy(self):
        d = self.data
        if self.ld != self.C:
            d = d[..., :self.C]
        return d.float().cpu().numpy().reshape(self.shape)

    def __repr__(self):
        return 'Var(shape=%s, ld=%d, lazy=%s)' % (self.shape, self.ld, self._lazy is not None)


class Tape:
    def __init__(self):
        self.nodes = []
        self.post = []      # weight-norm backward hooks run once after all nodes

    def backward(self):
        """Run the recorded closures in reverse, then the once-per-phase weight-norm backward hooks.  The tape is
        made current for the duration, so it can be differentiated outside the `recording()` block that built it."""
        old, ctx.tape = ctx.tape, self
        try:
            from . import ops
            for fn in reversed(self.nodes):
                fn()
            ops.join_side_if_pending()      # filter gradients parked on the side stream (ops.defer_join)
            for fn in self.post:
                fn()
        finally:
            ctx.tape = old
        self.nodes, self.post = [], []


@contextlib.contextmanager
def recording():
    old, ctx.tape = ctx.tape, Tape()
    try:
        yield ctx.tape
    finally:
        ctx.tape = old


@contextlib.contextmanager
def no_grad():
    old, ctx.tape = ctx.tape, None
    try:
        yield
    finally:
        ctx.tape = old


def add_grad(v, g):
    """Accumulate gradient tensor g into Var v."""
    from . import ops
    if v.grad is not None and v.aux and v.aux.get('du_ready') and not getattr(g, '_tgan_du', False):
        # (a consumer already delivered dy * lrelu'(y) for this tensor; a plain dy must not be mixed into it)
        raise RuntimeError('add_grad: a plain gradient meets a leaky-ReLU-fused one on the same tensor')
    if v.grad is None:
        v.grad = g
    else:
        ops.accumulate_(v.grad, g)


# ----------------------------------------------------------------------------------------------
# Variables
# ----------------------------------------------------------------------------------------------


class Param:
    """A TF variable: fp32 tensor + gradient (views into the per-network flat buffers once the
    store is finalized)."""
    __slots__ = ('name', 'shape', 'trainable', 'data', 'grad', 'requires_grad', 'init_value', 'group', 'cache')

    def __init__(self, name, shape, trainable, init_value):
        self.name, self.shape, self.trainable = name, tuple(shape), trainable
        self.init_value = init_value      # numpy float32
        self.data = None
        self.grad = None
        self.requires_grad = False
        self.group = name.split('/')[0]
        self.cache = {}

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1


class VariableStore:
    """Name -> Param with TF variable_scope(reuse) semantics; flat per-network fp32 buffers
    (theta / grad / Adam m,v) so one NCCL all-reduce and one Adam launch serve a whole network."""
    GROUPS = ('good_generator', 'discriminator', 'classifier')

    def __init__(self, seed=1234):
        self.vars = {}
        self.order = []
        self.scope = []
        self.reuse = [False]
        self.np_rng = np.random.default_rng(seed)
        self.flat = {}          # group -> dict(theta, grad, m, v, ema, params)
        self.version = 0        # bumped on every optimiser step (invalidates weight-norm caches)
        self.versions = {}      # per network: only the updated network's cached weights are rebuilt
        self.finalized = False

    # -- scopes --
    @contextlib.contextmanager
    def variable_scope(self, name, reuse=None):
        self.scope.append(name)
        self.reuse.append(self.reuse[-1] if reuse is None else bool(reuse))
        try:
            yield
        finally:
            self.scope.pop()
            self.reuse.pop()

    def full_name(self, name):
        return '/'.join(self.scope + [name])

    def get_variable(self, name, shape, initializer, trainable=True):
        full = self.full_name(name)
        if full in self.vars:
            if not self.reuse[-1]:
                raise ValueError('Variable %s already exists, disallowed. Did you mean to set reuse=True?' % full)
            p = self.vars[full]
            if tuple(shape) != p.shape:
                raise ValueError('Trying to share variable %s, but specified shape %s and found shape %s.'
                                 % (full, tuple(shape), p.shape))
            return p
        if self.reuse[-1]:
            raise ValueError('Variable %s does not exist, or was not created with tf.get_variable().' % full)
        if self.finalized:
            raise RuntimeError('cannot create variable %s after VariableStore.finalize()' % full)
        val = np.asarray(initializer(self.np_rng, tuple(shape)), dtype=np.float32).reshape(shape)
        p = Param(full, shape, trainable, val)
        self.vars[full] = p
        self.order.append(full)
        return p

    def has(self, full):
        return full in self.vars

    def trainable(self, substr=None):
        return [self.vars[n] for n in self.order if self.vars[n].trainable and (substr is None or substr in n)]

    # -- materialisation --
    def load_numpy(self, P, S=None):
        """Inject initial values keyed by TF variable names (oracle init / checkpoint import)."""
        for src in (P, S or {}):
            for k, v in src.items():
                if k not in self.vars:
                    raise KeyError('unknown variable %s' % k)
                if tuple(np.shape(v)) != self.vars[k].shape:
                    raise ValueError('shape mismatch for %s' % k)
                self.vars[k].init_value = np.asarray(v, np.float32)
                if self.vars[k].data is not None:
                    self.vars[k].data.copy_(torch.from_numpy(self.vars[k].init_value))

    def finalize(self, device):
        """Allocate the flat buffers and bind every Param to a 16-byte aligned view."""
        if self.finalized:
            return
        for grp in self.GROUPS + ('_state',):
            ps = [self.vars[n] for n in self.order
                  if (self.vars[n].trainable and self.vars[n].group == grp)
                  or (grp == '_state' and not self.vars[n].trainable)]
            if not ps:
                continue
            offs, tot = [], 0
            for p in ps:
                offs.append(tot)
                tot += (p.size + 3) // 4 * 4
            host = np.zeros(tot, np.float32)
            for p, o in zip(ps, offs):
                host[o:o + p.size] = p.init_value.reshape(-1)
            # data-parallel runs place parameters and gradients in symmetric (peer-addressable) memory: ddp.FusedUpdate
            alloc = self.alloc if (getattr(self, 'alloc', None) is not None and grp != '_state') else None
            if alloc is not None:
                theta = alloc(tot)
                theta.copy_(torch.from_numpy(host))
            else:
                theta = torch.from_numpy(host).to(device)
            fb = dict(theta=theta, params=ps, offsets=offs, n=tot)
            if grp != '_state':
                if alloc is not None:
                    fb['grad'] = alloc(tot)
                    fb['grad'].zero_()
                else:
                    fb['grad'] = torch.zeros_like(theta)
                fb['m'] = torch.zeros_like(theta)
                fb['v'] = torch.zeros_like(theta)
            for p, o in zip(ps, offs):
                p.data = theta[o:o + p.size].view(p.shape)
                if grp != '_state':
                    p.grad = fb['grad'][o:o + p.size].view(p.shape)
            self.flat[grp] = fb
        self.finalized = True

    def to_numpy(self):
        return {n: self.vars[n].data.detach().cpu().numpy().copy() for n in self.order}

    def bump(self, group=None):
        self.version += 1
        if group is None:
            for k in list(self.versions) + list(self.GROUPS):
                self.versions[k] = self.versions.get(k, 0) + 1
        else:
            self.versions[group] = self.versions.get(group, 0) + 1

    def group_version(self, group):
        return self.versions.get(group, 0)


def variable_scope(name, reuse=None):
    return ctx.store.variable_scope(name, reuse)


def get_variable(name, shape, initializer, trainable=True):
    return ctx.store.get_variable(name, shape, initializer, trainable)


# ----------------------------------------------------------------------------------------------
# initializers: f(np_rng, shape) -> ndarray   (mirror the tf initializers the reference passes)
# ----------------------------------------------------------------------------------------------


def random_normal_initializer(mean=0.0, stddev=1.0):
    return lambda rng, shape: mean + stddev * rng.standard_normal(shape)


def truncated_normal_initializer(mean=0.0, stddev=1.0):
    def f(rng, shape):
        x = rng.standard_normal(shape)
        bad = np.abs(x) > 2
        while bad.any():
            x[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(x) > 2
        return mean + stddev * x
    return f


def constant_initializer(v=0.0):
    return lambda rng, shape: np.full(shape, v, np.float32)


def zeros_initializer():
    return constant_initializer(0.0)


def ones_initializer():
    return constant_initializer(1.0)


def variance_scaling_initializer(factor=2.0):
    """tf.contrib.layers.variance_scaling_initializer() (FAN_IN, normal): the he_init of
    Good_GAN_cifar10.py:8-9.  fan_in = prod(shape[:-1])."""
    def f(rng, shape):
        fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else shape[0]
        return rng.standard_normal(shape) * math.sqrt(factor / max(fan_in, 1))
    return f
