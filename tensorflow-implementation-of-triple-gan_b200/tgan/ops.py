"""Differentiable ops of the Triple-GAN path on `Var`s, each a thin orchestration of C-ABI calls
(include/tgan.h).  Forward launches the sm_100a kernels and pushes a backward closure on the tape.
No torch compute op is used on the data path (torch only allocates memory / provides the stream).
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .core import BF16, F32, Var, add_grad, ctx, dt_code

ACT = {'none': 0, 'relu': 1, 'lrelu': 2, 'tanh': 3, 'sigmoid': 4, 'softplus': 5}
STATS_PARTS = 256


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _new(shape, dtype=None):
    return torch.empty(tuple(int(s) for s in shape), dtype=dtype or ctx.act_dtype, device=ctx.device)


def _zeros(shape, dtype=torch.float32):
    t = _new(shape, dtype)
    if dtype == torch.float32:
        _lib.call('tgan_fill_f32', _p(t), 0.0, t.numel(), _st())
    else:
        t.zero_()
    return t


def arena_reset():
    """Zero the statistics arena (one launch) -- called at the start of every phase of the step."""
    a = ctx.arena()
    _lib.call('tgan_fill_f32', _p(a), 0.0, a.numel(), _st())
    ctx.arena_off = 0


def arena_take(n):
    """n zero-initialised floats that a GEMM epilogue may atomically accumulate into."""
    a = ctx.arena()
    n4 = (n + 3) // 4 * 4
    if ctx.arena_off + n4 > a.numel():
        # a reset here would zero sums that kernels already in flight still accumulate into / read
        raise RuntimeError('statistics arena exhausted (%d floats): raise Context.arena_floats or call ops.arena_reset() '
                           'between passes' % a.numel())
    o = ctx.arena_off
    ctx.arena_off += n4
    return a[o:o + n]


class side_stream:
    """`with ops.side_stream(): ...` enqueues the enclosed launches on the context's second stream, ordered after
    everything already enqueued on the current stream (fork); `ops.join_side()` makes the current stream wait for them.
    Under CUDA-graph capture the two become parallel branches of the graph.  Used for work that is independent of what
    the main stream does meanwhile: the generator forward beside the classifier forward of phase D, filter gradients
    beside input gradients of small layers."""

    def __enter__(self):
        self._cm = None
        if os.environ.get('TGAN_NO_SIDE'):      # experiments / race hunting: everything on one stream
            return self
        main = torch.cuda.current_stream()
        side = ctx.side_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        self._cm = torch.cuda.stream(side)
        self._cm.__enter__()
        ctx.on_side = True
        return self

    def __exit__(self, *a):
        if self._cm is None:
            return
        ctx.on_side = False
        self._cm.__exit__(*a)


def join_side():
    ctx.side_keep, ctx.side_dirty = [], False
    if ctx.side is None:
        return
    ev = torch.cuda.Event()
    ev.record(ctx.side)
    torch.cuda.current_stream().wait_event(ev)


def defer_join(keep):
    """A filter gradient was enqueued on the side stream; nothing on the main stream needs its result (dW) before the
    backward pass ends, so the join is postponed to Tape.backward's tail (the side stream's filter gradients then
    overlap the whole chain of input gradients instead of one layer each).  `keep`: every tensor the side-stream
    kernels still read -- held alive until the join so the caching allocator cannot hand their memory to the main
    stream meanwhile."""
    if os.environ.get('TGAN_EAGER_JOIN'):
        join_side()
        return
    ctx.side_keep = getattr(ctx, 'side_keep', []) + [keep]
    ctx.side_dirty = True


def join_side_if_pending():
    if getattr(ctx, 'side_dirty', False):
        join_side()


class TagList:
    """RNG tags of several network calls grouped into one batch: [(tag, samples), ...].  `tags + '/noise'` appends
    the suffix to every tag, so builder code written for one call works unchanged on a grouped batch."""

    def __init__(self, items):
        self.items = list(items)

    def __add__(self, suffix):
        return TagList([(t + suffix, n) for t, n in self.items])

    def __str__(self):
        return '|'.join(t for t, _ in self.items)


def _segs(x):
    return x.aux.get('segs') if x.aux else None


def _prop(out, x):
    """carry the batch segmentation (samples per grouped network call) from x to out"""
    sg = _segs(x)
    if sg is not None:
        out.aux = dict(out.aux or {}, segs=sg)
    return out


def group_batch(vs, tags=None):
    """Concatenate the inputs of several calls of the same network along the batch axis.  Per-sample work
    (convolutions, pooling, dropout, noise) then runs once on the whole group, while batch statistics stay per call
    (the segments).  No gradient flows to the inputs."""
    shape = (sum(v.shape[0] for v in vs),) + tuple(vs[0].shape[1:])
    if ctx.building:
        out = Var(None, shape)
    else:
        t = torch.empty(shape, dtype=vs[0].data.dtype, device=ctx.device)
        per = int(np.prod(shape[1:]))
        for v in vs:      # same elements per sample in any view ([N,28,28,1] vs [N,784]); dtype converted on the fly
            assert v.ld == v.C and int(np.prod(v.shape[1:])) == per
        o = 0
        for i in range(0, len(vs), 8):      # ONE launch per 8 sources
            part = vs[i:i + 8]
            n = len(part)
            srcs = (ctypes.c_void_p * n)(*[_p(v.data) for v in part])
            dts = (ctypes.c_int * n)(*[dt_code(v.data) for v in part])
            cnt = (ctypes.c_int64 * n)(*[v.shape[0] * per for v in part])
            _lib.call('tgan_gather_rows', srcs, dts, cnt, n, t.data_ptr() + o * t.element_size(), dt_code(t), _st())
            o += sum(v.shape[0] * per for v in part)
        out = Var(t, shape)
    out.aux = {'segs': [v.shape[0] for v in vs]}
    return out


def _rng_normal(tag, shape):
    rng = ctx.rng
    if isinstance(tag, TagList):
        return torch.cat([rng.normal(t, (n,) + tuple(shape[1:])) for t, n in tag.items], 0)
    return rng.normal(tag, shape)


def _rng_mask(tag, shape, rate):
    rng = ctx.rng
    if isinstance(tag, TagList):
        return torch.cat([rng.keep_mask(t, (n,) + tuple(shape[1:]), rate) for t, n in tag.items], 0)
    return rng.keep_mask(tag, shape, rate)


def _on():
    return ctx.tape is not None


def _out_dtype(C):
    """Skinny outputs (logits, RGB images) stay fp32 even in bf16 mode: they feed the loss kernels /
    leave the network."""
    return torch.float32 if C < 16 else ctx.act_dtype


def same_pad(n, k, s):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return out, tot // 2


def accumulate_(y, x):
    assert y.dtype == x.dtype and y.numel() == x.numel()
    _lib.call('tgan_accumulate', _p(y), _p(x), dt_code(y), y.numel(), _st())


def _to_f32(t, rows, C, ld):
    """[rows, ld] (any dtype, logical C) -> contiguous fp32 [rows, C]"""
    if t.dtype == torch.float32 and ld == C:
        return t
    o = _new((rows, C), torch.float32)
    _lib.call('tgan_copy_channels', _p(t), dt_code(t), ld, _p(o), F32, C, rows, C, _st())
    return o


def _cast(t, dtype):
    if t.dtype == dtype:
        return t
    o = _new(t.shape, dtype)
    n = t.numel()
    _lib.call('tgan_copy_channels', _p(t), dt_code(t), n, _p(o), dt_code(o), n, 1, n, _st())
    return o


# ----------------------------------------------------------------------------------------------
# weights: plain variables and weight-normalised (V, g) pairs
# ----------------------------------------------------------------------------------------------


class PlainWeight:
    """A tf.layers kernel used as is (modle_base.py:40,161,250)."""

    def __init__(self, param):
        self.param = param
        self.scale = None

    @property
    def requires_grad(self):
        return self.param.requires_grad

    @property
    def key(self):
        return self.param

    def value(self):
        return self.param.data

    def grad_target(self):
        return self.param.grad


class WNWeight:
    """W = g * V/||V|| per output channel (nn.py:502,554; modle_base.py:66,101,148).  V is viewed as
    [A, Co, B].  The effective weight is computed once per optimiser version and shared by every call
    in a phase; dW accumulates over the calls and ONE weight-norm backward runs per phase."""

    def __init__(self, V, g, A, Co, B, eps_mode):
        self.V, self.g, self.A, self.Co, self.B, self.eps_mode = V, g, A, Co, B, eps_mode

    @property
    def requires_grad(self):
        return self.V.requires_grad or self.g.requires_grad

    @property
    def key(self):
        return self.V

    def _c(self):
        c = self.V.cache
        ver = ctx.store.group_version(self.V.group)
        if c.get('version') != ver:
            c.clear()
            c['version'] = ver
            c['W'] = _new(self.V.shape, torch.float32)
            c['inv'] = _new((self.Co,), torch.float32)
            c['scale'] = _new((self.Co,), torch.float32)
            _lib.call('tgan_weightnorm_fwd', _p(self.V.data), _p(self.g.data), _p(c['W']), _p(c['inv']),
                      _p(c['scale']), self.A, self.Co, self.B, self.eps_mode, _p(ctx.ws()), _st())
        return c

    def value(self):
        """fp32 effective weight (CUDA-core GEMM paths; the tensor-core path folds the scale while packing)"""
        return self._c()['W']

    def grad_target(self):
        if ctx.math == 'bf16':      # one zero-fill / one weight-norm backward launch for the whole network (prep.py)
            from .prep import WNGroup
            return WNGroup.of(self.V.group).grad_target(self, ctx.tape)
        c, tape = self._c(), ctx.tape
        if c.get('tape') is not tape:
            c['tape'] = tape
            c['dW'] = _zeros(self.V.shape)
            V, g, me = self.V, self.g, self

            def post():
                _lib.call('tgan_weightnorm_bwd', _p(V.data), _p(g.data), _p(c['inv']), _p(c['dW']), _p(V.grad),
                          _p(g.grad), me.A, me.Co, me.B, 1.0, _p(ctx.ws()), _st())
                c['tape'] = None
            tape.post.append(post)
        return c['dW']


# ----------------------------------------------------------------------------------------------
# dense contractions.  fp32 mode / skinny layers: im2col + SIMT GEMM.  bf16 mode: tcgen05 (tc.py).
# ----------------------------------------------------------------------------------------------


def _sgemm(ta, tb, M, N, K, A, lda, B, ldb, C, ldc, beta=0.0, alpha=1.0):
    tiles = -(-M // 64) * -(-N // 64)
    splits = 1
    if tiles < 148 and K >= 2048:
        splits = max(1, min(64, (2 * 148) // tiles, K // 512))
    ws = ctx.ws()
    if splits > 1 and splits * M * N > ws.numel():
        splits = max(1, ws.numel() // (M * N))
    _lib.call('tgan_sgemm', ta, tb, M, N, K, alpha, _p(A), lda, _p(B), ldb, beta, _p(C), ldc, splits, _p(ws), _st())


def _im2col(x, N, H, W, C, ld, kh, kw, s, pt, pl, Ho, Wo):
    col = _new((N * Ho * Wo, kh * kw * C), torch.float32)
    _lib.call('tgan_im2col', _p(x), dt_code(x), N, H, W, C, ld, kh, kw, s, s, pt, pl, Ho, Wo, _p(col), _st())
    return col


class Deferred:
    """A tensor-core contraction whose launch is deferred until its consumer is known, so that the bias, the
    nonlinearity and the output row stride (a label-concatenated successor, modle_base.py:239-244) are fused into the
    GEMM epilogue instead of running as separate passes over the activation."""
    __slots__ = ('run', 'b', 'act', 'alpha')

    def __init__(self, run):
        self.run, self.b, self.act, self.alpha = run, None, 'none', 0.2


def _deferred(v):
    return v._data is None and isinstance(v._lazy, Deferred)


class ConcatDeferred:
    """an elementwise producer (dropout, batch-norm apply) waiting to learn whether a label concat follows: if so it writes
    straight into the concatenated tensor and fills the label planes in the same launch.  run(ld, lab, K, rps); ld = None
    materialises the plain tensor."""
    __slots__ = ('run',)

    def __init__(self, run):
        self.run = run


class MobnDeferred:
    """a mean-only-BN apply waiting to learn whether a 2x2 max pool (+ dropout) follows: if so the three run as one pass
    and the full-resolution activation is never written (ops.mobn_act / ops.max_pool2)"""
    __slots__ = ('run_plain', 'run_pooled')

    def __init__(self, run_plain, run_pooled):
        self.run_plain, self.run_pooled = run_plain, run_pooled


class PoolDeferred:
    """a 2x2 max pool waiting to learn whether a dropout follows (the classifier's max_pool -> dropout pairs run as one
    kernel forward and one backward)"""
    __slots__ = ('run',)

    def __init__(self, run):
        self.run = run


def _epilogue_bwd(out, b, act, alpha):
    """gradient through y = act(z + b) fused in a GEMM epilogue: returns dz (bf16 [rows, C]) and accumulates db"""
    dy = out.grad
    C, rows = out.C, out.rows
    a = ACT[act]
    need_db = b is not None and b.requires_grad
    if a == 0:
        if need_db:
            _lib.call('tgan_channel_stats', _p(dy), dt_code(dy), rows, C, _p(b.grad), None, 1.0, _p(ctx.ws()), _st())
        return dy
    du = _new((rows, C), out.data.dtype)
    _lib.call('tgan_act_bwd_ld', _p(dy), dt_code(dy), _p(out.data), dt_code(out.data), out.ld, _p(du), dt_code(du),
              rows, C, a, alpha, None, _p(b.grad) if need_db else None, _p(ctx.ws()), _st())
    return du


def conv2d(x, w, kh, kw, stride=1, padding='SAME', colsum=False):
    """tf.nn.conv2d (NHWC x HWIO).  x may be 2-D [rows, C] with kh = kw = 1 (tf.matmul).  In tensor-core mode the
    launch is deferred (see Deferred) unless the caller wants the epilogue's channel sums (mean-only BN)."""
    Cout = w.key.shape[-1]
    C = x.C
    if len(x.shape) == 2:
        N, H, W = x.shape[0], 1, 1
        assert kh == 1 and kw == 1
    else:
        N, H, W = x.shape[0], x.shape[1], x.shape[2]
    if padding.upper() == 'SAME':
        Ho, pt = same_pad(H, kh, stride)
        Wo, pl = same_pad(W, kw, stride)
    else:
        Ho, pt, Wo, pl = (H - kh) // stride + 1, 0, (W - kw) // stride + 1, 0
    oshape = (N, Cout) if len(x.shape) == 2 else (N, Ho, Wo, Cout)
    rg = _on() and (x.requires_grad or w.requires_grad)
    if ctx.building:
        return _prop(Var(None, oshape, requires_grad=rg), x)
    from . import tc
    geom = dict(N=N, H=H, W=W, C=C, kh=kh, kw=kw, s=stride, pt=pt, pl=pl, Ho=Ho, Wo=Wo, Cout=Cout)
    use_tc = ctx.math == 'bf16' and tc.conv_eligible(geom, x)
    rows = N * Ho * Wo
    K = kh * kw * C
    direct = (kh == 1 and kw == 1 and stride == 1 and pt == 0 and pl == 0)
    segs = _segs(x)
    if use_tc:
        tape = ctx.tape
        out = _prop(Var(None, oshape, requires_grad=rg), x)

        def run(b, act, alpha, out_ld):
            ld = out_ld or Cout
            cs = None
            if colsum and (segs is None or len(segs) <= 4):
                cs = arena_take(2 * Cout * (len(segs) if segs else 1))      # int64 Q24 entries (two floats each)
            z = tc.conv_fwd(x, w, geom, cs, segs if cs is not None else None, bias=None if b is None else b.data,
                            act=ACT[act], ldo=ld)
            out._data, out.ld, out._lazy = z.view(tuple(out.shape[:-1]) + (ld,)), ld, None
            if cs is not None:
                out.aux = dict(out.aux or {}, colsum=cs)
            if out.requires_grad and tape is not None:
                def bwd():
                    if out.grad is not None:
                        tc.conv_bwd(x, w, geom, _epilogue_bwd(out, b, act, alpha))
                tape.nodes.append(bwd)

        if colsum:
            run(None, 'none', 0.2, None)
        else:
            out._lazy = Deferred(run)
        return out
    xd = x.data
    Wt = w.value()
    if direct:
        a = _to_f32(xd, rows, C, x.ld)
        lda = C
    else:
        a = _im2col(xd, N, H, W, C, x.ld, kh, kw, stride, pt, pl, Ho, Wo)
        lda = K
    z = _new((rows, Cout), torch.float32)
    _sgemm(0, 0, rows, Cout, K, a, lda, Wt, Cout, z, Cout)
    del a
    out = _prop(Var(z.view(oshape), oshape, requires_grad=rg), x)
    if rg:
        tape = ctx.tape

        def bwd():
            if out.grad is None:
                return
            dz = out.grad
            dzf = _to_f32(dz, rows, Cout, Cout)
            Wt = w.value()
            if w.requires_grad:
                if direct:
                    a, lda = _to_f32(x.data, rows, C, x.ld), C
                else:
                    a, lda = _im2col(x.data, N, H, W, C, x.ld, kh, kw, stride, pt, pl, Ho, Wo), K
                _sgemm(1, 0, K, Cout, rows, a, lda, dzf, Cout, w.grad_target(), Cout, beta=1.0)
                del a
            if x.requires_grad:
                dcol = _new((rows, K), torch.float32)
                _sgemm(0, 1, rows, K, Cout, dzf, Cout, Wt, Cout, dcol, K)
                if direct:
                    dx = dcol if x.data.dtype == torch.float32 else _cast(dcol, x.data.dtype)
                    dx = dx.view(x.shape)
                else:
                    dx = _new(x.shape, x.data.dtype)
                    _lib.call('tgan_col2im', _p(dcol), N, H, W, C, kh, kw, stride, stride, pt, pl, Ho, Wo, _p(dx),
                              dt_code(dx), C, C, _st())
                add_grad(x, dx)
        tape.nodes.append(bwd)
    return out


def conv2d_transpose(x, w, kh, kw, stride=2):
    """tf.nn.conv2d_transpose 'SAME' (modle_base.py:149,250); filter [kh,kw,Cout,Cin]."""
    Cout, Cin = w.key.shape[2], w.key.shape[3]
    N, h, wd = x.shape[0], x.shape[1], x.shape[2]
    assert x.C == Cin
    Ho, Wo = h * stride, wd * stride
    _, pt = same_pad(Ho, kh, stride)
    _, pl = same_pad(Wo, kw, stride)
    oshape = (N, Ho, Wo, Cout)
    rg = _on() and (x.requires_grad or w.requires_grad)
    if ctx.building:
        return Var(None, oshape, requires_grad=rg)
    from . import tc
    geom = dict(N=N, h=h, w=wd, Cin=Cin, Cout=Cout, kh=kh, kw=kw, s=stride, pt=pt, pl=pl, Ho=Ho, Wo=Wo)
    use_tc = ctx.math == 'bf16' and tc.deconv_eligible(geom, x) and not (
        tc._skinny(geom) and isinstance(w, WNWeight) and os.environ.get('TGAN_NO_WN_SKINNY_TC'))
    rows, KK = N * h * wd, kh * kw * Cout
    if use_tc:
        tape = ctx.tape
        out = Var(None, oshape, requires_grad=rg)

        def run(b, act, alpha, out_ld):
            ld = out_ld or Cout
            y = tc.deconv_fwd(x, w, geom, bias=None if b is None else b.data, act=ACT[act], ldo=ld)
            out._data, out.ld, out._lazy = y.view(tuple(out.shape[:-1]) + (ld,)), ld, None
            if out.requires_grad and tape is not None:
                def bwd():
                    if out.grad is not None:
                        tc.deconv_bwd(x, w, geom, _epilogue_bwd(out, b, act, alpha))
                tape.nodes.append(bwd)

        out._lazy = Deferred(run)
        return out
    xf = _to_f32(x.data, rows, Cin, x.ld)
    dcol = _new((rows, KK), torch.float32)
    _sgemm(0, 1, rows, KK, Cin, xf, Cin, w.value(), Cin, dcol, KK)
    y = _new(oshape, _out_dtype(Cout) if ctx.math == 'bf16' else torch.float32)
    _lib.call('tgan_col2im', _p(dcol), N, Ho, Wo, Cout, kh, kw, stride, stride, pt, pl, h, wd, _p(y), dt_code(y),
              Cout, Cout, _st())
    del dcol
    out = Var(y, oshape, requires_grad=rg)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            col = _im2col(dy, N, Ho, Wo, Cout, Cout, kh, kw, stride, pt, pl, h, wd)      # [rows, KK]
            if w.requires_grad:
                xf = _to_f32(x.data, rows, Cin, x.ld)
                _sgemm(1, 0, KK, Cin, rows, col, KK, xf, Cin, w.grad_target(), Cin, beta=1.0)
            if x.requires_grad:
                dxf = _new((rows, Cin), torch.float32)
                _sgemm(0, 0, rows, Cin, KK, col, KK, w.value(), Cin, dxf, Cin)
                add_grad(x, _cast(dxf, x.data.dtype).view(x.shape))
        ctx.tape.nodes.append(bwd)
    return out


# ----------------------------------------------------------------------------------------------
# epilogues: bias / mean-only BN / BN, fused with the nonlinearity
# ----------------------------------------------------------------------------------------------


def _affine_act(x, rows, C, scale, shift, act, alpha, out_dtype):
    y = _new(x.shape, out_dtype)
    _lib.call('tgan_affine_act', _p(x), dt_code(x), _p(y), dt_code(y), rows, C, _p(scale), _p(shift), act, alpha,
              _st())
    return y


def _colsum_into(param_grad, colsum):
    accumulate_(param_grad.view(-1), colsum.view(-1))


def bias_act(z, b, act='none', alpha=0.2):
    """y = act(z + b)  (tf.nn.bias_add + nonlinearity: nn.py:508,517; tf.layers bias)."""
    a = ACT[act]
    C, rows = z.C, z.rows
    rg = _on() and (z.requires_grad or (b is not None and b.requires_grad))
    if ctx.building:
        return Var(None, z.shape, requires_grad=rg)
    y = _affine_act(z.data, rows, C, None, None if b is None else b.data, a, alpha, _out_dtype(C))
    out = _prop(Var(y, z.shape, requires_grad=rg), z)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            need_db = b is not None and b.requires_grad
            if a == 0:
                du = dy
                if need_db:      # db += column sums of dy, accumulated in place by the reduction's last CTA
                    _lib.call('tgan_channel_stats', _p(dy), dt_code(dy), rows, C, _p(b.grad), None, 1.0, _p(ctx.ws()),
                              _st())
            else:
                du = _new(z.shape, z.data.dtype)
                _lib.call('tgan_act_bwd', _p(dy), dt_code(dy), _p(out.data), dt_code(out.data), _p(du), dt_code(du),
                          rows, C, a, alpha, None, _p(b.grad) if need_db else None, _p(ctx.ws()), _st())
            if z.requires_grad:
                add_grad(z, du if du.dtype == z.data.dtype else _cast(du, z.data.dtype))
        ctx.tape.nodes.append(bwd)
    return out


def lazy_bias(z, b):
    """Defers `z + b` so that a following activation fuses into ONE epilogue
    (tf.layers.dense / conv2d return value followed by tf.nn.relu / leakyReLu in the builders): the epilogue of the
    deferred tensor-core GEMM itself, or one bias+activation kernel behind an already computed z."""
    if _deferred(z) and z._lazy.b is None and z._lazy.act == 'none':
        z._lazy.b = b
        z.requires_grad = z.requires_grad or (_on() and b.requires_grad)
        return z
    v = _prop(Var(None, z.shape, requires_grad=_on() and (z.requires_grad or b.requires_grad)), z)
    v._lazy = (z, b)
    return v


def _materialize(v, out_ld=None):
    if isinstance(v._lazy, ConcatDeferred):
        v._lazy.run(None, None, 0, 0)
        return
    if isinstance(v._lazy, MobnDeferred):
        v._lazy.run_plain()
        return
    if isinstance(v._lazy, PoolDeferred):
        v._lazy.run(0.0, None)
        return
    if isinstance(v._lazy, Deferred):
        d = v._lazy
        d.run(d.b, d.act, d.alpha, out_ld)
        return
    z, b = v._lazy
    y = bias_act(z, b, 'none')
    v._lazy = None
    if ctx.building:
        return
    v._data = y.data
    if y.requires_grad:
        def bwd():
            if v.grad is not None:
                add_grad(y, v.grad)
        ctx.tape.nodes.append(bwd)


def activation(x, act, alpha=0.2):
    """tf.nn.relu / leaky_relu / tanh / sigmoid / softplus; fuses into a pending GEMM epilogue / pending bias."""
    if _deferred(x) and x._lazy.act == 'none' and (act != 'lrelu' or abs(alpha - 0.2) < 1e-12):
        x._lazy.act, x._lazy.alpha = act, alpha
        return x
    if x._data is None and isinstance(x._lazy, tuple):
        z, b = x._lazy
        return bias_act(z, b, act, alpha)
    return bias_act(x, None, act, alpha)


def mobn_act(z, b, pop_mean, train, act='none', alpha=0.2, decay=0.9):
    """mean_only_batch_norm_impl + nonlinearity (nn.py:147-187, :517).
    train: y = act(z - mean_{rows}(z) + b), pop_mean <- decay*pop_mean + (1-decay)*mean
    eval : y = act(z - pop_mean + b).
    A grouped batch (ops.group_batch) keeps one mean per call: statistics, apply and their backward run per segment
    of rows, everything else of the layer ran once on the whole group."""
    a = ACT[act]
    C, rows = z.C, z.rows
    rg = _on() and (z.requires_grad or b.requires_grad)
    if ctx.building:
        return _prop(Var(None, z.shape, requires_grad=rg), z)
    segs = _segs(z) or [z.shape[0]]
    rps = rows // sum(segs)                      # rows per sample
    bounds, r0 = [], 0
    for n in segs:
        bounds.append((r0, n * rps))
        r0 += n * rps
    ends = [b0 + nr for b0, nr in bounds[:-1]] + [0, 0, 0]
    zd = z.data
    ez = zd.element_size()
    cs_all = z.aux.get('colsum') if z.aux else None      # [nseg, C] accumulated by the tcgen05 GEMM epilogue
    ydt = _out_dtype(C)
    fused = zd.dtype == torch.bfloat16 and ydt == torch.bfloat16 and C % 8 == 0 and len(segs) <= 4
    # the 10 logits: every segment in one single-CTA launch (tgan_mobn_small_*) instead of two launches per segment
    small = (not fused and zd.dtype == torch.float32 and ydt == torch.float32 and C <= 32 and rows * C <= (1 << 16)
             and len(segs) <= 4 and z.ld == C and not os.environ.get('TGAN_NO_SMALL_MOBN'))
    tape = ctx.tape
    out = _prop(Var(None, z.shape, requires_grad=rg), z)

    def seg_stats():
        s_all = _new((len(segs), C), torch.float32)
        for i, (b0, nr) in enumerate(bounds):
            _lib.call('tgan_channel_stats', zd.data_ptr() + b0 * C * ez, dt_code(zd), nr, C, s_all.data_ptr() + 4 * i * C,
                      None, 0.0, _p(ctx.ws()), _st())
        return s_all

    def run_plain():
        y = _new(z.shape, ydt)
        ey = y.element_size()
        if fused:       # one launch for the whole grouped batch, nonlinearity included
            sums = (cs_all if cs_all is not None else seg_stats()) if train else None
            _lib.call('tgan_mobn_apply_seg', _p(zd), _p(y), rows, C, len(segs), ends[0], ends[1], ends[2], _p(sums),
                      1 if cs_all is not None else 0, _p(b.data), _p(pop_mean.data), decay, 1 if train else 0, a, alpha,
                      _st())
        elif small:
            _lib.call('tgan_mobn_small_fwd', _p(zd), _p(y), rows, C, len(segs), ends[0], ends[1], ends[2], _p(b.data),
                      _p(pop_mean.data), decay, 1 if train else 0, a, alpha, _st())
        else:
            for i, (b0, nr) in enumerate(bounds):
                s = None
                if train:       # (the epilogue's fixed-point sums are only consumed by the fused kernels)
                    s = _new((C,), torch.float32)
                    _lib.call('tgan_channel_stats', zd.data_ptr() + b0 * C * ez, dt_code(zd), nr, C, _p(s), None, 0.0,
                              _p(ctx.ws()), _st())
                _lib.call('tgan_mobn_apply', zd.data_ptr() + b0 * C * ez, dt_code(zd), y.data_ptr() + b0 * C * ey, dt_code(y),
                          nr, C, _p(s), _p(b.data), _p(pop_mean.data), decay, 1 if train else 0, a, alpha, _st())
        out._data, out._lazy = y, None
        if out.requires_grad and tape is not None:
            def bwd():
                if out.grad is None:
                    return
                dy = out.grad
                du = _new(z.shape, zd.dtype)
                ed, eu = dy.element_size(), du.element_size()
                if fused:
                    cs = _new((4, C), torch.float32)
                    _lib.call('tgan_act_bwd_seg', _p(dy), dt_code(dy), _p(y), dt_code(y), _p(du), dt_code(du), rows, C,
                              len(segs), ends[0], ends[1], ends[2], a, alpha, _p(cs),
                              _p(b.grad) if b.requires_grad else None, _p(ctx.ws()), _st())
                    if z.requires_grad and train:
                        _lib.call('tgan_sub_channel_mean_seg', _p(du), _p(du), rows, C, len(segs), ends[0], ends[1], ends[2],
                                  _p(cs), _st())
                elif small and dy.dtype == torch.float32:
                    _lib.call('tgan_mobn_small_bwd', _p(dy), _p(y), _p(du) if z.requires_grad else None, rows, C, len(segs),
                              ends[0], ends[1], ends[2], a, alpha, 1 if train else 0,
                              _p(b.grad) if b.requires_grad else None, _st())
                else:
                    for b0, nr in bounds:
                        cs = _new((C,), torch.float32)
                        _lib.call('tgan_act_bwd', dy.data_ptr() + b0 * C * ed, dt_code(dy), y.data_ptr() + b0 * C * ey,
                                  dt_code(y), du.data_ptr() + b0 * C * eu, dt_code(du), nr, C, a, alpha, _p(cs),
                                  _p(b.grad) if b.requires_grad else None, _p(ctx.ws()), _st())
                        if z.requires_grad and train:
                            _lib.call('tgan_sub_channel_mean', du.data_ptr() + b0 * C * eu, dt_code(du),
                                      du.data_ptr() + b0 * C * eu, dt_code(du), nr, C, _p(cs), _st())
                if z.requires_grad:
                    add_grad(z, du)
            tape.nodes.append(bwd)

    poolable = (fused and len(z.shape) == 4 and z.shape[1] % 2 == 0 and z.shape[2] % 2 == 0 and C >= 64 and 2048 % C == 0
                and z.ld == C)
    if not poolable:
        run_plain()
        return out

    def run_pooled(pout, rate, tag):
        """mean-only BN + nonlinearity + 2x2 max pool (+ dropout) in one pass; `out` (full resolution) is never written"""
        N, H, W = z.shape[0], z.shape[1], z.shape[2]
        iends = [sum(segs[:i + 1]) for i in range(len(segs) - 1)] + [0, 0, 0]       # segment ends in images
        y, code = _new(pout.shape, torch.bfloat16), _new(pout.shape, torch.uint8)
        sums = (cs_all if cs_all is not None else seg_stats()) if train else None
        # border-class sums of the pooled output: the next convolution's fused mean-only BN reads them (conv2d_mobn)
        clo = None
        if (train and W // 2 in (16, 32) and (H // 2) & (H // 2 - 1) == 0 and (H // 2) * C <= 8192
                and not os.environ.get('TGAN_NO_MOBN_FUSION')):      # (only where a fusable convolution can follow)
            clo = arena_take(2 * 9 * C * len(segs))
        rng = ctx.rng
        mask, seed, sid, ctr = None, 0, 0, None
        if rate > 0 and rng.injected:
            mask = _rng_mask(tag, pout.shape, rate)
        elif rate > 0:
            seed, sid, ctr = rng.seed, rng.stream_id(str(tag)), rng.counter()
        _lib.call('tgan_mobn_pool_dropout_fwd', _p(zd), _p(y), _p(code), N, H, W, C, len(segs), iends[0], iends[1], iends[2],
                  _p(sums), 1 if cs_all is not None else 0, _p(b.data), _p(pop_mean.data), decay, 1 if train else 0, a, alpha,
                  float(rate), _p(mask), seed, sid, _p(ctr), _st())
        pout._data, pout._lazy = y, None
        if clo is not None:      # one pass over the (L2-resident) pooled tensor: one CTA per image, fixed summation order
            se = (ctypes.c_int * 3)(*iends[:3])
            _lib.call('tgan_class_sums', _p(y), BF16, N, H // 2, W // 2, C, C, len(segs), se, _p(clo), _st())
            pout.aux = dict(pout.aux or {}, cls=clo)
        out._lazy = None                 # consumed: the full-resolution activation does not exist
        if pout.requires_grad and tape is not None:
            def bwd():
                if pout.grad is None:
                    return
                du = _new(z.shape, torch.bfloat16)
                cs = _new((4, C), torch.float32)
                _lib.call('tgan_mobn_pool_dropout_bwd', _p(_cast(pout.grad, torch.bfloat16)), _p(y), _p(code), _p(du), N, H, W,
                          C, len(segs), iends[0], iends[1], iends[2], a, alpha, float(rate), _p(cs),
                          _p(b.grad) if b.requires_grad else None, _p(ctx.ws()), _st())
                if z.requires_grad:
                    if train:
                        _lib.call('tgan_sub_channel_mean_seg', _p(du), _p(du), rows, C, len(segs), ends[0], ends[1], ends[2],
                                  _p(cs), _st())
                    add_grad(z, du)
            tape.nodes.append(bwd)

    out._lazy = MobnDeferred(run_plain, run_pooled)
    return out


def conv2d_mobn(x, w, kh, kw, stride, padding, b, pop_mean, train, act='none', alpha=0.2, decay=0.9):
    """tf.nn.conv2d + mean_only_batch_norm_impl + nonlinearity of nn.conv2d_WN (nn.py:504-518) as ONE unit.

    Tensor-core mode, training, 3x3 / stride 1 / SAME + leaky ReLU on 16- or 32-pixel-wide images (conv1_1, conv1_2,
    conv2_1, conv2_2 of the CIFAR-10 classifier): the batch mean of the convolution output is a linear function of
    border-class sums of its INPUT, which the producer of x emitted (GEMM epilogue / pooling kernel / tgan_class_sums);
    tgan_mobn_mean_from_sums turns them into the per-segment bias b - mean BEFORE the contraction runs, whose epilogue
    then applies `- mean + b`, the leaky ReLU, the bf16 store, the class sums for the NEXT layer and a 1-bit-per-element
    mask of the leaky-ReLU side.  No separate pass over the activation in forward; in backward the consumer's
    input-gradient epilogue applies lrelu' through the mask and accumulates the per-segment sums of du
    (tc.conv_bwd), so the activation-backward pass disappears as well.  When a 2x2 max pool follows (conv1_3, conv2_3)
    the one-pass pooled apply of mobn_act is used instead.  Everything else takes conv2d(colsum) + mobn_act."""
    def legacy():
        return mobn_act(conv2d(x, w, kh, kw, stride, padding, colsum=True), b, pop_mean, train, act, alpha, decay)

    if (ctx.building or ctx.math != 'bf16' or not train or act != 'lrelu' or abs(alpha - 0.2) > 1e-12 or kh != 3 or kw != 3
            or stride != 1 or padding.upper() != 'SAME' or len(x.shape) != 4 or os.environ.get('TGAN_NO_MOBN_FUSION')):
        return legacy()
    N, H, W = x.shape[0], x.shape[1], x.shape[2]
    C, Cout = x.C, w.key.shape[-1]
    segs = _segs(x) or [N]
    if not (W in (16, 32) and H >= 2 and H & (H - 1) == 0 and (H * W) % 256 == 0 and len(segs) <= 4 and Cout >= 64
            and 2048 % Cout == 0 and x.ld == C and (C <= 16 or C % 8 == 0)):
        return legacy()
    nseg = len(segs)
    rows, rps = N * H * W, H * W
    rg = _on() and (x.requires_grad or w.requires_grad or b.requires_grad)
    out = _prop(Var(None, (N, H, W, Cout), requires_grad=rg), x)
    tape = ctx.tape
    geom = dict(N=N, H=H, W=W, C=C, kh=3, kw=3, s=1, pt=1, pl=1, Ho=H, Wo=W, Cout=Cout)
    ends_img = [sum(segs[:i + 1]) for i in range(nseg - 1)] + [0, 0, 0]
    ends = [e * rps for e in ends_img]

    def forward_to(y):
        """`out` stands for y (the two-kernel path computed it): share the tensor, route the gradient"""
        out._data, out.ld, out._lazy = y.data, y.ld, None
        out.aux = dict(y.aux or {}, **{k: v for k, v in (out.aux or {}).items() if k == 'segs'})
        if out.requires_grad and tape is not None and y.requires_grad:
            def bwd():
                if out.grad is not None:
                    add_grad(y, out.grad)
            tape.nodes.append(bwd)

    def run_pooled(pout, rate, tag):
        y = legacy()      # (the eligibility test above implies mobn_act's own pooled form exists for this geometry)
        assert y._data is None and isinstance(y._lazy, MobnDeferred)
        y._lazy.run_pooled(pout, rate, tag)
        out._lazy = None

    def run_plain():
        from . import tc
        if x._data is None:      # tell a still-pending producer that THIS consumer reads its border-class sums
            x.aux = dict(x.aux or {}, want_cls=True)
        xd = x.data                                   # materialises the producer; its class sums are in x.aux now
        cls = (x.aux or {}).get('cls')
        if cls is None and C <= 16:                   # the network input: sums by a small kernel of its own
            cls = arena_take(2 * 9 * C * nseg)
            se = (ctypes.c_int * 3)(*ends_img[:3])
            _lib.call('tgan_class_sums', _p(xd), dt_code(xd), N, H, W, C, x.ld, nseg, se, _p(cls), _st())
        if cls is None or not tc.conv_eligible(geom, x):
            forward_to(legacy())
            return
        wp, Kpad, w_ts, w_cs = tc.fprop_operand(w, geom)
        shift = _new((4 * Cout,), torch.float32)
        cnt = (ctypes.c_int64 * nseg)(*[n * rps for n in segs])
        _lib.call('tgan_mobn_mean_from_sums', _p(cls), nseg, cnt, _p(wp), 9, Cout, C, w_ts, w_cs, _p(b.data),
                  _p(pop_mean.data), decay, _p(shift), _st())
        # class sums of the output only when the consumer asked for them (a fused layer behind this one); a consumer that
        # pools (conv1_3, conv2_3) does not, and the heavy layers' epilogues stay lean
        clo = arena_take(2 * 9 * Cout * nseg) if (out.aux or {}).get('want_cls') else None
        need_mask = out.requires_grad and tape is not None
        mask = _new((rows // 32, Cout), torch.int32) if need_mask else None
        y = tc.conv_fwd(x, w, geom, None, segs, bias=shift, act=ACT['lrelu'],
                        fused=dict(bias_seg=True, clsum=clo, cls_hw=(H, W), mask_out=mask))
        out._data, out.ld, out._lazy = y.view(N, H, W, Cout), Cout, None
        if clo is not None:
            out.aux = dict(out.aux or {}, cls=clo)
        if need_mask:
            out.aux = dict(out.aux or {}, mask=mask, mask_alpha=alpha)

            def bwd():
                if out.grad is None:
                    return
                cs = _new((4, Cout), torch.float32)
                if out.aux.get('du_ready'):           # the consumer's input-gradient epilogue already applied lrelu'
                    du = out.grad
                    _lib.call('tgan_seg_sums_finalize', _p(out.aux['du_q24']), nseg, Cout, _p(cs),
                              _p(b.grad) if b.requires_grad else None, _st())
                else:
                    dy = _cast(out.grad, torch.bfloat16)
                    du = _new(out.shape, torch.bfloat16)
                    _lib.call('tgan_act_bwd_seg', _p(dy), BF16, _p(y), BF16, _p(du), BF16, rows, Cout, nseg, ends[0],
                              ends[1], ends[2], ACT['lrelu'], alpha, _p(cs), _p(b.grad) if b.requires_grad else None,
                              _p(ctx.ws()), _st())
                if x.requires_grad or w.requires_grad:
                    _lib.call('tgan_sub_channel_mean_seg', _p(du), _p(du), rows, Cout, nseg, ends[0], ends[1], ends[2], _p(cs),
                              _st())
                    tc.conv_bwd(x, w, geom, du)
            tape.nodes.append(bwd)

    out._lazy = MobnDeferred(run_plain, run_pooled)
    return out


def batch_norm(x, gamma, beta, mm, mv, train, eps=1e-5, decay=0.9, unbiased=True):
    """tf.contrib.layers.batch_norm(scale=True, updates_collections=None) (modle_base.py:229-237); unbiased=False:
    the moving variance follows the biased batch variance (nn.batch_norm_impl, nn.py:207-214)."""
    C, rows = x.C, x.rows
    rg = _on() and (x.requires_grad or gamma.requires_grad or beta.requires_grad)
    if ctx.building:
        return _prop(Var(None, x.shape, requires_grad=rg), x)
    segs = _segs(x)
    if train and segs and len(segs) > 1:
        return _batch_norm_segments(x, gamma, beta, mm, mv, eps, decay, unbiased, segs, rg)
    xd = x.data
    scale, shift = _new((C,), torch.float32), _new((C,), torch.float32)
    mean, rstd = _new((C,), torch.float32), _new((C,), torch.float32)
    if train:
        s, ss = _new((C,), torch.float32), _new((C,), torch.float32)
        _lib.call('tgan_channel_stats', _p(xd), dt_code(xd), rows, C, _p(s), _p(ss), 0.0, _p(ctx.ws()), _st())
        _lib.call('tgan_bn_finalize', _p(s), _p(ss), rows, C, _p(gamma.data), _p(beta.data), eps, decay,
                  1 if unbiased else 0, None if mm is None else _p(mm.data), None if mv is None else _p(mv.data), _p(mean), _p(rstd), _p(scale), _p(shift), _st())
    else:
        _lib.call('tgan_bn_eval_affine', _p(gamma.data), _p(beta.data), _p(mm.data), _p(mv.data), eps, C, _p(scale),
                  _p(shift), _st())
    out = _prop(Var(None, x.shape, requires_grad=rg), x)

    def run(ld, lab, K, rps):
        if ld is None:
            out._data = _affine_act(xd, rows, C, scale, shift, 0, 0.0, _out_dtype(C))
        else:       # normalise straight into the label-concatenated tensor
            y = _new(tuple(out.shape[:-1]) + (ld,))
            _lib.call('tgan_affine_concat', _p(xd), dt_code(xd), _p(y), dt_code(y), ld, rows, C, _p(scale), _p(shift),
                      _p(lab), K, rps, _st())
            out._data, out.ld = y, ld
        out._lazy = None

    if ctx.math == 'bf16' and _out_dtype(C) == torch.bfloat16 and x.ld == C:
        out._lazy = ConcatDeferred(run)
    else:
        run(None, None, 0, 0)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            if not train:
                if x.requires_grad:
                    add_grad(x, _affine_act(dy, rows, C, scale, None, 0, 0.0, xd.dtype))
                return
            dx = _new(x.shape, xd.dtype)
            _lib.call('tgan_bn_bwd', _p(dy), dt_code(dy), _p(xd), dt_code(xd), _p(dx), dt_code(dx), rows, C, _p(mean),
                      _p(rstd), _p(gamma.data), _p(gamma.grad) if gamma.requires_grad else None,
                      _p(beta.grad) if beta.requires_grad else None, 1.0, _p(ctx.ws()), _st())
            if x.requires_grad:
                add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


def _batch_norm_segments(x, gamma, beta, mm, mv, eps, decay, unbiased, segs, rg):
    """Training-mode batch norm of a GROUPED batch (ops.group_batch): every network call keeps its own batch statistics
    (the rows of a call are contiguous), the moving statistics are updated once per call in call order -- what the
    separate calls of the reference graph compute -- while the convolutions around it run once on the whole group."""
    C, rows = x.C, x.rows
    assert x.ld == C, 'segmented batch norm of a channel-padded tensor'
    xd = x.data
    ex = xd.element_size()
    rps = rows // sum(segs)
    y = _new(x.shape, _out_dtype(C))
    ey = y.element_size()
    if len(segs) <= 4 and not os.environ.get('TGAN_NO_BN_SEG'):
        # all segments in three launches per direction (csrc/bn_seg.cu) instead of 3 + 4 launches per segment
        ends = [sum(segs[:i + 1]) * rps for i in range(len(segs) - 1)] + [0, 0, 0]
        mean, rstd = _new((len(segs), C), torch.float32), _new((len(segs), C), torch.float32)
        _lib.call('tgan_bn_fwd_seg', _p(xd), dt_code(xd), _p(y), dt_code(y), rows, C, len(segs), ends[0], ends[1], ends[2],
                  _p(gamma.data), _p(beta.data), eps, decay, 1 if unbiased else 0, None if mm is None else _p(mm.data),
                  None if mv is None else _p(mv.data), _p(mean), _p(rstd), _p(ctx.ws()), _st())
        out = _prop(Var(y, x.shape, requires_grad=rg), x)
        if rg:
            def bwd_seg():
                if out.grad is None:
                    return
                dy = out.grad
                dx = _new(x.shape, xd.dtype)
                _lib.call('tgan_bn_bwd_seg', _p(dy), dt_code(dy), _p(xd), dt_code(xd), _p(dx), dt_code(dx), rows, C, len(segs),
                          ends[0], ends[1], ends[2], _p(mean), _p(rstd), _p(gamma.data),
                          _p(gamma.grad) if gamma.requires_grad else None, _p(beta.grad) if beta.requires_grad else None,
                          1.0, _p(ctx.ws()), _st())
                if x.requires_grad:
                    add_grad(x, dx)
            ctx.tape.nodes.append(bwd_seg)
        return out
    stats, r0 = [], 0
    for n in segs:
        nr = n * rps
        scale, shift = _new((C,), torch.float32), _new((C,), torch.float32)
        mean, rstd = _new((C,), torch.float32), _new((C,), torch.float32)
        s, ss = _new((C,), torch.float32), _new((C,), torch.float32)
        xp = xd.data_ptr() + r0 * C * ex
        _lib.call('tgan_channel_stats', xp, dt_code(xd), nr, C, _p(s), _p(ss), 0.0, _p(ctx.ws()), _st())
        _lib.call('tgan_bn_finalize', _p(s), _p(ss), nr, C, _p(gamma.data), _p(beta.data), eps, decay, 1 if unbiased else 0,
                  None if mm is None else _p(mm.data), None if mv is None else _p(mv.data), _p(mean), _p(rstd), _p(scale),
                  _p(shift), _st())
        _lib.call('tgan_affine_act', xp, dt_code(xd), y.data_ptr() + r0 * C * ey, dt_code(y), nr, C, _p(scale), _p(shift), 0, 0.0,
                  _st())
        stats.append((r0, nr, mean, rstd))
        r0 += nr
    out = _prop(Var(y, x.shape, requires_grad=rg), x)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            ed = dy.element_size()
            dx = _new(x.shape, xd.dtype)
            for b0, nr, mean, rstd in stats:
                _lib.call('tgan_bn_bwd', dy.data_ptr() + b0 * C * ed, dt_code(dy), xd.data_ptr() + b0 * C * ex, dt_code(xd),
                          dx.data_ptr() + b0 * C * ex, dt_code(dx), nr, C, _p(mean), _p(rstd), _p(gamma.data),
                          _p(gamma.grad) if gamma.requires_grad else None, _p(beta.grad) if beta.requires_grad else None,
                          1.0, _p(ctx.ws()), _st())
            if x.requires_grad:
                add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


# ----------------------------------------------------------------------------------------------
# pointwise layers
# ----------------------------------------------------------------------------------------------


def add_noise(x, std, tag):
    """x + N(0, std) (modle_base.py:193-202; Good_GAN_cifar10.py:29-31).  Active in eval too."""
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, x.shape, requires_grad=rg)
    xd = x.data
    n = xd.numel()
    y = _new(x.shape)
    rng = ctx.rng
    if rng.injected:
        noise = _rng_normal(tag, x.shape)
        _lib.call('tgan_add_noise', _p(xd), dt_code(xd), _p(y), dt_code(y), n, std, _p(noise), 0, 0, None, _st())
    else:
        _lib.call('tgan_add_noise', _p(xd), dt_code(xd), _p(y), dt_code(y), n, std, None, rng.seed,
                  rng.stream_id(str(tag)), _p(rng.counter()), _st())
    out = _prop(Var(y, x.shape, requires_grad=rg), x)
    if rg:
        def bwd():
            if out.grad is not None:
                add_grad(x, out.grad if out.grad.dtype == xd.dtype else _cast(out.grad, xd.dtype))
        ctx.tape.nodes.append(bwd)
    return out


def dropout(x, rate, tag, training=True):
    """tf.layers.dropout (modle_base.py:190-191; Good_GAN_cifar10.py:124,143)."""
    if not training:
        return x
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, x.shape, requires_grad=rg)
    if x._data is None and isinstance(x._lazy, PoolDeferred):
        x._lazy.run(rate, tag)           # fused into the pending max pool
        return x
    xd = x.data
    n = xd.numel()
    assert x.ld == x.C, 'dropout of a channel-padded tensor'
    rng = ctx.rng
    mask = _rng_mask(tag, x.shape, rate) if rng.injected else _new(x.shape, torch.uint8)
    gen = 0 if rng.injected else 1
    seed, sid, ctr = (0, 0, None) if rng.injected else (rng.seed, rng.stream_id(str(tag)), rng.counter())
    out = _prop(Var(None, x.shape, requires_grad=rg), x)

    def run(ld, lab, K, rps):
        if ld is None:
            y = _new(x.shape)
            _lib.call('tgan_dropout', _p(xd), dt_code(xd), _p(y), dt_code(y), _p(mask), n, rate, gen, seed, sid, _p(ctr), _st())
            out._data = y
        else:       # drop straight into the label-concatenated tensor (modle_base.py:190-191 + :239-244)
            y = _new(tuple(out.shape[:-1]) + (ld,))
            _lib.call('tgan_dropout_concat', _p(xd), dt_code(xd), _p(y), dt_code(y), ld, _p(mask), x.rows, x.C, rate, gen,
                      seed, sid, _p(ctr), _p(lab), K, rps, _st())
            out._data, out.ld = y, ld
        out._lazy = None

    if ctx.math == 'bf16':
        out._lazy = ConcatDeferred(run)
    else:
        run(None, None, 0, 0)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dx = _new(x.shape, xd.dtype)
            _lib.call('tgan_dropout', _p(out.grad), dt_code(out.grad), _p(dx), dt_code(dx), _p(mask), n, rate, 0, 0, 0,
                      None, _st())
            add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


def max_pool2(x):
    """2x2/s2 max pool (Good_GAN_cifar10.py:123,142; Good_GAN.py:223,233,264,279)."""
    N, H, W, C = x.shape
    oshape = (N, H // 2, W // 2, C)
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, oshape, requires_grad=rg)
    if x._data is None and isinstance(x._lazy, MobnDeferred):
        out = _prop(Var(None, oshape, requires_grad=rg), x)
        fuse = x._lazy.run_pooled
        out._lazy = PoolDeferred(lambda rate, tag: fuse(out, rate, tag))
        return out
    xd = x.data
    if ctx.math == 'bf16' and xd.dtype == torch.bfloat16 and C % 8 == 0 and x.ld == C:
        out = _prop(Var(None, oshape, requires_grad=rg), x)
        tape = ctx.tape

        def run(rate, tag):
            y, code = _new(oshape, torch.bfloat16), _new(oshape, torch.uint8)
            rng = ctx.rng
            if rate > 0 and rng.injected:
                mask = _rng_mask(tag, oshape, rate)
                _lib.call('tgan_maxpool2_dropout_fwd', _p(xd), _p(y), _p(code), N, H, W, C, rate, _p(mask), 0, 0, None, _st())
            elif rate > 0:
                _lib.call('tgan_maxpool2_dropout_fwd', _p(xd), _p(y), _p(code), N, H, W, C, rate, None, rng.seed,
                          rng.stream_id(str(tag)), _p(rng.counter()), _st())
            else:
                _lib.call('tgan_maxpool2_dropout_fwd', _p(xd), _p(y), _p(code), N, H, W, C, 0.0, None, 0, 0, None, _st())
            out._data, out._lazy = y, None
            if out.requires_grad and tape is not None:
                def bwd():
                    if out.grad is None:
                        return
                    dx = _new(x.shape, torch.bfloat16)
                    _lib.call('tgan_maxpool2_dropout_bwd', _p(_cast(out.grad, torch.bfloat16)), _p(code), _p(dx), N, H, W,
                              C, rate, _st())
                    add_grad(x, dx)
                tape.nodes.append(bwd)

        out._lazy = PoolDeferred(run)
        return out
    y, idx = _new(oshape, xd.dtype), _new(oshape, torch.uint8)
    _lib.call('tgan_maxpool2_fwd', _p(xd), dt_code(xd), _p(y), _p(idx), N, H, W, C, _st())
    out = _prop(Var(y, oshape, requires_grad=rg), x)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dx = _new(x.shape, xd.dtype)
            _lib.call('tgan_maxpool2_bwd', _p(out.grad), dt_code(out.grad), _p(idx), _p(dx), N, H, W, C, _st())
            add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


def global_pool(x, mode):
    """mode 'max': tf.layers.max_pooling2d(pool=H) ('avg_pool_0', Good_GAN_cifar10.py:163);
    mode 'mean': average_pooling2d(8,1) (:94) / reduce_mean([1,2]) (Good_GAN.py:157,243,295)."""
    N, H, W, C = x.shape
    m = 0 if mode == 'max' else 1
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, (N, C), requires_grad=rg)
    xd = x.data
    y = _new((N, C), xd.dtype)
    idx = _new((N, C), torch.uint8) if m == 0 else None
    _lib.call('tgan_global_pool_fwd', _p(xd), dt_code(xd), _p(y), dt_code(y), _p(idx), N, H * W, C, m, _st())
    out = _prop(Var(y, (N, C), requires_grad=rg), x)
    if rg:
        def bwd():
            if out.grad is None:
                return
            dx = _new(x.shape, xd.dtype)
            _lib.call('tgan_global_pool_bwd', _p(out.grad), dt_code(out.grad), _p(idx), _p(dx), dt_code(dx), N, H * W,
                      C, m, _st())
            add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


def concat_label(x, y):
    """_conv_cond_concat (modle_base.py:239-244) and tf.concat([h, y], 1): appends the K label
    channels.  The result is padded to a multiple of 8 channels (16-byte pixel stride for TMA); the
    pad is zero and invisible to consumers (they use the logical C)."""
    K = y.shape[-1]
    C = x.C
    ld = (C + K + 7) // 8 * 8 if ctx.math == 'bf16' else C + K
    oshape = tuple(x.shape[:-1]) + (C + K,)
    rows = x.rows
    rps = rows // x.shape[0]
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, oshape, ld=ld, requires_grad=rg)
    lab = y.data if isinstance(y, Var) else y
    assert lab.dtype == torch.float32
    if x._data is None and isinstance(x._lazy, ConcatDeferred) and ctx.math == 'bf16':
        x._lazy.run(ld, lab, K, rps)           # the producer writes the concatenated tensor itself, labels included
        o = x.data
    elif _deferred(x) and ctx.math == 'bf16':
        # the producing GEMM writes channels [0, C) straight into the concatenated buffer; only the label planes
        # (and the zero pad) are filled here
        _materialize(x, out_ld=ld)
        o = x.data
        _lib.call('tgan_fill_label', _p(lab), K, rps, _p(o), dt_code(o), rows, C, ld, _st())
    else:
        xd = x.data
        o = _new(tuple(x.shape[:-1]) + (ld,))
        _lib.call('tgan_concat_label', _p(xd), dt_code(xd), rows, C, x.ld, _p(lab), K, rps, _p(o), dt_code(o), ld, _st())
    out = Var(o, oshape, ld=ld, requires_grad=rg)
    # tensor-core consumers compute the input gradient of the first C channels only and hand it to x directly
    out.aux = {'concat_src': x, 'C0': C}
    if rg:
        def bwd():
            if out.grad is None:
                return
            g = out.grad          # [rows, C+K] contiguous (producers write logical channels densely)
            dx = _new(x.shape, x.data.dtype)
            _lib.call('tgan_copy_channels', _p(g), dt_code(g), C + K, _p(dx), dt_code(dx), C, rows, C, _st())
            add_grad(x, dx)
        ctx.tape.nodes.append(bwd)
    return out


def reshape(x, shape):
    shape = list(shape)
    if -1 in shape:
        known = int(np.prod([s for s in shape if s != -1]))
        shape[shape.index(-1)] = int(np.prod(x.shape)) // known
    shape = tuple(int(s) for s in shape)
    assert x.ld == x.C, 'reshape of a channel-padded tensor'
    rg = _on() and x.requires_grad
    if ctx.building:
        return Var(None, shape, requires_grad=rg)
    out = _prop(Var(x.data.view(shape), shape, requires_grad=rg), x)
    if rg:
        def bwd():
            if out.grad is not None:
                add_grad(x, out.grad.view(x.shape))
        ctx.tape.nodes.append(bwd)
    return out


def reshape_pending(x, shape):
    """reshape that keeps a pending epilogue (deferred GEMM / pending bias) pending"""
    shape = tuple(int(v) for v in shape)
    if _deferred(x):
        assert int(np.prod(shape)) == int(np.prod(x.shape))
        x.shape = shape           # the launch views its output with the shape current at materialisation
        return x
    if x._data is None and isinstance(x._lazy, tuple):
        z, b = x._lazy
        return lazy_bias(reshape(z, shape), b)
    return reshape(x, shape)


def force(v):
    """apply whatever is pending on v now (logits that are consumed as they are)"""
    if v._data is None and isinstance(v._lazy, tuple):
        return bias_act(*v._lazy, 'none')
    if not ctx.building:
        v.data
    return v


def constant(t, dtype=torch.float32):
    """Host/device tensor or ndarray -> Var (no gradient)."""
    if ctx.building:
        return Var(None, tuple(np.shape(t)))
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    return Var(t.to(ctx.device, dtype).contiguous(), tuple(t.shape))


# ----------------------------------------------------------------------------------------------
# pseudo-labels and losses
# ----------------------------------------------------------------------------------------------


def argmax_onehot(logits, depth):
    """tf.argmax(axis=1) + tf.one_hot(depth) (Good_GAN_cifar10.py:232,259): (int64 idx, fp32 one-hot)."""
    N, K = logits.shape
    if ctx.building:
        return Var(None, (N,)), Var(None, (N, depth))
    assert K == depth and logits.data.dtype == torch.float32
    idx, oh = _new((N,), torch.int64), _new((N, depth), torch.float32)
    _lib.call('tgan_argmax_onehot', _p(logits.data), N, K, _p(idx), _p(oh), _st())
    return Var(idx, (N,)), Var(oh, (N, depth))


class Loss:
    """A scalar loss left on the device + the closed-form dlogits computed by the same kernel."""

    def __init__(self, value, pairs):
        self.value = value          # fp32 [1] device tensor
        self._pairs = pairs         # [(logits Var, dlogits tensor)]

    def seed(self):
        for v, g in self._pairs:
            if v is not None and v.requires_grad:
                add_grad(v, g)

    def item(self):
        return float(self.value.item())


def _f32logits(v):
    assert v.data.dtype == torch.float32 and v.ld == v.C, 'loss kernels take fp32 logits'
    return v.data


def loss_d(dr, df, du):
    """train_base.py:123-126."""
    if ctx.building:
        return Loss(None, [])
    val = _new((1,), torch.float32)
    g = [_new(v.shape, torch.float32) for v in (dr, df, du)]
    _lib.call('tgan_loss_d', _p(_f32logits(dr)), dr.rows, _p(_f32logits(df)), df.rows, _p(_f32logits(du)), du.rows,
              _p(val), _p(g[0]), _p(g[1]), _p(g[2]), _st())
    return Loss(val, list(zip((dr, df, du), g)))


def loss_g(df, out=None):
    """train_base.py:128.  out: optional fp32 [1] device view that receives the scalar."""
    if ctx.building:
        return Loss(None, [])
    val, g = (out if out is not None else _new((1,), torch.float32)), _new(df.shape, torch.float32)
    _lib.call('tgan_loss_g', _p(_f32logits(df)), df.rows, _p(val), _p(g), _st())
    return Loss(val, [(df, g)])


def loss_c(c_real, y_l_c, c_unl, c_rep, d_unl_logits, c_fake, y_g, lambdas, out=None):
    """train_base.py:130-152.  lambdas: device fp32 [2] = {lambda_1, lambda_2}."""
    if ctx.building:
        return Loss(None, [])
    K = c_real.shape[1]
    val = out if out is not None else _new((1,), torch.float32)
    g_real, g_unl, g_fake = (_new(v.shape, torch.float32) for v in (c_real, c_unl, c_fake))
    g_rep = _new(c_rep.shape, torch.float32) if c_rep is not None else None
    _lib.call('tgan_loss_c', _p(_f32logits(c_real)), _p(y_l_c.data), c_real.rows, _p(_f32logits(c_unl)),
              None if c_rep is None else _p(_f32logits(c_rep)), _p(_f32logits(d_unl_logits)), c_unl.rows,
              _p(_f32logits(c_fake)), _p(y_g.data), c_fake.rows, K, _p(lambdas), _p(val), _p(g_real), _p(g_unl),
              _p(g_rep), _p(g_fake), _st())
    return Loss(val, [(c_real, g_real), (c_unl, g_unl), (c_rep, g_rep), (c_fake, g_fake)])


def loss_d_grouped(logits, nr, nf, nu, out=None):
    """d_loss on one grouped D pass: rows [0,nr) real, [nr,nr+nf) fake, [nr+nf, ...) unlabelled (train_base.py:123-126)"""
    if ctx.building:
        return Loss(None, [])
    x = _f32logits(logits)
    val, g = (out if out is not None else _new((1,), torch.float32)), _new(logits.shape, torch.float32)
    _lib.call('tgan_loss_d', x.data_ptr(), nr, x.data_ptr() + 4 * nr, nf, x.data_ptr() + 4 * (nr + nf), nu, _p(val),
              g.data_ptr(), g.data_ptr() + 4 * nr, g.data_ptr() + 4 * (nr + nf), _st())
    return Loss(val, [(logits, g)])


def loss_c_grouped(logits, segs, has_rep, y_l_c, d_unl_logits, y_g, lambdas, out=None):
    """c_loss on one grouped C pass with rows ordered [real | unl | (rep) | fake] (train_base.py:130-152)"""
    if ctx.building:
        return Loss(None, [])
    K = logits.shape[1]
    x = _f32logits(logits)
    offs = [0]
    for n in segs:
        offs.append(offs[-1] + n)
    val, g = (out if out is not None else _new((1,), torch.float32)), _new(logits.shape, torch.float32)
    at = lambda t, i: t.data_ptr() + 4 * K * offs[i]
    i_fake = 3 if has_rep else 2
    _lib.call('tgan_loss_c', at(x, 0), _p(y_l_c.data), segs[0], at(x, 1), at(x, 2) if has_rep else None,
              _p(_f32logits(d_unl_logits)), segs[1], at(x, i_fake), _p(y_g.data), segs[i_fake], K, _p(lambdas), _p(val),
              at(g, 0), at(g, 1), at(g, 2) if has_rep else None, at(g, i_fake), _st())
    return Loss(val, [(logits, g)])


def backward(loss):
    """tf.gradients(loss, var_list) for the ops recorded on the current tape."""
    loss.seed()
    ctx.tape.backward()
