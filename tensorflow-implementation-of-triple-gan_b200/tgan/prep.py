"""Multi-tensor weight preparation for the tensor-core mode: per network (variable group) and optimiser version ONE
launch computes every weight-norm scale (nn.py:502,554; modle_base.py:66,101,148), ONE launch packs every bf16
operand the tcgen05 kernels read (fprop / dgrad / parity-class layouts, weight-norm scale folded in), and ONE launch
per backward pass turns the accumulated dW of all weight-normalised layers into dV / dg.

Entries register themselves the first time a layer runs (the warm-up steps); from then on the descriptor tables are
static device arrays -- parameter pointers are views into the flat per-network buffers and never change, and the
packed / scratch buffers are allocated once -- so CUDA-graph replays find everything in place.
"""
import ctypes

import torch

from . import _lib
from .core import ctx


def _st():
    return torch.cuda.current_stream().cuda_stream


def to_device_table(host):
    """small host table -> device, legal under CUDA-graph capture (the first step of a trainer may itself be captured):
    staged through PINNED memory that stays alive with the device tensor, so the copy is a replayable memcpy node"""
    pinned = host.contiguous().pin_memory()
    dev = torch.empty(pinned.shape, dtype=pinned.dtype, device=ctx.device)
    dev.copy_(pinned, non_blocking=True)
    _keep_alive.append((pinned, dev))
    return dev


_keep_alive = []


def _table(descs):
    raw = b''.join(bytes(d) for d in descs)
    return to_device_table(torch.frombuffer(bytearray(raw), dtype=torch.uint8))


def _groups(kind):
    reg = getattr(ctx.store, '_prep', None)
    if reg is None:
        reg = ctx.store._prep = {'wn': {}, 'pack': {}}
    return reg[kind]


class WNGroup:
    """all weight-normalised tensors of one network"""

    def __init__(self, group):
        self.group, self.entries, self.version = group, {}, None
        self.table, self.max_co, self.bwd_tape, self.bwd_tables = None, 0, None, {}
        # dW accumulators mirror the network's flat parameter buffer (same offsets as V), so one fill zeroes them all
        fb = ctx.store.flat.get(group) if ctx.store is not None else None
        self.fb = fb
        self.dWflat = torch.zeros(fb['n'], dtype=torch.float32, device=ctx.device) if fb is not None else None

    @staticmethod
    def of(group):
        g = _groups('wn')
        if group not in g:
            g[group] = WNGroup(group)
        return g[group]

    def entry(self, w):
        e = self.entries.get(w.V)
        if e is None:
            n = w.V.size
            off = None
            if self.fb is not None:
                off = (w.V.data.data_ptr() - self.fb['theta'].data_ptr()) // 4
                if not (0 <= off and off + n <= self.fb['n']):
                    off = None
            dW = self.dWflat[off:off + n] if off is not None else torch.zeros(n, dtype=torch.float32, device=ctx.device)
            e = dict(w=w, inv=torch.empty(w.Co, dtype=torch.float32, device=ctx.device),
                     scale=torch.empty(w.Co, dtype=torch.float32, device=ctx.device), dW=dW.view(w.V.shape),
                     flat=off is not None, off=off)
            self.entries[w.V] = e
            self.table = None
            if self.version == ctx.store.group_version(self.group):      # the group was prepared without this tensor
                self._launch('tgan_weightnorm_fwd_multi', [e])
        return e

    def _desc(self, e):
        w = e['w']
        d = _lib.TganWnDesc()
        d.V, d.g, d.inv_norm, d.scale = w.V.data.data_ptr(), w.g.data.data_ptr(), e['inv'].data_ptr(), e['scale'].data_ptr()
        d.dW = e['dW'].data_ptr()
        d.dV = w.V.grad.data_ptr() if w.V.grad is not None else 0
        d.dg = w.g.grad.data_ptr() if w.g.grad is not None else 0
        d.A, d.Co, d.B, d.eps_mode = w.A, w.Co, w.B, w.eps_mode
        return d

    def _launch(self, fn, entries=None):
        if entries is None:
            if self.table is None:
                es = list(self.entries.values())
                self.table, self.max_co = _table([self._desc(e) for e in es]), max(e['w'].Co for e in es)
            t, n, mc = self.table, len(self.entries), self.max_co
        else:
            t, n, mc = _table([self._desc(e) for e in entries]), len(entries), max(e['w'].Co for e in entries)
            self._keep = t
        assert _lib.load().tgan_weightnorm_bwd_multi_ws_floats(n, mc) <= ctx.ws().numel()
        _lib.call(fn, t.data_ptr(), n, mc, ctx.ws().data_ptr(), _st())

    def prepare(self):
        """scales of every registered tensor for the current optimiser version (one launch)"""
        ver = ctx.store.group_version(self.group)
        if self.version != ver and self.entries:
            self._launch('tgan_weightnorm_fwd_multi')
        self.version = ver

    def grad_target(self, w, tape):
        """dW accumulator of w for this backward pass; the first request of a pass zeroes the accumulators of the
        whole network and schedules the one weight-norm backward launch that closes the pass"""
        e = self.entry(w)
        self.prepare()
        if self.bwd_tape is not tape:
            self.bwd_tape = tape
            if self.dWflat is not None:
                _lib.call('tgan_fill_f32', self.dWflat.data_ptr(), 0.0, self.dWflat.numel(), _st())
            for x in self.entries.values():
                if not x['flat']:
                    _lib.call('tgan_fill_f32', x['dW'].data_ptr(), 0.0, x['dW'].numel(), _st())
            me = self
            self.done = set()

            def post():
                # every tensor registered by now (tensors that join during the pass start from zeroed accumulators),
                # minus the ones an early gradient bucket already converted (flush_bucket)
                me._run_bwd([x for x in me.entries.values() if id(x) not in me.done])
                me.bwd_tape = None
            tape.post.append(post)
        hook = getattr(self, 'bucket_hook', None)
        if hook is not None:
            hook(self, e, tape)
        return e['dW']

    def _run_bwd(self, entries):
        """dV / dg of `entries` from their accumulated dW: one multi-tensor launch"""
        live = [x for x in entries if x['w'].V.grad is not None and x['w'].g.grad is not None]
        if not live:
            return
        key = tuple(id(x['w'].V) for x in live)
        tbl = self.bwd_tables.get(key)
        if tbl is None:
            tbl = self.bwd_tables[key] = (_table([self._desc(x) for x in live]), len(live), max(x['w'].Co for x in live))
        assert _lib.load().tgan_weightnorm_bwd_multi_ws_floats(tbl[1], tbl[2]) <= ctx.ws().numel()
        _lib.call('tgan_weightnorm_bwd_multi', tbl[0].data_ptr(), tbl[1], tbl[2], ctx.ws().data_ptr(), _st())

    def flush_bucket(self, pred):
        """convert the dW of the entries selected by pred NOW (their filter gradients are complete): the data-parallel
        step all-reduces that part of the flat gradient buffer while the rest of the backward pass still runs"""
        sel = [x for x in self.entries.values() if id(x) not in self.done and pred(x)]
        self._run_bwd(sel)
        self.done.update(id(x) for x in sel)


class PackGroup:
    """all packed bf16 operands of one network"""

    def __init__(self, group):
        self.group, self.entries, self.version, self.table = group, {}, None, None

    @staticmethod
    def of(group):
        g = _groups('pack')
        if group not in g:
            g[group] = PackGroup(group)
        return g[group]

    def get(self, w, key, T, Nrows, K, st, sn, sk, taps_dev, scale_mod=0):
        from .ops import WNWeight
        ver = ctx.store.group_version(self.group)
        wn = isinstance(w, WNWeight)
        k = (w.key, key)      # keyed by the Param object itself (ids are recycled)
        e = self.entries.get(k)
        if e is None:
            Kpad = (K + 7) // 8 * 8
            dst = torch.empty((T, Nrows, Kpad), dtype=torch.bfloat16, device=ctx.device)
            d = _lib.TganPackDesc()
            d.src, d.dst, d.taps = w.key.data.data_ptr(), dst.data_ptr(), (0 if taps_dev is None else taps_dev.data_ptr())
            d.st, d.sn, d.sk, d.T, d.Nr, d.K, d.Kpad = st, sn, sk, T, Nrows, K, Kpad
            if wn:
                we = WNGroup.of(self.group).entry(w)
                WNGroup.of(self.group).prepare()
                d.scale = we['scale'].data_ptr()
                assert (sn == w.B) != (sk == w.B), 'cannot tell which packed axis is the output channel'
                d.scale_on = 1 if sn == w.B else 2
                d.scale_mod = scale_mod      # k = (tap, co) flattened: the per-channel scale is scale[k % Co]
            e = self.entries[k] = (d, dst, Kpad, taps_dev)
            self.table = None
            if self.version == ver:        # the group was packed without this operand
                one = _table([d])
                self._keep = one
                _lib.call('tgan_pack_weight_multi', one.data_ptr(), 1, _st())
        if self.version != ver:
            if wn or any(x[0].scale_on for x in self.entries.values()):
                WNGroup.of(self.group).prepare()
            if self.table is None:
                self.table = _table([x[0] for x in self.entries.values()])
            _lib.call('tgan_pack_weight_multi', self.table.data_ptr(), len(self.entries), _st())
            self.version = ver
        return e[1], e[2]
