"""Train: the Triple-GAN training step of the reference's Training/Train_goodGAN.py (class Train :26-454)
on the sm_100a path.

What is mirrored: graph construction `_build_train_graph` (:400-426), variable partition by name
substring + three Adam optimisers + EMA over the classifier (:78-103), and the D -> G -> C step
(:249-276) with the per-`sess.run` pruning TF applies: each phase executes only the
passes its loss depends on, draws fresh noise per pass, and updates only its own `var_list`.
Epoch bookkeeping, summaries, checkpoints and the tf.data pipeline are out of scope.

Data parallelism (new; the reference is single-GPU): one process per GPU, each rank runs the full
per-rank batch tuple, the flat fp32 gradient buffer of the phase's network is all-reduced (NCCL, sum)
and 1/world is folded into the fused Adam kernel.
"""
import os

import numpy as np
import torch

from . import _lib, ddp, ops
from .core import VariableStore, building, ctx, no_grad, recording
from .train_base import AdamOptimizer, Train_base

INPUT_NAMES = ('z_g', 'y_g', 'x_l_c', 'y_l_c', 'x_l_d', 'y_l_d', 'x_u_d', 'x_u_c')


class ExponentialMovingAverage:
    """tf.train.ExponentialMovingAverage(decay).apply(c_vars) (Train_goodGAN.py:101-103): shadow buffer
    updated by the fused Adam kernel; no debias, no num_updates."""

    def __init__(self, decay):
        self.decay = decay
        self.shadow = None
        self._fb = None

    def bind(self, fb):
        self._fb = fb
        self.shadow = fb['theta'].clone()

    def average(self, param):
        i = self._fb['params'].index(param)
        o = self._fb['offsets'][i]
        return self.shadow[o:o + param.size].view(param.shape)


class HostFeed:
    """Double-buffered host -> device input feed for Train.step (the reference feeds numpy batches through feed_dict on
    every sess.run, Train_goodGAN.py:249-276).  `step(next_batch)` runs one iteration on the batch uploaded by the
    PREVIOUS call while `next_batch` travels to the device on a copy stream, and returns the losses of the previous
    iteration (their device-to-host copy finished long ago), so neither the H2D copy nor the loss read-back stalls the
    GPU.  Usage:
        feed = tr.host_feed(); feed.prime(batch0)
        for b in batches[1:]: losses_prev = feed.step(b, lambda_1=..., lambda_2=...)
        last = feed.drain()
    """

    def __init__(self, trainer):
        self.tr = trainer
        n = trainer.input_flat.numel()
        self.staging = [torch.empty(n, dtype=torch.float32, device=ctx.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream()
        self.uploaded = [torch.cuda.Event(), torch.cuda.Event()]      # staging[i] holds a complete batch
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]      # the step has copied staging[i] away
        self.loss_host = torch.zeros((2, 3), dtype=torch.float32).pin_memory()
        self.loss_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.cur, self.k, self.primed = 0, 0, False
        for e in self.consumed:
            e.record()

    def _upload(self, batch, slot):
        tr = self.tr
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])
            for name in INPUT_NAMES:
                t = batch[name]
                if isinstance(t, np.ndarray):
                    t = torch.from_numpy(t)
                o = tr.input_offsets[name]
                self.staging[slot][o:o + t.numel()].copy_(t.reshape(-1), non_blocking=True)
            self.uploaded[slot].record(self.copy_stream)

    def prime(self, batch):
        self._upload(batch, self.cur)
        self.primed = True

    def step(self, next_batch=None, **kw):
        """one iteration on the batch uploaded last; uploads `next_batch` meanwhile.  Returns the PREVIOUS iteration's
        [d_loss, g_loss, c_loss] as a host tensor (None on the first call)."""
        assert self.primed, 'HostFeed.prime(batch) first'
        tr, cur = self.tr, self.cur
        st = torch.cuda.current_stream()
        st.wait_event(self.uploaded[cur])
        tr.input_flat.copy_(self.staging[cur], non_blocking=True)      # one 2.5 MB device-to-device copy
        self.consumed[cur].record(st)
        if next_batch is not None:
            self._upload(next_batch, 1 - cur)
        out = tr.step(**kw)
        slot = self.k & 1
        prev = None
        if self.k > 0:
            self.loss_ready[1 - slot].synchronize()
            prev = self.loss_host[1 - slot].clone()
        self.loss_host[slot].copy_(out, non_blocking=True)
        self.loss_ready[slot].record(st)
        self.cur, self.k = 1 - cur, self.k + 1
        self.primed = next_batch is not None
        return prev

    def drain(self):
        """losses of the last iteration (synchronises)"""
        slot = (self.k - 1) & 1
        self.loss_ready[slot].synchronize()
        return self.loss_host[slot].clone()


class Train(Train_base):
    def __init__(self, config, log_dir=None, save_dir=None, **kwargs):
        super(Train, self).__init__()
        self.config = config
        self.comments = kwargs.get('comments', '')
        self.store = VariableStore(seed=kwargs.get('seed', 1234))
        self.model = None
        self.graph = None
        self.world = 1
        self.pg = None
        self.last_losses = None

    # ------------------------------------------------------------------ graph construction --------
    def input_shapes(self):
        c = self.config
        img = list(c.IMAGE_DIM)
        return dict(z_g=[c.BATCH_SIZE_G, c.Z_DIM], y_g=[c.BATCH_SIZE_G, c.NUM_CLASSES],
                    x_l_c=[c.BATCH_SIZE_L_C] + img, y_l_c=[c.BATCH_SIZE_L_C, c.NUM_CLASSES],
                    x_l_d=[c.BATCH_SIZE_L_D] + img, y_l_d=[c.BATCH_SIZE_L_D, c.NUM_CLASSES],
                    x_u_d=[c.BATCH_SIZE_U_D] + img, x_u_c=[c.BATCH_SIZE_U_C] + img)

    def _build_train_graph(self, Model):
        """Train_goodGAN.py:400-426: placeholders -> Model(config).forward_pass(...).  Runs in shape-only
        mode: creates every variable under its TF name, launches nothing (works without a GPU)."""
        ctx.store = self.store
        self.model = Model(self.config)
        with building(), no_grad():
            ph = {k: ops.Var(None, tuple(v)) for k, v in self.input_shapes().items()}
            G, D, C = self.model.forward_pass(ph['z_g'], ph['y_g'], ph['x_l_c'], ph['y_l_c'], ph['x_l_d'],
                                              ph['y_l_d'], ph['x_u_d'], ph['x_u_c'], True)
        t_vars = self.store.trainable()
        self.g_vars = [v for v in t_vars if 'good_generator' in v.name]
        self.d_vars = [v for v in t_vars if 'discriminator' in v.name]
        self.c_vars = [v for v in t_vars if 'classifier' in v.name]
        return ph, G, D, C, self.model

    def initialize(self, init=None):
        """global_variables_initializer (Train_goodGAN.py:154-156) or injected values
        (`init=(P, S)` numpy dicts keyed by TF names), then the optimisers of :85-103."""
        if self.model is None:
            raise RuntimeError('call _build_train_graph(Model) first')
        if init is not None:
            self.store.load_numpy(*init)
        # data parallel: parameters / gradients in peer-addressable memory for the fused update (ddp.FusedUpdate)
        self.store.alloc = ddp.symmetric_allocator(ctx.device) if ctx.device.type == 'cuda' else None
        self.store.finalize(ctx.device)
        c = self.config
        self.d_optimizer = self._Adam_optimizer(lr=c.LEARNING_RATE, beta1=c.BETA1)
        self.g_optimizer = self._Adam_optimizer(lr=c.LEARNING_RATE, beta1=c.BETA1)
        self.c_optimizer = self._Adam_optimizer(lr=c.CLA_LEARNINIG_RATE, beta1=0.5)
        self.ema = ExponentialMovingAverage(decay=0.9999)
        self.lambdas = torch.zeros(2, dtype=torch.float32, device=ctx.device)
        self._lam = (None, None)
        # the eight step inputs are views of ONE flat buffer (16-byte aligned each): a double-buffered host feed moves a
        # whole batch with one device-to-device copy (HostFeed)
        shp = self.input_shapes()
        offs, tot = {}, 0
        for k in INPUT_NAMES:
            offs[k] = tot
            tot += (int(np.prod(shp[k])) + 3) // 4 * 4
        self.input_flat = torch.zeros(tot, dtype=torch.float32, device=ctx.device)
        self.input_offsets = offs
        self.inputs = {k: self.input_flat[offs[k]:offs[k] + int(np.prod(shp[k]))].view(tuple(shp[k])) for k in INPUT_NAMES}
        # everything a CUDA-graph replay must NOT re-initialise is created here, outside any capture
        for o in (self.d_optimizer, self.g_optimizer, self.c_optimizer):
            o._state()
        self.loss_buf = torch.zeros(3, dtype=torch.float32, device=ctx.device)
        ctx.ws()
        ctx.arena()
        if not ctx.rng.injected:
            ctx.rng.counter()
        if c.DATA_NAME == 'cifar10' and hasattr(self.model, '_whitener'):
            self.model._whitener()._upload()
        self.world = ddp.world_size()
        ddp.broadcast_params(self.store, 0, self.pg)
        self.fused_dp = ddp.FusedUpdate(self.store, ctx.device, self.pg) if self.store.alloc is not None else None
        # the EMA shadow starts from the parameters every rank actually trains with (rank 0's after the broadcast)
        self.ema.bind(self.store.flat['classifier'])
        return self

    def load_state(self, P=None, S=None, adam=None):
        """Overwrite parameters (P), state variables (S: pop_mean / BN moving statistics) and Adam slots
        (adam = {network: (m, v)}), all dicts keyed by TF variable name.  Used for checkpoint import
        (Training/Saver.py:14-66 names) and by the teacher-forced multi-step parity tests."""
        self.store.load_numpy(P or {}, S or {})
        for grp, (m, v) in (adam or {}).items():
            fb = self.store.flat[grp]
            for p, o in zip(fb['params'], fb['offsets']):
                for slot, src in (('m', m), ('v', v)):
                    if p.name in src:
                        t = torch.as_tensor(np.asarray(src[p.name], np.float32)).reshape(-1)
                        fb[slot][o:o + p.size].copy_(t)
        self.store.bump()
        return self

    # ------------------------------------------------------------------ one step ------------------
    def _set_scalars(self, lambda_1, lambda_2, lr, cla_lr):
        if (lambda_1, lambda_2) != self._lam:
            self._lam = (lambda_1, lambda_2)
            _lib.call('tgan_fill_f32', self.lambdas.data_ptr(), float(lambda_1), 1, ops._st())
            _lib.call('tgan_fill_f32', self.lambdas.data_ptr() + 4, float(lambda_2), 1, ops._st())
        if lr is not None:
            self.d_optimizer.set_lr(lr)
            self.g_optimizer.set_lr(lr)
        if cla_lr is not None:
            self.c_optimizer.set_lr(cla_lr)

    def _begin(self, group, var_list):
        for p in self.store.vars.values():
            p.requires_grad = False
        for p in var_list:
            p.requires_grad = True
        fb = self.store.flat[group]
        _lib.call('tgan_fill_f32', fb['grad'].data_ptr(), 0.0, fb['n'], ops._st())
        ops.arena_reset()
        return fb

    def _apply(self, fb, opt, ema=None, group=None):
        if getattr(self, 'fused_dp', None) is not None:
            # reduce-scatter -> Adam on the owned shard -> all-gather of the parameters, one kernel over NVLink peer memory
            self.fused_dp.apply(group, fb, opt, ema.shadow if ema is not None else None, ema.decay if ema is not None else 0.0)
            self.store.bump(group)
            return
        if group == 'classifier' and self.world > 1 and ctx.math == 'bf16' and self.config.DATA_NAME == 'cifar10' \
                and hasattr(self.model, '_whitener'):
            # the classifier's gradient travels in two buckets; the tail one was started behind its own backward pass
            if getattr(self, '_c_buckets', None) is None:
                self._c_buckets = ddp.BucketedAllReduce(self.store, 'classifier', 'classifier/conv2_2/V', self.pg).install()
            scale = self._c_buckets.finish()
        else:
            scale = ddp.allreduce_grads(fb['grad'], self.pg)      # NCCL sum over NVLink; 1/world folded into Adam
        opt.apply_flat(fb, scale, ema.shadow if ema is not None else None, ema.decay if ema is not None else 0.0)
        self.store.bump(group)

    def _pre(self):
        m = self.model
        if self.config.DATA_NAME == 'cifar10' and hasattr(m, '_whitener'):
            return m._whitener().apply
        return lambda t: t

    def _step_impl(self, train=True, phases='DGC'):
        """The three phases with TF's per-`sess.run` pruning, scheduled for the GPU:

        * calls of the same network inside a phase are GROUPED into one batch (ops.group_batch): per-sample work
          runs once on the group; batch statistics (mean-only BN) stay per call through the batch segments, so every
          value equals the separate calls of the reference graph.  D has no batch statistics at all, so its three
          phase-D calls are one pass over 250 samples.  (Models with tf.contrib batch_norm in C keep separate calls.)
        * the generator forward of phase G is the phase-D forward: same weights G0, same z / y, no stochastic op, so
          the tensors are identical; it is recorded once on its own tape and differentiated in phase G.
        """
        m, c = self.model, self.config
        v = {k: ops.Var(t, tuple(t.shape)) for k, t in self.inputs.items()}
        pre, K = self._pre(), c.NUM_CLASSES
        cif = c.DATA_NAME == 'cifar10' and hasattr(m, '_whitener')
        # grouped classifier passes: segment-aware mean-only BN (cifar10) / per-segment tf.contrib batch norm (Good_GAN)
        grouped_c = not os.environ.get('TGAN_NO_GROUPING')
        nLD, nUD, nUC, nG, nLC = (c.BATCH_SIZE_L_D, c.BATCH_SIZE_U_D, c.BATCH_SIZE_U_C, c.BATCH_SIZE_G, c.BATCH_SIZE_L_C)
        TL = ops.TagList
        d_loss = g_loss = None
        if phases == 'C':
            self.aux = {}
            _lib.call('tgan_fill_f32', self.loss_buf.data_ptr(), 0.0, 2, ops._st())      # no d / g loss in a PRE_TRAIN iteration
        if 'D' in phases or 'G' in phases:
            # ---- phase D: sess.run([d_solver, d_loss]) (:267) ----
            ops.arena_reset()
            # the generator forward (a chain of small kernels) runs on the side stream beside the classifier forward
            for p in self.g_vars:                # record G(z, y) once, on its own tape, for phase G's backward
                p.requires_grad = True
            with ops.side_stream(), recording() as tape_g:
                G = m.good_generator(v['z_g'], v['y_g'], reuse=True, tag='D/G')
                G.data
            with no_grad():
                if grouped_c:
                    lg, _ = m.classifier(pre(ops.group_batch([v['x_u_d'], v['x_u_c']])), train, reuse=True,
                                         tag=TL([('D/C_unl_d', nUD), ('D/C_unl', nUC)]))
                    idx, oh = ops.argmax_onehot(lg, K)
                    idx_d, idx_u = ops.Var(idx.data[:nUD], (nUD,)), ops.Var(idx.data[nUD:], (nUC,))
                    oh_d, oh_u = ops.Var(oh.data[:nUD], (nUD, K)), ops.Var(oh.data[nUD:], (nUC, K))
                else:
                    c_unl_d, _ = m.classifier(pre(v['x_u_d']), train, reuse=True, tag='D/C_unl_d')
                    c_unl, _ = m.classifier(pre(v['x_u_c']), train, reuse=True, tag='D/C_unl')
                    idx_d, oh_d = ops.argmax_onehot(c_unl_d, K)
                    idx_u, oh_u = ops.argmax_onehot(c_unl, K)
            ops.join_side()
            G_const = ops.Var(G.data, G.shape)   # phase D sees the generated images as constants (var_list = d_vars)
            fb = self._begin('discriminator', self.d_vars)
            with recording():
                X = ops.group_batch([v['x_l_d'], v['x_u_d'], G_const, v['x_u_c']])
                Y = ops.group_batch([v['y_l_d'], oh_d, v['y_g'], oh_u])
                X.aux = None                     # D is per-sample: one plain batch of 250
                _, dl = m.discriminator(X, Y, reuse=True, tag=TL([('D/D_real', nLD + nUD), ('D/D_fake', nG), ('D/D_unl', nUC)]))
                d_loss = ops.loss_d_grouped(dl, nLD + nUD, nG, nUC, out=self.loss_buf[0:1])
                ops.backward(d_loss)
            self._apply(fb, self.d_optimizer, group='discriminator')
            self.aux = dict(idx_unl_d=idx_d, idx_unl=idx_u, G_phaseD=G, d_logits=dl)
            # ---- phase G: sess.run([g_solver, g_loss]) (:270) ----
            fb = self._begin('good_generator', self.g_vars)
            with recording() as tape_d:
                _, df = m.discriminator(G, v['y_g'], reuse=True, tag='G/D_fake')
                g_loss = ops.loss_g(df, out=self.loss_buf[1:2])
                g_loss.seed()
                tape_d.backward()                # d g_loss / d G through D1 (dgrad only)
            tape_g.backward()                    # ... and through the generator recorded in phase D
            self._apply(fb, self.g_optimizer, group='good_generator')
        # ---- phase C: sess.run([c_solver, c_loss]) (:275) ----
        fb = self._begin('classifier', self.c_vars)
        with no_grad():
            G = m.good_generator(v['z_g'], v['y_g'], reuse=True, tag='C/G')
        with recording():
            if grouped_c:
                Gv = ops.Var(G.data, G.shape)      # (group_batch takes any per-sample view: [N,784] beside [N,28,28,1])
                if cif:      # Good_GAN_cifar10.forward_pass: a second stochastic pass over x_u_c (:233)
                    segs = [nLC, nUC, nUC, nG]
                    xs = ops.group_batch([v['x_l_c'], v['x_u_c'], v['x_u_c'], Gv])
                    tags = TL([('C/C_real', nLC), ('C/C_unl', nUC), ('C/C_unl_rep', nUC), ('C/C_fake', nG)])
                else:
                    segs = [nLC, nUC, nG]
                    xs = ops.group_batch([v['x_l_c'], v['x_u_c'], Gv])
                    tags = TL([('C/C_real', nLC), ('C/C_unl', nUC), ('C/C_fake', nG)])
                lg, _ = m.classifier(pre(xs), train, reuse=True, tag=tags)
                c_unl_v = ops.Var(lg.data[nLC:nLC + nUC], (nUC, K))
                with no_grad():
                    idx_c, oh_u = ops.argmax_onehot(c_unl_v, K)
                    _, du = m.discriminator(v['x_u_c'], oh_u, reuse=True, tag='C/D_unl')
                c_loss = ops.loss_c_grouped(lg, segs, cif, v['y_l_c'], du, v['y_g'], self.lambdas, out=self.loss_buf[2:3])
                self.aux['c_logits'] = lg
            else:
                c_real, _ = m.classifier(pre(v['x_l_c']), train, reuse=True, tag='C/C_real')
                c_unl, _ = m.classifier(pre(v['x_u_c']), train, reuse=True, tag='C/C_unl')
                c_rep = m.classifier(pre(v['x_u_c']), train, reuse=True, tag='C/C_unl_rep')[0] if cif else None
                with no_grad():
                    idx_c, oh_u = ops.argmax_onehot(c_unl, K)
                    _, du = m.discriminator(v['x_u_c'], oh_u, reuse=True, tag='C/D_unl')
                c_fake, _ = m.classifier(pre(G), train, reuse=True, tag='C/C_fake')
                c_loss = ops.loss_c(c_real, v['y_l_c'], c_unl, c_rep, du, c_fake, v['y_g'], self.lambdas,
                                    out=self.loss_buf[2:3])
                self.aux['c_logits'] = (c_real, c_unl, c_fake, c_rep)
            self.aux['idx_unl_c'] = idx_c
            ops.backward(c_loss)
        self._apply(fb, self.c_optimizer, self.ema, group='classifier')
        if not ctx.rng.injected:
            _lib.call('tgan_counter_advance', ctx.rng.counter().data_ptr(), 1, ops._st())
        return d_loss, g_loss, c_loss      # the three loss kernels wrote their scalars straight into self.loss_buf

    def load_batch(self, batch):
        """Copy one step's inputs (numpy / CPU / device tensors) into the static device buffers."""
        for k in INPUT_NAMES:
            t = batch[k]
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(t)
            self.inputs[k].copy_(t.reshape(self.inputs[k].shape), non_blocking=True)

    def step(self, batch=None, lambda_1=None, lambda_2=0.0, lr=None, cla_lr=None, train=True, phases='DGC'):
        """One training iteration (Train_goodGAN.py:249-276).  Returns the device tensor
        [d_loss, g_loss, c_loss] (values before each phase's update); no host synchronisation.
        phases='C' is the PRE_TRAIN iteration (:182-224): only `sess.run([c_solver, c_loss])`."""
        ctx.store = self.store
        if batch is not None:
            self.load_batch(batch)
        lam1 = self.config.FAKE_G_LAMBDA if lambda_1 is None else lambda_1
        self._set_scalars(lam1, lambda_2, lr, cla_lr)
        if self.graph is not None and phases == 'DGC':
            self.graph.replay()
        else:
            self._step_impl(train, phases)
        return self.loss_buf

    def _snapshot(self):
        """every tensor a training step mutates: parameters, Adam slots, population / moving statistics, the optimisers'
        {lr, beta1^t, beta2^t} accumulators, the EMA shadow, the Philox step counter, the loss buffer"""
        ts = [fb[k] for fb in self.store.flat.values() for k in ('theta', 'm', 'v') if k in fb]
        ts += [o._state() for o in (self.d_optimizer, self.g_optimizer, self.c_optimizer)]
        ts += [self.ema.shadow, self.loss_buf]
        if not ctx.rng.injected:
            ts.append(ctx.rng.counter())
        return ts

    def host_feed(self):
        """-> HostFeed: the step driven from HOST batches without stalling the GPU (upload of batch k+1 overlaps step k,
        losses are read one step late)."""
        return HostFeed(self)

    def capture(self, warmup=3):
        """Capture the whole three-phase step into one CUDA graph (the step is a few hundred small
        launches, SURVEY.md §7 'Launch-bound regime').  Inputs are read from the static buffers.

        The `warmup` eager steps exist only to create every lazily allocated buffer and cache before the capture;
        they are REAL training steps, so everything they mutate (parameters, Adam slots and beta powers, pop_mean /
        BN moving statistics, EMA shadow, RNG counter) is saved before and restored after: capture() leaves the
        training state -- a fresh initialisation or a restored checkpoint -- bitwise unchanged, whatever the static input
        buffers hold at that moment (zeros before the first load_batch)."""
        assert not ctx.rng.injected, 'graph capture needs the in-kernel Philox RNG'
        ctx.store = self.store
        live = self._snapshot()
        saved = [t.clone() for t in live]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._step_impl(True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if warmup:
            for t, c in zip(live, saved):
                t.copy_(c)
            self.store.bump()                    # cached weight-norm scales / packed operands belong to the warm-up weights
            torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _lib.load().tgan_launch_count()
        with torch.cuda.graph(g):
            self._step_impl(True)
        self.launches_per_step = _lib.load().tgan_launch_count() - n0
        self.graph = g
        return g

    def close(self):
        """Orderly shutdown: drop the captured graph (it may hold collective / peer-memory launches) BEFORE the process
        group goes away, wait for the device, then leave the group.  Safe to call more than once."""
        self.graph = None
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.fused_dp = None
        self._c_buckets = None
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()

    # ------------------------------------------------------------------ evaluation --
    def evaluate(self, x, y, reset=True):
        """Validation pass of Train_goodGAN.py:296-351 with the metric of :428-447: the classifier with train=False
        (mean-only BN / BN use their population statistics, dropout is off, the Gaussian input-noise layer stays active as
        in the reference graph), streaming accuracy of argmax(C(x)) against argmax(y).  x, y: numpy / tensors of one
        validation batch.  Returns (streaming accuracy, predictions as an int64 device tensor)."""
        ctx.store = self.store
        if reset or not hasattr(self, '_acc'):
            self._acc = [0, 0]                  # tf.metrics.accuracy's (total, count) local variables
        tx = torch.as_tensor(np.asarray(x, np.float32)).to(ctx.device)
        ty = torch.as_tensor(np.asarray(y, np.float32)).to(ctx.device)
        ops.arena_reset()
        with no_grad():
            logits, _ = self.model.classifier(self._pre()(ops.Var(tx, tuple(tx.shape))), False, reuse=True, tag='V/C_real')
            idx, _ = ops.argmax_onehot(logits, self.config.NUM_CLASSES)
        self._acc[0] += int((idx.data == ty.argmax(dim=1)).sum().item())
        self._acc[1] += int(tx.shape[0])
        self.aux_val = dict(logits=logits)
        return self._acc[0] / max(1, self._acc[1]), idx.data

    def sample(self, z, y, grid=True):
        """The per-epoch sample of Train_goodGAN.py:353-364: `model.good_sampler(sample_z, sample_y)` (train=False) and,
        with grid=True, utils.save_images' 8x8 manifold of the first 64 images in [0, 1] (device tensor; writing the
        PNG is left to the caller).  z [n, Z_DIM], y [n, NUM_CLASSES]: numpy / tensors."""
        from . import pipeline
        ctx.store = self.store
        tz = torch.as_tensor(np.asarray(z, np.float32)).to(ctx.device)
        ty = torch.as_tensor(np.asarray(y, np.float32)).to(ctx.device)
        ops.arena_reset()
        with no_grad():
            g = self.model.good_sampler(ops.Var(tz, tuple(tz.shape)), ops.Var(ty, tuple(ty.shape)))
            img = ops.force(g).data.float().reshape([-1] + list(self.config.IMAGE_DIM)).contiguous()
        if not grid:
            return img
        n = min(64, int(img.shape[0]))
        return pipeline.image_grid(img[:n], pipeline.image_manifold_size(n))

    # ------------------------------------------------------------------ epoch driver --
    def train_epoch(self, batches, epoch, start_epoch=0):
        """One epoch of Train_goodGAN.py:160-276: the lambda / learning-rate schedule of :165-177, the classifier-only
        PRE_TRAIN iterations for the first 30 epochs when config.PRE_TRAIN is set (:182-224), the full D -> G -> C
        iteration otherwise.  `batches` yields dicts with the 8 step inputs; returns the mean (d, g, c) losses."""
        lam1, lam2, lr, cla_lr = self.schedule(epoch, start_epoch)
        pre = bool(getattr(self.config, 'PRE_TRAIN', False)) and (start_epoch + epoch <= 30)
        tot, n = None, 0
        for b in batches:
            out = self.step(b, lambda_1=lam1, lambda_2=lam2, lr=lr, cla_lr=cla_lr, phases='C' if pre else 'DGC')
            # epoch bookkeeping, accumulated where the losses live: no host synchronisation inside the epoch
            tot = out.detach().double() if tot is None else tot + out.detach().double()
            n += 1
        return [0.0, 0.0, 0.0] if tot is None else (tot.cpu() / n).tolist()

    # ------------------------------------------------------------------ schedules ------------------
    def schedule(self, epoch, start_epoch=0):
        """lambda_1 / lambda_2 / lr schedule of Train_goodGAN.py:165-177 for (1-based) `epoch`."""
        c = self.config
        lam1 = c.FAKE_G_LAMBDA if (start_epoch + epoch) > 200 else 0.
        lam2 = (0.5 if epoch > 67 else 0) if c.DATA_NAME == 'cifar10' else 0.0
        n = max(0, start_epoch + epoch - 299)
        return lam1, lam2, c.LEARNING_RATE * 0.995 ** n, c.CLA_LEARNINIG_RATE * 0.99 ** n
