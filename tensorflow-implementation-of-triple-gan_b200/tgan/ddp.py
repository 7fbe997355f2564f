"""Data parallelism for the Triple-GAN step (new: the reference is single-process, single-GPU --
Train_goodGAN.py:48, :724).

One process per GPU.  Every rank runs the full per-rank batch tuple (G 100, L_C 50, U_C 50, L_D 20, U_D 80) drawn
from its own stream; batch statistics (mean-only BN, BN, balance-entropy) stay per rank, which reproduces the
reference's semantics at batch 100 exactly (no sync-BN).  The only exchange is ONE all-reduce (sum) per phase over
the flat fp32 gradient buffer of the network being updated (D 1.3 MB, G 20.5 MB, C 12.5 MB); the 1/world average is
folded into the fused Adam kernel.  NCCL over NVLink on GPUs; the same code runs on gloo for CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Returns (rank, world, local)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_grads(flat_grad, group=None):
    """Sum the flat gradient buffer of one network over all ranks, in place.  Returns the factor the optimiser must
    apply to turn the sum into the data-parallel average (1/world)."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / w


def rank_seed(base, rank):
    """per-rank input / RNG stream (seed 1234 + rank)"""
    return int(base) + int(rank)


def broadcast_params(store, src=0, group=None):
    """Make every rank start from rank `src`'s variables (flat theta + running statistics)."""
    if world_size() > 1:
        for fb in store.flat.values():
            dist.broadcast(fb['theta'], src=src, group=group)


class BucketedAllReduce:
    """Two gradient buckets for the classifier (the network whose backward pass is long enough to hide a collective).

    The flat gradient buffer is ordered by variable creation, i.e. by layer; the backward pass produces it back to
    front.  Bucket A = the tail [split:] (conv2_2 ... output_dense: 81 % of the 12.5 MB) is complete as soon as the
    first filter gradient of an EARLIER layer is requested; at that moment the weight-norm backward of bucket A runs and
    its all-reduce starts on a communication stream, overlapped with the remaining backward work (conv2_1 ... conv1_1,
    ~0.4 ms).  Bucket B = the head [:split] is reduced when the pass ends; Adam waits for both.  Sums are identical to
    the single all-reduce (same elements, same reduction), only the timing changes."""

    def __init__(self, store, group, first_param_of_tail, pg=None):
        fb = store.flat[group]
        self.fb, self.pg, self.group = fb, pg, group
        self.split = None
        for p, o in zip(fb['params'], fb['offsets']):
            if p.name == first_param_of_tail:
                self.split = o
        if self.split is None:
            raise KeyError(first_param_of_tail)
        self.stream = None
        self.done_evt = None
        self.flushed = None

    def install(self):
        from .prep import WNGroup
        grp = WNGroup.of(self.group)
        grp.bucket_hook = self._hook
        return self

    def _hook(self, grp, entry, tape):
        if world_size() <= 1 or self.flushed is tape or entry['off'] is None or entry['off'] >= self.split:
            return
        # first filter gradient of a head layer: every tail layer's dW, bias gradient and g / V inputs are final
        from . import ops
        ops.join_side_if_pending()      # (small tail layers computed theirs on the side stream)
        self.flushed = tape
        grp.flush_bucket(lambda x: x['off'] is not None and x['off'] >= self.split)
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.fb['grad'][self.split:], op=dist.ReduceOp.SUM, group=self.pg)
            self.done_evt = torch.cuda.Event()
            self.done_evt.record(self.stream)

    def finish(self, tape_done=True):
        """all-reduce what is left and make the current stream wait for the early bucket; -> 1 / world"""
        w = world_size()
        if w <= 1:
            return 1.0
        if self.flushed is not None and self.done_evt is not None:
            dist.all_reduce(self.fb['grad'][:self.split], op=dist.ReduceOp.SUM, group=self.pg)
            torch.cuda.current_stream().wait_event(self.done_evt)
        else:
            dist.all_reduce(self.fb['grad'], op=dist.ReduceOp.SUM, group=self.pg)
        self.flushed, self.done_evt = None, None
        return 1.0 / w


def shard_len(n, world):
    """length of one rank's shard of a flat buffer of n elements (n % 4 == 0): a multiple of 4 elements (16-byte vectors),
    the same formula as tgan_dp_adam (csrc/dp_fused.cu); rank r owns [r * len, min(n, (r + 1) * len))"""
    return ((n // 4 + world - 1) // world) * 4


def symmetric_allocator(device):
    """-> alloc(n) returning fp32 tensors in symmetric memory (every rank allocates the same sizes in the same order), or
    None when this is a single-process run / symmetric memory is unavailable (the NCCL all-reduce path is used then)."""
    if world_size() <= 1 or os.environ.get('TGAN_DP_NCCL'):
        return None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        probe = symm_mem.empty(16, dtype=torch.float32, device=device)
        symm_mem.rendezvous(probe, dist.group.WORLD)
    except Exception as e:      # no peer access / old runtime: say so once, keep NCCL
        if dist.get_rank() == 0:
            print('tgan.ddp: symmetric memory unavailable (%s); gradients travel through ncclAllReduce' % str(e)[:200])
        return None
    return lambda n: symm_mem.empty(int(n), dtype=torch.float32, device=device)


class FusedUpdate:
    """reduce-scatter -> Adam on the owned shard -> all-gather of the parameters in ONE kernel over NVLink peer memory
    (csrc/dp_fused.cu) instead of ncclAllReduce + Adam.  Needs the store's theta / grad buffers in symmetric memory
    (VariableStore.alloc = ddp.symmetric_allocator(...) before finalize).  Adam's m / v slots are SHARDED: rank r keeps
    the slice [r*per, (r+1)*per) current; `gather_slots` makes them whole again (checkpoints)."""

    def __init__(self, store, device, pg=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        grp = dist.group.WORLD if pg is None else pg
        self.flags = symm_mem.empty(16, dtype=torch.int32, device=device)
        self.flags.zero_()
        self.epoch = torch.zeros(2, dtype=torch.int32, device=device)
        arr = lambda ptrs: (ctypes.c_uint64 * len(ptrs))(*[int(x) for x in ptrs])
        self.flag_ptrs = arr(symm_mem.rendezvous(self.flags, grp).buffer_ptrs)
        self.ptrs = {}
        for name, fb in store.flat.items():
            if 'grad' not in fb:
                continue
            hg, ht = symm_mem.rendezvous(fb['grad'], grp), symm_mem.rendezvous(fb['theta'], grp)
            mcg = mct = None
            try:      # NVSwitch multicast addresses (NVLS) when the fabric has them
                if not os.environ.get('TGAN_DP_NO_MULTIMEM') and int(hg.multicast_ptr) and int(ht.multicast_ptr):
                    mcg, mct = int(hg.multicast_ptr), int(ht.multicast_ptr)
            except Exception:
                mcg = mct = None
            self.ptrs[name] = (arr(hg.buffer_ptrs), arr(ht.buffer_ptrs), mcg, mct)
        self.multimem = all(p[2] is not None for p in self.ptrs.values())
        torch.cuda.synchronize()
        dist.barrier()

    def shard(self, n):
        return shard_len(n, self.world)

    def apply(self, name, fb, opt, ema=None, ema_decay=0.9999):
        from . import _lib
        st = torch.cuda.current_stream().cuda_stream
        g, t, mcg, mct = self.ptrs[name]
        state = opt._state()
        _lib.call('tgan_dp_barrier', self.flag_ptrs, self.rank, self.world, 0, self.epoch.data_ptr(), st)
        _lib.call('tgan_dp_adam', g, t, mcg, mct, fb['m'].data_ptr(), fb['v'].data_ptr(), fb['n'], self.rank, self.world,
                  state.data_ptr(), opt.beta1, opt.beta2, opt.eps, st)
        _lib.call('tgan_dp_barrier', self.flag_ptrs, self.rank, self.world, 1, self.epoch.data_ptr(), st)
        if ema is not None:
            _lib.call('tgan_ema', ema.data_ptr(), fb['theta'].data_ptr(), fb['n'], ema_decay, st)
        _lib.call('tgan_adam_advance', state.data_ptr(), opt.beta1, opt.beta2, st)

    def gather_slots(self, store):
        """make every rank's m / v complete (each rank only keeps its own shard current)"""
        for name, fb in store.flat.items():
            if 'm' not in fb:
                continue
            per = self.shard(fb['n'])
            for r in range(self.world):
                lo, hi = r * per, min(fb['n'], (r + 1) * per)
                if hi > lo:
                    dist.broadcast(fb['m'][lo:hi], src=r)
                    dist.broadcast(fb['v'][lo:hi], src=r)
