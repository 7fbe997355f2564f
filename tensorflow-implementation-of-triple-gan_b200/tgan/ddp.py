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
