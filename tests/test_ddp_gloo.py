"""CPU, world_size 2 over gloo: the host-side data-parallel plumbing of tgan.ddp -- rendezvous from the torchrun
environment, the per-network flat-gradient all-reduce, the 1/world factor handed to the optimiser, per-rank seeds
and the initial parameter broadcast.  (The kernels need a GPU; the N-rank path on GPUs is bench.py --gpus N.)"""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'tensorflow-implementation-of-triple-gan_b200')]
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    import tgan
    from tgan import ddp, synthetic
    r, w, _ = ddp.init_from_env('gloo')
    assert (r, w) == (rank, world) and ddp.world_size() == world
    tr = tgan.make_trainer('cifar10', build_only=True, seed=1234 + rank)      # different init per rank on purpose
    tr.store.finalize(torch.device('cpu'))
    ddp.broadcast_params(tr.store, 0)
    head = tr.store.flat['classifier']['theta'][:8].clone()
    # every rank contributes gradient = (rank+1) * ones; the all-reduced buffer times the returned factor = mean
    out = {}
    for grp in ('discriminator', 'good_generator', 'classifier'):
        g = tr.store.flat[grp]['grad']
        g.fill_(float(rank + 1))
        scale = ddp.allreduce_grads(g)
        out[grp] = (float(g[0]), float(g[-1]), scale, g.numel())
    b0 = synthetic.make_batch(tr.config, ddp.rank_seed(1234, rank))
    q.put((rank, out, head.numpy(), float(b0['z_g'][0, 0])))
    torch.distributed.destroy_process_group()


def test_two_rank_gradient_allreduce_and_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, head, z00 in res:
        for grp, (g0, g1, scale, n) in out.items():
            assert g0 == g1 == 3.0 and scale == 0.5            # 1 + 2 summed; Adam multiplies by 1/world -> 1.5
        assert out['discriminator'][3] >= 327467 and out['good_generator'][3] >= 5129201
    assert np.array_equal(res[0][2], res[1][2])                # parameters broadcast from rank 0
    assert res[0][3] != res[1][3]                              # per-rank input streams differ


def test_shard_partition_of_the_fused_update():
    """ddp.shard_len (the host mirror of tgan_dp_adam's partition): the shards of every rank are 16-byte aligned, disjoint
    and cover the flat buffer exactly, for the three networks' sizes and every world size the kernel accepts."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'tensorflow-implementation-of-triple-gan_b200')]
    from tgan import ddp
    for n in (327468, 5129204, 3121812, 4, 8, 64):
        assert n % 4 == 0
        for world in range(2, 9):
            per = ddp.shard_len(n, world)
            assert per % 4 == 0 and per * world >= n and per * (world - 1) < n + 4 * world
            covered = 0
            for r in range(world):
                lo, hi = r * per, min(n, (r + 1) * per)
                if hi > lo:
                    assert lo == covered and lo % 4 == 0
                    covered = hi
            assert covered == n
