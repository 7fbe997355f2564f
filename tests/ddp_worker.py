"""One rank of the 2-GPU data-parallel equivalence test (launched by tests/test_gpu_ddp.py through
`python -m torch.distributed.run --nproc-per-node 2`).  Runs `steps` Triple-GAN iterations on this rank's own shard
(batch, noise and dropout streams seeded by the rank) and saves losses, the all-reduced gradient buffers after the
first iteration, and all parameters after the last one."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tensorflow-implementation-of-triple-gan_b200')]


def shard_inputs(cfg, rank, step):
    from oracle import tgan_oracle as O
    return O.make_batch(cfg, seed=500 + 10 * step + rank), O.TagRNG(900 + 10 * step + rank)


def main():
    out_dir, math, scale, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    import tgan
    from tgan import core, ddp
    from oracle import tgan_oracle as O
    rank, world, local = ddp.init_from_env('nccl')
    tgan.init('cuda:%d' % local, math=math, seed=1234 + rank)
    P, S = O.init_params('cifar10', seed=5 + rank)            # different init per rank on purpose: rank 0's must win
    tr = tgan.make_trainer('cifar10', scale=scale, init=(P, S), zca=O.make_zca(3))
    res = {}
    for k in range(steps):
        batch, rng = shard_inputs(O.OracleConfig('cifar10', scale), rank, k)
        core.ctx.rng = core.InjectedSource(rng)
        res['loss%d' % k] = tr.step(batch, lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy()
        if k == 0:
            for grp in tr.store.GROUPS:
                fb = tr.store.flat[grp]
                for p, o in zip(fb['params'], fb['offsets']):
                    res['grad:' + p.name] = fb['grad'][o:o + p.size].cpu().numpy().reshape(p.shape)
                    res['theta1:' + p.name] = p.data.cpu().numpy().copy()
    for grp in tr.store.GROUPS:
        res['theta:' + grp] = tr.store.flat[grp]['theta'].cpu().numpy().copy()
    res['ema'] = tr.ema.shadow.cpu().numpy().copy()
    res['fused'] = np.array(1 if getattr(tr, 'fused_dp', None) is not None else 0)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), **res)
    tr.close()              # orderly: graph, device, process group


if __name__ == '__main__':
    main()
