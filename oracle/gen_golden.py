"""Generates tests/golden/*.npz from the oracle (run here, committed with the fixtures).

PARITY UNPINNED: the reference cannot be executed in this image (no TensorFlow) and ships no golden vectors, so
these fixtures are produced by the oracle itself.  They pin the oracle against silent drift (a refactor that
changes its numbers fails tests/test_oracle.py) and give the GPU tests reference-free known answers; the oracle's
semantics are pinned separately against the loop-level numpy restatement and hand-computed TF geometry cases.

    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tgan_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def step_losses(name, steps=3, scale=10):
    P, S = O.init_params(name, seed=5)
    zca = O.make_zca(3) if name == 'cifar10' else None
    tr = O.OracleTrainer(name, P, S, zca, dtype=torch.float64, scale=scale)
    out = []
    for s in range(steps):
        out.append(tr.step(O.make_batch(tr.cfg, seed=50 + s), O.TagRNG(100 + s), 0.3, 0.5))
    first = tr.c_vars[0]
    return np.array(out), tr.P[first].detach().numpy().ravel()[:16].copy()


def layer_kats():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 8, 8, 3))
    w = rng.standard_normal((3, 3, 3, 4))
    wt = rng.standard_normal((5, 5, 2, 3))
    d = dict(x=x, w=w, wt=wt)
    T = lambda a: torch.tensor(a)
    d['conv_same_s1'] = O.conv2d_tf(T(x), T(w), 1, 'SAME').numpy()
    d['conv_same_s2'] = O.conv2d_tf(T(x), T(w), 2, 'SAME').numpy()
    d['conv_valid'] = O.conv2d_tf(T(x), T(w), 1, 'VALID').numpy()
    d['deconv_same_s2'] = O.conv2d_transpose_tf(T(x), T(wt), 2).numpy()
    lg = rng.standard_normal((6, 10))
    d['logits'] = lg
    d['entropy'] = float(O.entropy(T(lg)))
    d['balance_entropy'] = float(O.balance_entropy(T(lg)))
    return d


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    for name in ('cifar10', 'svhn', 'mnist'):
        losses, head = step_losses(name)
        np.savez(os.path.join(OUT, 'step_%s.npz' % name), losses=losses, c_var_head=head)
        print(name, losses)
    np.savez(os.path.join(OUT, 'layers.npz'), **layer_kats())
