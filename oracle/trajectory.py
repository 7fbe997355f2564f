"""Free-running loss trajectories of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

`run(kind, lambdas, steps, scale)` steps the restated Training/Train_goodGAN.py:230-278 loop `steps` times from
the seeded initial state and returns the per-step (d_loss, g_loss, c_loss) of train_base.py:113-154:
  kind 'f64' : float64 golden run
  kind 'f32' : the same graph in float32 (what TF itself computes in) -> the fp32 divergence band
  kind 'q'   : float64 with the bf16 rounding points of the tensor-core path inserted -> the bf16 divergence band
Batches / noise of step k are seeded by k alone, so every run (oracle or CUDA) consumes identical inputs.
Runs are independent, so tests farm them out to worker processes (`run_many`).
"""
import contextlib

import numpy as np
import torch

from . import tgan_oracle as O

DATA = 'cifar10'


def step_inputs(cfg, k):
    return O.make_batch(cfg, seed=50 + k), O.TagRNG(100 + k)


def run(kind, lambdas, steps, scale, threads=2):
    torch.set_num_threads(threads)
    P, S = O.init_params(DATA, seed=5)
    zca = O.make_zca(3)
    tr = O.OracleTrainer(DATA, P, S, zca, dtype=torch.float32 if kind == 'f32' else torch.float64, scale=scale)
    out = []
    for k in range(steps):
        batch, rng = step_inputs(tr.cfg, k)
        with (O.quantized() if kind == 'q' else contextlib.nullcontext()):
            out.append(tr.step(batch, rng, lambdas[0], lambdas[1]))
    return np.asarray(out, np.float64)


def run_many(jobs, workers):
    """jobs: [(kind, lambdas, steps, scale)] -> {(kind, lambdas): [steps, 3] array}.  Each job is its own
    `python -m oracle.trajectory ...` child process (the calling pytest process may hold a CUDA context, which does
    not survive fork), at most `workers` at a time."""
    import os
    import subprocess
    import sys
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out, pending, running = {}, list(jobs), []
    with tempfile.TemporaryDirectory() as tmp:
        while pending or running:
            while pending and len(running) < workers:
                job = pending.pop(0)
                path = os.path.join(tmp, '%d.npy' % len(pending))
                cmd = [sys.executable, '-m', 'oracle.trajectory', job[0], repr(job[1][0]), repr(job[1][1]), str(job[2]),
                       str(job[3]), path]
                running.append((job, path, subprocess.Popen(cmd, cwd=root)))
            job, path, proc = running.pop(0)
            if proc.wait(timeout=1800) != 0:
                raise RuntimeError('oracle trajectory worker failed: %r' % (job,))
            out[(job[0], job[1])] = np.load(path)
    return out


if __name__ == '__main__':
    import sys
    kind, l1, l2, steps, scale, path = sys.argv[1:7]
    np.save(path, run(kind, (float(l1), float(l2)), int(steps), int(scale)))
