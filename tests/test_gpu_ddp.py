"""-m gpu, needs 2 GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`): data-parallel equivalence of the
Triple-GAN step (SURVEY.md 8e "Check").  Two ranks, one process per GPU over NCCL, each on its own shard:

  * after the first iteration every rank's all-reduced gradient buffer / world == the ORACLE's average of the two
    per-shard gradients, phase by phase with the sequential semantics of Train_goodGAN.py:266-276 (phase G sees the
    averaged-update D, phase C the averaged-update G and D), and the parameters equal the oracle's after one
    averaged Adam update per network (Train_goodGAN.py:85-103);
  * the initial broadcast makes rank 0's variables win; after 3 iterations the two ranks' parameters and EMA shadows
    are BIT-identical (both math modes) -- the replicas never diverge.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE, STEPS = 10, 3


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(tmp, math):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'ddp_worker.py'), str(tmp), math, str(SCALE), str(STEPS)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return [np.load(os.path.join(str(tmp), 'rank%d.npz' % k)) for k in range(2)]


def _oracle_dp_first_step(dtype=torch.float64):
    """one data-parallel iteration restated on the oracle: per phase, the gradients of the two shards at the SAME
    parameters are averaged and applied once"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from ddp_worker import shard_inputs
    P, S = O.init_params('cifar10', seed=5)               # rank 0's initial values
    orc = O.OracleTrainer('cifar10', P, S, O.make_zca(3), dtype=dtype, scale=SCALE)
    shards = [shard_inputs(orc.cfg, r, 0) for r in range(2)]
    lr, cla_lr = orc.cfg.LEARNING_RATE, orc.cfg.CLA_LEARNINIG_RATE
    avg = lambda gs: {n: (gs[0][n] + gs[1][n]) / 2 for n in gs[0]}
    losses, grads = np.zeros((2, 3)), {}
    out = [orc.phase_d(orc.tensors(b), rng, lr, True, update=False) for b, rng in shards]
    losses[:, 0] = [o[0] for o in out]
    grads.update(avg([o[1] for o in out]))
    orc.opt_d.apply(orc.P, avg([o[1] for o in out]), lr)
    out = [orc.phase_g(orc.tensors(b), rng, lr, update=False) for b, rng in shards]
    losses[:, 1] = [o[0] for o in out]
    grads.update(avg([o[1] for o in out]))
    orc.opt_g.apply(orc.P, avg([o[1] for o in out]), lr)
    gcs = []
    for r, (b, rng) in enumerate(shards):
        losses[r, 2] = orc._phase_c(orc.tensors(b), rng, 0.3, 0.5, cla_lr, True, False)[2]
        gcs.append(orc.last_grads['C'])
    grads.update(avg(gcs))
    orc.apply_c(avg(gcs), cla_lr)
    return orc, losses, grads


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_rank_step_equals_shard_averaged_oracle(tmp_path):
    ranks = _launch(tmp_path, 'fp32')
    orc, losses, grads = _oracle_dp_first_step()
    # the same restatement in float32: the arithmetic's own noise floor on these (heavily cancelling) gradients
    _, _, grads32 = _oracle_dp_first_step(torch.float32)
    for r in range(2):
        got = ranks[r]['loss0']
        assert np.allclose(got, losses[r], rtol=2e-5, atol=2e-5), (r, got, losses[r])
    scale = {ph: max(float(grads[n].abs().max()) for n in names)
             for ph, names in (('D', orc.d_vars), ('G', orc.g_vars), ('C', orc.c_vars))}
    worst = 0.0
    for ph, names in (('D', orc.d_vars), ('G', orc.g_vars), ('C', orc.c_vars)):
        dens = {n: max(np.abs(grads[n].numpy()).max(), 1e-3 * scale[ph]) for n in names}
        # the float32 noise of these heavily cancelling gradients is a property of the network, one sample per tensor:
        # a tensor is also allowed twice the worst floor any tensor of its network shows
        phase_floor = max(np.abs(grads32[n].double().numpy() - grads[n].numpy()).max() / dens[n] for n in names)
        for n in names:
            ref = grads[n].numpy()
            den = dens[n]
            floor = max(np.abs(grads32[n].double().numpy() - ref).max() / den, 0.25 * phase_floor)
            fused = int(ranks[0]['fused']) == 1
            for r in range(2):
                # NCCL path: the buffer holds the all-reduced SUM (Adam applies 1/world).  Fused peer-memory update: the
                # reduction happens inside the update kernel, the buffers keep the LOCAL gradients -> average them here
                got = (ranks[0]['grad:' + n] + ranks[1]['grad:' + n]) / 2.0 if fused else ranks[r]['grad:' + n] / 2.0
                e = np.abs(got - ref).max() / den
                worst = max(worst, e)
                # the single-GPU bound of tests/test_gpu_step.py: 2e-4, or 8x the float32 floor of the oracle itself
                assert e < max(2e-4, 8 * floor), (n, r, e, floor)
            # one averaged Adam update per network: |theta - oracle| is a small fraction of the step lr
            lr = orc.cfg.CLA_LEARNINIG_RATE if ph == 'C' else orc.cfg.LEARNING_RATE
            resolved = np.abs(ref) > 1e-3 * scale[ph]
            if resolved.any():
                d = np.abs(ranks[0]['theta1:' + n] - orc.P[n].detach().numpy())[resolved].max()
                assert d < 0.05 * lr, (n, d, lr)
    print('2-rank fp32: worst averaged-gradient error %.2e (relative to max-abs), fused peer-memory update: %s' % (worst, fused))
    for k in [k for k in ranks[0].files if k.startswith('theta') or k == 'ema' or (k.startswith('grad:') and not fused)]:
        assert np.array_equal(ranks[0][k], ranks[1][k]), k + ' differs between the ranks'


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_rank_replicas_stay_bit_identical_bf16(tmp_path):
    ranks = _launch(tmp_path, 'bf16')
    fused = int(ranks[0]['fused']) == 1
    for k in [k for k in ranks[0].files if k.startswith('theta') or k == 'ema' or (k.startswith('grad:') and not fused)]:
        assert np.array_equal(ranks[0][k], ranks[1][k]), k + ' differs between the ranks'
    # the shards differ, so the per-rank losses do
    assert not np.array_equal(ranks[0]['loss0'], ranks[1]['loss0'])
    assert all(np.isfinite(ranks[r]['loss%d' % k]).all() for r in range(2) for k in range(STEPS))
