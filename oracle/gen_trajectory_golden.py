"""Writes tests/golden/trajectory_r2.npz: the oracle's free-running 20-step (d, g, c)-loss trajectories
(oracle/trajectory.py) for lambda_1 in {0, 0.3} x lambda_2 in {0, 0.5}, CIFAR-10 tuple at scale 10
(G 10, L_C 5, U_C 5, L_D 2, U_D 8): float64 golden run, float32 run, float64 run with the bf16 rounding points.
The fixture is consumed by tests/test_gpu_trajectory.py (the GPU box does not spend minutes of host time on it) and
re-derived for its first steps by tests/test_oracle.py::test_trajectory_fixture_is_current.

    python -m oracle.gen_trajectory_golden          (about 10 minutes on 8 cores)
"""
import os

import numpy as np

from . import trajectory as TJ

LAMBDAS = [(0.0, 0.0), (0.3, 0.0), (0.0, 0.5), (0.3, 0.5)]
STEPS, SCALE = 20, 10


def key(kind, lam):
    return '%s_l1=%g_l2=%g' % (kind, lam[0], lam[1])


if __name__ == '__main__':
    jobs = [(kind, lam, STEPS, SCALE) for lam in LAMBDAS for kind in ('f64', 'f32', 'q')]
    res = TJ.run_many(jobs, workers=max(1, (os.cpu_count() or 4) // 2))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'trajectory_r2.npz')
    np.savez(out, steps=STEPS, scale=SCALE, **{key(k, lam): v for (k, lam), v in res.items()})
    print('wrote', out)
