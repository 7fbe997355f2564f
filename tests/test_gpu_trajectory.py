"""-m gpu: the free-running 20-step (d, g, c)-loss trajectory of SURVEY.md 8c / the north_star: 20 iterations of
Training/Train_goodGAN.py:230-278 from identical weights, inputs, noise and dropout masks, lambda_1 in {0, 0.3},
lambda_2 in {0, 0.5}, in both math modes, against the float64 oracle.

What can be asserted.  Step 0 is an exact single-step comparison (2e-5 fp32 / 2e-2 bf16, or 4x the error the
oracle itself shows at that precision: c_loss in float32 carries ~5e-5).  From step 1 on the
trajectories of ANY two arithmetics decorrelate: Adam's early updates are sign-like, so rounding noise on
barely-resolved gradient elements moves weights by +-lr, and the reference's own float32 arithmetic leaves its
float64 trajectory by 1e-2 ... 1e-1 within a few iterations (measured below, not assumed).  The band is therefore
taken from the ORACLE'S OWN divergence at the same precision: the float32 oracle for fp32 mode, the float64 oracle
with the bf16 rounding points inserted for bf16 mode.  Per loss, over steps 1..19, pooled over the four lambda
settings: rms |cuda - f64| <= BAND_RMS x rms |oracle_p - f64| and max |cuda - f64| <= BAND_MAX x max |oracle_p - f64|.
Per-step exactness over 20 steps with evolving optimiser state is the teacher-forced test
(tests/test_gpu_step.py::test_teacher_forced_20_steps).  The oracle's three trajectories per lambda setting are a committed fixture
(tests/golden/trajectory_r2.npz, generator oracle/gen_trajectory_golden.py).  Measured numbers ->
gpurun_out/parity_trajectory.json -> profiles/parity_r2.txt.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import tgan_oracle as O                 # noqa: E402
from oracle import trajectory as TJ                 # noqa: E402

LAMBDAS = [(0.0, 0.0), (0.3, 0.0), (0.0, 0.5), (0.3, 0.5)]
STEPS, SCALE = 20, 10
STEP0_TOL = {'fp32': 2e-5, 'bf16': 2e-2}
BAND_RMS, BAND_MAX = 2.5, 3.0
_CACHE = {}


def _oracle_runs():
    """the oracle's trajectories: committed fixture tests/golden/trajectory_r2.npz (generator:
    oracle/gen_trajectory_golden.py; kept current by tests/test_oracle.py::test_trajectory_fixture_is_current)"""
    if 'o' not in _CACHE:
        from oracle.gen_trajectory_golden import key
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'trajectory_r2.npz'))
        assert int(z['steps']) == STEPS and int(z['scale']) == SCALE
        _CACHE['o'] = {(kind, lam): z[key(kind, lam)] for lam in LAMBDAS for kind in ('f64', 'f32', 'q')}
    return _CACHE['o']


def _cuda_run(math, lam):
    import tgan
    from tgan import core
    P, S = O.init_params(TJ.DATA, seed=5)
    tgan.init('cuda:0', math=math)
    tr = tgan.make_trainer(TJ.DATA, scale=SCALE, init=(P, S), zca=O.make_zca(3))
    out = []
    for k in range(STEPS):
        batch, rng = TJ.step_inputs(O.OracleConfig(TJ.DATA, SCALE), k)
        core.ctx.rng = core.InjectedSource(rng)
        out.append(tr.step(batch, lambda_1=lam[0], lambda_2=lam[1]).cpu().numpy().astype(np.float64))
    return np.stack(out)


def _cuda_runs(math):
    if math not in _CACHE:
        _CACHE[math] = {lam: _cuda_run(math, lam) for lam in LAMBDAS}
    return _CACHE[math]


def _report(name, obj):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, 'parity_trajectory.json')
        old = json.load(open(path)) if os.path.exists(path) else {}
        old[name] = obj
        json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.parametrize('math', ['fp32', 'bf16'])
def test_loss_trajectory_20_steps(math):
    orc = _oracle_runs()
    mine = _cuda_runs(math)
    kind = 'f32' if math == 'fp32' else 'q'
    rep = {}
    dev_c = {lam: mine[lam] - orc[('f64', lam)] for lam in LAMBDAS}
    dev_o = {lam: orc[(kind, lam)] - orc[('f64', lam)] for lam in LAMBDAS}
    for lam in LAMBDAS:
        ref0 = orc[('f64', lam)][0]
        e0 = np.abs(dev_c[lam][0]) / np.maximum(1.0, np.abs(ref0))
        f0 = np.abs(dev_o[lam][0]) / np.maximum(1.0, np.abs(ref0))      # the oracle's own error at this precision
        rep['%s step0 rel.err' % (lam,)] = e0.tolist()
        rep['%s step0 oracle floor' % (lam,)] = f0.tolist()
        assert (e0 < np.maximum(STEP0_TOL[math], 4 * f0)).all(), (lam, mine[lam][0], ref0, f0)
        assert np.isfinite(mine[lam]).all()
    rms = lambda d: np.sqrt(np.mean(np.concatenate([d[lam][1:] for lam in LAMBDAS]) ** 2, axis=0))
    mx = lambda d: np.max(np.abs(np.concatenate([d[lam][1:] for lam in LAMBDAS])), axis=0)
    rep.update(rms_cuda_vs_f64=rms(dev_c).tolist(), rms_oracle_same_precision_vs_f64=rms(dev_o).tolist(),
               max_cuda_vs_f64=mx(dev_c).tolist(), max_oracle_same_precision_vs_f64=mx(dev_o).tolist(),
               band=dict(rms_factor=BAND_RMS, max_factor=BAND_MAX), steps=STEPS, scale=SCALE,
               per_lambda={str(lam): dict(cuda=mine[lam].tolist(), f64=orc[('f64', lam)].tolist(),
                                          oracle_same_precision=orc[(kind, lam)].tolist()) for lam in LAMBDAS})
    if math == 'bf16' and 'fp32' in _CACHE:      # and against the CUDA fp32 mode (reported)
        rep['rms_cuda_bf16_vs_cuda_fp32'] = np.sqrt(np.mean(np.concatenate(
            [mine[lam][1:] - _CACHE['fp32'][lam][1:] for lam in LAMBDAS]) ** 2, axis=0)).tolist()
    _report('trajectory ' + math, rep)
    print(math, {k: v for k, v in rep.items() if k != 'per_lambda'})
    assert (rms(dev_c) <= BAND_RMS * rms(dev_o)).all(), (rms(dev_c), rms(dev_o))
    assert (mx(dev_c) <= BAND_MAX * mx(dev_o)).all(), (mx(dev_c), mx(dev_o))
    # the losses stay in the regime of the golden run: mean over the last 10 iterations within the same band
    for lam in LAMBDAS:
        m_c, m_r = mine[lam][10:].mean(0), orc[('f64', lam)][10:].mean(0)
        assert (np.abs(m_c - m_r) <= BAND_MAX * mx(dev_o)).all(), (lam, m_c, m_r)
