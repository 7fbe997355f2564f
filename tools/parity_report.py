"""gpurun_out/parity_steps.json + gpurun_out/parity_trajectory.json (written by the -m gpu tests tests/test_gpu_step.py and
tests/test_gpu_trajectory.py on a B200) -> profiles/parity_r2.txt: the measured parity numbers the stated tolerances are
set against.

    python tools/parity_report.py > profiles/parity_r2.txt
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = json.load(open(os.path.join(ROOT, 'gpurun_out', 'parity_steps.json')))
traj = json.load(open(os.path.join(ROOT, 'gpurun_out', 'parity_trajectory.json')))

print('# Parity of the CUDA path against the oracle, measured on a B200 by the -m gpu tests (round 2).')
print('# All errors are relative to max(1, |reference|) for losses, to max-abs for gradients (fp32), cosine / relative L2')
print('# per parameter tensor for gradients (bf16, against the float64 oracle with the same bf16 rounding points and the')
print('# same pseudo-labels; cos64_* = against the plain float64 oracle).  Worst value over all steps of the run.')
print()
print('## whole-step tests (tests/test_gpu_step.py)')
for k in sorted(steps):
    v = steps[k]
    print('%s' % k)
    for grp in (('loss_d', 'loss_g', 'loss_c'), ('lfloor_d', 'lfloor_g', 'lfloor_c'), ('loss_vs_q_d', 'loss_vs_q_g', 'loss_vs_q_c'),
                ('grad_D', 'grad_G', 'grad_C'), ('floor_D', 'floor_G', 'floor_C'), ('cos_D', 'cos_G', 'cos_C'),
                ('relL2_D', 'relL2_G', 'relL2_C'), ('cos64_D', 'cos64_G', 'cos64_C'), ('labels_checked', 'labels_total')):
        if grp[0] in v:
            print('    ' + '  '.join('%s=%s' % (g, ('%.3g' % v[g]) if isinstance(v[g], float) else v[g]) for g in grp if g in v))
print()
print('## free-running 20-step (d, g, c)-loss trajectories, lambda_1 in {0, 0.3} x lambda_2 in {0, 0.5} (tests/test_gpu_trajectory.py)')
for k in sorted(traj):
    v = traj[k]
    print(k)
    for name in ('rms_cuda_vs_f64', 'rms_oracle_same_precision_vs_f64', 'max_cuda_vs_f64', 'max_oracle_same_precision_vs_f64',
                 'rms_cuda_bf16_vs_cuda_fp32'):
        if name in v:
            print('    %-36s d %.4f  g %.4f  c %.4f' % (name, *v[name]))
    for kk in sorted(v):
        if kk.endswith('step0 rel.err'):
            print('    %-36s d %.2e  g %.2e  c %.2e' % (kk, *v[kk]))
    lam = "(0.3, 0.5)"
    if 'per_lambda' in v and lam in v['per_lambda']:
        pl = v['per_lambda'][lam]
        print('    trajectory for lambda = %s: step | cuda (d g c) | float64 oracle | oracle at the same precision' % lam)
        for i, (a, b, c) in enumerate(zip(pl['cuda'], pl['f64'], pl['oracle_same_precision'])):
            print('      %2d | %.4f %.4f %.4f | %.4f %.4f %.4f | %.4f %.4f %.4f' % (i, *a, *b, *c))
