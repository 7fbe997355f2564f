"""run-to-run repeatability of the first steps (graph replay, in-kernel Philox with fixed seeds): differences beyond the
fp32 atomic-order noise of the epilogue statistics would point at a race between the two streams / PDL"""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import synthetic
tgan.init('cuda:0', math='fp32' if '--fp32' in sys.argv else 'bf16', seed=1234)
tr = tgan.make_trainer('cifar10', zca=synthetic.make_zca(1234), seed=1234)
tr.load_batch({k: torch.from_numpy(v) for k, v in synthetic.make_batch(tr.config, 1234).items()})
if '--eager' not in sys.argv:
    tr.capture(warmup=0)
out = []
for i in range(4):
    out.append(tr.step(lambda_1=0.3, lambda_2=0.5).cpu().numpy().copy())
for o in out:
    print(' '.join('%.7f' % v for v in o))
fb = tr.store.flat['classifier']
print('theta checksum %.12e' % float(fb['theta'].double().abs().sum()))
for g in ('discriminator', 'good_generator'):
    print(g, '%.12e' % float(tr.store.flat[g]['theta'].double().abs().sum()))
