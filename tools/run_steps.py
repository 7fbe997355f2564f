"""run N eager (no CUDA graph) CIFAR-10 training steps on synthetic data: the launch sequence of every step is identical
(66 igemm, 19 wgrad launches, profiles/r2_step_kernels_seq.txt), so `ncu -k regex:<kernel> -s <3 * per_step + index> -c 1`
captures one chosen launch of the fourth step"""
import os
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
tgan.init('cuda:0', math='bf16')
tr = tgan.make_trainer('cifar10', zca=synthetic.make_zca(1234))
tr.load_batch({k: torch.from_numpy(v) for k, v in synthetic.make_batch(tr.config, 1234).items()})
for _ in range(n):
    tr.step(lambda_1=tr.config.FAKE_G_LAMBDA, lambda_2=0.5)
torch.cuda.synchronize()
print('ok')
