"""two eager (no CUDA graph) CIFAR-10 steps at the benchmark batch -- for the ncu launch list"""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import synthetic
tgan.init('cuda:0', math='bf16')
tr = tgan.make_trainer('cifar10', zca=synthetic.make_zca(1234))
tr.load_batch({k: torch.from_numpy(v) for k, v in synthetic.make_batch(tr.config, 1234).items()})
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    tr.step(lambda_1=0.3, lambda_2=0.5)
torch.cuda.synchronize()
from tgan import _lib
print('launches', _lib.load().tgan_launch_count())
