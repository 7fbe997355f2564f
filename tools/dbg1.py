import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import numpy as np, torch
from oracle import tgan_oracle as O
from util_gpu import tnp
import test_gpu_nets as T
from tgan import core, ops
orc, tr, rng = T._setup('cifar10', 'fp32')
nrng = np.random.default_rng(2)
B=16
x = nrng.uniform(-1, 1, [B,32,32,3]).astype(np.float32)
y = np.eye(10, dtype=np.float32)[nrng.integers(0, 10, B)]
R = nrng.standard_normal((B, 1))
t64 = lambda a: torch.tensor(a, dtype=torch.float64)
_, lt = orc.model.discriminator(t64(x), t64(y), rng, 'T/D')
tr._begin('discriminator', tr.d_vars)
with core.recording():
    _, lv = tr.model.discriminator(ops.constant(x), ops.constant(y), reuse=True, tag='T/D')
    lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
    core.ctx.tape.backward()
gs = torch.autograd.grad((lt*t64(R)).sum(), [orc.P[n] for n in orc.d_vars])
fb = tr.store.flat['discriminator']
for (p,o),r in zip(zip(fb['params'], fb['offsets']), gs):
    g = tnp(fb['grad'][o:o+p.size]).reshape(p.shape); r=r.numpy()
    print(p.name, p.shape, 'g', np.abs(g).max(), 'r', np.abs(r).max(), 'diff', np.abs(g-r).max())
    if 'conv2d_00/kernel' in p.name:
        d = np.abs(g-r); print(' per-cin diff', d.max(axis=(0,1,3))); print(' per-tap diff', d.max(axis=(2,3)))
