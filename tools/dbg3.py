import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import numpy as np, torch
from oracle import tgan_oracle as O
from util_gpu import tnp
import test_gpu_nets as T
from tgan import core, ops
orc, o32, tr, rng = T._setup('svhn', 'fp32')
nrng = np.random.default_rng(1); B=16
x = nrng.uniform(-1, 1, [B,32,32,3]).astype(np.float32)
R = nrng.standard_normal((B, 10))
def fn(m, t):
    lt, _ = m.classifier(t(x), True, rng, 'T/C'); return lt, (lt * t(R)).sum()
lt, ref = T._oracle(orc, orc.c_vars, fn, 'fp32')
_, r32 = T._oracle(o32, orc.c_vars, fn, 'fp32')
tr._begin('classifier', tr.c_vars)
with core.recording():
    lv, _ = tr.model.classifier(ops.constant(x), True, reuse=True, tag='T/C')
    lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
    core.ctx.tape.backward()
fb = tr.store.flat['classifier']
scale = max(np.abs(r).max() for r in ref.values())
print('scale', scale)
for p,o in zip(fb['params'], fb['offsets']):
    if True:
        g = tnp(fb['grad'][o:o+p.size]).reshape(p.shape)
        print('%-40s |r|max %.3e  mine-r %.3e  o32-r %.3e' % (p.name, np.abs(ref[p.name]).max(), np.abs(g-ref[p.name]).max(), np.abs(r32[p.name]-ref[p.name]).max()))
