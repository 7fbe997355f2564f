"""layerwise fwd / bwd error of the cifar10 classifier in bf16 mode vs the float64 oracle"""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import numpy as np, torch
from oracle import tgan_oracle as O
from util_gpu import tnp
import test_gpu_nets as T
from tgan import core, ops, nn
math = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
orc, o32, tr, rng = T._setup('cifar10', math)
ctxq = O.quantized() if math == 'bf16' else __import__('contextlib').nullcontext()
ctxq.__enter__()
nrng = np.random.default_rng(1); B=16
x = nrng.uniform(-1, 1, [B,32,32,3]).astype(np.float32)
R = nrng.standard_normal((B, 10))
orc.model.acts = {}
lt, _ = orc.model.classifier(orc.model.zca_apply(torch.tensor(x, dtype=torch.float64)), True, rng, 'T/C')
for v in orc.model.acts.values(): v.retain_grad()
(lt*torch.tensor(R)).sum().backward()
ctxq.__exit__(None,None,None)
mine = {}
orig = nn.conv2d_WN
def rec(x, *a, **k):
    out = orig(x, *a, **k); mine['T/C/' + k['name']] = out; return out
nn.conv2d_WN = rec
tr._begin('classifier', tr.c_vars)
with core.recording():
    lv, _ = tr.model.classifier(tr._pre()(ops.constant(x)), True, reuse=True, tag='T/C')
    lv.grad = torch.tensor(R, dtype=torch.float32).cuda()
    core.ctx.tape.backward()
rel = lambda a,b: np.abs(a-b).max()/max(np.abs(b).max(),1e-30)
rms = lambda a,b: np.sqrt(((a-b)**2).sum()/max((b**2).sum(),1e-300))
frac = lambda a,b: float((np.abs(a-b) > 1e-3*np.abs(b).max()).mean())
for k,v in mine.items():
    r = orc.model.acts[k]
    a, b, ga, gb = tnp(v.data), r.detach().numpy(), tnp(v.grad), r.grad.numpy()
    print(k, 'fwd max %.2e rms %.2e frac %.4f | grad max %.2e rms %.2e frac %.4f' % (rel(a,b), rms(a,b), frac(a,b), rel(ga,gb), rms(ga,gb), frac(ga,gb)))
print('logits', rel(tnp(lv.data), lt.detach().numpy()))
fb = tr.store.flat['classifier']
for p,o in zip(fb['params'], fb['offsets']):
    g = tnp(fb['grad'][o:o+p.size]).reshape(p.shape); r = orc.P[p.name].grad.numpy()
    print(p.name, 'g %.3e r %.3e max %.3e rms %.3e' % (np.abs(g).max(), np.abs(r).max(), rel(g,r), rms(g,r)))
