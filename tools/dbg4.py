import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import numpy as np, torch
from oracle import tgan_oracle as O
from util_gpu import tnp
import tgan
from tgan import core
P, S = O.init_params('cifar10', seed=5); zca = O.make_zca(3)
orc = O.OracleTrainer('cifar10', P, S, zca, dtype=torch.float64, scale=10)
o32 = O.OracleTrainer('cifar10', P, S, zca, dtype=torch.float32, scale=10)
tgan.init('cuda:0', math='fp32')
tr = tgan.make_trainer('cifar10', scale=10, init=(P, S), zca=zca)
for step in range(3):
    rng = O.TagRNG(100 + step); core.ctx.rng = core.InjectedSource(rng)
    batch = O.make_batch(orc.cfg, seed=50 + step)
    ref = orc.step(batch, rng, 0.3, 0.5); r32 = o32.step(batch, rng, 0.3, 0.5)
    got = tr.step(batch, lambda_1=0.3, lambda_2=0.5).cpu().numpy()
    print('step', step, 'mine', got, 'ref', np.array(ref), 'o32', np.array(r32))
    for grp in ('discriminator','good_generator','classifier'):
        fb = tr.store.flat[grp]; lr = 3e-3 if grp=='classifier' else 3e-4
        dm = max(np.abs(tnp(p.data) - orc.P[p.name].detach().numpy()).max() for p in fb['params'])
        d32 = max(np.abs(o32.P[p.name].detach().double().numpy() - orc.P[p.name].detach().numpy()).max() for p in fb['params'])
        nm = sum(int((np.abs(tnp(p.data) - orc.P[p.name].detach().numpy()) > 0.5*lr).sum()) for p in fb['params'])
        n32 = sum(int((np.abs(o32.P[p.name].detach().double().numpy() - orc.P[p.name].detach().numpy()) > 0.5*lr).sum()) for p in fb['params'])
        print('   %-16s max|dtheta|/lr mine %.2f o32 %.2f ; #elements off by >0.5 lr: mine %d o32 %d' % (grp, dm/lr, d32/lr, nm, n32))
