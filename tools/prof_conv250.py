"""runs the dominant launch of the timed step a few times (for `ncu --set full`): conv 128 -> 128 @32x32 forward on the
grouped phase-C batch of 250 with four mean-only-BN segments and epilogue channel sums -- the launch bench.py's roofline
block measures"""
import sys
sys.path.insert(0, '/root/repo')
sys.path.insert(0, '/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch
import tgan
from tgan import core, ops
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
N, H, C = 250, 32, 128
p = core.Param('w', (3, 3, C, C), True, None)
p.data = torch.randn(3, 3, C, C, device='cuda') * 0.03
w = ops.PlainWeight(p)
xs = []
for _ in range(3):
    v = ops.Var(torch.randn(N, H, H, C, device='cuda').to(torch.bfloat16), (N, H, H, C))
    v.aux = {'segs': [50, 50, 50, 100]}
    xs.append(v)
for it in range(6):
    ops.arena_reset()
    ops.conv2d(xs[it % 3], w, 3, 3, 1, 'SAME', colsum=True).data
torch.cuda.synchronize()
print('ok')
