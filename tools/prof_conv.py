"""runs the dominant kernel (conv 128->128 @32x32, batch 100) fprop / dgrad / wgrad a few times (for ncu)"""
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import core, ops
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
N,H,C = 100,32,128
p = core.Param('w', (3,3,C,C), True, None); p.data = torch.randn(3,3,C,C, device='cuda')*0.03; p.grad = torch.zeros_like(p.data); p.requires_grad=True
for it in range(3):
    with core.recording():
        x = ops.Var(torch.randn(N,H,H,C, device='cuda').to(torch.bfloat16), (N,H,H,C), requires_grad=True)
        y = ops.conv2d(x, ops.PlainWeight(p), 3,3,1,'SAME'); y.data      # .data launches the deferred contraction
        y.grad = torch.randn(N,H,H,C, device='cuda').to(torch.bfloat16)
        core.ctx.tape.backward()
torch.cuda.synchronize()
print('ok')
