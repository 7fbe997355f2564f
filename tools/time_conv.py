"""times igemm fprop on the classifier's conv shapes: R launches back to back over a ring of distinct inputs whose
total size exceeds L2 (126 MB), captured in one CUDA graph, CUDA events around the replay (no host launch overhead in the number)"""
import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import core, ops, tc
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
def run(N,H,C,Co,k=3,pad='SAME', reps=20):
    nbuf = max(2, int(200e6 // (N*H*H*C*2)) + 1)
    xs = [ops.Var(torch.randn(N,H,H,C, device='cuda').to(torch.bfloat16), (N,H,H,C)) for _ in range(nbuf)]
    p = core.Param('w', (k,k,C,Co), True, None); p.data = torch.randn(k,k,C,Co, device='cuda')*0.03
    w = ops.PlainWeight(p)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3): ops.conv2d(xs[i % nbuf], w, k,k,1,pad).data
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [ops.conv2d(xs[i % nbuf], w, k,k,1,pad).data for i in range(reps)]
    g.replay(); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); e1.synchronize()
    t = e0.elapsed_time(e1)*1e3/reps
    Ho = H if pad=='SAME' else H-k+1
    fl = 2.0*N*Ho*Ho*k*k*C*Co
    print('N=%d H=%d C=%d Co=%d k=%d %s: %.1f us  %.0f TFLOP/s' % (N,H,C,Co,k,pad,t,fl/t/1e6), flush=True)
print('DBG', os.environ.get('TGAN_IGEMM_DBG'))
run(100,32,128,128)
run(100,16,256,256)
run(100,16,128,256)
run(100,8,256,512,pad='VALID')
run(50,32,128,128)
run(250,32,128,128)
run(250,16,256,256)
run(250,32,3,128)
