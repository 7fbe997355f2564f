"""times the flat tcgen05 GEMM (ops.conv2d on a 2-D input) for skinny-K shapes"""
import sys, os; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import core, ops
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
def run(rows, K, Co, reps=20):
    nbuf = max(2, int(200e6 // (rows*K*2)) + 1)
    xs = [ops.Var(torch.randn(rows, K, device='cuda').to(torch.bfloat16), (rows, K)) for _ in range(min(nbuf, 8))]
    nbuf = len(xs)
    p = core.Param('w%d_%d' % (K, Co), (K, Co), True, None); p.data = torch.randn(K, Co, device='cuda')*0.03
    w = ops.PlainWeight(p)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3): ops.conv2d(xs[i % nbuf], w, 1, 1).data
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [ops.conv2d(xs[i % nbuf], w, 1, 1).data for i in range(reps)]
    g.replay(); torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); e1.synchronize()
    t = e0.elapsed_time(e1)*1e3/reps
    print('rows=%d K=%d Co=%d: %.1f us  %.0f TFLOP/s  out %.0f GB/s' % (rows, K, Co, t, 2.0*rows*K*Co/t/1e6, rows*Co*2/t/1e3), flush=True)
print('DBG', os.environ.get('TGAN_IGEMM_DBG'))
run(256000, 32, 128)
run(256000, 128, 128)
run(64000, 512, 256)
