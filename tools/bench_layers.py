"""nn.py layer microbench sweep (BASELINE.json configs[4]): conv2d / deconv2d / dense forward, input-gradient and
filter-gradient of the Triple-GAN layer shapes at batch 100 .. 4096, in TFLOP/s and as a fraction of the measured bf16
tensor-pipe peak.  Every timing is a CUDA graph of R launches over a ring of inputs larger than L2, bracketed by CUDA
events on the launching stream (no host launch overhead, no profiler).

    python tools/bench_layers.py [--batches 100,256,1024] > profiles/r1_layer_sweep.txt
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tensorflow-implementation-of-triple-gan_b200'))
import torch
import tgan
from tgan import core, ops, tc

ap = argparse.ArgumentParser()
ap.add_argument('--batches', default='100,256,1024,4096')
args = ap.parse_args()
peak = 1639.2
pj = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pj):
    peak = json.load(open(pj))['bf16_tflops']
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
_id = [0]


def timed(fn, reps):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(0); fn(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return sorted(ts)[1]


def layer(kind, N, H, Cin, Cout, k=3, s=1, pad='SAME'):
    _id[0] += 1
    Ho = H if (kind == 'conv' and pad == 'SAME' and s == 1) else (H // s if kind == 'conv' and pad == 'SAME' else
                                                                   (H - k + 1 if kind == 'conv' else H * s))
    if kind == 'dense':
        xshape, flops = (N, Cin), 2.0 * N * Cin * Cout
    elif kind == 'conv':
        xshape, flops = (N, H, H, Cin), 2.0 * N * Ho * Ho * k * k * Cin * Cout
    else:
        xshape, flops = (N, H, H, Cin), 2.0 * N * H * H * k * k * Cin * Cout
    bytes_x = 2 * int(torch.tensor(xshape).prod())
    nbuf = max(2, min(8, int(300e6 // max(bytes_x, 1)) + 1))
    xs = [ops.Var(torch.randn(*xshape, device='cuda').to(torch.bfloat16), xshape, requires_grad=True) for _ in range(nbuf)]
    wshape = (Cin, Cout) if kind == 'dense' else ((k, k, Cin, Cout) if kind == 'conv' else (k, k, Cout, Cin))
    p = core.Param('w%d' % _id[0], wshape, True, None)
    p.data = torch.randn(*wshape, device='cuda') * 0.03
    p.grad = torch.zeros_like(p.data)
    w = ops.PlainWeight(p)
    reps = 12 if flops > 2e10 else 24

    def fwd(i):
        x = xs[i % nbuf]
        if kind == 'deconv':
            return ops.conv2d_transpose(x, w, k, k, s).data
        return (ops.conv2d(x, w, 1, 1) if kind == 'dense' else ops.conv2d(x, w, k, k, s, pad)).data
    out = {'fwd': timed(fwd, reps)}
    # backward: build one forward under a tape, then time wgrad-only and dgrad-only by toggling requires_grad
    for which in ('dgrad', 'wgrad'):
        p.requires_grad = which == 'wgrad'
        for x in xs:
            x.requires_grad = which == 'dgrad'
        ys = []
        with core.recording() as tape:
            for i in range(nbuf):
                x = xs[i]
                y = ops.conv2d_transpose(x, w, k, k, s) if kind == 'deconv' else (
                    ops.conv2d(x, w, 1, 1) if kind == 'dense' else ops.conv2d(x, w, k, k, s, pad))
                y.data
                ys.append(y)
            nodes = list(tape.nodes)
        gy = torch.randn(*ys[0].data.shape, device='cuda').to(torch.bfloat16)

        def bwd(i):
            y = ys[i % nbuf]
            y.grad = gy
            xs[i % nbuf].grad = None
            old, core.ctx.tape = core.ctx.tape, tape
            try:
                nodes[i % nbuf]()
            finally:
                core.ctx.tape = old
        out[which] = timed(bwd, reps)
        p.requires_grad = False
    return flops, out


SHAPES = [
    ('conv', 'C conv1_1 3->128 @32', dict(H=32, Cin=3, Cout=128)),
    ('conv', 'C conv1_2 128->128 @32', dict(H=32, Cin=128, Cout=128)),
    ('conv', 'C conv2_1 128->256 @16', dict(H=16, Cin=128, Cout=256)),
    ('conv', 'C conv2_2 256->256 @16', dict(H=16, Cin=256, Cout=256)),
    ('conv', 'C conv3 256->512 VALID @8', dict(H=8, Cin=256, Cout=512, pad='VALID')),
    ('conv', 'D conv2d_01 42->32 s2 @32', dict(H=32, Cin=42, Cout=32, s=2)),
    ('conv', 'D conv2d_21 138->128 @8', dict(H=8, Cin=138, Cout=128)),
    ('deconv', 'G dconv0 522->256 @4', dict(H=4, Cin=522, Cout=256, k=5, s=2)),
    ('deconv', 'G dconv1 266->128 @8', dict(H=8, Cin=266, Cout=128, k=5, s=2)),
    ('deconv', 'G dconv2 138->3 @16', dict(H=16, Cin=138, Cout=3, k=5, s=2)),
    ('dense', 'G fc 110->8192', dict(H=1, Cin=112, Cout=8192)),
    ('dense', 'C NiN1 [36N,512]->256', dict(H=1, Cin=512, Cout=256, rows=36)),
    ('dense', 'MNIST D 794->1000', dict(H=1, Cin=800, Cout=1000)),
]
print('# bf16 tensor-core mode, B200; TFLOP/s (fraction of the measured bf16 burst peak %.0f TFLOP/s); us per launch' % peak)
print('%-28s %6s | %-24s | %-24s | %-24s' % ('layer', 'batch', 'forward', 'input gradient', 'filter gradient'))
for kind, name, kw in SHAPES:
    for N in [int(b) for b in args.batches.split(',')]:
        kw2 = dict(kw)
        rows = kw2.pop('rows', 1)
        if N * rows * kw2['H'] ** 2 * max(kw2['Cin'], kw2['Cout']) * 2 > 6e9:
            continue
        try:
            flops, t = layer(kind, N * rows, **kw2)
        except Exception as e:
            print('%-28s %6d | failed: %s' % (name, N, str(e)[:80]))
            continue
        cells = ['%7.1f us %6.0f (%.2f)' % (t[k], flops / t[k] / 1e6, flops / t[k] / 1e6 / peak) for k in ('fwd', 'dgrad', 'wgrad')]
        print('%-28s %6d | %s | %s | %s' % (name, N, *cells), flush=True)
    torch.cuda.empty_cache()
