"""HBM-bound kernels of the Triple-GAN step (BASELINE.json configs[4], the 'HBM roofline' half of the layer sweep):
achieved GB/s = ALGORITHMIC bytes (tensors the op must read + write once) / time per launch, as a fraction of the measured
HBM copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  Every timing is a CUDA graph of R launches over a ring of buffers
larger than the 126 MB L2, bracketed by CUDA events on the launching stream.

    python tools/bench_hbm.py [--batches 100,250,1024,4096] > profiles/r2_hbm_sweep.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tensorflow-implementation-of-triple-gan_b200'))
import torch                                   # noqa: E402
import tgan                                    # noqa: E402
from tgan import _lib, core                    # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batches', default='100,250,1024,4096')
args = ap.parse_args()
peak = 6453.4
pj = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pj):
    peak = json.load(open(pj))['hbm_gbs']
tgan.init('cuda:0', math='bf16')
core.ctx.store = core.VariableStore()
ws = core.ctx.ws()
st = lambda: torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()


def timed(fn, reps):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(0)
        fn(1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    return sorted(ts)[1]


def ring(shape, dtype, nbytes_each):
    n = max(2, min(8, int(300e6 // max(nbytes_each, 1)) + 1))
    return [torch.randn(*shape, device='cuda').to(dtype) if dtype != torch.uint8 else
            torch.randint(0, 8, shape, device='cuda', dtype=torch.uint8) for _ in range(n)]


def row(name, N, nbytes, t):
    gbs = nbytes / t / 1e3
    print('%-46s %6d | %8.1f us | %8.1f MB | %7.0f GB/s (%.2f of %.0f)' % (name, N, t, nbytes / 1e6, gbs, gbs / peak, peak), flush=True)


print('# HBM-bound kernels, bf16 activations, B200; algorithmic bytes per launch / time; peak = measured copy bandwidth')
print('%-46s %6s | %11s | %11s | %s' % ('kernel (classifier tensor)', 'batch', 'time', 'bytes', 'achieved'))
for H, C in ((32, 128), (16, 256)):
    for N in [int(b) for b in args.batches.split(',')]:
        rows = N * H * H
        nb = rows * C * 2
        if nb * 3 > 20e9:
            continue
        segs = [N // 5, N // 5, N // 5, N - 3 * (N // 5)]
        ends = [segs[0] * H * H, (segs[0] + segs[1]) * H * H, (segs[0] + segs[1] + segs[2]) * H * H]
        iends = [segs[0], segs[0] + segs[1], segs[0] + segs[1] + segs[2]]
        xs, ys = ring((rows, C), torch.bfloat16, nb), ring((rows, C), torch.bfloat16, nb)
        k = len(xs)
        sums = torch.randn(4, C, device='cuda')
        b, pm = torch.randn(C, device='cuda'), torch.zeros(C, device='cuda')
        tag = '[N,%d,%d,%d]' % (H, H, C)
        reps = 12
        t = timed(lambda i: _lib.call('tgan_mobn_apply_seg', p(xs[i % k]), p(ys[i % k]), rows, C, 4, ends[0], ends[1], ends[2],
                                      p(sums), 0, p(b), p(pm), 0.9, 1, 2, 0.2, st()), reps)
        row('mobn_apply_seg (mean-only BN + lrelu) ' + tag, N, 2 * nb, t)
        cs = torch.zeros(4, C, device='cuda')
        gb = torch.zeros(C, device='cuda')
        dus = ring((rows, C), torch.bfloat16, nb)
        t = timed(lambda i: _lib.call('tgan_act_bwd_seg', p(xs[i % k]), 1, p(ys[i % k]), 1, p(dus[i % k]), 1, rows, C, 4, ends[0],
                                      ends[1], ends[2], 2, 0.2, p(cs), p(gb), p(ws), st()), reps)
        row('act_bwd_seg (du = dy*lrelu\', segment sums) ' + tag, N, 3 * nb, t)
        t = timed(lambda i: _lib.call('tgan_sub_channel_mean_seg', p(dus[i % k]), p(dus[i % k]), rows, C, 4, ends[0], ends[1],
                                      ends[2], p(cs), st()), reps)
        row('sub_channel_mean_seg (dz = du - mean) ' + tag, N, 2 * nb, t)
        po = [torch.empty(N, H // 2, H // 2, C, device='cuda', dtype=torch.bfloat16) for _ in range(k)]
        code = [torch.empty(N, H // 2, H // 2, C, device='cuda', dtype=torch.uint8) for _ in range(k)]
        ctr = torch.zeros(1, dtype=torch.int64, device='cuda')
        t = timed(lambda i: _lib.call('tgan_mobn_pool_dropout_fwd', p(xs[i % k]), p(po[i % k]), p(code[i % k]), N, H, H, C, 4,
                                      iends[0], iends[1], iends[2], p(sums), 0, p(b), p(pm), 0.9, 1, 2, 0.2, 0.5, None, 7, 3,
                                      p(ctr), st()), reps)
        row('mobn_pool_dropout_fwd (BN+lrelu+pool+drop) ' + tag, N, nb + nb // 4 + nb // 8, t)
        t = timed(lambda i: _lib.call('tgan_mobn_pool_dropout_bwd', p(po[i % k]), p(po[(i + 1) % k]), p(code[i % k]), p(dus[i % k]),
                                      N, H, H, C, 4, iends[0], iends[1], iends[2], 2, 0.2, 0.5, p(cs), p(gb), p(ws), st()), reps)
        row('mobn_pool_dropout_bwd ' + tag, N, nb + 2 * (nb // 4) + nb // 8, t)
        del xs, ys, dus, po, code
        torch.cuda.empty_cache()

for name, n in (('classifier 3.12 M', 3121812), ('generator 5.13 M', 5129201), ('64 M parameters', 64 << 20)):
    n4 = (n + 3) // 4 * 4
    k = max(2, min(8, int(400e6 // (n4 * 16)) + 1))
    th = [torch.randn(n4, device='cuda') for _ in range(k)]
    m = [torch.zeros(n4, device='cuda') for _ in range(k)]
    v = [torch.zeros(n4, device='cuda') for _ in range(k)]
    g = [torch.randn(n4, device='cuda') for _ in range(k)]
    ema = [torch.randn(n4, device='cuda') for _ in range(k)]
    state = torch.tensor([3e-4, 0.5, 0.999], device='cuda')
    t = timed(lambda i: _lib.call('tgan_adam', p(th[i % k]), p(m[i % k]), p(v[i % k]), p(g[i % k]), n4, p(state), 0.5, 0.999, 1e-8,
                                  1.0, p(ema[i % k]), 0.9999, st()), 12)
    row('adam + EMA (%s)' % name, 0, n4 * 36, t)
    t = timed(lambda i: _lib.call('tgan_adam', p(th[i % k]), p(m[i % k]), p(v[i % k]), p(g[i % k]), n4, p(state), 0.5, 0.999, 1e-8,
                                  1.0, None, 0.0, st()), 12)
    row('adam (%s)' % name, 0, n4 * 28, t)
    del th, m, v, g, ema
    torch.cuda.empty_cache()
