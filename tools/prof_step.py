"""warm per-kernel timings of the eager CIFAR-10 step via torch.profiler (CUPTI): aggregated by kernel name and a
chronological list (name, grid, us) of one step -> gpurun_out/prof_step_{agg,seq}.txt"""
import sys, collections, os
# per-kernel attribution needs plain stream order: with programmatic dependent launch a kernel starts (and is timed)
# while its predecessor is still running, waiting at griddepcontrol.wait
os.environ.setdefault('TGAN_NO_PDL', '1')
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tensorflow-implementation-of-triple-gan_b200')
import torch, tgan
from tgan import synthetic
from torch.profiler import profile, ProfilerActivity
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
wl = sys.argv[2] if len(sys.argv) > 2 else 'cifar10'
tgan.init('cuda:0', math='bf16')
tr = tgan.make_trainer(wl, zca=synthetic.make_zca(1234) if wl == 'cifar10' else None)
tr.load_batch({k: torch.from_numpy(v) for k, v in synthetic.make_batch(tr.config, 1234).items()})
for _ in range(3):
    tr.step(lambda_1=tr.config.FAKE_G_LAMBDA, lambda_2=0.5)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(lambda_1=tr.config.FAKE_G_LAMBDA, lambda_2=0.5)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
with open('gpurun_out/prof_step_seq_%s.txt' % tag, 'w') as f:
    for e in evs:
        d = e.time_range.end - e.time_range.start
        nm = e.name.split('(')[0][:70]
        agg[nm][0] += 1; agg[nm][1] += d; tot += d
        f.write('%-72s %9.1f\n' % (e.name[:72], d))
with open('gpurun_out/prof_step_agg_%s.txt' % tag, 'w') as f:
    f.write('kernels in one eager step: %d, summed warm GPU time %.0f us\n' % (len(evs), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-72s n=%4d %9.1f us %5.1f%% avg %7.1f\n' % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
print(open('gpurun_out/prof_step_agg_%s.txt' % tag).read())
