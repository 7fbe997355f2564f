#!/usr/bin/env python
"""bench.py -- CIFAR-10 Triple-GAN training throughput (images/sec) on 1..8 B200, one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (TF-equivalent restatement) on host cores

A step = phases D + G + C of Training/Train_goodGAN.py:266-276 on the per-rank batch tuple
(G 100, L_C 50, U_C 50, L_D 20, U_D 80 -> BATCH_SIZE = 100 images).  Weak scaling: per-rank batch fixed,
global batch = 100 * N.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'tensorflow-implementation-of-triple-gan_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# algorithmic dense-contraction FLOPs of one step, x1e12; the default workload is the north-star one
STEP_TFLOPS = {'cifar10': 1.465, 'svhn': 1.295, 'mnist': 0.0511}
WORKLOADS = {
    'cifar10': 'CIFAR-10 32x32x3 Triple-GAN (Good_GAN_cifar10: WN 9-layer conv C + conv D + deconv G), one D+G+C training '
               'step, batch 100 per GPU (G 100, L_C 50, U_C 50, L_D 20, U_D 80)',
    'svhn': 'SVHN 32x32x3 Triple-GAN (Good_GAN), one D+G+C training step, batch 100 per GPU',
    'mnist': 'MNIST 28x28x1 Triple-GAN (Good_GAN), one D+G+C training step, batch 100 per GPU'}
STEP_TFLOP = STEP_TFLOPS['cifar10']
IMAGES_PER_STEP = 100


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    hbm=d['hbm_gbs'], src='measured')
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, src='fallback')


class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
             'clocks_event_reasons.sw_power_cap'

    def __init__(self, idx):
        self.idx, self.p = idx, None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.FIELDS,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


LAMBDAS = {'cifar10': (0.3, 0.5), 'svhn': (0.03, 0.0), 'mnist': (0.1, 0.0)}     # FAKE_G_LAMBDA of each main; lambda_2 is cifar-only


def cpu_step_time(workload, scale, steps, warmup):
    """the reference's CPU path: float32 torch-CPU restatement of the identical three-phase step (TensorFlow is
    not installable in this image), all host threads.  -> (mean seconds per step, images per step, threads)"""
    import torch
    from oracle import tgan_oracle as O
    torch.set_num_threads(os.cpu_count())
    P, S = O.init_params(workload, seed=1234)
    tr = O.OracleTrainer(workload, P, S, O.make_zca(1234) if workload == 'cifar10' else None, dtype=torch.float32, scale=scale)
    batch = O.make_batch(tr.cfg, seed=1234)
    rng = O.TagRNG(0)
    l1, l2 = LAMBDAS[workload]
    for _ in range(warmup):
        tr.step(batch, rng, l1, l2)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(batch, rng, l1, l2)
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), tr.cfg.BATCH_SIZE, torch.get_num_threads()


def metric_name(wl):
    return ('CIFAR-10' if wl == 'cifar10' else wl.upper()) + ' Triple-GAN train images/sec'


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the path.  TensorFlow 1.x cannot be installed in
    this image, so this is the oracle's float32 restatement of the same graph on all host threads, at
    the FULL batch tuple of the workload (scale 1: the same 100 images per step the GPU arm processes per rank).
    Under torchrun only rank 0 works; the other ranks exit."""
    if rank != 0:
        return
    wl = args.workload
    t, imgs, threads = cpu_step_time(wl, 1, args.steps, max(1, min(args.warmup, 2)))
    v = imgs / t
    n = max(1, args.gpus)
    sample = '%s: the full per-rank batch tuple (%d images/step), %d timed steps, float32 torch-CPU restatement of the ' \
             'TF graph (TensorFlow not installable here), one process on %d host threads' % (WORKLOADS[wl], imgs, args.steps, threads)
    if n > 1:
        sample += '; the global batch of %d images is %d such steps on the same host cores, so images/s is unchanged' % (imgs * n, n)
    print(json.dumps({
        'impl': 'reference', 'metric': metric_name(wl), 'value': v, 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3 * n,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[wl], 'global_batch': IMAGES_PER_STEP * n, 'parallelism': 'dp%d' % n},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)


def _time_conv(torch, ops, core, N, segs, colsum, R, ring):
    """R launches of the conv 128 -> 128 @32x32 forward over a ring of `ring` distinct inputs (ring x input bytes > the
    126 MB L2, so no launch finds its input cached), captured in one CUDA graph; CUDA events on the launching stream
    bracket the replay, so no host launch overhead is in the number.  -> seconds per launch"""
    H, C = 32, 128
    xs = []
    for _ in range(ring):
        v = ops.Var(torch.randn(N, H, H, C, device='cuda').to(torch.bfloat16), (N, H, H, C))
        if segs:
            v.aux = {'segs': list(segs)}
        xs.append(v)
    p = core.Param('w', (3, 3, C, C), True, None)
    p.data = torch.randn(3, 3, C, C, device='cuda') * 0.03
    w = ops.PlainWeight(p)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            ops.arena_reset()
            ops.conv2d(xs[i % ring], w, 3, 3, 1, 'SAME', colsum=colsum).data      # .data launches a deferred contraction
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ops.arena_reset()
        keep = [ops.conv2d(xs[i % ring], w, 3, 3, 1, 'SAME', colsum=colsum).data for i in range(R)]
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / R)
    del keep, xs
    return sorted(ts)[len(ts) // 2]


def dominant_kernel_roofline(torch, tgan, pk):
    """The dominant kernel of the step is igemm_kernel on the classifier's 3x3 convolutions.  Its heaviest launches in
    the timed step are the phase-C passes over the GROUPED batch of 250 (C(x_l) 50 + C(x_u) 50 + C(x_u) 50 + C(G(z)) 100,
    Good_GAN_cifar10.py:228-240 as one pass with four mean-only-BN segments): conv1_2 / conv1_3, 128 -> 128 channels at
    32x32, with the per-segment channel sums of the mean-only batch norm accumulated in the epilogue.  That launch is
    measured here exactly as the step issues it; the single-call shape (batch 100, no statistics) is reported beside it."""
    from tgan import core, ops
    core.ctx.store = core.VariableStore()
    C, H = 128, 32
    t250 = _time_conv(torch, ops, core, 250, (50, 50, 50, 100), True, 12, 4)
    t100 = _time_conv(torch, ops, core, 100, None, False, 16, 8)
    f250, f100 = 2.0 * 250 * H * H * 9 * C * C, 2.0 * 100 * H * H * 9 * C * C
    achieved = f250 / t250 / 1e12
    traffic = None
    pj = os.path.join(ROOT, 'profiles', 'dominant_kernel.json')
    if os.path.exists(pj):
        traffic = json.load(open(pj)).get('dram_bytes_per_launch_b250')
    return {'bound': 'tensor',
            'kernel': 'igemm_kernel (conv 128->128 @32x32 fprop, grouped phase-C batch 250 with 4-segment channel sums: '
                      'the launch the timed step issues for conv1_2 / conv1_3)',
            'achieved': achieved, 'peak': pk['bf16'], 'unit': 'TFLOP/s', 'frac': achieved / pk['bf16'],
            'peak_source': pk['src'] + ' burst bf16 (kernel timed alone)', 'frac_of_sustained': achieved / pk['bf16_sustained'],
            'traffic': traffic, 'flops_per_launch': f250, 'us_per_launch': t250 * 1e6,
            'single_call_batch100': {'achieved': f100 / t100 / 1e12, 'frac': f100 / t100 / 1e12 / pk['bf16'],
                                     'flops_per_launch': f100, 'us_per_launch': t100 * 1e6}}


def _finish(world, dist, torch):
    """Leave without tearing the NCCL communicator down: destroy_process_group() with NCCL work captured in a live
    CUDA graph can block forever; all timing is done, so synchronise and exit the process directly."""
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--math', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='cifar10', choices=sorted(WORKLOADS),
                    help='cifar10 = the north-star configuration; svhn / mnist = BASELINE.json configs[1] / configs[0]')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    dbg = {k: v for k, v in os.environ.items() if k.startswith('TGAN_')}
    if 'TGAN_IGEMM_DBG' in dbg:
        raise SystemExit('bench.py: TGAN_IGEMM_DBG is set -- refusing to time a kernel with debug switches (they exist only '
                         'in builds made with -DTGAN_DEBUG_SWITCHES)')
    import numpy as np
    import torch
    import torch.distributed as dist
    import tgan
    from tgan import _lib, synthetic
    args.warmup = max(args.warmup, 3)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    tgan.init('cuda:%d' % local, math=args.math, seed=1234 + rank)
    wl = args.workload
    tr = tgan.make_trainer(wl, zca=synthetic.make_zca(1234) if wl == 'cifar10' else None, seed=1234)
    batch = {k: torch.from_numpy(v).pin_memory() for k, v in synthetic.make_batch(tr.config, 1234 + rank).items()}
    h2d = sum(v.numel() * v.element_size() for v in batch.values())
    tr.load_batch(batch)
    lam = dict(lambda_1=tr.config.FAKE_G_LAMBDA, lambda_2=0.5)
    graph = False
    if not args.no_graph:
        try:
            tr.capture(warmup=3)
            graph = True
        except Exception as e:
            if world > 1:                 # a silent eager fallback on some ranks would desynchronise the collectives
                raise
            tr.graph = None
            torch.cuda.synchronize()
            print('graph capture failed (%s); running the same launches eagerly (config.cuda_graph = false)'
                  % str(e)[:200], file=sys.stderr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tr.step(**lam)
    barrier()
    # ---- leg 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    _lib.load().tgan_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        tr.step(**lam)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    launches = _lib.load().tgan_launch_count()
    if graph:
        launches = tr.launches_per_step * args.steps
    clocks = sampler.stop()
    # ---- leg 2: end to end through the public API (Train.host_feed): every step uploads ITS OWN pinned host batch
    # (H2D on the copy stream, overlapped with the previous step) and the loss triple comes back D2H every step (read one
    # step late); both are inside the timed region ----
    feed = tr.host_feed()
    feed.prime(batch)
    for _ in range(2):
        feed.step(batch, **lam)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = None
    for _ in range(args.steps):
        last = feed.step(batch, **lam)
    last = feed.drain()
    f1.record()
    barrier()
    t_e2e = f0.elapsed_time(f1) * 1e-3
    # ---- leg 3: the same public step fed by the device-side input pipeline (tgan/pipeline.py): a uint8 dataset resident
    # in HBM, each step's eight inputs formed by gather / pixel-map / draw kernels, loss D2H every step, no H2D ----
    from tgan import pipeline
    rs = np.random.default_rng(99 + rank)
    shp = tuple(tr.config.IMAGE_DIM)
    n_unl, n_lab = 20000, 4000
    inp = pipeline.TripleGANInput(
        tr.config,
        pipeline.DeviceDataset(rs.integers(0, 256, (n_lab,) + shp, dtype=np.uint8), rs.integers(0, tr.config.NUM_CLASSES, n_lab),
                               tr.config.NUM_CLASSES, wl),
        pipeline.DeviceDataset(rs.integers(0, 256, (n_unl,) + shp, dtype=np.uint8), rs.integers(0, tr.config.NUM_CLASSES, n_unl),
                               tr.config.NUM_CLASSES, wl), seed=1234 + rank)

    def pipe_step():
        try:
            inp.next_into(tr.inputs)
        except StopIteration:
            inp.start_epoch()
            inp.next_into(tr.inputs)
        return tr.step(**lam).cpu()
    for _ in range(2):
        pipe_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        pipe_step()
    g1.record()
    barrier()
    t_pipe = g0.elapsed_time(g1) * 1e-3
    tt = torch.tensor([t_dev, t_e2e, t_pipe], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, t_pipe = float(tt[0]), float(tt[1]), float(tt[2])
    if rank != 0:
        _finish(world, dist, torch)
        return
    pk = peaks()
    imgs = IMAGES_PER_STEP * world * args.steps
    value = imgs / t_dev
    out = {
        'metric': metric_name(wl), 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16' if args.math == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOADS[wl],
                   'global_batch': IMAGES_PER_STEP * world, 'parallelism': 'dp%d' % world,
                   'cuda_graph': graph, 'debug_env': dbg,
                   'dp_update': ('n/a (1 GPU)' if world == 1 else ('fused reduce-scatter + Adam + all-gather kernel over NVLink peer memory'
                                  + (' (NVSwitch multimem.ld_reduce / multimem.st)' if tr.fused_dp.multimem else ' (peer loads / stores)'))
                                 if getattr(tr, 'fused_dp', None) is not None else 'ncclAllReduce + Adam'),
                   'l2': 'no explicit flush: one step streams > 1 GB of activations (>> 126 MB L2) between reuses'},
        'e2e': {'value': imgs / t_e2e, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 12,
                'ms_per_step': t_e2e / args.steps * 1e3},
        'e2e_device_pipeline': {'value': imgs / t_pipe, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 12,
                                'ms_per_step': t_pipe / args.steps * 1e3,
                                'dataset': 'synthetic uint8, %d unlabelled + %d labelled images resident in HBM per rank' % (n_unl, n_lab)},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'step_tensor_frac': STEP_TFLOPS[wl] * 1e12 * world * args.steps / t_dev / (pk['bf16_sustained'] * 1e12 * world),
        'losses': [float(x) for x in last],
    }
    if args.math == 'bf16':
        out['roofline'] = dominant_kernel_roofline(torch, tgan, pk)
    if world == 1 and not args.no_cpu_baseline:
        t, n, threads = cpu_step_time(wl, 1, 4, 1)
        out['cpu_baseline'] = {'value': n / t, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                               'sample': '%s step at the full batch tuple (%d images/step), 4 timed steps after 1 warm-up, '
                                         'float32 torch-CPU restatement of the TF graph' % (wl, n)}
    print(json.dumps(out), flush=True)
    _finish(world, dist, torch)


if __name__ == '__main__':
    main()
