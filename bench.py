#!/usr/bin/env python
"""bench.py -- images/sec of the fused SNGan training step (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cifar|stl|celeba|lsun]
                  [--batch B] [--scaling weak|strong] [--passes 3|1] [--no-cpu-baseline] [--no-roofline]

Own arm: workload = BASELINE.json configs[1] (CIFAR-10 32x32 SNGAN + repulsive MMD, batch 256 per GPU, spectral norm
on), synthetic data, random-init weights.  A "step" is one fused training step ([losses, dis_op, gen_op, UPDATE_OPS]).
`value` is measured with the batch resident in HBM (CUDA-graph replay); `e2e` goes through SNGanEngine.step() with
HOST tensors: pinned H2D of the batch and D2H of the two losses inside the timed region.  N > 1 (torchrun): the batch
is sharded, weak scaling (256 images per GPU), scores all-gathered + one gradient all-reduce over NCCL.
Besides `roofline` (the tcgen05 GEMM launches that carry the step's FLOPs) the N = 1 line carries `mmd_kernel`: BASELINE's second
metric, the fused MMD kernel's algorithmic bytes (20 * B * d) over its launch time against the measured HBM peak.

Reference arm (--impl reference): TensorFlow-1.8 cannot be installed in this image, so the reference's CPU path is the
oracle (PyTorch-CPU restatement of the TF1 step, oracle/net.py) timed on the host cores; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



def read_traffic(name, batch):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the top kernel, from the newest committed
    `ncu --set full` capture: profiles/traffic.json, written by scripts/ncu_traffic.py together with the commit it was taken
    at.  bench.py cannot measure it (no profiler inside a timed run); absent or for another workload -> None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None, None
    try:
        with open(path) as f:
            t = json.load(f)
        if t.get('workload') != name or int(t.get('batch', -1)) != int(batch):
            return None, None
        return float(t['dram_bytes']), t
    except (ValueError, KeyError, OSError):
        return None, None


def workload_string(name, arch, loss_type, batch):
    """The same string on both arms (the driver compares config.workload of the two lines)."""
    return '{} {}x{} SNGAN + {} MMD, batch {} per GPU, spectral norm on'.format(name, arch['input'][0][1], arch['input'][0][2], loss_type, batch)

WORKLOADS = {'cifar': ('cifar', 256, 'rep', (5e-4, 2e-4)), 'stl': ('stl', 128, 'rmb', (2e-4, 2e-4)),
             'celeba': ('celeba', 128, 'rep', (1e-4, 2e-4)), 'lsun': ('lsun', 128, 'rep', (2e-4, 1e-4))}
# algorithmic cost per (real, fake) pair = 3G + 7D forward-equivalents, in GFLOP
GFLOP_PER_PAIR = {'cifar': 3.642, 'stl': 8.195, 'celeba': 19.351, 'lsun': 19.351}


def read_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p.get('hbm_gbs', 6650.0), burst=p.get('bf16_tflops', 1590.0),
                    sustained=p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1590.0)), source='measured')
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, source='fallback')


class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def synthetic_host_batches(arch, batch, n, seed):
    import torch
    c, h, w = arch['input'][0]
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        data = (torch.rand(batch, c, h, w, generator=g) * 2.0 - 1.0).pin_memory()      # uniform [-1, 1], input_func.py:839
        code = torch.randn(batch, arch['code'][0][0], generator=g).pin_memory()
        out.append((data, code))
    return out


def cpu_reference_throughput(arch_name, batch, loss_type, lr_list, steps, warmup, budget_s=200.0):
    """The oracle (CPU restatement of the TF1 step) on the host cores.  Returns (img/s, ms/step, sample description, cores)."""
    import torch
    from oracle import architectures as oa
    from oracle import net as onet
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arch = oa.ARCHITECTURES[arch_name]()
    model = onet.OracleSNGan(arch, loss_type, lr_list=lr_list, dtype=torch.float32)
    bs = batch
    data, code = onet.synthetic_batch(arch, bs, seed=0)
    t0 = time.perf_counter()
    model.step(data, code)                       # probe (also the first warm-up step)
    t_probe = time.perf_counter() - t0
    while bs > 16 and t_probe * (bs / batch) * (steps + max(warmup - 1, 0)) > budget_s:
        bs //= 2
    if bs != batch:
        data, code = data[:bs], code[:bs]
    for _ in range(max(warmup - 1, 0)):
        model.step(data, code)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.step(data, code)
    dt = (time.perf_counter() - t0) / steps
    sample = '{} steps of the same fused step at batch {} (workload batch {}), fp32, {} torch threads'.format(steps, bs, batch, cores)
    return bs / dt, dt * 1e3, sample, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    name, batch, loss_type, lr = WORKLOADS[args.workload]
    batch = args.batch or batch
    ips, ms, sample, cores = cpu_reference_throughput(name, batch, loss_type, lr, args.steps, args.warmup)
    from oracle import architectures as oa
    line = {
        'impl': 'reference', 'metric': 'images/sec', 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_string(name, oa.ARCHITECTURES[name](), loss_type, batch),
                   'note': 'TensorFlow 1.8 is not installable here; the reference CPU path is the PyTorch-CPU restatement of the '
                           'TF1 step (oracle/net.py) on the host cores.  The reference is single-device: for --gpus N > 1 this arm '
                           'still runs ONE host at the workload batch (it does not scale with N), while the repo arm is weak-scaled'},
        'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mmdgan_b200 import experiments as oa
    from mmdgan_b200 import kernels as K
    from mmdgan_b200.engine import SNGanEngine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    name, batch, loss_type, lr = WORKLOADS[args.workload]
    batch = args.batch or batch
    if args.scaling == 'strong':       # the workload's batch is the GLOBAL batch, split over the ranks (default: per-GPU batch, weak scaling)
        if batch % world:
            raise SystemExit('bench.py: --scaling strong needs the batch ({}) to be a multiple of the number of GPUs ({})'.format(batch, world))
        batch //= world
    arch = oa.ARCHITECTURES[name]()
    eng = SNGanEngine(arch, batch, loss_type=loss_type, lr_list=lr, npass=args.passes, device=dev, world_size=world, rank=rank,
                      use_graph=True)
    host = synthetic_host_batches(arch, batch, 4, seed=100 + rank)
    pool = [(d.to(dev), c.to(dev)) for d, c in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput
    # the training step draws its codes on the device (SNGan.sample_codes -> tf.random_normal in the reference's graph): only
    # the image batch is an input
    for i in range(max(args.warmup, 3)):
        eng.stage(pool[i % len(pool)][0])
        eng.step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        eng.stage(pool[i % len(pool)][0])
        eng.step_device()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    # ---- end to end through the public API (host tensors in, host losses out)
    for i in range(3):
        eng.step(host[i % len(host)][0])
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    pending = None
    for i in range(args.steps):
        # the training loop's calls (Agent.train): this step's batch + the next one, whose H2D copy overlaps this step; step i + 1
        # is enqueued before the losses of step i are read back.  Every batch is copied from pinned host memory exactly once and
        # every step's losses are read on the host, all inside the timed region
        nxt = (host[(i + 1) % len(host)][0], None) if i + 1 < args.steps else None
        enq = eng.step_async(host[i % len(host)][0], prefetch=nxt)
        if pending is not None:
            last = eng.result(pending)
        pending = enq
    last = eng.result(pending)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    images = batch * world * args.steps
    launches_weak = eng.kernel_launches_per_step
    nvls_mode = eng.nvls
    value = images / (ms_dev / 1e3)
    e2e_value = images / (ms_e2e / 1e3)
    h2d = batch * (arch['input'][0][0] * arch['input'][0][1] * arch['input'][0][2]) * 4      # the image batch; codes are drawn on the device
    peaks = read_peaks()

    # ---- roofline of the dominant kernels: every tcgen05 GEMM launch of one step, timed with CUDA events (eager step)
    roof = None
    if not args.no_roofline:
        evs = []
        orig_gemm, orig_wgrad, orig_direct = K.LinearOp._gemm, K.LinearOp.wgrad, K.LinearOp._direct

        depth = [0]

        def timed(fn):
            def wrapper(*a, **kw):
                # batch-1 spectral-norm launches: side streams, not part of the 3G+7D FLOP count.  Nested calls (an image layer
                # runs as im2col / tap-sum + an inner dense GEMM) are timed once, at the outermost level
                if a[3] == 1 or depth[0] > 0:
                    return fn(*a, **kw)
                depth[0] += 1
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                try:
                    r = fn(*a, **kw)
                finally:
                    depth[0] -= 1
                e.record()
                evs.append((s, e))
                return r
            return wrapper
        K.LinearOp._gemm, K.LinearOp.wgrad, K.LinearOp._direct = timed(orig_gemm), timed(orig_wgrad), timed(orig_direct)
        forks = (eng.sn_fork, eng.grad_fork)
        eng.sn_fork = eng.grad_fork = False      # one stream: every launch is timed alone, not against a concurrent kernel
        try:
            eng.stage(pool[0][0])
            torch.cuda.synchronize(dev)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            eng._run_phases()
            t1.record()
            torch.cuda.synchronize(dev)
        finally:
            K.LinearOp._gemm, K.LinearOp.wgrad, K.LinearOp._direct = orig_gemm, orig_wgrad, orig_direct
            eng.sn_fork, eng.grad_fork = forks
        gemm_ms = sum(s.elapsed_time(e) for s, e in evs)
        traffic, traffic_src = read_traffic(name, batch)
        flop_step = GFLOP_PER_PAIR[name] * 1e9 * batch
        achieved = flop_step / (gemm_ms / 1e3) / 1e12
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['sustained'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['sustained'],
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peaks['source'] + ' bf16 sustained (MEASURED_PEAKS.json)',
                'kernel': 'conv_gemm(_pair)_kernel + wgrad_gemm_kernel (tcgen05) + conv3x3 direct kernels of the image layers: the {} batch-sized '
                          'launches that carry the 3G+7D FLOPs of one step'.format(len(evs)),
                'gemm_ms_per_step': gemm_ms, 'eager_single_stream_step_ms': t0.elapsed_time(t1), 'share_of_eager_step': gemm_ms / t0.elapsed_time(t1),
                'note': ('algorithmic FLOPs = (3G+7D) x 2 x B; parity mode: every tensor-core launch issues 3 plane-pair MMAs per '
                         'algorithmic FLOP (forward: two fp16 planes per operand, gradients: two bf16 planes), so frac <= 0.333 by '
                         'construction in this precision mode') if args.passes == 3 else
                        'algorithmic FLOPs = (3G+7D) x 2 x B; single bf16 pass (speed mode, not parity grade)'}

    # ---- the fused MMD kernel alone (BASELINE's second metric): algorithmic bytes / launch time against the measured HBM peak
    mmd_roof = None
    if not args.no_roofline and world == 1:
        try:
            Bm = eng.B
            sc = eng.D.layers[-1].a[0]                   # [2B, d] scores of the last step
            sd = eng.D.layers[-1].dz_f32                 # [3B, d] score gradients
            reps, per_graph = 200, 20
            for _ in range(10):
                eng.mmd(sc[Bm:], sc[:Bm], sd[2 * Bm:], sd[Bm:2 * Bm], sd[:Bm])
            torch.cuda.synchronize(dev)
            # 20 launches per CUDA graph, 10 replays: the device time of a launch, not the host's ctypes call rate
            gmm = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=dev)
            with torch.cuda.graph(gmm, stream=side):
                for _ in range(per_graph):
                    eng.mmd(sc[Bm:], sc[:Bm], sd[2 * Bm:], sd[Bm:2 * Bm], sd[:Bm])
            gmm.replay()
            torch.cuda.synchronize(dev)
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            for _ in range(reps // per_graph):
                gmm.replay()
            m1.record()
            torch.cuda.synchronize(dev)
            us = m0.elapsed_time(m1) * 1e3 / reps
            dsc = int(sc.shape[1])
            alg_bytes = 20 * Bm * dsc                    # read 2 x [B, d] fp32, write 3 x [B, d] fp32 gradients
            gbs = alg_bytes / (us * 1e-6) / 1e9
            mmd_roof = {'kernel': 'mmd_fused_kernel<{}>'.format(dsc), 'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s',
                        'frac': gbs / peaks['hbm'], 'algorithmic_bytes': alg_bytes, 'us_per_launch': us, 'launches_timed': reps,
                        'note': 'back-to-back launches replayed from a CUDA graph (device time per launch, launch latency included); the {} KB '
                                'of scores are L2-resident here as in the step, where the preceding launch writes them; at this size the '
                                'kernel is launch/latency bound, not HBM bound (DESIGN.md section 5, profiles/r2_mmd_sweep.txt)'.format(alg_bytes // 1024)}
        except Exception as exc:                         # the extra measurement must never cost the bench line
            mmd_roof = {'error': '{}: {}'.format(type(exc).__name__, exc)}

    # ---- N > 1: (1) a driver-run witness that the data-parallel step equals the single-process step on the concatenated batch,
    # (2) the strong-scaling point (the workload batch as the GLOBAL batch, split over the ranks) next to the weak one
    dp_check = strong = None
    if world > 1 and not args.no_dp_check:
        from mmdgan_b200 import parallel
        try:
            dp_check = parallel.dp_equals_single(world, rank, dev)
        except Exception as exc:                          # the check must never cost the bench line; a failure is reported as such
            dp_check = {'ok': False, 'error': '{}: {}'.format(type(exc).__name__, exc)}
    if world > 1 and args.scaling == 'weak' and not args.no_strong and WORKLOADS[args.workload][1] % world == 0 and not args.batch:
        try:
            bs = WORKLOADS[args.workload][1] // world
            del eng
            torch.cuda.empty_cache()
            eng_s = SNGanEngine(arch, bs, loss_type=loss_type, lr_list=lr, npass=args.passes, device=dev, world_size=world, rank=rank,
                                use_graph=True)
            pool_s = [(d[:bs].contiguous(), c[:bs].contiguous()) for d, c in pool]
            for i in range(max(args.warmup, 3)):
                eng_s.stage(pool_s[i % len(pool_s)][0])
                eng_s.step_device()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.steps):
                eng_s.stage(pool_s[i % len(pool_s)][0])
                eng_s.step_device()
            s1.record()
            barrier()
            t = torch.tensor([s0.elapsed_time(s1)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_s = float(t[0])
            strong = {'scaling': 'strong', 'global_batch': bs * world, 'per_gpu_batch': bs, 'value': bs * world * args.steps / (ms_s / 1e3),
                      'unit': 'images/s', 'ms_per_step': ms_s / args.steps, 'steps': args.steps,
                      'note': 'BASELINE metric (1) read literally: global batch {} split over {} GPUs; device-resident, max over ranks'.format(bs * world, world)}
            eng = eng_s
        except Exception as exc:
            strong = {'error': '{}: {}'.format(type(exc).__name__, exc)}
    if rank == 0:
        line = {
            'metric': 'images/sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'fp16x3 fwd / bf16x3 grad (two 16-bit planes per operand, three products, fp32 accumulate)' if args.passes == 3 else 'bf16', 'data': 'synthetic',
            'config': {'workload': workload_string(name, arch, loss_type, batch),
                'global_batch': batch * world, 'parallelism': 'dp{}'.format(world),
                'collectives': None if world == 1 else ('nvswitch multicast kernels (MMDGAN_NVLS_ADAM=1)' if nvls_mode else 'nccl all-gather + all-reduce'),
                'l2': 'per-step working set (activations + gradients, > 1 GB) exceeds the 126 MB L2; no explicit flush',
                'tensor_passes': args.passes, 'cuda_graph': True, 'loss_last_step': last},
            'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 12,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches_weak * args.steps),
            'gpu_launches_per_step': int(launches_weak),
            'clocks': clocks,
        }
        if roof:
            line['roofline'] = roof
            if mmd_roof:
                roof['mmd'] = mmd_roof       # BASELINE's second metric, kept inside `roofline` so that parsers of the contract keys see it
        if mmd_roof:
            line['mmd_kernel'] = mmd_roof
        if dp_check is not None:
            line['dp_equals_single'] = bool(dp_check.get('ok'))
            line['config']['dp_check'] = dp_check
        if strong is not None:
            line['config']['strong'] = strong
        if world == 1 and not args.no_cpu_baseline:
            ips, ms, sample, cores = cpu_reference_throughput(name, batch, loss_type, lr, steps=2, warmup=1, budget_s=40.0)
            line['cpu_baseline'] = {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample,
                                    'ms_per_step': ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cifar', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: the workload batch per GPU (default, what the driver runs); strong: the workload batch is global and split over the GPUs')
    ap.add_argument('--passes', type=int, default=3, choices=[1, 3])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-dp-check', action='store_true', help='N > 1: skip the data-parallel == single-process equality check')
    ap.add_argument('--no-strong', action='store_true', help='N > 1: skip the extra strong-scaling measurement (config.strong)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
