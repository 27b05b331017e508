"""Fused MMD kernel (mmdgan_mmd_fwd_bwd) across batch sizes and score widths: time per launch (CUDA events, 200 launches
after warm-up), achieved ALGORITHMIC HBM bandwidth (20*B*d bytes per launch: two [B,d] reads, three [B,d] gradient writes,
SURVEY.md section 8d) against the measured HBM peak, and the pair rate (3*B^2 kernel evaluations per launch).
At the configuration size (B = 256, d = 16: 81 920 bytes) a launch is latency bound; the sweep shows where the kernel
leaves that regime -- it becomes bound by the B^2 pairwise work (exp + 2d FMAs per pair), never by HBM."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdgan_b200 import kernels as K

peak = 6550.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
dev = torch.device('cuda')
print('loss  %6s %4s %10s %12s %10s %14s' % ('B', 'd', 'us/launch', 'alg GB/s', 'frac HBM', 'Gpairs/s'))
for loss in ('rep', 'rmb'):
    for B in (64, 256, 1024, 4096, 16384):
        for d in (16, 64):
            g = torch.Generator().manual_seed(B + d)
            gen = (torch.randn(B, d, generator=g) * 0.35).to(dev)
            real = (torch.randn(B, d, generator=g) * 0.35 + 0.1).to(dev)
            mk = K.MmdKernel(loss, (0.0, -1.0), b=B)
            out = [torch.zeros(B, d, device=dev) for _ in range(3)]
            per, reps = (20, 10) if B <= 4096 else (4, 5)
            for _ in range(5):
                mk(gen, real, out[0], out[1], out[2])
            torch.cuda.synchronize()
            # launches replayed from a CUDA graph: device time per launch, not the host's call rate
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=torch.cuda.Stream()):
                for _ in range(per):
                    mk(gen, real, out[0], out[1], out[2])
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (per * reps)
            gbs = 20.0 * B * d / (us * 1e-6) / 1e9
            print('%-5s %6d %4d %10.2f %12.2f %10.5f %14.2f' % (loss, B, d, us, gbs, gbs / peak, 3.0 * B * B / (us * 1e-6) / 1e9))
