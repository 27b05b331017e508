"""Where the step goes, phase by phase: every phase of the single-GPU step captured as its own CUDA graph and timed with CUDA
events over N replays of the sequence (forward | loss | discriminator backward | generator backward | update), once with the
engine's stream forks (the shipped schedule) and once on a single stream.  The sum exceeds the one-graph step by the overlap the
whole-step graph has across phase boundaries (the discriminator's update under the generator's backward pass)."""
import sys, torch
sys.path.insert(0, '.')
from mmdgan_b200 import experiments as oa
from mmdgan_b200.engine import SNGanEngine
name = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = 20
arch = oa.ARCHITECTURES[name]()
g = torch.Generator().manual_seed(0)
data = (torch.rand(B, *arch['input'][0], generator=g) * 2 - 1).cuda()
code = torch.randn(B, 128, generator=g).cuda()
for forks in (True, False):
    eng = SNGanEngine(arch, B, loss_type="rep", npass=3, use_graph=True)
    if not forks:
        eng.sn_fork = eng.grad_fork = False
    for it in range(3):
        eng.stage(data, code); eng.step_device()
    torch.cuda.synchronize()
    phases = [('forward', eng._phase_forward), ('loss', eng._phase_loss), ('bwd_dis', lambda: eng._phase_backward('dis')),
              ('bwd_gen', lambda: eng._phase_backward('gen')), ('update', eng._phase_update)]
    graphs = []
    for nm, fn in phases:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=eng._stream):
            fn()
        graphs.append(gr)
    tot = [0.0] * len(phases)
    with torch.cuda.stream(eng._stream):
        for it in range(N + 3):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 1)]
            evs[0].record()
            for i, gr in enumerate(graphs):
                gr.replay(); evs[i + 1].record()
            torch.cuda.synchronize()
            if it >= 3:
                for i in range(len(phases)):
                    tot[i] += evs[i].elapsed_time(evs[i + 1])
    # the whole step as one graph, for reference
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.stage(data, code); eng.step_device(); eng.step_device()
    torch.cuda.synchronize()
    t0.record(eng._stream if False else None)
    for it in range(N):
        eng.step_device()
    t1.record()
    torch.cuda.synchronize()
    print('forks=%s  ' % forks + '  '.join('%s %.3f' % (nm, tot[i] / N) for i, (nm, _) in enumerate(phases)) +
          '  | sum %.3f  one-graph step %.3f ms' % (sum(tot) / N, t0.elapsed_time(t1) / N))
    del eng
