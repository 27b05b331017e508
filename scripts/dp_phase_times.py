"""torchrun --nproc-per-node N scripts/dp_phase_times.py : the data-parallel step's segments timed with CUDA events on the
compute stream (forward graph | score all-gather | loss + D backward graph | G backward graph (D all-reduce + D optimiser
running beside it) | G all-reduce | join + G optimiser graph), averaged over 20 steps, rank 0 prints.  Compare with
scripts/phase_times.py (single GPU)."""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdgan_b200 import experiments as oa      # noqa: E402
from mmdgan_b200.engine import SNGanEngine     # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
B, N = 256, 20
arch = oa.ARCHITECTURES['cifar']()
eng = SNGanEngine(arch, B, loss_type='rep', npass=3, device=dev, world_size=world, rank=rank, use_graph=True)
g = torch.Generator().manual_seed(rank)
data = (torch.rand(B, *arch['input'][0], generator=g) * 2 - 1).to(dev)
code = torch.randn(B, 128, generator=g).to(dev)
for it in range(4):
    eng.stage(data, code); eng.step_device()
torch.cuda.synchronize()
graphs = eng._graph_cache[eng._graph_key()]
assert len(graphs) == 5, len(graphs)
names = ['forward', 'gather', 'loss+bwd_dis', 'bwd_gen(+AR D, D update)', 'allreduce G', 'join+G update']
tot = [0.0] * len(names)
main = torch.cuda.current_stream(dev)
for it in range(N + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    graphs[0].replay(); ev[1].record()
    eng._gather_scores(); ev[2].record()
    graphs[1].replay(); ev[3].record()
    eng._allreduce_dis_async()
    with torch.cuda.stream(eng._upd_stream):
        eng._join_dis_allreduce()
        graphs[4].replay()
    graphs[2].replay(); ev[4].record()
    eng._allreduce_grads(); ev[5].record()
    main.wait_stream(eng._upd_stream)
    graphs[3].replay(); ev[6].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i in range(len(names)):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
if rank == 0:
    print('world=%d  ' % world + '  '.join('%s %.3f' % (names[i], tot[i] / N) for i in range(len(names))) + '  | sum %.3f ms' % (sum(tot) / N), flush=True)
dist.destroy_process_group()
