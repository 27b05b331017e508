"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : N-rank data-parallel step == single-rank step on the
concatenated batch (a generator WITHOUT batch norm, so that per-rank BN statistics do not enter), and replicas stay
bit-identical.  Prints 'MULTI_GPU_OK' on rank 0.  With MMDGAN_NVLS_ADAM=1 in the environment the same checks run on the fused
NVSwitch-multicast all-reduce + Adam kernel (csrc/nvls.cu); 'NVLS_UNAVAILABLE' if the fabric has no multicast."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmdgan_b200 import experiments as oa   # noqa: E402
from mmdgan_b200.engine import SNGanEngine      # noqa: E402


def arch_no_bn():
    a = oa.tiny(act_k=2.6)
    for ly in a['generator']:
        if ly.get('act_nm') == 'bn':
            ly['act_nm'] = None
    return a


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    # MMDGAN_SYNC_BN=1 (with MMDGAN_NVLS_ADAM=1): batch-norm statistics over the global batch -> the generator keeps its batch norm
    arch, b = (oa.tiny(act_k=2.6) if os.environ.get('MMDGAN_SYNC_BN') == '1' else arch_no_bn()), 8
    g = torch.Generator().manual_seed(7)
    steps = 3
    data = torch.rand(steps, world * b, 3, 8, 8, generator=g) * 2 - 1
    code = torch.randn(steps, world * b, arch['code'][0][0], generator=g)
    for use_graph in (False, True):
        try:
            eng = SNGanEngine(arch, b, loss_type='rmb', device=dev, world_size=world, rank=rank, use_graph=use_graph, seed=5)
        except RuntimeError as exc:
            if os.environ.get('MMDGAN_NVLS_ADAM') == '1' and 'NVLS' in str(exc):     # no multicast on this fabric: nothing to check
                if rank == 0:
                    print('NVLS_UNAVAILABLE {}'.format(exc), flush=True)
                dist.destroy_process_group()
                return
            raise
        assert eng.nvls == (os.environ.get('MMDGAN_NVLS_ADAM') == '1')
        assert eng.sync_bn == (eng.nvls and os.environ.get('MMDGAN_SYNC_BN') == '1')
        ref = SNGanEngine(arch, world * b, loss_type='rmb', device=dev, use_graph=False, seed=5) if rank == 0 else None
        for it in range(steps):
            sl = slice(rank * b, (rank + 1) * b)
            lg, ld = eng.step(data[it, sl], code[it, sl])
            if rank == 0:
                # the single-process engine sees real rows [r0 | r1 | ...] and the same codes in the same global order
                lg1, ld1 = ref.step(data[it], code[it])
                assert abs(lg - lg1) <= 1e-4 * abs(lg1) + 1e-6, (it, lg, lg1)
                assert abs(ld - ld1) <= 1e-4 * abs(ld1) + 1e-6, (it, ld, ld1)
        # replicas identical; and equal to the single-process weights up to summation order
        for net in (eng.D, eng.G):
            w0 = net.w.clone()
            dist.broadcast(w0, 0)
            assert torch.equal(w0, net.w), 'replicas diverged'
        if rank == 0:
            for net, rnet in ((eng.D, ref.D), (eng.G, ref.G)):
                num = float((net.w - rnet.w).norm())
                den = float(rnet.w.norm())
                assert num <= 2e-3 * den, (net.name, num, den)
        dist.barrier()
    if rank == 0:
        print('MULTI_GPU_OK world={}'.format(world), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
