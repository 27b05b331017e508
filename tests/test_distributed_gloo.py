"""CPU, world_size 2, gloo: the data-parallel host logic (score all-gather layout, row-block decomposition of the MMD
estimate and of its gradients, sum all-reduce of gradients and kernel sums).  The per-rank row-block maths is done by
the oracle here (the CUDA row-block kernel is checked against the same identity on the GPU in test_gpu_kernels.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _row_block_oracle(gen_all, real_all, r0, r1, loss_type):
    """Kernel sums (x 1/(B(B-1))) over the ordered pairs whose FIRST index is a local row; gradients of local rows."""
    from oracle import mmd as omm
    g = gen_all.clone().requires_grad_(True)
    r = real_all.clone().requires_grad_(True)
    lg, ld = omm.gan_loss(g, r, loss_type, batch_size=g.shape[0])
    dlg = torch.autograd.grad(lg, g, retain_graph=True)[0][r0:r1]
    dld_g, dld_r = torch.autograd.grad(ld, [g, r])
    return lg.detach(), ld.detach(), dlg, dld_g[r0:r1], dld_r[r0:r1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mmdgan_b200 import parallel
    b, d = 8, 16
    gen = torch.Generator().manual_seed(100 + rank)
    s_local = torch.randn(2 * b, d, generator=gen, dtype=torch.float64) * 0.4
    buf = torch.zeros(world, 2 * b, d, dtype=torch.float64)
    gen_all, real_all = torch.zeros(world * b, d, dtype=torch.float64), torch.zeros(world * b, d, dtype=torch.float64)
    parallel.gather_scores(s_local, b, buf, gen_all, real_all)
    r0, r1 = parallel.row_block(rank, b)
    assert torch.equal(real_all[r0:r1], s_local[:b]) and torch.equal(gen_all[r0:r1], s_local[b:])
    # every rank sees the same global matrices
    chk = gen_all.clone()
    dist.broadcast(chk, 0)
    assert torch.equal(chk, gen_all)
    # gradients of the local rows from the global estimate == what a single process computes for those rows
    lg, ld, dlg, dld_g, dld_r = _row_block_oracle(gen_all, real_all, r0, r1, 'rmb')
    # a "parameter gradient" that is linear in the local score gradients: summing it over ranks gives the global one
    w = torch.arange(d, dtype=torch.float64)
    local_param_grad = (dld_g * w).sum(0) + (dld_r * w).sum(0)
    t = local_param_grad.clone()
    parallel.allreduce_sum([t])
    torch.save({'lg': lg, 'ld': ld, 'sum_grad': t, 'gen_all': gen_all, 'real_all': real_all}, os.path.join(out_dir, 'r{}.pt'.format(rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_row_block_data_parallel_identities(tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import mmd as omm
    from mmdgan_b200 import parallel
    outs = [torch.load(os.path.join(str(tmp_path), 'r{}.pt'.format(r))) for r in range(world)]
    assert torch.equal(outs[0]['gen_all'], outs[1]['gen_all']) and torch.equal(outs[0]['sum_grad'], outs[1]['sum_grad'])
    ref = omm.gan_loss_with_grads(outs[0]['gen_all'].numpy(), outs[0]['real_all'].numpy(), 'rmb')
    w = np.arange(16, dtype=np.float64)
    expect = (ref['dLd_dgen'] * w).sum(0) + (ref['dLd_ddata'] * w).sum(0)
    assert np.allclose(outs[0]['sum_grad'].numpy(), expect, atol=1e-12)
    assert abs(float(outs[0]['ld']) - float(ref['loss_dis'])) < 1e-13
    # losses from all-reduced kernel sums
    sums = torch.tensor([0.5, 0.25, 0.75, 0.4, 0.2, 0.6], dtype=torch.float64)
    lg, ld = parallel.losses_from_sums(sums, [-1.0, 0.0, 1.0])
    assert abs(float(lg) - (0.5 + 0.75 - 0.5)) < 1e-15 and abs(float(ld) - (0.6 - 0.4)) < 1e-15
