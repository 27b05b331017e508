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


def _sharded_worker(rank, world, port, out_dir):
    """The arithmetic of the opt-in NVLS mode, with gloo collectives standing in for the multicast loads / stores:
    (1) shard-owner Adam: rank r updates elements [begin, end) from the SUMMED gradient and publishes w, m, v;
    (2) global-batch batch norm: forward sums (x, x^2) and backward sums (dy, dy * xhat) added over ranks; the input gradient
        computed with LOCAL row count from the global sums divided by the world size (what engine.py passes to bn_bwd_apply)."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mmdgan_b200 import parallel
    # ---- (1)
    n, lr, b1, b2, eps, t = 64 * 5, 2e-4, 0.5, 0.999, 1e-8, 3
    g0 = torch.Generator().manual_seed(1)
    w, m, v = (torch.randn(n, generator=g0, dtype=torch.float64), torch.randn(n, generator=g0, dtype=torch.float64) * 1e-2,
               torch.rand(n, generator=g0, dtype=torch.float64) * 1e-3)
    g_local = torch.randn(n, generator=torch.Generator().manual_seed(10 + rank), dtype=torch.float64)
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)                                   # = multimem.ld_reduce of the shard
    lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
    ref_m = b1 * m + (1 - b1) * g_sum
    ref_v = b2 * v + (1 - b2) * g_sum * g_sum
    ref_w = w - lr_t * ref_m / (ref_v.sqrt() + eps)
    begin, end = parallel.shard_range(n, rank, world)
    pub = torch.zeros(3, n, dtype=torch.float64)             # = multimem.st of w, m, v: every element written by exactly one rank
    pub[1, begin:end] = b1 * m[begin:end] + (1 - b1) * g_sum[begin:end]
    pub[2, begin:end] = b2 * v[begin:end] + (1 - b2) * g_sum[begin:end] * g_sum[begin:end]
    pub[0, begin:end] = w[begin:end] - lr_t * pub[1, begin:end] / (pub[2, begin:end].sqrt() + eps)
    dist.all_reduce(pub)
    assert torch.equal(pub[0], ref_w) and torch.equal(pub[1], ref_m) and torch.equal(pub[2], ref_v)
    # ---- (2)
    rows, c, eps_bn = 12, 8, 1e-3
    gx = torch.Generator().manual_seed(5)
    x_all = torch.randn(world * rows, c, generator=gx, dtype=torch.float64) * 2 + 0.5
    dy_all = torch.randn(world * rows, c, generator=gx, dtype=torch.float64)
    gamma = torch.rand(c, generator=gx, dtype=torch.float64) + 0.5
    xg = x_all.clone().requires_grad_(True)
    mu, var = xg.mean(0), xg.var(0, unbiased=False)
    y = (xg - mu) / (var + eps_bn).sqrt() * gamma
    (dx_ref,) = torch.autograd.grad(y, xg, dy_all)
    x, dy = x_all[rank * rows:(rank + 1) * rows], dy_all[rank * rows:(rank + 1) * rows]
    fwd = torch.stack([x.sum(0), (x * x).sum(0)])
    dist.all_reduce(fwd)
    n_glob = world * rows
    mean = fwd[0] / n_glob
    invstd = 1.0 / (fwd[1] / n_glob - mean * mean + eps_bn).sqrt()
    xhat = (x - mean) * invstd
    bwd = torch.stack([dy.sum(0), (dy * xhat).sum(0)])
    dgamma_local = bwd[1].clone()
    dist.all_reduce(bwd)
    db, dg = bwd[0] / world, bwd[1] / world                  # engine.py: tot.mul_(1 / world_size)
    dx = gamma * invstd * (dy - db / rows - xhat * dg / rows)     # bn_bwd_apply_kernel with the LOCAL row count
    assert torch.allclose(dx, dx_ref[rank * rows:(rank + 1) * rows], rtol=1e-10, atol=1e-12)
    tot = dgamma_local.clone()
    dist.all_reduce(tot)                                     # parameter gradient: local sums, added by the gradient reduction
    xhat_all = (x_all - x_all.mean(0)) / (x_all.var(0, unbiased=False) + eps_bn).sqrt()
    assert torch.allclose(tot, (dy_all * xhat_all).sum(0), rtol=1e-10, atol=1e-12)
    open(os.path.join(out_dir, 'ok{}'.format(rank)), 'w').close()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_owner_adam_and_global_batch_norm_identities(tmp_path):
    world, port = 2, 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
