"""Data-parallel plumbing of the SNGan step (one process per GPU, torch.distributed; NCCL on NVLink, gloo in CPU tests).

The reference is single-GPU (FLAGS.num_gpus is never read, misc_fun.py:28); this is new functionality specified in
SURVEY.md section 8(e).  The batch dimension is sharded; the only batch-coupled quantities are
  * the three B x B kernel matrices of the MMD loss -> every rank needs ALL scores: one all-gather of the [2b, d] local
    score block (<= 128 KiB at global batch 1024), after which rank r evaluates rows [r*b, (r+1)*b) of each matrix
    (mmdgan_mmd_fwd_bwd row-block form) and gets the exact gradients of its own rows;
  * the parameter gradients and the six kernel sums -> one sum all-reduce per flat buffer (the local gradients are
    already normalised by the GLOBAL 1/(B(B-1)), so the sum is the global-batch gradient; no averaging).
Batch-norm statistics stay per rank (per-GPU batch), as documented in DESIGN.md.

Opt-in (MMDGAN_NVLS_ADAM=1): the gradient all-reduce and the Adam update become ONE kernel over NVSwitch multicast
(csrc/nvls.cu).  SymmetricFlat below owns the [g | w | m | v] allocation of one network, mapped on every rank and bound to a
multicast object through torch's symmetric-memory rendezvous (plumbing only: allocation, address exchange, stream barriers).
"""
import torch
import torch.distributed as dist


def gather_scores(s_local, b, gather_buf, gen_all, real_all, group=None):
    """s_local [2b, d] (rows [0,b) real, [b,2b) generated) -> real_all / gen_all [world*b, d] in global row order."""
    world = dist.get_world_size(group)
    d = s_local.shape[1]
    try:
        # straight into the global row order: rank r's block lands at rows [r*b, (r+1)*b) of each matrix -- no re-ordering copies
        dist.all_gather_into_tensor(real_all, s_local[:b], group=group)
        dist.all_gather_into_tensor(gen_all, s_local[b:2 * b], group=group)
    except (RuntimeError, NotImplementedError):       # backends without the flat variant
        parts = [torch.empty_like(s_local) for _ in range(world)]
        dist.all_gather(parts, s_local.contiguous(), group=group)
        g = torch.stack(parts, 0)
        real_all.copy_(g[:, :b, :].reshape(world * b, d))
        gen_all.copy_(g[:, b:, :].reshape(world * b, d))
    return gen_all, real_all


def allreduce_sum(tensors, group=None):
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def losses_from_sums(sums, cD):
    """sums = [e_gg, e_gr, e_rr, e_gg^b, e_gr^b, e_rr^b] (global, after the all-reduce) -> (loss_gen, loss_dis)."""
    loss_gen = sums[0] + sums[2] - 2.0 * sums[1]
    loss_dis = cD[0] * sums[3] + cD[1] * sums[4] + cD[2] * sums[5]
    return loss_gen, loss_dis


def row_block(rank, b):
    """Global row range owned by `rank`."""
    return rank * b, (rank + 1) * b


def shard_range(n_flat, rank, world):
    """[begin, end) of `rank`'s share of a flat buffer of n_flat floats, in whole float4 groups."""
    n4 = n_flat // 4
    assert n4 * 4 == n_flat, 'flat parameter buffers are padded to whole float4 groups'
    return (n4 * rank // world) * 4, (n4 * (rank + 1) // world) * 4


def _symmetric(numel, device, group):
    """(tensor, handle, multicast address) of a zero-filled symmetric float32 allocation of `numel` elements."""
    import torch.distributed._symmetric_memory as symm_mem
    group = dist.group.WORLD if group is None else group
    buf = symm_mem.empty(int(numel), dtype=torch.float32, device=device)
    buf.zero_()
    try:                                   # older releases need the group registered first; newer ones deprecate the call
        symm_mem.enable_symm_mem_for_group(group.group_name)
    except Exception:
        pass
    handle = symm_mem.rendezvous(buf, group.group_name)
    mc = int(getattr(handle, 'multicast_ptr', 0) or 0)
    if mc:
        mc += int(getattr(handle, 'offset', 0) or 0)      # position of this tensor inside the mapped block
    if mc == 0:
        raise RuntimeError('MMDGAN_NVLS_ADAM=1 needs NVSwitch multicast (NVLS); this fabric / driver offers none')
    return buf, handle, mc


class SymmetricScores(object):
    """gen_all / real_all [world * b, d] of the MMD loss in one symmetric allocation; scatter() replaces gather_scores()."""

    def __init__(self, b, d, device, group=None, stat_slots=0, stat_width=0):
        world = dist.get_world_size(group)
        n = world * b * d
        assert stat_width % 4 == 0
        # + one 8-float slot for the kernel sums + `stat_slots` slots of `stat_width` floats (batch-norm statistics)
        self.buf, self.handle, mc = _symmetric(2 * n + 8 + stat_slots * stat_width, device, group)
        self.world, self.stat_width = world, stat_width
        self._stat0, self._stat0_mc = 2 * n + 8, mc + 4 * (2 * n + 8)
        self.stat_out = torch.zeros(max(stat_slots * stat_width, 4), dtype=torch.float32, device=device)
        self.b, self.rank = b, self.handle.rank
        self.gen_all, self.real_all = self.buf[:n].view(world * b, d), self.buf[n:2 * n].view(world * b, d)
        self.gen_mc, self.real_mc = mc, mc + 4 * n
        self.sums_slot, self.sums_mc = self.buf[2 * n:2 * n + 8], mc + 8 * n
        self.sums_out = torch.zeros(8, dtype=torch.float32, device=device)

    def scatter(self, s_local):
        from . import kernels as K
        self.handle.barrier(channel=0)         # every rank has finished reading the previous step's scores
        K.scatter_scores_nvls(s_local, self.b, self.rank, self.gen_mc, self.real_mc)
        self.handle.barrier(channel=0)         # every rank's block has landed everywhere
        return self.gen_all, self.real_all

    def allreduce_sums(self, sums):
        """sums [6] (this rank's partial kernel sums) -> the global sums, in place, reduced inside the switch.  The slot is
        rewritten only after the next step's scatter() barriers, i.e. after every rank has read it."""
        from . import kernels as K
        self.sums_slot[:sums.numel()].copy_(sums)
        self.handle.barrier(channel=0)
        K.allreduce_small_nvls(self.sums_out, self.sums_mc, 8)
        sums.copy_(self.sums_out[:sums.numel()])

    def stat_slot(self, i):
        """This rank's slot i (stat_width floats): write the local partial sums here, then allreduce_stats(i, width)."""
        return self.buf[self._stat0 + i * self.stat_width:self._stat0 + (i + 1) * self.stat_width]

    def allreduce_stats(self, i, width):
        """The first `width` floats of slot i summed over all ranks (switch-reduced), as a local tensor.  A slot is rewritten
        one step later, after several barriers: every rank has read it by then."""
        from . import kernels as K
        out = self.stat_out[i * self.stat_width:i * self.stat_width + width]
        self.handle.barrier(channel=0)
        K.allreduce_small_nvls(out, self._stat0_mc + 4 * i * self.stat_width, width)
        return out


class SymmetricFlat(object):
    """g, w, m, v of one network in one symmetric allocation (4 x n_flat float32), the same size on every rank."""

    def __init__(self, n_flat, device, group=None):
        self.n = int(n_flat)
        self.buf, self.handle, mc = _symmetric(4 * self.n, device, group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.g, self.w, self.m, self.v = (self.buf[i * self.n:(i + 1) * self.n] for i in range(4))
        self.g_mc, self.w_mc, self.m_mc, self.v_mc = (mc + 4 * i * self.n for i in range(4))
        self.begin, self.end = shard_range(self.n, self.rank, self.world)

    def barrier(self):
        """Cross-rank barrier enqueued on the current stream (device side, through the allocation's signal pads)."""
        self.handle.barrier(channel=0)


def dp_equals_single(world, rank, device, arch_name='cifar', b=8, steps=3, loss_type='rep', seed=5):
    """Witness that the N-rank data-parallel step IS the single-process step: the same `steps` fused steps are run (a) by a
    `world`-rank engine on `b` samples per rank (CUDA graphs + collectives, the path bench.py times) and (b), on rank 0, by a
    single-process engine on the concatenated batch of world * b samples.  Checked: both losses of every step (1e-4 relative +
    1e-6), the updated flat parameter buffers of both networks (2e-3 normwise: Adam's first steps are sign-like, so summation
    order moves a few entries by 2 lr), and bit-identical replicas (torch.equal against rank 0's parameters, Adam slots and
    spectral-norm state).  The generator's batch norm is removed for this check: batch-norm statistics are per rank by design
    (DESIGN.md section 6), which is the one documented difference from the single-process step.  Collective: every rank calls."""
    from . import experiments as oa
    from .engine import SNGanEngine
    arch = oa.ARCHITECTURES[arch_name]()
    for ly in arch['generator']:
        if ly.get('act_nm') == 'bn':
            ly['act_nm'] = None
    g = torch.Generator().manual_seed(7)
    c, h, w = arch['input'][0]
    data = torch.rand(steps, world * b, c, h, w, generator=g) * 2 - 1
    code = torch.randn(steps, world * b, arch['code'][0][0], generator=g)
    eng = SNGanEngine(arch, b, loss_type=loss_type, device=device, world_size=world, rank=rank, use_graph=True, seed=seed)
    ref = SNGanEngine(arch, world * b, loss_type=loss_type, device=device, use_graph=False, seed=seed) if rank == 0 else None
    out = {'ok': True, 'arch': arch_name + ' (generator batch norm removed: BN statistics are per rank by design)', 'ranks': world,
           'per_rank_batch': b, 'steps': steps, 'loss_type': loss_type, 'max_loss_rel_diff': 0.0, 'param_rel_diff': {}}
    sl = slice(rank * b, (rank + 1) * b)
    for it in range(steps):
        lg, ld = eng.step(data[it, sl], code[it, sl])
        if rank == 0:
            lg1, ld1 = ref.step(data[it], code[it])
            for a, r in ((lg, lg1), (ld, ld1)):
                diff = abs(a - r)
                out['max_loss_rel_diff'] = max(out['max_loss_rel_diff'], diff / max(abs(r), 1e-30))
                if diff > 1e-4 * abs(r) + 1e-6:
                    out['ok'] = False
    same = torch.ones(1, device=device)
    for net in (eng.D, eng.G):
        tensors = [net.w, net.m, net.v] + [L.sn_x.view(torch.int16).float() for L in net.layers if L.has_sn]
        for t in tensors:
            t0 = t.clone()
            dist.broadcast(t0, 0)
            if not torch.equal(t0, t):
                same.zero_()
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out['replicas_bit_identical'] = bool(same.item() == 1.0)
    if rank == 0:
        for net, rnet in ((eng.D, ref.D), (eng.G, ref.G)):
            rel = float((net.w - rnet.w).norm()) / max(float(rnet.w.norm()), 1e-30)
            out['param_rel_diff'][net.name] = rel
            if rel > 2e-3:
                out['ok'] = False
    out['ok'] = bool(out['ok'] and out['replicas_bit_identical'])
    flag = torch.tensor([1.0 if out['ok'] else 0.0], device=device)
    dist.broadcast(flag, 0)
    out['ok'] = bool(flag.item() == 1.0) and out['replicas_bit_identical']
    dist.barrier()
    return out
