"""SNGanEngine -- the fused SNGan training step on one B200 (one process per GPU).

Implements what one `sess.run([loss_list, op_list, UPDATE_OPS, global_step])` of the reference does
(GeneralTools/graph_func.py:851-854 driving DeepLearning/my_sngan.py:259-323, 412-426): from ONE forward pass at the
current weights it produces loss_gen / loss_dis, the discriminator gradients of loss_dis, the generator gradients of
loss_gen (through D), applies both TF-Adam updates simultaneously, updates the batch-norm moving statistics and
every spectral-norm `in_rand`.

Data layout in HBM: GEMM operands NHWC as 16-bit planes [npl][N*H*W][C] whose sum is the (scaled) fp32 value: two fp16
planes for everything that feeds a forward launch (activations, forward weights), two bf16 planes for gradients
(include/mmdgan_b200.h); pre-batch-norm outputs, scores and
reductions raw fp32; parameters, gradients and Adam slots fp32 as one flat buffer per net in the reference's canonical
variable layouts; packed GEMM operands per layer, refreshed after every update.  The generator's last conv writes straight into rows
[B, 2B) of the discriminator input (no tf.concat copy); the discriminator backward runs on a 3B "virtual batch"
(rows: dL_D/d s_real, dL_D/d s_gen, dL_G/d s_gen) so that one dgrad chain serves both losses, while the weight
gradients use the first 2B rows only.

Multi-GPU (torch.distributed, NCCL): the batch is sharded; scores are all-gathered, each rank evaluates its row block
of the kernel matrices (exact gradients for its own rows, global 1/(B(B-1)) normalisation), and ONE all-reduce sums
the flat gradient buffers (+ the six kernel sums).  Batch-norm statistics stay per rank (documented in DESIGN.md).
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import kernels as K
from .GeneralTools.layer_func import Net, Routine
from .GeneralTools.misc_fun import FLAGS

ALIGN = 64  # floats; every variable starts on a 256-byte boundary inside the flat buffers


def _trunc_normal(gen, shape, std):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return t * std


def _fans(shape):
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def weight_initializer(gen, shape, act_fun):
    """GeneralTools/layer_func.py:14-66 with FLAGS.WEIGHT_INITIALIZER == 'default' (TF-1.8 variance scaling)."""
    if FLAGS.WEIGHT_INITIALIZER != 'default':
        raise NotImplementedError('The initializer {} is not implemented.'.format(FLAGS.WEIGHT_INITIALIZER))
    fan_in, fan_out = _fans(shape)
    if act_fun == 'relu':
        return _trunc_normal(gen, shape, math.sqrt(2.0 / fan_in))
    if act_fun == 'lrelu':
        return _trunc_normal(gen, shape, math.sqrt(2.0 / 1.01 / fan_in))
    limit = math.sqrt(3.0 / ((fan_in + fan_out) / 2.0))
    return (torch.rand(shape, dtype=torch.float64, generator=gen) * 2.0 - 1.0) * limit


class _LayerRT(object):
    """Run-time record of one layer: geometry, parameter views, packed operands and work buffers."""
    pass


class NetRuntime(object):
    """Parameters + per-layer kernel plans of one Net for a fixed number of samples per forward pass."""

    def __init__(self, routine, nimg_fwd, nimg_bwd, npass, device, gen, flat_alloc=None):
        self.routine = routine
        self.net = routine.net
        self.name = self.net.net_name
        self.npass = npass
        self.device = device
        self.layers = []
        self._jobs = None
        layers = routine.ordered_layers()
        # ---- variables in creation order: kernel, bias, gamma, beta per layer
        self.var_offsets = OrderedDict()
        off = 0
        inits = OrderedDict()
        self.state_init = OrderedDict()
        for ly in layers:
            d = ly.design
            ks = ly.kernel_shape
            inits[ly.kernel_name] = weight_initializer(gen, ks, d['act'])
            if 'bias' in ly.ops:
                inits[ly.bias_name] = _trunc_normal(gen, [ly.op_output_shape[1]], 1e-5)      # layer_func.py:745-747
            if 'BN' in ly.ops:
                c = ly.op_output_shape[1]
                inits[ly.bn_name('gamma')] = torch.ones(c, dtype=torch.float64)
                inits[ly.bn_name('beta')] = torch.zeros(c, dtype=torch.float64)
                self.state_init[ly.bn_name('moving_mean')] = torch.zeros(c, dtype=torch.float64)
                self.state_init[ly.bn_name('moving_variance')] = torch.ones(c, dtype=torch.float64)
            if d.get('w_nm') == 's':
                self.state_init[ly.sn_name] = _trunc_normal(gen, ly.sn_x_shape, 1.0)          # math_func.py:565-567
        for name, t in inits.items():
            self.var_offsets[name] = (off, list(t.shape))
            off += (t.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.n_flat = off
        # flat_alloc (data-parallel NVLS mode): g, w, m, v live in one symmetric allocation bound to an NVSwitch multicast
        # object (parallel.SymmetricFlat); every view / pointer below is taken from these tensors, so nothing else changes
        self.flat = flat_alloc(off) if flat_alloc is not None else None
        if self.flat is not None:
            self.w, self.g, self.m, self.v = self.flat.w, self.flat.g, self.flat.m, self.flat.v
        else:
            self.w = torch.zeros(off, dtype=torch.float32, device=device)
            # ALIGN spare floats behind the gradients: the generator's tail carries the six MMD kernel sums, so that the data-parallel
            # step reduces them in the SAME all-reduce as the gradients (the optimiser only ever sees the first n_flat entries)
            self.g_all = torch.zeros(off + ALIGN, dtype=torch.float32, device=device)
            self.g = self.g_all[:off]
            self.m = torch.zeros(off, dtype=torch.float32, device=device)
            self.v = torch.zeros(off, dtype=torch.float32, device=device)
        for name, t in inits.items():
            self.view(self.w, name).copy_(t.reshape(-1).float())
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        # ---- per-layer runtime
        for idx, ly in enumerate(layers):
            self.layers.append(self._build_layer(idx, ly, nimg_fwd, nimg_bwd))
        for name, t in self.state_init.items():
            self.set_state(name, t.float())
        self.refresh()

    # -------------------------------------------------------------------------------------------- variables
    def view(self, flat, name):
        off, shape = self.var_offsets[name]
        return flat[off:off + int(np.prod(shape))]

    def num_params(self):
        return sum(int(np.prod(s)) for _, s in self.var_offsets.values())

    def get_variable(self, name):
        """Canonical (reference-layout) copy of a trainable variable."""
        off, shape = self.var_offsets[name]
        return self.view(self.w, name).reshape(shape).clone()

    def get_grad(self, name):
        off, shape = self.var_offsets[name]
        return self.view(self.g, name).reshape(shape).clone()

    def set_variable(self, name, value):
        self.view(self.w, name).copy_(torch.as_tensor(value).reshape(-1).float())

    def _feat_perm(self, L):
        return L.lop.out_flat if L.op == 'd' else (L.Cout, 1)

    def get_state(self, name):
        """BN moving statistics / SN in_rand in the reference's layout ([C] / NCHW)."""
        for L in self.layers:
            if L.has_bn and name in (L.ly.bn_name('moving_mean'), L.ly.bn_name('moving_variance')):
                src = L.mm if name.endswith('moving_mean') else L.mv
                out = torch.empty(L.Cout, device=self.device)
                c, hw = self._feat_perm(L)
                K.permute_features(src, out, L.Cout, c, hw, inverse=True)
                return out
            if L.has_sn and name == L.ly.sn_name:
                shp = L.ly.sn_x_shape
                if len(shp) == 2:
                    c, hw = L.sn_flat
                    out = torch.empty(shp[1], device=self.device)
                    K.permute_features(K.planes_value(L.sn_x).reshape(-1)[:shp[1]].contiguous(), out, shp[1], c, hw, inverse=True)
                    return out.reshape(shp)
                return K.planes_to_nchw(L.sn_x, 1, shp[1], shp[2], shp[3])
        raise KeyError(name)

    def set_state(self, name, value):
        value = torch.as_tensor(value).float().to(self.device).contiguous()
        for L in self.layers:
            if L.has_bn and name in (L.ly.bn_name('moving_mean'), L.ly.bn_name('moving_variance')):
                dst = L.mm if name.endswith('moving_mean') else L.mv
                c, hw = self._feat_perm(L)
                K.permute_features(value.reshape(-1), dst, L.Cout, c, hw)
                return
            if L.has_sn and name == L.ly.sn_name:
                shp = L.ly.sn_x_shape
                if len(shp) == 2:
                    c, hw = L.sn_flat
                    tmp = torch.zeros(L.sn_x.shape[1] * L.sn_x.shape[2], device=self.device)
                    K.permute_features(value.reshape(-1), tmp, shp[1], c, hw)
                    K.to_planes(tmp, L.sn_x)
                else:
                    K.nchw_to_planes(value.reshape(shp), L.sn_x)
                return
        raise KeyError(name)

    def state_names(self):
        return list(self.state_init.keys())

    # -------------------------------------------------------------------------------------------- layers
    def _build_layer(self, idx, ly, nf, nb):
        L = _LayerRT()
        d = ly.design
        dev, npass = self.device, self.npass
        L.ly, L.idx, L.op = ly, idx, d['op']
        L.act, L.act_code = d['act'], K.ACT[d['act']]
        L.has_bias, L.has_bn, L.has_sn = 'bias' in ly.ops, 'BN' in ly.ops, d.get('w_nm') == 's'
        L.act_k = float(d['act_k']) if L.has_sn else 1.0
        ins, outs = ly.op_input_shape[1:], ly.op_output_shape[1:]
        in_flat = out_flat = None
        if L.op == 'd':
            prev = self.layers[idx - 1] if idx > 0 else None
            # the input features of a dense layer that follows a conv layer are an NCHW-flatten of an NHWC buffer
            if prev is not None and prev.op != 'd':
                in_flat = (prev.Cout, prev.Hout * prev.Wout)
            if d['out_reshape'] is not None and len(d['out_reshape']) == 3:
                c, h, w = d['out_reshape']
                out_flat = (c, h * w)
        L.lop = K.LinearOp(L.op, ins, outs, d.get('kernel', 1), d.get('strides', 1), npass=npass, in_flat=in_flat,
                           out_flat=out_flat, device=dev)
        lop = L.lop
        L.Cin, L.Cout, L.Hin, L.Win, L.Hout, L.Wout = lop.Cin, lop.Cout, lop.Hin, lop.Win, lop.Hout, lop.Wout
        L.rows_in, L.rows_out = L.Hin * L.Win, L.Hout * L.Wout      # per sample
        L.Cs_in, L.Cs_out = lop.Cs_in, lop.Cs_out
        # ---- parameter-derived internal vectors (channel padded, feature permuted)
        z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
        if L.has_bias:
            L.bias_int = z(L.Cs_out)
        if L.has_bn:
            L.gamma_int, L.beta_int = z(L.Cs_out), z(L.Cs_out)
            L.mm, L.mv, L.mean, L.invstd = z(L.Cs_out), z(L.Cs_out), z(L.Cs_out), z(L.Cs_out)
            L.dgamma_int, L.dbeta_int = z(L.Cs_out), z(L.Cs_out)
        # ---- activations (forward batch nf) and gradients (backward batch nb)
        L.a = None            # output activation planes; set by the owner (may alias the D input buffer)
        if L.has_bn:
            L.zraw = torch.zeros((1, nf * L.rows_out, L.Cs_out), dtype=torch.float32, device=dev)
            Tt = lop.fwd_tiles(nf)
            L.ps, L.pq, L.T_fwd = z(Tt * L.Cs_out), z(Tt * L.Cs_out), Tt
            L.da_raw = torch.zeros((1, nb * L.rows_out, L.Cs_out), dtype=torch.float32, device=dev)
            L.rpb = max(1, (nb * L.rows_out + 295) // 296)
            nblk = (nb * L.rows_out + L.rpb - 1) // L.rpb
            L.bp1, L.bp2, L.nblk = z(nblk * L.Cs_out), z(nblk * L.Cs_out), nblk
        L.dz = None           # gradient w.r.t. the op output (pre-bias / pre-BN), planes; set by the owner
        # weight-gradient workspace
        R, NC, bn, splits, P = lop.wgrad_plan(nf)
        L.wg_splits = splits
        L.wg_parts = z(splits * R * NC)
        L.n_dots = K.lib().mmdgan_wgrad_reduce_blocks(R * NC)
        L.dots = torch.zeros(L.n_dots, dtype=torch.float64, device=dev)
        # column-sum workspace of the dgrad that PRODUCES this layer's dz (bias gradient), sized by the owner
        L.cs = None
        L.cs_T = 0
        # ---- spectral norm state (batch-1 power iteration)
        if L.has_sn:
            L.sn_lop = lop
            if ly.sn_pim:
                # PIM ('sn_paper'): power iteration on the canonical kernel viewed as the [k*k*C, C'] matrix -- the same memory,
                # so d(sigma)/dW lands in the canonical layout and the gradient combine is unchanged
                rows, cols = int(np.prod(ly.kernel_shape[:3])), int(ly.kernel_shape[3])
                L.sn_lop = K.LinearOp('d', [rows], [cols], npass=npass, device=dev)
                x_is_input = ly.use_u
                r_x, c_x = (1, L.sn_lop.Cs_in) if x_is_input else (1, L.sn_lop.Cs_out)
                r_y, c_y = (1, L.sn_lop.Cs_out) if x_is_input else (1, L.sn_lop.Cs_in)
            else:
                x_is_input = ly.use_u if L.op != 'tc' else (not ly.use_u)
                r_x, c_x = (L.rows_in, L.Cs_in) if x_is_input else (L.rows_out, L.Cs_out)
                r_y, c_y = (L.rows_out, L.Cs_out) if x_is_input else (L.rows_in, L.Cs_in)
            L.sn_x_is_input = x_is_input
            # the vector that enters the layer op is a forward operand (value planes: fp16 in the parity mode); the one that
            # enters the adjoint meets the bf16 input-gradient weights with six plane pairs
            npl = K.mode_planes(npass)
            vp = lambda r, c: K.new_value_planes(r, c, npass, dev)
            bp = lambda r, c: K.new_planes(r, c, npl, dev)
            L.sn_x = (vp if x_is_input else bp)(r_x, c_x)
            L.sn_xnew = (vp if x_is_input else bp)(r_x, c_x)
            L.sn_y = (bp if x_is_input else vp)(r_y, c_y)
            # bf16 re-split of the fp16 vector for the d(sigma)/dW weight-gradient GEMM (an MMA cannot mix fp16 with bf16)
            L.sn_b16 = K.new_planes(r_x if x_is_input else r_y, c_x if x_is_input else c_y, 2, dev)
            L.sn_v = torch.zeros((1, r_y, c_y), dtype=torch.float32, device=dev)
            L.sn_w = torch.zeros((1, r_x, c_x), dtype=torch.float32, device=dev)
            L.sigma = torch.ones(1, dtype=torch.float32, device=dev)
            Rs, NCs, _, sps, _ = L.sn_lop.wgrad_plan(1)
            L.sn_splits = sps
            L.sn_parts = z(sps * Rs * NCs)
            L.sn_S = z(lop.canon_numel)
            if L.op == 'd' or ly.sn_pim:
                L.sn_flat = (L.sn_lop.in_flat if x_is_input else L.sn_lop.out_flat)
        return L

    def refresh(self):
        """canonical parameters -> packed GEMM operands and internal (padded / permuted) vectors, ONE launch per net."""
        if self._jobs is None:
            packs, perms = [], []
            for L in self.layers:
                packs += L.lop.pack_descs(self.view(self.w, L.ly.kernel_name))
                if L.has_sn and L.sn_lop is not L.lop:
                    packs += L.sn_lop.pack_descs(self.view(self.w, L.ly.kernel_name))
                c, hw = self._feat_perm(L)
                if L.has_bias:
                    perms.append((self.view(self.w, L.ly.bias_name), L.bias_int, L.Cout, c, hw))
                if L.has_bn:
                    perms.append((self.view(self.w, L.ly.bn_name('gamma')), L.gamma_int, L.Cout, c, hw))
                    perms.append((self.view(self.w, L.ly.bn_name('beta')), L.beta_int, L.Cout, c, hw))
            self._jobs = K.build_refresh_jobs(packs, perms, self.device)
        for L in self.layers:
            L.lop.pre_refresh()
        K.refresh(*self._jobs)


class SNGanEngine(object):
    def __init__(self, architecture, batch_size, loss_type='rep', rep_weights=(0.0, -1.0), lr_list=(5e-4, 2e-4),
                 seed=2, npass=None, device='cuda', world_size=1, rank=0, process_group=None, use_graph=True):
        from ._lib import check as _check
        _check(K.lib().mmdgan_check_device())
        self.arch = architecture
        self.B = int(batch_size)                 # per-GPU batch of real images (= batch of codes)
        self.loss_type = loss_type
        self.rep_weights = list(rep_weights)
        self.lr_dis, self.lr_gen = float(lr_list[0]), float(lr_list[1])
        self.npass = FLAGS.TENSOR_PASSES if npass is None else npass
        self.device = torch.device(device)
        self.world_size, self.rank, self.pg = world_size, rank, process_group
        self.use_graph = use_graph
        if self.npass not in (1, 3):
            raise ValueError('TENSOR_PASSES must be 3 (parity: fp16x3 forward / bf16x3 gradients) or 1 (single bf16 pass)')
        self.om = 0                                # GEMM outputs that feed another GEMM are written as bf16 planes
        self.code_size = architecture['code'][0][0]
        self.channels, self.height, self.width = architecture['input'][0]
        self.score_size = architecture['discriminator'][-1]['out']
        B = self.B
        # ---- nets as the reference builds them (my_sngan.py:85-108)
        g_net = Net(architecture['generator'], net_name='gen', data_format=FLAGS.IMAGE_FORMAT, num_class=0)
        self.Gen = Routine(g_net)
        self.Gen.add_input_layers([64, self.code_size], [0])
        self.Gen.seq_links(list(range(g_net.num_layers)))
        self.Gen.add_output_layers([g_net.num_layers - 1])
        d_net = Net(architecture['discriminator'], net_name='dis', data_format=FLAGS.IMAGE_FORMAT, num_class=0)
        self.Dis = Routine(d_net)
        self.Dis.add_input_layers([64] + list(architecture['input'][0]), [0])
        self.Dis.seq_links(list(range(d_net.num_layers)))
        self.Dis.add_output_layers([d_net.num_layers - 1])
        gen = torch.Generator().manual_seed(seed)
        # opt-in: gradient all-reduce fused with Adam through NVSwitch multicast (csrc/nvls.cu) instead of NCCL + adam_kernel
        self.nvls = world_size > 1 and os.environ.get('MMDGAN_NVLS_ADAM', '0') == '1'
        # opt-in on top of it: batch-norm statistics over the GLOBAL batch (sums exchanged through the same multicast loads), so
        # that the data-parallel step equals the single-GPU step at the global batch size including the generator's batch norm
        self.sync_bn = self.nvls and os.environ.get('MMDGAN_SYNC_BN', '0') == '1'
        flat_alloc = None
        if self.nvls:
            from . import parallel
            flat_alloc = lambda n: parallel.SymmetricFlat(n, self.device, process_group)      # noqa: E731
        # The fused backward pass implements what the shipped experiments use: spectral normalisation in the discriminator only,
        # batch norm in the generator only.  The layer DSL accepts the other combinations; training them
        # here would silently use wrong gradients, so they are refused up front.
        for ly in self.Gen.ordered_layers():
            if ly.design.get('w_nm') == 's':
                raise NotImplementedError('{}: spectral normalisation in the generator is not on the fused path'.format(ly.layer_scope))
        for ly in self.Dis.ordered_layers():
            if 'BN' in ly.ops:
                raise NotImplementedError('{}: batch normalisation in the discriminator is not on the fused path'.format(ly.layer_scope))
        self.G = NetRuntime(self.Gen, B, B, self.npass, self.device, gen, flat_alloc)
        self.D = NetRuntime(self.Dis, 2 * B, 3 * B, self.npass, self.device, gen, flat_alloc)
        self._alloc_buffers()
        self.mmd = K.MmdKernel(loss_type, self.rep_weights, b=B, device=self.device)
        if world_size > 1 and not self.nvls:
            self.mmd.sums = self.G.g_all[self.G.n_flat:self.G.n_flat + 6]      # reduced together with the generator's gradients
        self._comm_work = None
        self.global_step = 0
        self.nan_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        # set by any producer whose value does not fit the fp16 forward planes (|activation| >= 4094): reported by step()
        self.sat_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        for L in self.G.layers + self.D.layers:
            L.lop.sat_flag = self.sat_flag
        self._graph_cache = {}                   # update mask -> captured CUDA graphs
        self.update_mask = (True, True)          # (run dis_op, run gen_op) of the step being enqueued
        self._warm = False
        self._stream = torch.cuda.Stream(device=self.device)
        self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(8)]
        # weight-gradient GEMMs + gradient finalisation.  Its kernel nodes get the higher priority: the optimiser updates wait for
        # the LAST weight gradient, so this stream -- not the input-gradient chain on the capture stream -- ends the step
        # (measured: 4.40 ms/step with it, 4.49 without, 4.60 with the priorities the other way round)
        self._grad_stream = torch.cuda.Stream(device=self.device,
                                              priority=-1 if os.environ.get('MMDGAN_GRAD_PRIORITY', '1') == '1' else 0)
        self._upd_stream = torch.cuda.Stream(device=self.device)    # the generator's Adam + refresh
        self.sn_fork = True
        self.grad_fork = True
        self._dis_updated = False
        # pinned staging for the end-to-end path
        self._pin_data = torch.empty((B, self.channels, self.height, self.width), dtype=torch.float32).pin_memory()
        self._pin_code = torch.empty((B, self.code_size), dtype=torch.float32).pin_memory()
        self._pin_loss = torch.empty(2, dtype=torch.float32).pin_memory()
        self._pin_sat = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._dev_data = torch.empty_like(self._pin_data, device=self.device)
        self._dev_code = torch.empty_like(self._pin_code, device=self.device)
        # host -> device prefetch of the NEXT batch (step(..., prefetch=...)): copy stream + two staging buffers
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._stage_bufs = [(torch.empty_like(self._dev_data), torch.empty_like(self._dev_code)) for _ in range(2)]
        self._pin_bufs = [(torch.empty_like(self._pin_data).pin_memory(), torch.empty_like(self._pin_code).pin_memory()) for _ in range(2)]
        # codes drawn on the device (tf.random_normal inside the graph, my_sngan.py:122-124): Philox keyed by torch's seed at
        # construction, one draw counter per step in device memory so that the captured graph stays replayable
        self.code_seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * int(rank)) & 0xFFFFFFFFFFFFFFFF     # every rank its own stream
        self.code_draw = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.device_codes = False                # set per step: True when step() / stage() is given no codes
        self._pin_res = [torch.zeros(2, dtype=torch.float32).pin_memory() for _ in range(2)]      # results of the two steps in flight
        self._pin_flag = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]
        self._res_next = 0
        self._prefetched = None                  # (data tensor, code tensor, buffer index, event) of the batch in flight
        self._slot_free = [None, None]           # event: the last device-to-device read of a staging slot
        self._stage_next = 0
        self.kernel_launches_per_step = None

    # -------------------------------------------------------------------------------------------- buffers
    def _alloc_buffers(self):
        B, dev, npass = self.B, self.device, self.npass
        HW = self.height * self.width
        ng = K.mode_planes(npass, 'grad')
        vplanes = lambda rows, c: K.new_value_planes(rows, c, npass, dev)
        self.code_planes = vplanes(B, K.pad_c(self.code_size))
        # D input: rows [0, B*HW) real, [B*HW, 2B*HW) generated
        self.x_all = vplanes(2 * B * HW, K.pad_c(self.channels))
        # generator activations
        for i, L in enumerate(self.G.layers):
            last = i == len(self.G.layers) - 1
            L.a = self.x_all[:, B * HW:, :] if last else vplanes(B * L.rows_out, L.Cs_out)
            L.dz = K.new_planes(B * L.rows_out, L.Cs_out, ng, dev)
        # discriminator activations (2B) and gradients (3B virtual batch)
        for i, L in enumerate(self.D.layers):
            last = i == len(self.D.layers) - 1
            if last:
                L.a = torch.zeros((1, 2 * B, L.Cs_out), dtype=torch.float32, device=dev)      # scores, fp32
                L.raw_out = True
            else:
                L.a = vplanes(2 * B * L.rows_out, L.Cs_out)
            L.dz = K.new_planes(3 * B * L.rows_out, L.Cs_out, ng, dev)
            if last:
                L.dz_f32 = torch.zeros((3 * B, L.Cs_out), dtype=torch.float32, device=dev)    # score gradients from the MMD kernel
        # column-sum workspaces: the dgrad of layer i+1 produces dz of layer i
        for net, nb in ((self.G, B), (self.D, 3 * B)):
            for i, L in enumerate(net.layers[:-1]):
                nxt = net.layers[i + 1]
                L.cs_T = nxt.lop.dgrad_tiles(nb)
                L.cs = torch.zeros(L.cs_T * nxt.lop.d['ncols'], dtype=torch.float32, device=dev)
        # the generator's last layer gets its dz (and bias column sums) from the dgrad of D's first layer on B fakes
        gl, d0 = self.G.layers[-1], self.D.layers[0]
        gl.cs_T = d0.lop.dgrad_tiles(B)
        gl.cs = torch.zeros(gl.cs_T * d0.lop.d['ncols'], dtype=torch.float32, device=dev)
        # bf16 re-split of a layer's fp16 input activation for its weight-gradient GEMM; one buffer: those GEMMs run in order on
        # the gradient stream
        acts = [self.x_all, self.code_planes] + [L.a for L in self.G.layers + self.D.layers if L.a.dtype == torch.float16]
        need = max(t.shape[1] * t.shape[2] for t in acts)
        for net, nimg in ((self.G, B), (self.D, 2 * B)):        # image layers as 27-column dense products: [pixels][32] operands
            for L in net.layers:
                if L.lop.img_op is not None:
                    need = max(need, nimg * L.rows_in * 32)
        self.wg_scratch = torch.zeros(2 * need, dtype=torch.bfloat16, device=dev)
        # bf16 re-split ("twin") of every fp16 activation that is a weight-gradient operand (an MMA cannot mix fp16 with bf16).  The
        # twins are written DURING THE FORWARD PASS on the conversion stream, where the device has a free stream, instead of in
        # front of each weight-gradient GEMM on the gradient stream, which is the critical path of the backward pass (measured:
        # the 28 conversions cost 0.35 ms of step time there).
        self._twins = {}
        if os.environ.get('MMDGAN_TWINS', '1') == '1':
            operands = [self.x_all, self.code_planes] + [L.a for L in self.G.layers[:-1] + self.D.layers[:-1] if L.a.dtype == torch.float16]
            for t in operands:
                self._twins[t.data_ptr()] = torch.zeros((2, t.shape[1], t.shape[2]), dtype=torch.bfloat16, device=dev)
        self._cvt_stream = torch.cuda.Stream(device=dev)
        self._cvt_event = None
        self.tmp_vec = torch.zeros(max(max(L.Cs_out for L in self.G.layers), max(L.Cs_out for L in self.D.layers)) + 64,
                                   dtype=torch.float32, device=dev)
        if self.world_size > 1:
            d = self.D.layers[-1].Cs_out
            self.s_gather = torch.zeros((self.world_size, 2 * B, d), dtype=torch.float32, device=dev)
            self.gen_all = torch.zeros((self.world_size * B, d), dtype=torch.float32, device=dev)
            self.real_all = torch.zeros((self.world_size * B, d), dtype=torch.float32, device=dev)
            if self.nvls:        # the gathered score matrices live in a symmetric allocation: peers multicast their blocks into it
                from . import parallel
                bn_layers = [L for L in self.G.layers if L.has_bn] if self.sync_bn else []
                for k, L in enumerate(bn_layers):
                    L.bn_slot = 2 * k            # slot 2k: sum x | sum x^2 (forward); slot 2k + 1: sum dy | sum dy * xhat (backward)
                width = 2 * max([L.Cs_out for L in bn_layers] + [2])
                self.sym_scores = parallel.SymmetricScores(B, d, dev, self.pg, stat_slots=2 * len(bn_layers), stat_width=width)
                self.gen_all, self.real_all = self.sym_scores.gen_all, self.sym_scores.real_all
            # The all-reduce of the discriminator's gradients runs BESIDE the generator's backward pass.  With NCCL's default
            # channel count its CTAs (one SM each, scattered over the TPCs) keep the persistent CTA-pair GEMMs off part of the
            # GPU for the length of the transfer: measured on 2 B200 (scripts/dp_phase_times.py), the generator's backward graph
            # takes 0.74-0.76 ms next to the all-reduce against 0.58 ms alone; with that one all-reduce on its own communicator
            # capped at 4 CTAs it takes 0.65 ms (sum of segments 4.14 -> 3.97 ms).  EXPERIMENT, off by default (MMDGAN_AR_CTAS=<n>
            # turns it on): in the whole bench line the gain did not show (4.00 / 4.05 ms against 3.98-4.09 ms), DESIGN.md section 6.
            self.pg_overlap = self.pg
            ar_ctas = int(os.environ.get('MMDGAN_AR_CTAS', '0'))
            if not self.nvls and ar_ctas > 0:
                import torch.distributed as dist
                try:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = ar_ctas
                    opts.config.min_ctas = 1
                    ranks = dist.get_process_group_ranks(self.pg) if self.pg is not None else list(range(dist.get_world_size()))
                    self.pg_overlap = dist.new_group(ranks=ranks, backend='nccl', pg_options=opts)
                except (AttributeError, RuntimeError, TypeError):      # a backend without these options (gloo in the CPU tests)
                    self.pg_overlap = self.pg

    # -------------------------------------------------------------------------------------------- forward passes
    @staticmethod
    def _as_rows(planes, rows, c):
        """Reinterpret [npl, r0, c0] planes as [npl, rows, c] (same memory; NHWC flatten / unflatten)."""
        npl = planes.shape[0]
        assert planes.shape[1] * planes.shape[2] == rows * c
        return planes.as_strided((npl, rows, c), (planes.stride(0), c, 1), planes.storage_offset())

    def _net_forward(self, net, src, nimg, is_training=True, sigma_on=True, update_moving=True):
        """update_moving=False: training-mode batch norm WITHOUT its UPDATE_OPS (the moving averages are assigned only by the
        training sess.run, graph_func.py:848-854; a sampling call between steps must not advance them)."""
        for L in net.layers:
            mm, mv = (L.mm, L.mv) if (L.has_bn and update_moving) else (None, None)
            bessel = L.op != 'd'      # rank-2 batch norm (dense layer): TF 1.8 falls back from the fused kernel to nn.moments
            lop = L.lop
            src = self._as_rows(src, nimg * L.rows_in, L.Cs_in)
            sig = L.sigma if (L.has_sn and sigma_on) else None
            if L.has_bn:
                lop.forward(src, nimg, L.zraw, sigma=sig, alpha_k=L.act_k, out_mode=2, colsum=L.ps, colsumsq=L.pq)
                if is_training and self.sync_bn:
                    c = L.Cs_out
                    slot = self.sym_scores.stat_slot(L.bn_slot)
                    K.reduce_tiles(L.ps, L.T_fwd, c, slot[:c])
                    K.reduce_tiles(L.pq, L.T_fwd, c, slot[c:2 * c])
                    tot = self.sym_scores.allreduce_stats(L.bn_slot, 2 * c)
                    K.bn_finalize(tot[:c], tot[c:2 * c], 1, c, nimg * L.rows_out * self.world_size, L.mean, L.invstd, mm, mv, bessel=bessel)
                    K.bn_apply(L.zraw, L.mean, L.invstd, L.gamma_int, L.beta_int, L.Cs_out, nimg * L.rows_out * L.Cs_out,
                               L.act_code, L.a, sat_flag=self.sat_flag)
                elif is_training:
                    K.bn_finalize(L.ps, L.pq, L.T_fwd, L.Cs_out, nimg * L.rows_out, L.mean, L.invstd, mm, mv, bessel=bessel)
                    K.bn_apply(L.zraw, L.mean, L.invstd, L.gamma_int, L.beta_int, L.Cs_out, nimg * L.rows_out * L.Cs_out,
                               L.act_code, L.a, sat_flag=self.sat_flag)
                else:                            # tf.layers.batch_normalization(training=False): the moving averages normalise
                    K.bn_inference_stats(L.mm, L.mv, L.Cs_out, L.mean, L.invstd)
                    K.bn_apply(L.zraw, L.mean, L.invstd, L.gamma_int, L.beta_int, L.Cs_out, nimg * L.rows_out * L.Cs_out,
                               L.act_code, L.a, sat_flag=self.sat_flag)
            else:
                out_mode = 2 if getattr(L, 'raw_out', False) else self.om
                lop.forward(src, nimg, L.a, sigma=sig, alpha_k=L.act_k, bias=L.bias_int if L.has_bias else None,
                            act=L.act_code, out_mode=out_mode)
            src = L.a
            if is_training and update_moving:
                self._make_twin(L.a)
        return src

    def _make_twin(self, t):
        """bf16 twin of an fp16 activation, on the conversion stream (forked from the current stream, joined by the gradient
        stream before its first weight-gradient GEMM)."""
        tw = self._twins.get(t.data_ptr()) if t.dtype == torch.float16 else None
        if tw is None:
            return
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._cvt_stream.wait_event(ev)
        with torch.cuda.stream(self._cvt_stream):
            K.convert_planes(t, tw)
            self._cvt_event = torch.cuda.Event()
            self._cvt_event.record(self._cvt_stream)

    def _sn_layer(self, L):
        """One PICO power iteration of one spectrally-normalised layer (math_func.py:661-672): sigma = ||F(x)||,
        x' = l2n(F^T(l2n(F(x)))) and S = d(sigma)/dW = wgrad(x, u)."""
        lop = L.sn_lop
        if L.sn_x_is_input:
            lop.forward(L.sn_x, 1, L.sn_v, out_mode=2)
            K.sn_normalize(L.sn_v, L.sn_v.numel(), L.sn_y, sigma_out=L.sigma, eps=FLAGS.EPSI)
            lop.dgrad(L.sn_y, 1, L.sn_w, out_mode=2, npass=lop.adj_npass)
            K.sn_normalize(L.sn_w, L.sn_w.numel(), L.sn_xnew, eps=FLAGS.EPSI)
            sx = K.convert_planes(L.sn_x, L.sn_b16) if L.sn_x.dtype == torch.float16 else L.sn_x
            lop.wgrad(sx, L.sn_y, 1, L.sn_parts, L.sn_splits)
        else:
            lop.dgrad(L.sn_x, 1, L.sn_v, out_mode=2, npass=lop.adj_npass)    # the adjoint is the FORWARD operator here: sigma = ||F^T x||
            K.sn_normalize(L.sn_v, L.sn_v.numel(), L.sn_y, sigma_out=L.sigma, eps=FLAGS.EPSI)
            lop.forward(L.sn_y, 1, L.sn_w, out_mode=2)
            K.sn_normalize(L.sn_w, L.sn_w.numel(), L.sn_xnew, eps=FLAGS.EPSI)
            sy = K.convert_planes(L.sn_y, L.sn_b16) if L.sn_y.dtype == torch.float16 else L.sn_y
            lop.wgrad(sy, L.sn_x, 1, L.sn_parts, L.sn_splits)
        lop.wgrad_reduce(L.sn_parts, L.sn_splits, 1, L.sn_S)

    def _sn_power_iteration(self, fork=True):
        """All spectral-norm layers; the per-layer chains are independent (batch-1, latency-bound), so they are forked
        onto side streams and joined before the discriminator forward (they overlap the generator forward)."""
        layers = [L for L in self.D.layers + self.G.layers if L.has_sn]
        if not fork or not layers:
            for L in layers:
                self._sn_layer(L)
            return []
        main = torch.cuda.current_stream(self.device)
        ev0 = torch.cuda.Event()
        ev0.record(main)
        joins = []
        for i, L in enumerate(layers):
            st = self._side_streams[i % len(self._side_streams)]
            st.wait_event(ev0)
            with torch.cuda.stream(st):
                self._sn_layer(L)
        for st in self._side_streams[:len(layers)]:
            ev = torch.cuda.Event()
            ev.record(st)
            joins.append(ev)
        return joins

    # -------------------------------------------------------------------------------------------- step pieces
    def _phase_forward(self):
        B, HW = self.B, self.height * self.width
        if self.device_codes:                     # SNGan.sample_codes: z ~ N(0, 1) drawn here, no host round trip
            K.sample_normal(self._dev_code, self.code_seed, self.code_draw)
            K.incr_counter(self.code_draw)
        K.nchw_to_planes(self._dev_code, self.code_planes)
        K.nchw_to_planes(self._dev_data, self.x_all[:, :B * HW, :])
        joins = self._sn_power_iteration(fork=self.sn_fork)
        self._make_twin(self.code_planes)
        self._net_forward(self.G, self.code_planes, B)
        self._make_twin(self.x_all)              # real rows from the input conversion, generated rows from G's last layer
        main = torch.cuda.current_stream(self.device)
        for ev in joins:
            main.wait_event(ev)
        self._net_forward(self.D, self.x_all, 2 * B)
        if self._cvt_event is not None:          # join: the step's later phases (and the end of a captured graph) see the twins
            main.wait_event(self._cvt_event)
            self._cvt_event = None

    def _phase_loss(self):
        B = self.B
        s = self.D.layers[-1].a[0]                      # [2B, d]: rows [0,B) real, [B,2B) generated (my_sngan.py:279)
        seed = self.D.layers[-1].dz_f32                 # fp32 [3B, d]: dL_D/ds_real, dL_D/ds_gen, dL_G/ds_gen
        if self.world_size > 1:
            self.mmd(s[B:], s[:B], seed[2 * B:], seed[B:2 * B], seed[:B], gen_all=self.gen_all, real_all=self.real_all,
                     row0=self.rank * B)
        else:
            self.mmd(s[B:], s[:B], seed[2 * B:], seed[B:2 * B], seed[:B])
        K.to_planes(seed, self.D.layers[-1].dz)         # operand planes of the first input-gradient GEMM

    def _bias_grad_from_colsum(self, net, L, ncols_per_row_group):
        """bias gradient of layer L from the per-tile column sums of the dgrad that produced L.dz."""
        groups = ncols_per_row_group // L.Cs_out          # > 1 when dz was produced as a flattened dense input
        K.reduce_tiles(L.cs, L.cs_T * groups, L.Cs_out, self.tmp_vec)
        c, hw = net._feat_perm(L)
        K.permute_features(self.tmp_vec, net.view(net.g, L.ly.bias_name), L.Cout, c, hw, inverse=True)

    def _weight_grad(self, net, L, x_in, dz, nimg):
        lop = L.lop
        x_in = self._as_rows(x_in, x_in.shape[1] * x_in.shape[2] // L.Cs_in, L.Cs_in)
        tw = self._twins.get(x_in.data_ptr()) if x_in.dtype == torch.float16 else None
        if tw is not None:                       # written during the forward pass
            x_in = self._as_rows(tw, x_in.shape[1], x_in.shape[2])
        lop.wgrad(x_in, dz, nimg, L.wg_parts, L.wg_splits, scratch=self.wg_scratch)

    def _reduce_net(self, net, nimg):
        """Split-K partials of EVERY layer of `net` -> canonical gradients (+ the spectral-norm combine) in two batched launches,
        issued once after the net's last weight-gradient GEMM.  (One reduction per layer right behind its GEMM -- 28 latency-bound
        launches per step -- cost 0.3 ms of step time.)"""
        if getattr(net, '_red_jobs', None) is None:
            descs, combos = [], []
            for L in net.layers:
                gview = net.view(net.g, L.ly.kernel_name)
                wv = net.view(net.w, L.ly.kernel_name) if L.has_sn else None
                descs.append(L.lop.wgrad_reduce_desc(L.wg_parts, L.wg_splits, nimg, gview, w_canon=wv, dots=L.dots if L.has_sn else None))
                if L.has_sn:
                    combos.append((gview, L.sn_S, L.dots, descs[-1][1], L.sigma, L.act_k, L.lop.canon_numel))
            net._red_jobs = K.build_wred_jobs(descs, self.device)
            net._cmb_jobs = K.build_sn_combine_jobs(combos, self.device) if combos else None
        K.wgrad_reduce_batched(*net._red_jobs)
        if net._cmb_jobs is not None:
            K.sn_grad_combine_batched(*net._cmb_jobs)

    def _phase_backward(self, part='all'):
        """Input-gradient chain on the main stream; everything that only FINALISES gradients (weight-gradient GEMMs, split-K
        reductions, the spectral-norm combine, bias / batch-norm parameter reductions) is forked onto one in-order side
        stream and joined before the optimiser, so the ~100 small launches overlap the GEMM chain instead of serialising it.
        part = 'dis' / 'gen': the data-parallel step captures the discriminator half and the generator half as separate graphs,
        so that the all-reduce of the discriminator's gradients runs while the generator's are still being computed."""
        B, HW = self.B, self.height * self.width
        D, G = self.D, self.G
        main = torch.cuda.current_stream(self.device)
        side = self._grad_stream if self.grad_fork else main
        pending = []          # closures that may run once everything issued so far on the main stream has completed

        def flush():
            if not pending:
                return
            if side is not main:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
            with torch.cuda.stream(side):
                for fn in pending:
                    fn()
            del pending[:]

        # ================= discriminator: loss_dis -> D variables (rows [0,2B)), loss_gen -> dx_fake (rows [2B,3B))
        last = D.layers[-1]
        if part != 'gen' and last.has_bias:
            def last_bias():
                K.colsum_small(last.dz_f32, 2 * B, last.Cs_out, self.tmp_vec)
                c, hw = D._feat_perm(last)
                K.permute_features(self.tmp_vec, D.view(D.g, last.ly.bias_name), last.Cout, c, hw, inverse=True)
            pending.append(last_bias)
        for i in (range(len(D.layers) - 1, -1, -1) if part != 'gen' else ()):
            L = D.layers[i]
            x_in = D.layers[i - 1].a if i > 0 else self.x_all
            pending.append(lambda L=L, x_in=x_in: self._weight_grad(D, L, x_in, L.dz, 2 * B))
            flush()
            sig = L.sigma if L.has_sn else None
            if i > 0:
                P = D.layers[i - 1]
                dzp = self._as_rows(P.dz, 3 * B * L.rows_in, L.Cs_in)
                aux = self._as_rows(P.a, 2 * B * L.rows_in, L.Cs_in)
                L.lop.dgrad(L.dz, 3 * B, dzp, sigma=sig, alpha_k=L.act_k, aux=aux, aux_mode=P.act_code,
                            aux_wrap=(2 * B * L.rows_in, B * L.rows_in), colsum=P.cs if P.has_bias else None,
                            colsum_rows=2 * B * L.rows_in, out_mode=self.om)
                if P.has_bias:
                    pending.append(lambda P=P, L=L: self._bias_grad_from_colsum(D, P, L.Cs_in))
            else:
                gl = G.layers[-1]
                L.lop.dgrad(L.dz[:, 2 * B * L.rows_out:, :], B, gl.dz, sigma=sig, alpha_k=L.act_k,
                            aux=self.x_all[:, B * HW:, :], aux_mode=gl.act_code, colsum=gl.cs if gl.has_bias else None,
                            out_mode=self.om)
                if gl.has_bias:
                    pending.append(lambda gl=gl, L=L: self._bias_grad_from_colsum(G, gl, L.Cs_in))
        if part != 'gen':
            pending.append(lambda: self._reduce_net(D, 2 * B))      # every D layer's split-K partials -> gradients, one launch
        # The discriminator's gradients are complete once its finalisation work has drained; its Adam update and operand
        # refresh then run on the update stream WHILE the generator's backward pass proceeds (nothing below reads a
        # discriminator weight).  Multi-GPU: the update has to wait for the gradient all-reduce, so it stays in _phase_update.
        self._dis_updated = False
        if part == 'dis':
            flush()
            if side is not main:
                ev = torch.cuda.Event()
                ev.record(side)
                main.wait_event(ev)
            return
        if side is not main and self.world_size == 1 and self.update_mask[0] and part == 'all':
            flush()
            ev_side, ev_main = torch.cuda.Event(), torch.cuda.Event()
            ev_side.record(side)
            ev_main.record(main)       # the last kernel that reads a packed / canonical D weight (D's first layer input gradient)
            self._upd_stream.wait_event(ev_side)
            self._upd_stream.wait_event(ev_main)
            with torch.cuda.stream(self._upd_stream):
                K.incr_step(D.step)
                K.adam(D.w, D.m, D.v, D.g, D.n_flat, self.lr_dis, D.step)
                D.refresh()
            self._dis_updated = True
        # ================= generator: loss_gen -> G variables
        for i in range(len(G.layers) - 1, -1, -1):
            L = G.layers[i]
            x_in = G.layers[i - 1].a if i > 0 else self.code_planes
            pending.append(lambda L=L, x_in=x_in: self._weight_grad(G, L, x_in, L.dz, B))
            flush()
            if i == 0:
                break
            P = G.layers[i - 1]
            if P.has_bn:
                da = self._as_rows(P.da_raw, B * L.rows_in, L.Cs_in)
                L.lop.dgrad(L.dz, B, da, out_mode=2)
                rows = B * P.rows_out
                K.bn_bwd_reduce(P.da_raw, P.zraw, P.mean, P.invstd, P.gamma_int, P.beta_int, P.Cs_out, rows, P.rpb, P.act_code,
                                P.bp1, P.bp2)
                K.reduce_tiles(P.bp1, P.nblk, P.Cs_out, P.dbeta_int)
                K.reduce_tiles(P.bp2, P.nblk, P.Cs_out, P.dgamma_int)
                db, dg = P.dbeta_int, P.dgamma_int
                if self.sync_bn:
                    # dx needs the GLOBAL means of dy and dy * xhat: global sums / (world * rows) = (global sums / world) / rows;
                    # the parameter gradients keep the local sums (the gradient reduction adds the ranks up)
                    c = P.Cs_out
                    slot = self.sym_scores.stat_slot(P.bn_slot + 1)
                    slot[:c].copy_(P.dbeta_int)
                    slot[c:2 * c].copy_(P.dgamma_int)
                    tot = self.sym_scores.allreduce_stats(P.bn_slot + 1, 2 * c)
                    tot.mul_(1.0 / self.world_size)
                    db, dg = tot[:c], tot[c:2 * c]
                K.bn_bwd_apply(P.da_raw, P.zraw, P.mean, P.invstd, P.gamma_int, P.beta_int, db, dg, P.Cs_out,
                               rows, P.act_code, P.dz)

                def bn_params(P=P):
                    c, hw = G._feat_perm(P)
                    K.permute_features(P.dbeta_int, G.view(G.g, P.ly.bn_name('beta')), P.Cout, c, hw, inverse=True)
                    K.permute_features(P.dgamma_int, G.view(G.g, P.ly.bn_name('gamma')), P.Cout, c, hw, inverse=True)
                pending.append(bn_params)
            else:
                dzp = self._as_rows(P.dz, B * L.rows_in, L.Cs_in)
                aux = self._as_rows(P.a, B * L.rows_in, L.Cs_in) if P.act_code != 0 else None
                fused_cs = P.has_bias and P.op != 'd'
                L.lop.dgrad(L.dz, B, dzp, aux=aux, aux_mode=P.act_code, colsum=P.cs if fused_cs else None, out_mode=self.om)
                if fused_cs:
                    pending.append(lambda P=P, L=L: self._bias_grad_from_colsum(G, P, L.Cs_in))
                elif P.has_bias:
                    # a dense layer's bias is per FEATURE: sum its [B, F] gradient over the batch only
                    def dense_bias(P=P):
                        K.colsum_planes(P.dz, B, P.Cs_out, self.tmp_vec)
                        c, hw = G._feat_perm(P)
                        K.permute_features(self.tmp_vec, G.view(G.g, P.ly.bias_name), P.Cout, c, hw, inverse=True)
                    pending.append(dense_bias)
        pending.append(lambda: self._reduce_net(G, B))
        flush()
        if side is not main:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

    def _phase_update_dis(self):
        """Data-parallel step only: the discriminator's optimiser + operand refresh as a graph of its own, replayed on the update
        stream as soon as its gradients' all-reduce has completed -- concurrently with the generator's backward pass."""
        if not self.update_mask[0]:
            return
        K.incr_step(self.D.step)
        K.adam(self.D.w, self.D.m, self.D.v, self.D.g, self.D.n_flat, self.lr_dis, self.D.step)
        self.D.refresh()

    def _phase_update(self, skip_dis=False):
        """Both Adam updates from the same forward pass, then UPDATE_OPS (my_sngan.py:424-426; graph_func.py:848-854)."""
        main = torch.cuda.current_stream(self.device)
        side = self._upd_stream if self.grad_fork else main
        if side is not main:                      # the two optimisers touch disjoint buffers: run them concurrently
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        for net, lr, st, scheduled in ((self.D, self.lr_dis, side, self.update_mask[0] and not skip_dis), (self.G, self.lr_gen, main, self.update_mask[1])):
            if not scheduled:                              # imbalanced update: this optimiser is not run on this step (graph_func.py:885-886)
                continue
            if net is self.D and self._dis_updated:        # already enqueued on the update stream during the backward pass
                continue
            if self.nvls:
                st = main        # cross-rank barriers: one stream, one order on every rank (no inversion through shared hardware queues)
            with torch.cuda.stream(st):
                K.incr_step(net.step)
                if self.nvls:
                    # every rank's gradients are complete -> switch-reduced gradients, Adam on this rank's shard, new w / m / v
                    # multicast to all replicas -> every replica complete before the packed operands are rebuilt from w
                    f = net.flat
                    f.barrier()
                    K.adam_allreduce_nvls(net.w, net.m, net.v, f.g_mc, f.w_mc, f.m_mc, f.v_mc, f.begin, f.end, lr, net.step)
                    f.barrier()
                else:
                    K.adam(net.w, net.m, net.v, net.g, net.n_flat, lr, net.step)
                net.refresh()
        self._dis_updated = False
        with torch.cuda.stream(side):
            for L in self.D.layers + self.G.layers:
                if L.has_sn:
                    L.sn_x.copy_(L.sn_xnew)
        if side is not main:
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)
        K.nan_flag(self.mmd.losses, 2, self.nan_flag)

    # -------------------------------------------------------------------------------------------- collectives
    def _gather_scores(self):
        from . import parallel
        if self.nvls:            # one multicast kernel between two device-side barriers instead of all-gather + two copies
            self.sym_scores.scatter(self.D.layers[-1].a[0].contiguous())
            return
        parallel.gather_scores(self.D.layers[-1].a[0], self.B, self.s_gather, self.gen_all, self.real_all, self.pg)

    def _allreduce_dis_async(self):
        """The discriminator's gradients are final after the first backward graph: their all-reduce is started now, on NCCL's
        own stream, and overlaps the generator's backward pass (joined in _allreduce_grads)."""
        import torch.distributed as dist
        if not self.nvls:
            self._comm_work = dist.all_reduce(self.D.g, op=dist.ReduceOp.SUM, group=self.pg_overlap, async_op=True)

    def _join_dis_allreduce(self):
        """Make the CURRENT stream wait for the discriminator's all-reduce (issued by _allreduce_dis_async)."""
        if self._comm_work is not None:
            self._comm_work.wait()
            self._comm_work = None

    def _allreduce_grads(self):
        from . import parallel
        if self.nvls:    # the parameter gradients are reduced inside the optimiser kernel, the six kernel sums by a multicast load
            self.sym_scores.allreduce_sums(self.mmd.sums)
        else:
            # ONE call for the generator's gradients and the six kernel sums (the tail of the same buffer)
            parallel.allreduce_sum([self.G.g_all], self.pg)
            self._join_dis_allreduce()
        K.losses_from_sums(self.mmd.sums, [float(c) for c in self.mmd.desc.cD], self.mmd.losses)

    # -------------------------------------------------------------------------------------------- public API
    def _run_phases(self):
        self._phase_forward()
        if self.world_size > 1:
            self._gather_scores()
        self._phase_loss()
        if self.world_size > 1:
            self._phase_backward('dis')
            self._allreduce_dis_async()
            self._phase_backward('gen')
            self._allreduce_grads()
        else:
            self._phase_backward()
        self._phase_update()

    def _run_phases_overlapped(self):
        """The data-parallel step as it is captured into ONE graph: like _run_phases, with the discriminator's optimiser forked
        onto the update stream behind its all-reduce (concurrent with the generator's backward pass)."""
        main = torch.cuda.current_stream(self.device)
        self._phase_forward()
        self._gather_scores()
        self._phase_loss()
        self._phase_backward('dis')
        self._allreduce_dis_async()
        ev = torch.cuda.Event()
        ev.record(main)
        self._upd_stream.wait_event(ev)                # fork (capture: the update stream joins the graph here)
        with torch.cuda.stream(self._upd_stream):
            self._join_dis_allreduce()
            self._phase_update_dis()
        self._phase_backward('gen')
        self._allreduce_grads()
        main.wait_stream(self._upd_stream)
        self._phase_update(skip_dis=True)

    def _replay_data_parallel(self, graphs):
        """The captured data-parallel step: four collective-free graphs around the collectives; the discriminator's optimiser is a
        fifth graph replayed on the update stream behind its all-reduce, while the generator's backward pass runs."""
        main = torch.cuda.current_stream(self.device)
        if len(graphs) == 1:                       # the whole step, collectives included, is one graph
            graphs[0].replay()
            return
        graphs[0].replay()                         # forward
        self._gather_scores()
        graphs[1].replay()                         # loss + discriminator backward
        self._allreduce_dis_async()
        if self.nvls:
            graphs[2].replay()                     # generator backward
            self._allreduce_grads()
            self._phase_update()                   # holds cross-rank barriers: enqueued eagerly
            return
        with torch.cuda.stream(self._upd_stream):
            self._join_dis_allreduce()             # the update stream (not the compute stream) waits for the all-reduce
            graphs[4].replay()                     # discriminator optimiser + operand refresh
        graphs[2].replay()                         # generator backward
        self._allreduce_grads()                    # generator gradients + the six kernel sums
        main.wait_stream(self._upd_stream)
        graphs[3].replay()                         # generator optimiser, UPDATE_OPS

    def _capture(self):
        """Capture the step (or, multi-GPU, its three collective-free segments) into CUDA graphs."""
        torch.cuda.synchronize(self.device)
        segs = ([[self._phase_forward], [self._phase_loss, lambda: self._phase_backward('dis')], [lambda: self._phase_backward('gen')],
                 [lambda: self._phase_update(skip_dis=True)], [self._phase_update_dis]]
                if self.world_size > 1 else [[self._phase_forward, self._phase_loss, self._phase_backward, self._phase_update]])
        if self.nvls:
            segs = segs[:3]      # the update phase holds cross-rank barriers: it is enqueued eagerly after the third graph
        self._dp_one_graph = False
        if self.world_size > 1 and not self.nvls and os.environ.get('MMDGAN_DP_ONE_GRAPH', '0') == '1':
            # EXPERIMENT, off by default: the whole data-parallel step INCLUDING its NCCL collectives as one CUDA graph (no host
            # launch gaps at the graph boundaries; the all-reduce of the discriminator's gradients forked onto NCCL's stream inside
            # the graph).  Measured on 2 B200: 4.35 instead of 4.42 ms per step, but the processes then hung at teardown /
            # when eager collectives followed (two 420 s timeouts) -- not shipped.  Falls back if the capture is refused.
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    self._run_phases_overlapped()
                self._graph_cache[self._graph_key()] = [g]
                self._dp_one_graph = True
                return
            except Exception as exc:      # noqa: BLE001 -- any capture failure: use the segmented form
                self._dp_one_graph_error = '{}: {}'.format(type(exc).__name__, exc)
                torch.cuda.synchronize(self.device)
        if self.nvls and os.environ.get('MMDGAN_NVLS_ONE_GRAPH', '0') == '1':
            # EXPERIMENT: the NVSwitch-multicast step has no NCCL call at all -- its exchanges are kernels with device-side barriers
            # on the symmetric allocation's signal pads -- so the whole data-parallel step can be ONE graph.
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    self._run_phases()
                self._graph_cache[self._graph_key()] = [g]
                self._dp_one_graph = True
                return
            except Exception as exc:      # noqa: BLE001
                self._dp_one_graph_error = '{}: {}'.format(type(exc).__name__, exc)
                torch.cuda.synchronize(self.device)
        graphs = []
        for fns in segs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                for fn in fns:
                    fn()
            graphs.append(g)
        self._graph_cache[self._graph_key()] = graphs

    def _graph_key(self):
        return self.update_mask + (self.device_codes,)

    @property
    def _graphs(self):
        """CUDA graphs of the ordinary step (both optimisers run); None until captured."""
        return self._graph_cache.get((True, True, False)) or self._graph_cache.get((True, True, True))

    def step_device(self, update=(True, True)):
        """One training step on the batch already staged in self._dev_data / self._dev_code (device resident).
        update = (run_dis, run_gen): which optimisers this step applies (Agent(imbalanced_update=...), graph_func.py:876-908);
        losses, both gradient sets, UPDATE_OPS and the global step are the same either way.  One CUDA graph per mask."""
        self.update_mask = (bool(update[0]), bool(update[1]))
        if not self.use_graph or not self._warm:
            # the first step always runs eagerly (it also sets the kernels' shared-memory attributes)
            n0 = K.LAUNCHES[0]
            self._run_phases()
            self.kernel_launches_per_step = K.LAUNCHES[0] - n0
            self._warm = True
        else:
            if self._graph_key() not in self._graph_cache:
                self._capture()
            graphs = self._graph_cache[self._graph_key()]
            if self.world_size > 1:
                self._replay_data_parallel(graphs)
            else:
                graphs[0].replay()
        self.global_step += 1

    def stage(self, data_x, code_x=None):
        """Put one batch on the device ({'x': NCHW float32 in [-1, 1]} contract of input_func.py:837-868).  code_x = None: the
        step draws its codes on the device (SNGan.sample_codes with code_x=None, my_sngan.py:122-124)."""
        self._dev_data.copy_(data_x, non_blocking=True)
        self.device_codes = code_x is None
        if code_x is not None:
            self._dev_code.copy_(code_x, non_blocking=True)

    def prefetch(self, data_x, code_x):
        """Start the host -> device copy of a FUTURE batch on the copy stream (it overlaps the step that is running); the
        step() call that is later given the same two tensors picks the staged copy up instead of copying again."""
        k = self._stage_next
        self._stage_next ^= 1
        dd, dc = self._stage_bufs[k]
        if not (data_x.is_pinned() and (code_x is None or code_x.is_pinned())):      # pageable host memory: through this slot's pinned buffers
            pd, pc = self._pin_bufs[k]
            pd.copy_(data_x)
            if code_x is not None:
                pc.copy_(code_x)
            src_d, src_c = pd, (pc if code_x is not None else None)
        else:
            src_d, src_c = data_x, code_x
        ev = torch.cuda.Event()
        # the staging slot was last read by the device-to-device copy of two steps ago: wait for THAT copy only (waiting for the
        # compute stream would put the transfer behind the step it is meant to overlap)
        if self._slot_free[k] is not None:
            self._copy_stream.wait_event(self._slot_free[k])
        with torch.cuda.stream(self._copy_stream):
            dd.copy_(src_d, non_blocking=True)
            if src_c is not None:
                dc.copy_(src_c, non_blocking=True)
            ev.record(self._copy_stream)
        self._prefetched = (data_x, code_x, k, ev)

    def step_async(self, data_x, code_x=None, update=(True, True), prefetch=None):
        """Enqueue one end-to-end step from HOST tensors -- H2D of the batch (picked up from an earlier prefetch() of the same
        tensors, else started now), the fused step, D2H of [loss_gen, loss_dis] and of the saturation flag -- WITHOUT waiting for
        it; result() returns the losses.  A training loop that enqueues step i + 1 before it reads the losses of step i keeps
        the device busy across the host's per-step work (Agent.train does; the reference's per-step NaN assert then fires one step
        later).  code_x = None: the codes are drawn on the device inside the step (the reference's in-graph tf.random_normal).
        prefetch = (next_data, next_code): the copy of the NEXT batch is started behind this step's launch and overlaps it."""
        pf = self._prefetched
        if not (pf is not None and pf[0] is data_x and pf[1] is code_x):
            self.prefetch(data_x, code_x)
            pf = self._prefetched
        self._prefetched = None
        main = torch.cuda.current_stream(self.device)
        dd, dc = self._stage_bufs[pf[2]]
        main.wait_event(pf[3])
        self._dev_data.copy_(dd, non_blocking=True)          # device-to-device, ~2 us
        self.device_codes = code_x is None
        if code_x is not None:
            self._dev_code.copy_(dc, non_blocking=True)
        self._slot_free[pf[2]] = torch.cuda.Event()
        self._slot_free[pf[2]].record(main)
        self.step_device(update)
        if prefetch is not None:
            self.prefetch(prefetch[0], prefetch[1])
        k = self._res_next
        self._res_next ^= 1
        self._pin_res[k].copy_(self.mmd.losses, non_blocking=True)
        self._pin_flag[k].copy_(self.sat_flag, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        return (k, ev, self.global_step)

    def result(self, pending, check_nan=True):
        """[loss_gen, loss_dis] of a step enqueued by step_async (waits for it)."""
        k, ev, gs = pending
        ev.synchronize()
        lg, ld = float(self._pin_res[k][0]), float(self._pin_res[k][1])
        if int(self._pin_flag[k][0]) != 0:
            raise FloatingPointError('an activation exceeded the range of the fp16 forward planes (|x| >= 4094) at step {}: the parity '
                                     'mode is not valid for this model state; run with MMDGAN_F16_FORWARD=0 (three bf16 planes)'.format(gs))
        if check_nan:
            assert not (math.isnan(lg) or math.isnan(ld)), \
                'Model diverged with loss = {} at step {}'.format([lg, ld], gs)    # graph_func.py:856
        return lg, ld

    def step(self, data_x, code_x=None, check_nan=True, update=(True, True), prefetch=None):
        """One end-to-end step from HOST tensors, synchronously: step_async + result."""
        return self.result(self.step_async(data_x, code_x, update=update, prefetch=prefetch), check_nan=check_nan)

    def losses(self):
        return self.mmd.losses.clone()

    def _check_saturation(self):
        if int(self.sat_flag.item()) != 0:
            raise FloatingPointError('an activation exceeded the range of the fp16 forward planes (|x| >= 4094); run with '
                                     'MMDGAN_F16_FORWARD=0 (three bf16 planes)')

    def generate(self, code_x, is_training=True):
        """Generated images NCHW in [-1, 1] for `code_x` [n, code_size] (mdl.Gen(code_batch, is_training), my_sngan.py:270 and
        :533).  is_training=True is the step's own forward (batch-statistics batch norm; n must be the engine's batch);
        is_training=False is the eval_sampling graph: moving-average batch norm, per-sample, any n <= batch_size.
        (With MMDGAN_SYNC_BN=1 the is_training=True form is a collective: every rank has to call it.)"""
        B, HW = self.B, self.height * self.width
        code_x = torch.as_tensor(code_x, dtype=torch.float32)
        n = code_x.shape[0]
        if code_x.dim() != 2 or code_x.shape[1] != self.code_size:
            raise ValueError('code_x must be [n, {}], got {}'.format(self.code_size, list(code_x.shape)))
        if n > B or (is_training and n != B):
            raise ValueError('{} codes for an engine of batch size {} (training-mode batch norm needs exactly the batch)'.format(n, B))
        if n < B:
            self._dev_code.zero_()
        self._dev_code[:n].copy_(code_x)
        K.nchw_to_planes(self._dev_code, self.code_planes)
        for L in self.G.layers:                 # sigma of the current weights and in_rand (in_rand itself is not advanced)
            if L.has_sn:
                self._sn_layer(L)
        self._net_forward(self.G, self.code_planes, B, is_training=is_training, update_moving=False)
        self._check_saturation()
        return K.planes_to_nchw(self.x_all[:, B * HW:, :], B, self.channels, self.height, self.width)[:n]

    def discriminate(self, x):
        """Discriminator scores [n, d] of NCHW images `x`, n <= 2 * batch_size (mdl.Dis(batch, is_training=False),
        my_sngan.py:548-551).  sigma of every spectrally-normalised layer comes from one power iteration on the stored
        in_rand, which -- as in the reference's eval graph, where UPDATE_OPS are not run -- is left unchanged."""
        B, HW = self.B, self.height * self.width
        x = torch.as_tensor(x, dtype=torch.float32).to(self.device)
        n = x.shape[0]
        if list(x.shape[1:]) != [self.channels, self.height, self.width] or n > 2 * B:
            raise ValueError('x must be [n <= {}, {}, {}, {}], got {}'.format(2 * B, self.channels, self.height, self.width, list(x.shape)))
        buf = torch.zeros((2 * B, self.channels, self.height, self.width), device=self.device)
        buf[:n].copy_(x)
        K.nchw_to_planes(buf, self.x_all)
        for L in self.D.layers:
            if L.has_sn:
                self._sn_layer(L)
        self._net_forward(self.D, self.x_all, 2 * B, is_training=False)
        self._check_saturation()
        last = self.D.layers[-1]
        return last.a[0][:n, :last.Cout].clone()
