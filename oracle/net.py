"""Oracle (test infrastructure): layer DSL, spectral norm and the SNGan step on the CPU.

Pinned against outputs of the reference's own Python run on oracle/tfshim (tests/golden/ref_*.npz; see oracle/__init__.py).  PyTorch-CPU restatement of:
  update_layer_design / Layer default / Net / Routine   layer_func.py:1189-1275, 1611-1685, 2111-2150, 2207-2494
  ParametricOperation ops d / c / tc / bias / bn        layer_func.py:566-600, 709-783, 870-966
  weight_initializer / bias_initializer                 layer_func.py:14-80
  leaky_relu (alpha 0.1) / activations                  layer_func.py:104-167
  spatial_shape_after_conv / _transpose_conv            math_func.py:172-216
  SpectralNorm (PICO, num_iter=1) and PIM mode          math_func.py:397-749, layer_func.py:785-825
  SNGan.__gpu_task__ / training (simultaneous update)   my_sngan.py:259-323, 412-426
  opt_config Adam(beta1 .5, beta2 .999, eps 1e-8)       graph_func.py:518-527
  UPDATE_OPS semantics (BN moving stats, in_rand)       graph_func.py:848-854

TF-1.8 semantics hard-coded here (third-party dependency, not vendored in the
reference: tensorflow==1.8.0 per misc_fun.py:33): SAME padding is symmetric 1
for k3/s1 and k4/s2 on even sizes; conv2d is a cross-correlation with HWIO
filters; conv2d_transpose is the input-gradient of conv2d with filter
[h, w, out, in]; tf.layers.batch_normalization defaults momentum 0.99, eps 1e-3,
the fused kernel normalises with the biased variance and feeds the Bessel
corrected one to the moving average; Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
theta -= lr_t*m/(sqrt(v)+eps).

Parameters are kept in the reference's variable names and layouts
(`dis/l1_f32/kernel/kernel` [k,k,Cin,Cout], `.../kernel/SN/in_rand` [1,C,H,W], ...).
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import mmd as omm

EPSI = 1e-10          # misc_fun.py:29
BN_EPS = 1e-3         # tf.layers.batch_normalization default
BN_MOMENTUM = 0.99    # tf.layers.batch_normalization default
LRELU_ALPHA = 0.1     # layer_func.py:112


# ----------------------------------------------------------------------------------------------
# tf32 emulation (used only to size the error budget of the tensor-core path)
class _RoundTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return round_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


def round_tf32(x):
    """Round-to-nearest-even of an fp32 tensor to the 10-bit tf32 mantissa (cvt.rna.tf32.f32 ties away;
    the difference only shows on exact ties)."""
    if x.dtype != torch.float32:
        return x
    i = x.detach().contiguous().view(torch.int32)
    r = ((i + 0x00001000) & ~0x00001FFF)        # round half away from zero in magnitude (sign-magnitude repr)
    return r.view(torch.float32).reshape(x.shape)


def _maybe_tf32(x, on):
    return _RoundTF32.apply(x) if on else x


# ----------------------------------------------------------------------------------------------
def update_layer_design(layer_design):
    """layer_func.py:1189-1275."""
    template = {'name': None, 'type': 'default', 'op': 'c', 'out': None, 'bias': 'b',
                'act': 'linear', 'act_nm': None, 'act_k': False,
                'w_nm': None, 'w_p': None,
                'kernel': 3, 'strides': 1, 'dilation': 1, 'padding': 'SAME', 'scale': None,
                'in_reshape': None, 'out_reshape': None, 'aux': None}
    for key in layer_design:
        template[key] = layer_design[key]
    if template['act_nm'] in {'bn', 'BN'} and template['bias'] in {'b', 'bias'}:
        template['bias'] = None                                          # layer_func.py:1241-1242
    if template['op'] in {'tc'}:
        template['scale'] = None
    if template['op'] not in {'d', 'c', 'tc'}:
        raise AttributeError('layer op {} not supported.'.format(template['op']))
    return template


def spatial_shape_after_conv(size, kernel, strides, dilation, padding):
    """math_func.py:172-191."""
    if padding in ['same', 'SAME']:
        return int(np.ceil(size / strides))
    return int(np.ceil((size - (kernel - 1) * dilation) / strides))


def spatial_shape_after_transpose_conv(size, kernel, strides, dilation, padding):
    """math_func.py:194-216."""
    if padding in ['same', 'SAME']:
        return int(size * strides)
    return int(size * strides + (kernel - 1) * dilation)


def _fans(shape):
    """TF variance_scaling fan computation (_compute_fans)."""
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def _trunc_normal(gen, shape, std):
    t = torch.empty(shape, dtype=torch.float64)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return t * std


def weight_initializer(gen, shape, act_fun):
    """layer_func.py:14-66, FLAGS.WEIGHT_INITIALIZER == 'default' (TF-1.8 variance_scaling_initializer)."""
    fan_in, fan_out = _fans(shape)
    if act_fun == 'relu':
        return _trunc_normal(gen, shape, math.sqrt(2.0 / fan_in))
    if act_fun == 'lrelu':
        return _trunc_normal(gen, shape, math.sqrt(2.0 / 1.01 / fan_in))
    scale = 16.0 if act_fun == 'sigmoid' else 1.0
    limit = math.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
    return (torch.rand(shape, dtype=torch.float64, generator=gen) * 2.0 - 1.0) * limit


class LayerSpec(object):
    """Static description of one default layer after shape inference (layer_func.py:2043-2076)."""

    def __init__(self, design, in_shape, scope):
        self.design = design
        self.scope = scope                      # e.g. 'dis/l2_ds'
        self.op = design['op']
        self.in_shape = list(in_shape)          # without batch, after in_reshape
        if design['in_reshape'] is not None:
            self.in_shape = list(design['in_reshape'])
        if self.op == 'd':
            assert len(self.in_shape) == 1, '{}: dense layer needs a flat input'.format(scope)
            self.kernel_shape = [self.in_shape[0], design['out']]           # layer_func.py:576-578
            self.op_out_shape = [design['out']]
        elif self.op == 'c':
            c, h, w = self.in_shape
            k, s = design['kernel'], design['strides']
            self.kernel_shape = [k, k, c, design['out']]                     # layer_func.py:584
            self.op_out_shape = [design['out'],
                                 spatial_shape_after_conv(h, k, s, design['dilation'], design['padding']),
                                 spatial_shape_after_conv(w, k, s, design['dilation'], design['padding'])]
        else:
            c, h, w = self.in_shape
            k, s = design['kernel'], design['strides']
            self.kernel_shape = [k, k, design['out'], c]                     # layer_func.py:595
            self.op_out_shape = [design['out'],
                                 spatial_shape_after_transpose_conv(h, k, s, design['dilation'], design['padding']),
                                 spatial_shape_after_transpose_conv(w, k, s, design['dilation'], design['padding'])]
        self.out_shape = list(design['out_reshape']) if design['out_reshape'] is not None else self.op_out_shape
        self.has_bias = design.get('bias') is not None
        self.has_bn = design['act_nm'] in {'bn', 'BN'}
        self.has_sn = design.get('w_nm') in {'s'}
        self.act = design['act']
        self.act_k = design['act_k']
        # spectral-norm routing (math_func.py:470-528); integer compare, must be bit-exact
        if self.has_sn:
            if self.op == 'd':
                num_in, num_out = self.kernel_shape
                self.use_u = num_in <= num_out
                self.x_shape = [1, num_in] if self.use_u else [1, num_out]
            elif self.op == 'c':
                self.use_u = int(np.prod(self.in_shape)) <= int(np.prod(self.op_out_shape))
                self.x_shape = [1] + (self.in_shape if self.use_u else self.op_out_shape)
            else:
                self.use_u = int(np.prod(self.in_shape)) <= int(np.prod(self.op_out_shape))
                self.x_shape = [1] + (self.op_out_shape if self.use_u else self.in_shape)

    # variable names as the reference's checkpoint would hold them
    @property
    def kernel_name(self):
        return self.scope + '/kernel/kernel'

    @property
    def bias_name(self):
        return self.scope + '/bias/bias'

    @property
    def sn_name(self):
        return self.scope + '/kernel/SN/in_rand'

    def bn_name(self, what):
        return self.scope + '/BN/BN/' + what


def build_net(net_design, net_name, in_shape):
    """Net.__init__ + Routine.seq_links shape inference (layer_func.py:2118-2150, 2349-2376)."""
    specs = []
    shape = list(in_shape)
    for d in net_design:
        design = update_layer_design(d)
        spec = LayerSpec(design, shape, net_name + '/' + design['name'])
        specs.append(spec)
        shape = spec.out_shape
    return specs


def init_params(specs, gen, dtype=torch.float32):
    """Variables of one net with the reference initialisers.

    kernel: weight_initializer(act) (layer_func.py:719-721); bias: truncated normal sd 1e-5 (:745-747);
    BN gamma 1 / beta 0 / moving_mean 0 / moving_variance 1; in_rand: truncated normal sd 1, NOT
    normalised (math_func.py:565-567).
    """
    params, state = OrderedDict(), OrderedDict()
    for sp in specs:
        params[sp.kernel_name] = weight_initializer(gen, sp.kernel_shape, sp.act).to(dtype)
        if sp.has_bias:
            params[sp.bias_name] = _trunc_normal(gen, [sp.op_out_shape[0]], 1e-5).to(dtype)
        if sp.has_bn:
            c = sp.op_out_shape[0]
            params[sp.bn_name('gamma')] = torch.ones(c, dtype=dtype)
            params[sp.bn_name('beta')] = torch.zeros(c, dtype=dtype)
            state[sp.bn_name('moving_mean')] = torch.zeros(c, dtype=dtype)
            state[sp.bn_name('moving_variance')] = torch.ones(c, dtype=dtype)
        if sp.has_sn:
            state[sp.sn_name] = _trunc_normal(gen, sp.x_shape, 1.0).to(dtype)
    return params, state


# ----------------------------------------------------------------------------------------------
def _op_forward(sp, x, kernel, tf32=False):
    """ParametricOperation ops d / c / tc (layer_func.py:909-928) on NCHW tensors."""
    x = _maybe_tf32(x, tf32)
    kernel = _maybe_tf32(kernel, tf32)
    if sp.op == 'd':
        return x @ kernel
    s = sp.design['strides']
    k = sp.design['kernel']
    pad = (k - s) // 2 if s > 1 else (k - 1) // 2
    assert (k - s) % 2 == 0 or s == 1, 'asymmetric SAME padding is not on the hot path'
    if sp.op == 'c':
        return F.conv2d(x, kernel.permute(3, 2, 0, 1), stride=s, padding=pad)
    return F.conv_transpose2d(x, kernel.permute(3, 2, 0, 1), stride=s, padding=pad)


def _op_adjoint(sp, y, kernel, tf32=False):
    """Adjoint of _op_forward w.r.t. its input (SpectralNorm._dense_t_ / _conv_t_ / _conv_, math_func.py:583-637)."""
    y = _maybe_tf32(y, tf32)
    kernel = _maybe_tf32(kernel, tf32)
    if sp.op == 'd':
        return y @ kernel.t()
    s = sp.design['strides']
    k = sp.design['kernel']
    pad = (k - s) // 2 if s > 1 else (k - 1) // 2
    if sp.op == 'c':
        return F.conv_transpose2d(y, kernel.permute(3, 2, 0, 1), stride=s, padding=pad)
    return F.conv2d(y, kernel.permute(3, 2, 0, 1), stride=s, padding=pad)


def _l2n(w):
    """math_func.py:653-659."""
    return w / (torch.linalg.vector_norm(w) + EPSI)


def spectral_norm(sp, kernel, x, mode='default', tf32=False):
    """One PICO power iteration (math_func.py:661-672, 739-744).  Returns (sigma, x_update).

    sigma is differentiable w.r.t. kernel through forward(x) only; x is a non-trainable variable.
    mode 'sn_paper' (PIM, layer_func.py:811-814) treats a conv kernel as the [k*k*Cin, Cout] matrix.
    """
    if mode in {'sn_paper', 'PIM', 'pim'} and sp.op in {'c', 'tc'}:
        w2 = kernel.reshape(-1, kernel.shape[3])
        use_u = w2.shape[0] <= w2.shape[1]
        fwd = (lambda v: v @ w2) if use_u else (lambda v: v @ w2.t())
        bwd = (lambda v: v @ w2.t()) if use_u else (lambda v: v @ w2)
    else:
        # for 'c' forward is the conv when use_u else its adjoint; for 'tc' likewise with the tc op
        # (math_func.py:524-525: forward = _conv_ if use_u else _conv_t_; for a 'tc' kernel _conv_ applied
        # to an output-shaped x IS the adjoint of the layer op, hence the swap below).
        if sp.op == 'tc':
            a = lambda v: _op_adjoint(sp, v, kernel, tf32)      # tf.nn.conv2d with the tc kernel
            b = lambda v: _op_forward(sp, v, kernel, tf32)      # tf.nn.conv2d_transpose
        else:
            a = lambda v: _op_forward(sp, v, kernel, tf32)
            b = lambda v: _op_adjoint(sp, v, kernel, tf32)
        fwd, bwd = (a, b) if sp.use_u else (b, a)
    x = x.detach()
    v = fwd(x)
    y = _l2n(v)
    x_update = _l2n(bwd(y)).detach()
    sigma = torch.linalg.vector_norm(v)
    return sigma, x_update


def sn_pim_x_shape(sp):
    w_rows = int(np.prod(sp.kernel_shape[:3]))
    return [1, w_rows] if w_rows <= sp.kernel_shape[3] else [1, sp.kernel_shape[3]]


class _KinkWithMask(torch.autograd.Function):
    """relu / leaky-relu whose DERIVATIVE branch is chosen by an external boolean mask instead of sign(x).

    Test helper: for an input within rounding noise of the kink, an fp32 implementation and this float64 oracle may
    legitimately land on different sides (either is correct to the working precision).  Parity tests pass the
    sign pattern of the implementation under test so that such ties do not masquerade as errors."""

    @staticmethod
    def forward(ctx, x, mask, slope):
        ctx.save_for_backward(mask)
        ctx.slope = slope
        return torch.where(x > 0, x, x * slope)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return torch.where(mask, g, g * ctx.slope), None, None


def activation(x, act, mask=None):
    """layer_func.py:104-167."""
    if act == 'linear':
        return x
    if mask is not None and act in ('relu', 'lrelu'):
        return _KinkWithMask.apply(x, mask.reshape(x.shape), 0.0 if act == 'relu' else LRELU_ALPHA)
    if act == 'relu':
        return F.relu(x)
    if act == 'lrelu':
        return F.leaky_relu(x, LRELU_ALPHA)
    if act == 'tanh':
        return torch.tanh(x)
    raise NotImplementedError('Function {} is not implemented.'.format(act))


def net_forward(specs, params, state, x, is_training=True, sn_mode='default', tf32=False, collect=None, act_masks=None):
    """Routine.__call__ over a sequential net of default layers (layer_func.py:1646-1685, 2451-2491).

    Returns (output, updates) where updates holds the UPDATE_OPS results (BN moving stats, SN in_rand).
    `collect`, if a dict, receives per-layer sigma and activations for the layer-level parity tests.
    """
    updates = OrderedDict()
    n = x.shape[0]
    for sp in specs:
        if sp.design['in_reshape'] is not None:
            x = x.reshape([n] + list(sp.design['in_reshape']))
        kernel = params[sp.kernel_name]
        if sp.has_sn:
            sigma, x_upd = spectral_norm(sp, kernel, state[sp.sn_name], sn_mode, tf32)
            updates[sp.sn_name] = x_upd
            # layer_func.py:884-887: multiplier = act_k / sigma when act_k is a number, else 1 / sigma
            mult = (sp.act_k if isinstance(sp.act_k, (float, int)) else 1.0) / sigma
            kernel = kernel * mult
            if collect is not None:
                collect[sp.scope + '/sigma'] = sigma
        x = _op_forward(sp, x, kernel, tf32)
        if sp.has_bias:
            b = params[sp.bias_name]
            x = x + (b if x.dim() == 2 else b.view(1, -1, 1, 1))
        if sp.has_bn:
            g, b = params[sp.bn_name('gamma')], params[sp.bn_name('beta')]
            mm, mv = state[sp.bn_name('moving_mean')], state[sp.bn_name('moving_variance')]
            if is_training:
                red = [0] if x.dim() == 2 else [0, 2, 3]
                cnt = x.numel() // x.shape[1]
                mean = x.mean(red)
                var_b = x.var(red, unbiased=False)
                shp = (1, -1) if x.dim() == 2 else (1, -1, 1, 1)
                x = (x - mean.view(shp)) / torch.sqrt(var_b.view(shp) + BN_EPS) * g.view(shp) + b.view(shp)
                # TF 1.8 uses the fused kernel (Bessel-corrected moving variance) for rank-4 inputs only; rank-2 inputs fall
                # back to nn.moments and feed the biased variance into the moving average
                var_u = var_b.detach() * (cnt / max(cnt - 1.0, 1.0)) if x.dim() == 4 else var_b.detach()
                updates[sp.bn_name('moving_mean')] = mm * BN_MOMENTUM + mean.detach() * (1.0 - BN_MOMENTUM)
                updates[sp.bn_name('moving_variance')] = mv * BN_MOMENTUM + var_u * (1.0 - BN_MOMENTUM)
            else:
                shp = (1, -1) if x.dim() == 2 else (1, -1, 1, 1)
                x = (x - mm.view(shp)) / torch.sqrt(mv.view(shp) + BN_EPS) * g.view(shp) + b.view(shp)
        x = activation(x, sp.act, None if act_masks is None else act_masks.get(sp.scope))
        if sp.design['out_reshape'] is not None:
            x = x.reshape([n] + list(sp.design['out_reshape']))
        if collect is not None:
            collect[sp.scope + '/out'] = x
    return x, updates


# ----------------------------------------------------------------------------------------------
class TFAdam(object):
    """tf.train.AdamOptimizer as configured at graph_func.py:518-527 (beta1 .5, beta2 .999, eps 1e-8)."""

    def __init__(self, params, lr, beta1=0.5, beta2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.t = 0
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())

    def apply(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k in params:
            g = grads[k]
            self.m[k] = self.b1 * self.m[k] + (1.0 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1.0 - self.b2) * g * g
            params[k] = params[k] - lr_t * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)


class OracleSNGan(object):
    """SNGan restated: init_net (my_sngan.py:85-108), __gpu_task__ (259-323), training update (412-426)."""

    def __init__(self, architecture, loss_type='rep', rep_weights=(0.0, -1.0), lr_list=(5e-4, 2e-4),
                 seed=2, dtype=torch.float32, sn_mode='default', tf32=False):
        self.arch = architecture
        self.loss_type = loss_type
        self.rep_weights = list(rep_weights)
        self.dtype = dtype
        self.sn_mode = sn_mode
        self.tf32 = tf32
        self.code_size = architecture['code'][0][0]
        self.gen_specs = build_net(architecture['generator'], 'gen', [self.code_size])
        self.dis_specs = build_net(architecture['discriminator'], 'dis', list(architecture['input'][0]))
        gen = torch.Generator().manual_seed(seed)
        self.gen_params, self.gen_state = init_params(self.gen_specs, gen, dtype)
        self.dis_params, self.dis_state = init_params(self.dis_specs, gen, dtype)
        if sn_mode in {'sn_paper', 'PIM', 'pim'}:
            for sp in self.dis_specs:
                if sp.has_sn and sp.op in {'c', 'tc'}:
                    self.dis_state[sp.sn_name] = _trunc_normal(gen, sn_pim_x_shape(sp), 1.0).to(dtype)
        self.opt_dis = TFAdam(self.dis_params, lr_list[0])
        self.opt_gen = TFAdam(self.gen_params, lr_list[1])
        self.global_step = 0

    def forward_losses(self, data_x, code_x, collect=None, act_masks=None):
        gp = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in self.gen_params.items())
        dp = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in self.dis_params.items())
        b = data_x.shape[0]
        x_gen, upd_g = net_forward(self.gen_specs, gp, self.gen_state, code_x, True, self.sn_mode, self.tf32, collect, act_masks)
        both = torch.cat([data_x, x_gen], 0)                                  # my_sngan.py:244-256, 278
        s_all, upd_d = net_forward(self.dis_specs, dp, self.dis_state, both, True, self.sn_mode, self.tf32, collect, act_masks)
        s_x, s_gen = s_all[:b], s_all[b:]                                     # my_sngan.py:279
        loss_gen, loss_dis = omm.gan_loss(s_gen, s_x, self.loss_type, batch_size=b, rep_weights=self.rep_weights)
        if collect is not None:
            collect['x_gen'] = x_gen
            collect['s_x'] = s_x
            collect['s_gen'] = s_gen
        return loss_gen, loss_dis, gp, dp, upd_g, upd_d

    def grads(self, data_x, code_x, collect=None, act_masks=None):
        """Both gradient sets evaluated at the pre-update weights (my_sngan.py:301-304)."""
        loss_gen, loss_dis, gp, dp, upd_g, upd_d = self.forward_losses(data_x, code_x, collect, act_masks)
        g_dis = torch.autograd.grad(loss_dis, list(dp.values()), retain_graph=True)
        g_gen = torch.autograd.grad(loss_gen, list(gp.values()))
        grads_dis = OrderedDict(zip(dp.keys(), [g.detach() for g in g_dis]))
        grads_gen = OrderedDict(zip(gp.keys(), [g.detach() for g in g_gen]))
        return loss_gen.detach(), loss_dis.detach(), grads_gen, grads_dis, upd_g, upd_d

    def eval_sampling(self, code_x, data_x=None):
        """The eval_sampling graph (my_sngan.py:533-551): generator with is_training=False (moving-average batch norm), samples
        clipped to [-1, 1]; with real samples, the discriminator (is_training=False) on [real; clipped generated].  The
        spectral-norm power iteration still runs (it is part of the kernel's graph) but its UPDATE_OPS are not applied."""
        with torch.no_grad():
            x_gen, _ = net_forward(self.gen_specs, self.gen_params, self.gen_state, code_x, False, self.sn_mode, self.tf32)
            x_gen = x_gen.clamp(-1.0, 1.0)
            out = {'x_gen': x_gen}
            if data_x is not None:
                b = data_x.shape[0]
                s_all, _ = net_forward(self.dis_specs, self.dis_params, self.dis_state, torch.cat([data_x, x_gen], 0), False,
                                       self.sn_mode, self.tf32)
                out['s_x'], out['s_gen'] = s_all[:b], s_all[b:]
        return out

    def step(self, data_x, code_x, imbalanced_update=None):
        """One fused sess.run of [losses, dis_op, gen_op, UPDATE_OPS, global_step] (graph_func.py:851-854).
        imbalanced_update = (k_dis, k_gen), one of them 1 (my_sngan.py:427-439): optimiser i is run only on the steps whose
        global step is a multiple of k_i (graph_func.py:885-886); an optimiser that is not run keeps its slots and beta powers;
        UPDATE_OPS and the global step (carried by the optimiser that runs every step) advance on every step."""
        run_dis, run_gen = imbalanced_schedule(self.global_step, imbalanced_update)
        loss_gen, loss_dis, grads_gen, grads_dis, upd_g, upd_d = self.grads(data_x, code_x)
        if run_dis:
            self.opt_dis.apply(self.dis_params, grads_dis)
        if run_gen:
            self.opt_gen.apply(self.gen_params, grads_gen)
        for k, v in upd_g.items():
            self.gen_state[k] = v
        for k, v in upd_d.items():
            self.dis_state[k] = v
        self.global_step += 1
        assert not (math.isnan(float(loss_gen)) or math.isnan(float(loss_dis))), \
            'Model diverged with loss = {} at step {}'.format([float(loss_gen), float(loss_dis)], self.global_step)
        return float(loss_gen), float(loss_dis)


def imbalanced_schedule(global_step, imbalanced_update):
    """Which of (dis_op, gen_op) a step runs (graph_func.py:876-886; my_sngan.py:423-439)."""
    if imbalanced_update is None:
        return True, True
    assert len(imbalanced_update) == 2, 'Imbalanced_update length does not match that of op_list. Expected 2 got {}.'.format(
        len(imbalanced_update))
    if 1 not in tuple(imbalanced_update):
        raise AttributeError('One of the imbalanced_update must be 1.')
    return tuple(int(global_step) % int(k) == 0 for k in imbalanced_update)


def synthetic_batch(arch, batch_size, seed=0, dtype=torch.float32):
    """Synthetic inputs of SURVEY.md section 8(d): data uniform in [-1, 1] (input_func.py:839), codes N(0,1)."""
    c, h, w = arch['input'][0]
    g0 = torch.Generator().manual_seed(seed)
    g1 = torch.Generator().manual_seed(seed + 1)
    data = (torch.rand(batch_size, c, h, w, generator=g0, dtype=torch.float64) * 2.0 - 1.0).to(dtype)
    code = torch.randn(batch_size, arch['code'][0][0], generator=g1, dtype=torch.float64).to(dtype)
    return data, code


def warm_spectral_norm(model, n_iter=1):
    """Apply only the in_rand UPDATE_OPS n_iter times (no weight update).

    The reference's first step runs with an un-normalised in_rand (math_func.py:565-567), which makes
    sigma ~ ||x0|| times too large and the step degenerate; parity tests use this to reach the regime
    every later step is in.  Test helper, not a reference function.
    """
    for _ in range(n_iter):
        for sp in model.dis_specs + model.gen_specs:
            if sp.has_sn:
                params = model.dis_params if sp.scope.startswith('dis/') else model.gen_params
                state = model.dis_state if sp.scope.startswith('dis/') else model.gen_state
                _, x_upd = spectral_norm(sp, params[sp.kernel_name].detach(), state[sp.sn_name], model.sn_mode)
                state[sp.sn_name] = x_upd
