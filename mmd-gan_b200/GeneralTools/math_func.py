"""Losses and spectral normalisation of the reference, rebuilt over the fused B200 kernels.

Mirrors the public surface of GeneralTools/math_func.py for the hot path:
  spatial_shape_after_conv / _transpose_conv   math_func.py:172-216
  get_squared_dist(x, y, mode=...)             math_func.py:767-858   (diagnostic: materialises the B x B matrices)
  mmd_g / mmd_g_bounded / mixture_mmd_g        math_func.py:1288-1473 (from distance matrices, as in the reference)
  GANLoss(do_summary).apply(score_gen, score_data, loss_type, batch_size=, d=, rep_weights=) -> (loss_gen, loss_dis)
                                               math_func.py:2088-2116, 2505-2550, 2556-2658
  SpectralNorm(sn_def, name_scope, scope_prefix, num_iter).apply(kernel) -> sigma      math_func.py:397-749

GANLoss runs the single fused kernel mmdgan_mmd_fwd_bwd (pairwise distances, kernels, losses AND the score gradients
in one launch) and returns autograd-connected scalars, so `loss.backward()` feeds the kernel's gradients to whatever
produced the scores.  The matrix-level functions (get_squared_dist, mmd_g, ...) are kept for API parity and
diagnostics; they are thin torch expressions on the device and are NOT used by the training step.
"""
import numpy as np
import torch

from .misc_fun import FLAGS


def spatial_shape_after_conv(input_spatial_shape, kernel_size, strides, dilation, padding):
    """math_func.py:172-191."""
    if isinstance(input_spatial_shape, (list, tuple)):
        return [spatial_shape_after_conv(s, kernel_size, strides, dilation, padding) for s in input_spatial_shape]
    if padding in ['same', 'SAME']:
        return int(np.ceil(input_spatial_shape / strides))
    return int(np.ceil((input_spatial_shape - (kernel_size - 1) * dilation) / strides))


def spatial_shape_after_transpose_conv(input_spatial_shape, kernel_size, strides, dilation, padding):
    """math_func.py:194-216."""
    if isinstance(input_spatial_shape, (list, tuple)):
        return [spatial_shape_after_transpose_conv(s, kernel_size, strides, dilation, padding) for s in input_spatial_shape]
    if padding in ['same', 'SAME']:
        return int(input_spatial_shape * strides)
    return int(input_spatial_shape * strides + (kernel_size - 1) * dilation)


# ------------------------------------------------------------------------------------------------ matrix-level API
class MeshCode(object):
    """Latent codes laid out on a (rows, columns) mesh for sample sprites (math_func.py:220-335).  Host-side: the codes are a
    [rows * columns, code_length] float32 tensor handed to SNGanEngine.generate."""

    def __init__(self, code_length, mesh_num=None):
        self.D = int(code_length)
        self.mesh_num = (10, 10) if mesh_num is None else tuple(int(m) for m in mesh_num)

    def get_batch(self, mesh_mode, name=None):
        if mesh_mode == 0 or mesh_mode == 'random':
            return self.by_random()
        if mesh_mode == 1 or mesh_mode == 'sine':
            return self.by_sine()
        if mesh_mode == 2 or mesh_mode == 'feature':
            return self.by_feature()
        raise AttributeError('mesh_mode is not supported.')

    def by_random(self, name=None):
        return torch.randn(self.mesh_num[0] * self.mesh_num[1], self.D)

    def by_sine(self, z_support=None, name=None):
        """Great-arc interpolation between four supporting codes (math_func.py:257-291): with psi indexed by the mesh's second
        axis (slow) and phi by its first (fast), both over [0, pi/4],
        z = (cos(psi) z0 + sin(psi) z1) cos(phi) + (cos(psi) z2 + sin(psi) z3) sin(phi)."""
        z = torch.randn(4, self.D) if z_support is None else torch.as_tensor(np.asarray(z_support), dtype=torch.float32)
        assert tuple(z.shape) == (4, self.D), 'z_support must be [4, {}]'.format(self.D)
        quarter = np.float32(np.pi / 4.0)
        phi = torch.from_numpy(np.float32(quarter * np.linspace(0.0, 1.0, self.mesh_num[0])))
        psi = torch.from_numpy(np.float32(quarter * np.linspace(0.0, 1.0, self.mesh_num[1])))
        near = torch.cos(psi)[:, None] * z[0] + torch.sin(psi)[:, None] * z[1]           # [m1, D]
        far = torch.cos(psi)[:, None] * z[2] + torch.sin(psi)[:, None] * z[3]
        mesh = near[:, None, :] * torch.cos(phi)[None, :, None] + far[:, None, :] * torch.sin(phi)[None, :, None]
        return mesh.reshape(self.mesh_num[0] * self.mesh_num[1], self.D)

    def by_feature(self, grid=2.0, name=None):
        """One code coordinate varies over [-grid, grid] per mesh row, all others zero; which coordinates is a random draw
        (math_func.py:293-316: the kron(eye, mesh) pattern with its columns shuffled)."""
        m0, m1 = self.mesh_num
        assert m0 <= self.D, 'cannot mesh {} features of a {}-dimensional code'.format(m0, self.D)
        steps = torch.from_numpy(np.float32(np.linspace(-grid, grid, m1)))
        cols = torch.randperm(self.D)[:m0]
        z = torch.zeros(m0, m1, self.D)
        z[torch.arange(m0), :, cols] = steps
        return z.reshape(m0 * m1, self.D)

    def simple_grid(self, grid=None):
        """Regular grid over a two-dimensional code, first coordinate slow (math_func.py:318-335); numpy."""
        if self.D != 2:
            raise AttributeError('Code length has to be two')
        if grid is None:
            grid = np.array([[-1.0, 1.0], [-1.0, 1.0]], dtype=np.float32)
        x = np.linspace(grid[0][0], grid[0][1], self.mesh_num[0])
        y = np.linspace(grid[1][0], grid[1][1], self.mesh_num[1])
        xx, yy = np.meshgrid(x, y, indexing='ij')
        return np.stack([xx.reshape(-1), yy.reshape(-1)], axis=1)


def get_squared_dist(x, y=None, scale=None, z_score=False, mode='xxxyyy', name='squared_dist', do_summary=False,
                     scope_prefix=''):
    """Pairwise squared distances with the Gram trick and the clamp at zero (math_func.py:767-858)."""
    if x.dim() > 2:
        raise AttributeError('get_dist: Input must be a matrix.')
    if y is None:
        mode = 'xx'
    if z_score:
        mu = x.mean(0, keepdim=True) if y is None else torch.cat((x, y), 0).mean(0, keepdim=True)
        x = x - mu
        y = None if y is None else y - mu
    xs = x if scale is None else x * scale
    if mode in ['xx', 'xxxy', 'xxxyyy']:
        xxt = xs @ x.t()
        dx = torch.diagonal(xxt)
        dist_xx = torch.clamp(dx[:, None] - 2.0 * xxt + dx[None, :], min=0.0)
        if mode == 'xx':
            return dist_xx
        ys = y if scale is None else y * scale
        xyt = xs @ y.t()
        if mode == 'xxxy':
            dy = (ys * y).sum(1)
            return dist_xx, torch.clamp(dx[:, None] - 2.0 * xyt + dy[None, :], min=0.0)
        yyt = ys @ y.t()
        dy = torch.diagonal(yyt)
        dist_xy = torch.clamp(dx[:, None] - 2.0 * xyt + dy[None, :], min=0.0)
        dist_yy = torch.clamp(dy[:, None] - 2.0 * yyt + dy[None, :], min=0.0)
        return dist_xx, dist_xy, dist_yy
    if mode == 'xy':
        ys = y if scale is None else y * scale
        dx = (xs * x).sum(1)
        dy = (ys * y).sum(1)
        return torch.clamp(dx[:, None] - 2.0 * (xs @ y.t()) + dy[None, :], min=0.0)
    raise AttributeError('Mode {} not supported'.format(mode))


def matrix_mean_wo_diagonal(matrix, num_row, num_col=None, name='mu_wo_diag'):
    """math_func.py:1048-1069."""
    if num_col is None:
        return (matrix.sum() - torch.diagonal(matrix).sum()) / (num_row * (num_row - 1.0))
    return (matrix.sum() - torch.diagonal(matrix).sum()) / (num_row * num_col - min(num_col, num_row))


def mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma=1.0, var_target=None, upper_bound=None, lower_bound=None,
          name='mmd', do_summary=False, scope_prefix='', custom_weights=None):
    """math_func.py:1288-1352."""
    if var_target is not None:
        raise NotImplementedError('mmd_g: trainable sigma (var_target) is not on the hot path')
    s2 = 2.0 * sigma ** 2
    lb = (lambda d: d) if lower_bound is None else (lambda d: torch.clamp(d, min=lower_bound))
    ub = (lambda d: d) if upper_bound is None else (lambda d: torch.clamp(d, max=upper_bound))
    m = float(batch_size)
    e_kxx = matrix_mean_wo_diagonal(torch.exp(-lb(dist_xx) / s2), m)
    e_kxy = matrix_mean_wo_diagonal(torch.exp(-ub(dist_xy) / s2), m)
    e_kyy = matrix_mean_wo_diagonal(torch.exp(-lb(dist_yy) / s2), m)
    if custom_weights is None:
        return e_kxx + e_kyy - 2.0 * e_kxy
    assert custom_weights[0] - custom_weights[1] == 1.0, 'w[0]-w[1] must be 1'
    return e_kxx + e_kyy - 2.0 * e_kxy, custom_weights[0] * e_kxy - e_kxx - custom_weights[1] * e_kyy


def mmd_g_bounded(dist_xx, dist_xy, dist_yy, batch_size, sigma=1.0, var_target=None, upper_bound=None, lower_bound=None,
                  name='mmd', do_summary=False, scope_prefix='', custom_weights=None):
    """math_func.py:1356-1431 (the 1387-1390 / 1402 sign quirk is reproduced: e_kxy_b is never the bounded kernel)."""
    if var_target is not None:
        raise NotImplementedError('mmd_g_bounded: trainable sigma (var_target) is not on the hot path')
    s2 = 2.0 * sigma ** 2
    m = float(batch_size)
    k_xy = torch.exp(-dist_xy / s2)
    e_kxx = matrix_mean_wo_diagonal(torch.exp(-dist_xx / s2), m)
    e_kxy = matrix_mean_wo_diagonal(k_xy, m)
    e_kyy = matrix_mean_wo_diagonal(torch.exp(-dist_yy / s2), m)
    if custom_weights is None:
        return e_kxx + e_kyy - 2.0 * e_kxy
    k_xy_b = torch.exp(-torch.clamp(dist_xy, max=upper_bound) / s2) if custom_weights[0] > 0 else k_xy
    if custom_weights[1] > 0:
        k_yy_b = torch.exp(-torch.clamp(dist_yy, min=lower_bound) / s2)
    else:
        k_yy_b = torch.exp(-torch.clamp(dist_yy, max=upper_bound) / s2)
    e_kxx_b = matrix_mean_wo_diagonal(torch.exp(-torch.clamp(dist_xx, min=lower_bound) / s2), m)
    e_kyy_b = matrix_mean_wo_diagonal(k_yy_b, m)
    e_kxy_b = matrix_mean_wo_diagonal(k_xy_b, m) if custom_weights[0] < 0 else e_kxy
    assert custom_weights[0] - custom_weights[1] == 1.0, 'w[0]-w[1] must be 1'
    return e_kxx + e_kyy - 2.0 * e_kxy, custom_weights[0] * e_kxy_b - e_kxx_b - custom_weights[1] * e_kyy_b


def mixture_mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma=None, var_targets=None, name='mmd_g', do_summary=False,
                  scope_prefix=''):
    """math_func.py:1435-1473 (fixed bandwidth list)."""
    if var_targets is not None:
        raise NotImplementedError('mixture_mmd_g: trainable sigma (var_targets) is not on the hot path')
    mmd = 0.0
    for s in sigma:
        mmd = mmd + mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma=s)
    return mmd


# ------------------------------------------------------------------------------------------------ fused loss
class _FusedMmdLoss(torch.autograd.Function):
    """(score_gen, score_data) -> (loss_gen, loss_dis) through ONE launch of the fused kernel; the backward pass only
    combines the four gradient matrices the same launch already produced."""

    @staticmethod
    def forward(ctx, score_gen, score_data, kern):
        from .. import kernels as K
        b, d = score_gen.shape
        dp = d if d in (4, 8, 16, 32, 64) else next((s for s in (4, 8, 16, 32, 64) if s >= d), None)
        if dp is None:
            raise NotImplementedError('GANLoss: score size {} > 64 is not supported by the fused kernel'.format(d))
        g = score_gen.detach().float().contiguous()
        r = score_data.detach().float().contiguous()
        if dp != d:   # zero padding leaves every pairwise distance unchanged
            g = torch.nn.functional.pad(g, (0, dp - d)).contiguous()
            r = torch.nn.functional.pad(r, (0, dp - d)).contiguous()
        grads = [torch.empty(b, dp, device=g.device, dtype=torch.float32) for _ in range(4)]
        losses = kern(g, r, grads[0], grads[1], grads[2], dLg_dreal=grads[3]).clone()
        ctx.save_for_backward(*[t[:, :d] for t in grads])
        return losses[0], losses[1]

    @staticmethod
    def backward(ctx, g_lg, g_ld):
        dLg_dgen, dLd_dgen, dLd_dreal, dLg_dreal = ctx.saved_tensors
        return g_lg * dLg_dgen + g_ld * dLd_dgen, g_lg * dLg_dreal + g_ld * dLd_dreal, None


class GANLoss(object):
    """math_func.py:2088.  Hot-path loss types: 'rep', 'rmb' (and the siblings 'mmd_g' / 'fixed_g', 'mgb', 'mmd_t' / 'fixed_t')."""

    FUSED = {'rep': 'rep', 'rep_mmd_g': 'rep', 'rmb': 'rmb', 'rep_b': 'rmb', 'rep_mmd_b': 'rmb',
             'mmd_g': 'mmd_g', 'fixed_g': 'mmd_g', 'mgb': 'mgb', 'mmd_t': 'mmd_t', 'fixed_t': 'mmd_t'}

    def __init__(self, do_summary=False):
        self.do_summary = do_summary
        self.score_gen = None
        self.score_data = None
        self.batch_size = None
        self.num_scores = None
        self.loss_gen = None
        self.loss_dis = None
        self.dis_penalty = None
        self.dis_scale = None
        self.debug_register = None
        self.sigma = [1.0, np.sqrt(2.0), 2.0, np.sqrt(8.0), 4.0]       # math_func.py:2108
        self.alpha = [0.2, 0.5, 1, 2, 5.0]                              # math_func.py:2109
        self.beta = 2.0                                                 # math_func.py:2110
        self.repulsive_weights = [0.0, -1.0]                            # math_func.py:2115
        self._kernels = {}

    def _kernel(self, loss_type, b, device):
        from .. import kernels as K
        key = (loss_type, tuple(self.repulsive_weights), b, str(device), tuple(float(s) for s in self.sigma),
               tuple(float(a) for a in self.alpha), float(self.beta))
        if key not in self._kernels:
            self._kernels[key] = K.MmdKernel(loss_type, self.repulsive_weights, b=b, device=device, sigma=self.sigma,
                                             alpha=self.alpha, beta=self.beta)
        return self._kernels[key]

    def __call__(self, score_gen, score_data, loss_type='logistic', **kwargs):
        self.score_gen, self.score_data = score_gen, score_data
        if 'batch_size' in kwargs:
            self.batch_size = kwargs['batch_size']
        if 'd' in kwargs:
            self.num_scores = kwargs['d']
        if 'dis_penalty' in kwargs:
            self.dis_penalty = kwargs['dis_penalty']
        if 'dis_scale' in kwargs:
            self.dis_scale = kwargs['dis_scale']
        if 'sigma' in kwargs:
            self.sigma = kwargs['sigma']
        if 'alpha' in kwargs:
            self.alpha = kwargs['alpha']
        if 'beta' in kwargs:
            self.beta = kwargs['beta']
        if 'rep_weights' in kwargs:
            self.repulsive_weights = list(kwargs['rep_weights'])
        if loss_type in {'fixed_g', 'mmd_g', 'fixed_t', 'mmd_t', 'rep', 'rep_gp', 'rmb', 'rmb_gp', 'mgb'}:
            assert self.batch_size is not None, 'GANLoss: batch_size must be provided'       # math_func.py:2589-2592
        if loss_type in {'rep_gp', 'rmb_gp', 'wasserstein'}:
            assert self.dis_penalty is not None, 'Discriminator penalty must be provided.'
        if loss_type not in self.FUSED:
            raise NotImplementedError('Not implemented.')                                    # math_func.py:2651
        if not score_gen.is_cuda:
            raise RuntimeError('GANLoss: scores must live on the GPU (this framework has no CPU path)')
        assert score_gen.shape == score_data.shape and score_gen.shape[0] == self.batch_size, \
            'GANLoss: scores must both be [batch_size, d]'
        kern = self._kernel(self.FUSED[loss_type], score_gen.shape[0], score_gen.device)
        self.loss_gen, self.loss_dis = _FusedMmdLoss.apply(score_gen, score_data, kern)
        if self.dis_penalty is not None:
            self.loss_dis = self.loss_dis + self.dis_penalty
        if self.dis_scale is not None and self.FUSED[loss_type] in ('rep', 'rmb'):
            # math_func.py:2524-2525 (rep) and :2546-2547 (rmb); _mmd_g_ / _mmd_g_bound_ / _mmd_t_ never scale
            self.loss_dis = (self.loss_dis - 1.0) * self.dis_scale if self.FUSED[loss_type] == 'rep' else self.loss_dis * self.dis_scale
        return self.loss_gen, self.loss_dis

    def apply(self, score_gen, score_data, loss_type='logistic', **kwargs):
        return self.__call__(score_gen, score_data, loss_type=loss_type, **kwargs)


# ------------------------------------------------------------------------------------------------ spectral norm
class SpectralNorm(object):
    """math_func.py:397-749, ops 'd' / 'c' / 'tc', PICO with a persistent `in_rand`.

    sn_def keys as in the reference (math_func.py:431-443): 'op', and for conv ops 'strides', 'dilation', 'padding',
    'data_format', 'input_shape', 'output_shape'.  apply(kernel) runs num_iter power iterations on the GPU kernels
    (batch-1 gather-GEMMs + l2-normalise), stores the new `x` (the reference's UPDATE_OPS assign) and returns sigma as
    a 0-d CUDA tensor.  The gradient d(sigma)/dW is produced inside the training engine (mmdgan_wgrad_gemm on
    (x, u)); this standalone class is the forward/inference mirror.
    """

    def __init__(self, sn_def, name_scope='SN', scope_prefix='', num_iter=1):
        self.sn_def = dict(sn_def)
        self.name_scope = name_scope
        self.scope_prefix = scope_prefix
        self.name_in_err = scope_prefix + name_scope
        self.num_iter = num_iter
        self.x = None
        self.use_u = None
        self.is_initialized = False
        self._lop = None
        op = self.sn_def['op']
        if op in {'c', 'tc'}:
            if self.sn_def.get('data_format', 'NCHW') not in ['NCHW', 'channels_first']:
                raise NotImplementedError('{}: NHWC is not on the hot path'.format(self.name_in_err))
            assert 'output_shape' in self.sn_def, '{}: for conv, output_shape must be provided.'.format(self.name_in_err)
        elif op not in {'d'}:
            raise NotImplementedError('{}: {} is not implemented.'.format(self.name_in_err, op))

    def _init_routine(self, kernel):
        from .. import kernels as K
        op = self.sn_def['op']
        ks = list(kernel.shape)
        if op == 'd':
            assert len(ks) == 2, '{}: kernel shape {} does not have length 2'.format(self.name_in_err, ks)
            self.use_u = True if ks[0] <= ks[1] else False
            x_shape = [1, ks[0]] if self.use_u else [1, ks[1]]
            self._lop = K.LinearOp('d', [ks[0]], [ks[1]], npass=FLAGS.TENSOR_PASSES, device=kernel.device)
        else:
            assert len(ks) == 4, '{}: kernel shape {} does not have length 4'.format(self.name_in_err, ks)
            ins, outs = list(self.sn_def['input_shape']), list(self.sn_def['output_shape'])
            self.use_u = True if np.prod(ins[1:]) <= np.prod(outs[1:]) else False
            if op == 'c':
                x_shape = [1] + (ins[1:] if self.use_u else outs[1:])
            else:
                x_shape = [1] + (outs[1:] if self.use_u else ins[1:])
            self._lop = K.LinearOp(op, ins[1:], outs[1:], ks[0], self.sn_def['strides'], npass=FLAGS.TENSOR_PASSES,
                                   device=kernel.device)
        x = torch.empty(x_shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0)              # math_func.py:565-567 (not normalised)
        self.x = x.to(kernel.device)
        self.is_initialized = True

    def __call__(self, kernel, **kwargs):
        from .. import kernels as K
        if 'num_iter' in kwargs:
            self.num_iter = kwargs['num_iter']
        if self.sn_def['op'] == 'd' and 1 in list(kernel.shape):
            return torch.linalg.vector_norm(kernel)                     # closed form, math_func.py:700-702
        if not self.is_initialized:
            self._init_routine(kernel)
        lop, npass = self._lop, FLAGS.TENSOR_PASSES
        lop.pack(kernel.detach().float().contiguous())
        # 'forward' of the power iteration is the layer op when x is input-shaped, its adjoint otherwise; for a 'tc'
        # kernel tf.nn.conv2d with that kernel IS the adjoint of the layer op (math_func.py:516-525)
        x_is_input = self.use_u if self.sn_def['op'] != 'tc' else (not self.use_u)
        n_in = lop.Hin * lop.Win
        n_out = lop.Hout * lop.Wout
        npl = K.mode_planes(npass)
        raw = lambda rows, c: torch.zeros((1, rows, c), dtype=torch.float32, device=kernel.device)
        vp = lambda rows, c: K.new_value_planes(rows, c, npass, kernel.device)       # operand of the layer op (forward launch)
        bp = lambda rows, c: K.new_planes(rows, c, npl, kernel.device)               # operand of the adjoint (bf16 x 6)
        xp = (vp if x_is_input else bp)(n_in if x_is_input else n_out, lop.Cs_in if x_is_input else lop.Cs_out)
        K.nchw_to_planes(self.x, xp)
        sigma = torch.zeros(1, device=kernel.device)
        for _ in range(self.num_iter):
            # both operators act as FORWARD maps here (sigma = ||F x||): fp32-grade products for either direction
            if x_is_input:
                v = raw(n_out, lop.Cs_out)
                lop.forward(xp, 1, v, out_mode=2)
                y = bp(n_out, lop.Cs_out)
                K.sn_normalize(v, v.numel(), y, sigma_out=sigma, eps=FLAGS.EPSI)
                w = raw(n_in, lop.Cs_in)
                lop.dgrad(y, 1, w, out_mode=2, npass=lop.adj_npass)
            else:
                v = raw(n_in, lop.Cs_in)
                lop.dgrad(xp, 1, v, out_mode=2, npass=lop.adj_npass)
                y = vp(n_in, lop.Cs_in)
                K.sn_normalize(v, v.numel(), y, sigma_out=sigma, eps=FLAGS.EPSI)
                w = raw(n_out, lop.Cs_out)
                lop.forward(y, 1, w, out_mode=2)
            K.sn_normalize(w, w.numel(), xp, eps=FLAGS.EPSI)
        shp = list(self.x.shape)
        if len(shp) == 2:
            self.x = K.planes_to_nchw(xp, 1, shp[1], 1, 1).reshape(shp)
        else:
            self.x = K.planes_to_nchw(xp, 1, shp[1], shp[2], shp[3])
        return sigma[0]

    def apply(self, kernel, **kwargs):
        return self.__call__(kernel, **kwargs)
