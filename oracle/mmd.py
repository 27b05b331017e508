"""Oracle (test infrastructure): rep / rmb MMD losses.  Pinned against outputs of the reference's own Python run on oracle/tfshim (tests/golden/ref_*.npz; see oracle/__init__.py).

Restates GeneralTools/math_func.py of the reference:
  get_squared_dist          math_func.py:767-858   (mode 'xxxyyy', Gram trick, clamp at 0)
  matrix_mean_wo_diagonal   math_func.py:1048-1069
  mmd_g                     math_func.py:1288-1352
  mmd_g_bounded             math_func.py:1356-1431
  mixture_mmd_g             math_func.py:1435-1473
  GANLoss rep / rmb         math_func.py:2505-2550, 2556-2658

TF semantics hard-coded here (not visible in the reference source): tf.maximum /
tf.minimum pass the gradient to the first argument on ties -> torch.clamp, which
passes the gradient at equality as well.
"""
import numpy as np
import torch


def get_squared_dist(x, y=None, mode='xxxyyy'):
    """math_func.py:767-858.  x, y: [B, d].  Returns the pairwise squared distances."""
    if x.dim() > 2:
        raise AttributeError('get_dist: Input must be a matrix.')
    if y is None:
        mode = 'xx'
    if mode in ['xx', 'xxxy', 'xxxyyy']:
        xxt = x @ x.t()                                   # math_func.py:800
        dx = torch.diagonal(xxt)                          # math_func.py:803
        dist_xx = torch.clamp(dx[:, None] - 2.0 * xxt + dx[None, :], min=0.0)  # :804
        if mode == 'xx':
            return dist_xx
        if mode == 'xxxy':
            xyt = x @ y.t()
            dy = (y * y).sum(1)
            dist_xy = torch.clamp(dx[:, None] - 2.0 * xyt + dy[None, :], min=0.0)
            return dist_xx, dist_xy
        xyt = x @ y.t()                                   # :826
        yyt = y @ y.t()                                   # :827
        dy = torch.diagonal(yyt)                          # :831
        dist_xy = torch.clamp(dx[:, None] - 2.0 * xyt + dy[None, :], min=0.0)  # :832
        dist_yy = torch.clamp(dy[:, None] - 2.0 * yyt + dy[None, :], min=0.0)  # :833
        return dist_xx, dist_xy, dist_yy
    if mode == 'xy':
        dx = (x * x).sum(1)
        dy = (y * y).sum(1)
        xyt = x @ y.t()
        return torch.clamp(dx[:, None] - 2.0 * xyt + dy[None, :], min=0.0)
    raise AttributeError('Mode {} not supported'.format(mode))


def matrix_mean_wo_diagonal(matrix, num_row, num_col=None):
    """math_func.py:1048-1069."""
    if num_col is None:
        return (matrix.sum() - torch.diagonal(matrix).sum()) / (num_row * (num_row - 1.0))
    return (matrix.sum() - torch.diagonal(matrix).sum()) / (num_row * num_col - min(num_col, num_row))


def mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma=1.0, upper_bound=None, lower_bound=None,
          custom_weights=None):
    """math_func.py:1288-1352 (var_target branch not on the hot path)."""
    if lower_bound is None:
        k_xx = torch.exp(-dist_xx / (2.0 * sigma ** 2))
        k_yy = torch.exp(-dist_yy / (2.0 * sigma ** 2))
    else:
        k_xx = torch.exp(-torch.clamp(dist_xx, min=lower_bound) / (2.0 * sigma ** 2))
        k_yy = torch.exp(-torch.clamp(dist_yy, min=lower_bound) / (2.0 * sigma ** 2))
    if upper_bound is None:
        k_xy = torch.exp(-dist_xy / (2.0 * sigma ** 2))
    else:
        k_xy = torch.exp(-torch.clamp(dist_xy, max=upper_bound) / (2.0 * sigma ** 2))
    m = float(batch_size)
    e_kxx = matrix_mean_wo_diagonal(k_xx, m)
    e_kxy = matrix_mean_wo_diagonal(k_xy, m)
    e_kyy = matrix_mean_wo_diagonal(k_yy, m)
    if custom_weights is None:
        return e_kxx + e_kyy - 2.0 * e_kxy
    assert custom_weights[0] - custom_weights[1] == 1.0, 'w[0]-w[1] must be 1'
    mmd1 = e_kxx + e_kyy - 2.0 * e_kxy
    mmd2 = custom_weights[0] * e_kxy - e_kxx - custom_weights[1] * e_kyy
    return mmd1, mmd2


def mmd_g_bounded(dist_xx, dist_xy, dist_yy, batch_size, sigma=1.0, upper_bound=None, lower_bound=None,
                  custom_weights=None):
    """math_func.py:1356-1431.  Reproduces the sign quirk at 1387-1390 vs 1402 literally."""
    k_xx = torch.exp(-dist_xx / (2.0 * sigma ** 2))
    k_yy = torch.exp(-dist_yy / (2.0 * sigma ** 2))
    k_xy = torch.exp(-dist_xy / (2.0 * sigma ** 2))
    k_xx_b = torch.exp(-torch.clamp(dist_xx, min=lower_bound) / (2.0 * sigma ** 2))
    if custom_weights[0] > 0:
        k_xy_b = torch.exp(-torch.clamp(dist_xy, max=upper_bound) / (2.0 * sigma ** 2))
    else:
        k_xy_b = k_xy
    if custom_weights[1] > 0:
        k_yy_b = torch.exp(-torch.clamp(dist_yy, min=lower_bound) / (2.0 * sigma ** 2))
    else:
        k_yy_b = torch.exp(-torch.clamp(dist_yy, max=upper_bound) / (2.0 * sigma ** 2))
    m = float(batch_size)
    e_kxx = matrix_mean_wo_diagonal(k_xx, m)
    e_kxy = matrix_mean_wo_diagonal(k_xy, m)
    e_kyy = matrix_mean_wo_diagonal(k_yy, m)
    e_kxx_b = matrix_mean_wo_diagonal(k_xx_b, m)
    e_kyy_b = matrix_mean_wo_diagonal(k_yy_b, m)
    e_kxy_b = matrix_mean_wo_diagonal(k_xy_b, m) if custom_weights[0] < 0 else e_kxy
    if custom_weights is None:
        return e_kxx + e_kyy - 2.0 * e_kxy
    assert custom_weights[0] - custom_weights[1] == 1.0, 'w[0]-w[1] must be 1'
    mmd1 = e_kxx + e_kyy - 2.0 * e_kxy
    mmd2 = custom_weights[0] * e_kxy_b - e_kxx_b - custom_weights[1] * e_kyy_b
    return mmd1, mmd2


def mixture_mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma):
    """math_func.py:1435-1473 (fixed sigma list)."""
    mmd = 0.0
    for s in sigma:
        mmd = mmd + mmd_g(dist_xx, dist_xy, dist_yy, batch_size, sigma=s)
    return mmd


MIX_SIGMA = [1.0, float(np.sqrt(2.0)), 2.0, float(np.sqrt(8.0)), 4.0]   # math_func.py:2108
MIX_ALPHA = [0.2, 0.5, 1.0, 2.0, 5.0]                                   # math_func.py:2109 (mmd-t kernel scales)
MIX_BETA = 2.0                                                          # math_func.py:2110


def mmd_t(dist_xx, dist_xy, dist_yy, batch_size, alpha=1.0, beta=2.0):
    """math_func.py:1087-1142: t-distribution kernel k = exp(-alpha * log(d / (beta * alpha) + 1))."""
    k = lambda d: torch.exp(-alpha * torch.log(d / (beta * alpha) + 1.0))
    m = float(batch_size)
    return (matrix_mean_wo_diagonal(k(dist_xx), m) + matrix_mean_wo_diagonal(k(dist_yy), m)
            - 2.0 * matrix_mean_wo_diagonal(k(dist_xy), m))


def mixture_mmd_t(dist_xx, dist_xy, dist_yy, batch_size, alpha, beta=2.0):
    """math_func.py:1145-1184 (fixed alpha list)."""
    mmd = 0.0
    for a in alpha:
        mmd = mmd + mmd_t(dist_xx, dist_xy, dist_yy, batch_size, alpha=a, beta=beta)
    return mmd


def gan_loss(score_gen, score_data, loss_type, batch_size=None, rep_weights=(0.0, -1.0), sigma=None):
    """GANLoss.apply for the hot-path loss types (math_func.py:2556-2658).

    Returns (loss_gen, loss_dis).  x = generated scores, y = real scores (math_func.py:2511).
    """
    if batch_size is None:
        batch_size = score_gen.shape[0]
    rep_weights = list(rep_weights)
    d_gg, d_gd, d_dd = get_squared_dist(score_gen, score_data)
    if loss_type in {'rep', 'rep_mmd_g'}:                 # math_func.py:2505-2528
        return mmd_g(d_gg, d_gd, d_dd, batch_size, sigma=1.0, custom_weights=rep_weights)
    if loss_type in {'rmb', 'rep_b', 'rep_mmd_b'}:        # math_func.py:2530-2550
        return mmd_g_bounded(d_gg, d_gd, d_dd, batch_size, sigma=1.0, lower_bound=0.25, upper_bound=4.0,
                             custom_weights=rep_weights)
    if loss_type in {'fixed_g', 'mmd_g'}:                 # math_func.py:2160-2173
        lg = mixture_mmd_g(d_gg, d_gd, d_dd, batch_size, sigma=MIX_SIGMA if sigma is None else sigma)
        return lg, -lg
    if loss_type in {'fixed_t', 'mmd_t'}:                 # math_func.py:2263-2275
        lg = mixture_mmd_t(d_gg, d_gd, d_dd, batch_size, alpha=MIX_ALPHA, beta=MIX_BETA)
        return lg, -lg
    if loss_type == 'mgb':                                # math_func.py:2175-2193
        lg = mmd_g(d_gg, d_gd, d_dd, batch_size, sigma=1.0)
        mmd_b = mmd_g(d_gg, d_gd, d_dd, batch_size, sigma=1.0, upper_bound=4, lower_bound=0.25)
        return lg, -mmd_b
    raise NotImplementedError('Not implemented.')


def gan_loss_with_grads(score_gen, score_data, loss_type, rep_weights=(0.0, -1.0), dtype=torch.float64,
                        batch_size=None):
    """Losses plus the four score gradients, through autograd, in `dtype`.

    Returns dict(loss_gen, loss_dis, dLg_dgen, dLg_ddata, dLd_dgen, dLd_ddata) as numpy arrays.
    """
    g = torch.as_tensor(np.asarray(score_gen)).to(dtype).clone().requires_grad_(True)
    r = torch.as_tensor(np.asarray(score_data)).to(dtype).clone().requires_grad_(True)
    lg, ld = gan_loss(g, r, loss_type, batch_size=batch_size, rep_weights=rep_weights)
    dlg = torch.autograd.grad(lg, [g, r], retain_graph=True, allow_unused=True)
    dld = torch.autograd.grad(ld, [g, r], allow_unused=True)
    z = lambda t, ref: (torch.zeros_like(ref) if t is None else t).detach().numpy()
    return dict(loss_gen=lg.detach().numpy(), loss_dis=ld.detach().numpy(),
                dLg_dgen=z(dlg[0], g), dLg_ddata=z(dlg[1], r),
                dLd_dgen=z(dld[0], g), dLd_ddata=z(dld[1], r))


def rep_loss_loops(score_gen, score_data, loss_type='rep', rep_weights=(0.0, -1.0)):
    """Pure-Python double loop of the same losses, no Gram trick (exact pairwise differences).

    Used only by the oracle self-tests on tiny inputs as an independent cross-check
    (SURVEY.md section 8c item (v)).
    """
    g = np.asarray(score_gen, dtype=np.float64)
    r = np.asarray(score_data, dtype=np.float64)
    b = g.shape[0]
    lo, hi = 0.25, 4.0
    w0, w1 = rep_weights
    s = dict(gg=0.0, gr=0.0, rr=0.0, gg_b=0.0, gr_b=0.0, rr_b=0.0)
    for i in range(b):
        for j in range(b):
            if i == j:
                continue
            dgg = float(((g[i] - g[j]) ** 2).sum())
            dgr = float(((g[i] - r[j]) ** 2).sum())
            drr = float(((r[i] - r[j]) ** 2).sum())
            s['gg'] += np.exp(-dgg / 2.0)
            s['gr'] += np.exp(-dgr / 2.0)
            s['rr'] += np.exp(-drr / 2.0)
            s['gg_b'] += np.exp(-max(dgg, lo) / 2.0)
            s['gr_b'] += np.exp(-min(dgr, hi) / 2.0) if w0 > 0 else np.exp(-dgr / 2.0)
            s['rr_b'] += np.exp(-max(drr, lo) / 2.0) if w1 > 0 else np.exp(-min(drr, hi) / 2.0)
    c = 1.0 / (b * (b - 1.0))
    e = {k: v * c for k, v in s.items()}
    loss_gen = e['gg'] + e['rr'] - 2.0 * e['gr']
    if loss_type == 'rep':
        loss_dis = w0 * e['gr'] - e['gg'] - w1 * e['rr']
    else:
        e_gr_b = e['gr_b'] if w0 < 0 else e['gr']
        loss_dis = w0 * e_gr_b - e['gg_b'] - w1 * e['rr_b']
    return loss_gen, loss_dis
