"""CPU: the oracle against closed forms, an independent double-loop restatement and float64 finite differences.
The reference has no golden vectors, so this is what pins the oracle ("parity unpinned")."""
import math

import numpy as np
import pytest
import torch

from oracle import architectures as oa
from oracle import mmd as omm
from oracle import net as onet


@pytest.mark.parametrize('loss_type', ['rep', 'rmb'])
@pytest.mark.parametrize('w', [(0.0, -1.0), (1.0, 0.0), (-1.0, -2.0)])
def test_losses_match_double_loop(loss_type, w):
    rng = np.random.RandomState(1)
    g, r = rng.randn(7, 5) * 0.6, rng.randn(7, 5) * 0.6 + 0.2
    lg, ld = omm.gan_loss(torch.from_numpy(g), torch.from_numpy(r), loss_type, rep_weights=w)
    lg2, ld2 = omm.rep_loss_loops(g, r, loss_type, w)
    assert abs(float(lg) - lg2) < 1e-12 and abs(float(ld) - ld2) < 1e-12


def test_identical_sets_and_diagonal_exclusion():
    x = torch.randn(9, 4, dtype=torch.float64)
    lg, ld = omm.gan_loss(x, x.clone(), 'rep')
    # with G == R: e_gg == e_rr == e_gr (the i == j pair is dropped from k_xy too, math_func.py:1326) -> both losses 0
    assert abs(float(lg)) < 1e-14 and abs(float(ld)) < 1e-14
    dxx, dxy, dyy = omm.get_squared_dist(x, x.clone())
    assert float(torch.diagonal(dxx).abs().max()) == 0.0          # diagonal exactly 0 (math_func.py:804)
    assert float(dxx.min()) >= 0.0


def test_rep_weights_assert():
    x = torch.randn(4, 3, dtype=torch.float64)
    with pytest.raises(AssertionError):
        omm.gan_loss(x, x + 1, 'rep', rep_weights=(0.0, 0.0))


@pytest.mark.parametrize('loss_type', ['rep', 'rmb', 'mmd_g', 'mgb', 'mmd_t'])
def test_score_gradients_finite_differences(loss_type):
    rng = np.random.RandomState(3)
    g, r = rng.randn(6, 4) * 0.5, rng.randn(6, 4) * 0.5 + 0.1
    ref = omm.gan_loss_with_grads(g, r, loss_type)
    eps = 1e-6
    for key, which, idx in [('dLg_dgen', 0, 0), ('dLd_dgen', 0, 1), ('dLd_ddata', 1, 1), ('dLg_ddata', 1, 0)]:
        num = np.zeros_like(g)
        for i in range(g.shape[0]):
            for j in range(g.shape[1]):
                a = [g.copy(), r.copy()]
                b = [g.copy(), r.copy()]
                a[which][i, j] += eps
                b[which][i, j] -= eps
                fa = omm.gan_loss(torch.from_numpy(a[0]), torch.from_numpy(a[1]), loss_type)[idx]
                fb = omm.gan_loss(torch.from_numpy(b[0]), torch.from_numpy(b[1]), loss_type)[idx]
                num[i, j] = (float(fa) - float(fb)) / (2 * eps)
        assert np.abs(num - ref[key]).max() < 1e-7, key


def test_rmb_bounds_gate_the_gradient():
    # all generated pairs closer than the lower bound 0.25 -> loss_dis has no gradient w.r.t. them (math_func.py:1386)
    g = np.zeros((4, 2)) + np.arange(4)[:, None] * 0.01
    r = np.arange(8, dtype=np.float64).reshape(4, 2) * 3.0        # all real pairs farther than the upper bound 4.0
    out = omm.gan_loss_with_grads(g, r, 'rmb')
    assert np.abs(out['dLd_dgen']).max() == 0.0 and np.abs(out['dLd_ddata']).max() == 0.0
    assert np.abs(out['dLg_dgen']).max() > 0.0


def test_spectral_norm_converges_to_matrix_two_norm():
    g = torch.Generator().manual_seed(0)
    for cin, cout in [(24, 10), (10, 24)]:
        design = onet.update_layer_design({'name': 't', 'op': 'd', 'out': cout, 'w_nm': 's', 'act_k': 1.0})
        sp = onet.LayerSpec(design, [cin], 'n/t')
        assert sp.use_u == (cin <= cout)
        w = torch.randn(cin, cout, generator=g, dtype=torch.float64)
        x = torch.randn(sp.x_shape, generator=g, dtype=torch.float64)
        for _ in range(300):
            sigma, x = onet.spectral_norm(sp, w, x)
        assert abs(float(sigma) - float(torch.linalg.matrix_norm(w, 2))) < 1e-8
    # 1x1 convolution: the operator norm is the 2-norm of the [Cin, Cout] matrix
    design = onet.update_layer_design({'name': 't', 'op': 'c', 'out': 6, 'kernel': 1, 'strides': 1, 'w_nm': 's', 'act_k': 1.0})
    sp = onet.LayerSpec(design, [5, 4, 4], 'n/t')
    w = torch.randn(1, 1, 5, 6, generator=g, dtype=torch.float64)
    x = torch.randn(sp.x_shape, generator=g, dtype=torch.float64)
    for _ in range(400):
        sigma, x = onet.spectral_norm(sp, w, x)
    assert abs(float(sigma) - float(torch.linalg.matrix_norm(w[0, 0], 2))) < 1e-7


def test_spectral_norm_first_iteration_uses_unnormalised_x_and_pre_update_x():
    design = onet.update_layer_design({'name': 't', 'op': 'd', 'out': 3, 'w_nm': 's', 'act_k': 1.0})
    sp = onet.LayerSpec(design, [8], 'n/t')                      # 8 > 3 -> use_u False, x in R^3, forward = x W^T
    w = torch.randn(8, 3, dtype=torch.float64)
    x = torch.randn(1, 3, dtype=torch.float64) * 5.0
    sigma, x_upd = onet.spectral_norm(sp, w, x)
    v = x @ w.t()
    assert abs(float(sigma) - float(v.norm())) < 1e-12           # sigma = ||forward(x)|| with the PRE-update x
    y = v / (v.norm() + 1e-10)
    assert torch.allclose(x_upd, (y @ w) / ((y @ w).norm() + 1e-10))


def test_sn_routing_table_of_the_shipped_architectures():
    """CIFAR l1 use_u, l2/l4/l6 not, l3/l5/l7 use_u (equality), l8 dense 8192 -> 16 not."""
    specs = onet.build_net(oa.cifar()['discriminator'], 'dis', [3, 32, 32])
    assert [s.use_u for s in specs] == [True, False, True, False, True, False, True, False]
    assert [s.x_shape for s in specs][:3] == [[1, 3, 32, 32], [1, 128, 16, 16], [1, 128, 16, 16]] and specs[-1].x_shape == [1, 16]
    specs = onet.build_net(oa.celeba()['discriminator'], 'dis', [3, 64, 64])
    assert [s.use_u for s in specs] == [True, False, True, False, True, False, True, False, True, False]


def test_parameter_counts_and_shapes():
    for name, n_gen, n_dis in [('cifar', 3811907, 5983760)]:
        m = onet.OracleSNGan(oa.ARCHITECTURES[name](), 'rep')
        assert sum(v.numel() for v in m.gen_params.values()) == n_gen
        assert sum(v.numel() for v in m.dis_params.values()) == n_dis
    specs = onet.build_net(oa.stl()['generator'], 'gen', [128])
    assert specs[0].has_bn and not specs[0].has_bias              # BN removes the plain bias (layer_func.py:1241-1242)
    assert specs[-1].op_out_shape == [3, 48, 48]


def test_batch_norm_matches_torch_and_tf_moving_average():
    import torch.nn.functional as F
    arch = oa.tiny()
    specs = onet.build_net(arch['generator'], 'gen', [arch['code'][0][0]])
    g = torch.Generator().manual_seed(0)
    params, state = onet.init_params(specs, g, torch.float64)
    sp = specs[1]
    x = torch.randn(6, *specs[0].out_shape, generator=g, dtype=torch.float64)
    y, upd = onet.net_forward([sp], params, state, x, True)
    k = params[sp.kernel_name].permute(3, 2, 0, 1)
    z = F.conv_transpose2d(x, k, stride=2, padding=1)
    rm, rv = torch.zeros(z.shape[1], dtype=torch.float64), torch.ones(z.shape[1], dtype=torch.float64)
    ref = F.relu(F.batch_norm(z, rm, rv, params[sp.bn_name('gamma')], params[sp.bn_name('beta')], True, 0.01, 1e-3))
    assert torch.allclose(y, ref, atol=1e-12)
    assert torch.allclose(upd[sp.bn_name('moving_mean')], rm, atol=1e-12) and torch.allclose(upd[sp.bn_name('moving_variance')], rv, atol=1e-12)


def test_tf_adam_formula():
    p0, g = torch.tensor([1.0, -2.0], dtype=torch.float64), torch.tensor([0.5, -0.25], dtype=torch.float64)
    params = {'p': p0.clone()}
    opt = onet.TFAdam(params, 2e-4)
    opt.apply(params, {'p': g})
    m, v = 0.5 * g, 0.001 * g * g
    lr_t = 2e-4 * math.sqrt(1 - 0.999) / (1 - 0.5)
    assert torch.allclose(params['p'], p0 - lr_t * m / (v.sqrt() + 1e-8), atol=1e-15)


def test_full_step_gradients_finite_differences():
    """d(loss_dis)/d(D kernel) through the spectral norm and d(loss_gen)/d(G kernel) through D, float64 central differences."""
    arch = oa.tiny(channels=(8, 8), size=8, code=8, act_k=2.6)
    m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=1)
    onet.warm_spectral_norm(m, 4)
    data, code = onet.synthetic_batch(arch, 4, seed=2, dtype=torch.float64)
    lg, ld, gg, gd, _, _ = m.grads(data, code)
    eps = 1e-6
    for params, grads, idx, name in [(m.dis_params, gd, 1, 'dis/l2_ds/kernel/kernel'), (m.gen_params, gg, 0, 'gen/l2_up/kernel/kernel')]:
        flat = params[name].reshape(-1)
        for pos in (0, 7, flat.numel() - 1):
            old = float(flat[pos])
            flat[pos] = old + eps
            fa = m.forward_losses(data, code)[idx]
            flat[pos] = old - eps
            fb = m.forward_losses(data, code)[idx]
            flat[pos] = old
            num = (float(fa) - float(fb)) / (2 * eps)
            assert abs(num - float(grads[name].reshape(-1)[pos])) < 1e-6 * max(1.0, abs(num)), (name, pos)


def test_tf32_rounding_helper():
    x = torch.tensor([1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, -3.0 - 2.0 ** -10], dtype=torch.float32)
    r = onet.round_tf32(x)
    assert r.tolist() == [1.0 + 2.0 ** -10, 1.0, -3.0 - 2.0 ** -9 + 2.0 ** -9 - 2.0 ** -10 + 0.0] or float(r[0]) == 1.0 + 2.0 ** -10
    assert float(r[1]) == 1.0
