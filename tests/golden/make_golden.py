"""Generates the committed golden fixtures from the float64 CPU oracle (run: python tests/golden/make_golden.py).

The reference ships no golden vectors and TensorFlow 1.x cannot run here, so these fixtures are
authored from the oracle with fixed seeds; tests/test_oracle_*.py pin the oracle itself against closed forms and
finite differences.  Files: mmd_<loss>_<B>.npz, sn_<case>.npz, step_tiny_<loss>.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import architectures as oa      # noqa: E402
from oracle import mmd as omm               # noqa: E402
from oracle import net as onet              # noqa: E402


def mmd_case(loss_type, b, d=16, seed=0, w=(0.0, -1.0)):
    rng = np.random.RandomState(seed + b)
    gen = (rng.randn(b, d) * 0.35).astype(np.float32)
    real = (rng.randn(b, d) * 0.35 + 0.1).astype(np.float32)
    if b >= 4:
        gen[1] = gen[0]                          # duplicate rows: exact zero distance at the clamp
        real[2] = real[3] + np.float32(0.5) / 4  # a pair at distance^2 = 16 * (1/8)^2 = 0.25: exactly on the lower bound
    out = omm.gan_loss_with_grads(gen, real, loss_type, rep_weights=w)
    out.update(gen=gen, real=real, rep_weights=np.asarray(w))
    return out


def sn_case(op, cin, cout, hin, k, s, seed):
    design = onet.update_layer_design({'name': 't', 'op': op, 'out': cout, 'kernel': k, 'strides': s, 'w_nm': 's', 'act_k': 1.5})
    sp = onet.LayerSpec(design, [cin, hin, hin] if op != 'd' else [cin], 'n/t')
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.2).requires_grad_(True)
    x = torch.randn(sp.x_shape, generator=g, dtype=torch.float64)
    sigma, x_upd = onet.spectral_norm(sp, w, x)
    (dsdw,) = torch.autograd.grad(sigma, w)
    return dict(op=op, cin=cin, cout=cout, hin=hin, k=k, s=s, use_u=np.asarray(sp.use_u), w=w.detach().numpy(), x=x.numpy(),
                sigma=sigma.detach().numpy(), x_update=x_upd.numpy(), dsigma_dw=dsdw.numpy())


def step_case(loss_type, sn_mode='default'):
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    B = 8
    m = onet.OracleSNGan(arch, loss_type, dtype=torch.float64, seed=3, sn_mode=sn_mode)
    onet.warm_spectral_norm(m, 6)
    data, code = onet.synthetic_batch(arch, B, seed=5, dtype=torch.float32)       # fp32-representable inputs
    data, code = data.double(), code.double()
    out = {'data': data.numpy().astype(np.float32), 'code': code.numpy().astype(np.float32)}
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        out['before:' + k] = v.numpy()
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        out['state_before:' + k] = v.numpy()
    col = {}
    lg, ld, gg, gd, _, _ = m.grads(data, code, col)
    out['loss_gen'], out['loss_dis'] = lg.numpy(), ld.numpy()
    out['scores'] = torch.cat([col['s_x'], col['s_gen']], 0).detach().numpy()
    out['x_gen'] = col['x_gen'].detach().numpy()
    for k, v in list(gg.items()) + list(gd.items()):
        out['grad:' + k] = v.numpy()
    m.step(data, code)
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        out['after:' + k] = v.numpy()
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        out['state_after:' + k] = v.numpy()
    return out


IMBALANCED = {'d1g2': (1, 2), 'd3g1': (3, 1)}


def imbalanced_case(tag, steps=4):
    """`steps` consecutive steps of the tiny model under Agent(imbalanced_update=...) (graph_func.py:876-908)."""
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    B = 8
    m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=3)
    onet.warm_spectral_norm(m, 6)
    out = {'imbalanced_update': np.asarray(IMBALANCED[tag]), 'steps': np.asarray(steps)}
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        out['before:' + k] = v.numpy().copy()
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        out['state_before:' + k] = v.numpy().copy()
    prev = {k[7:]: v for k, v in out.items() if k.startswith('before:')}
    for t in range(steps):
        data, code = onet.synthetic_batch(arch, B, seed=5 + 10 * t, dtype=torch.float32)
        out['data_%d' % t], out['code_%d' % t] = data.numpy(), code.numpy()
        lg, ld = m.step(data.double(), code.double(), imbalanced_update=IMBALANCED[tag])
        out['losses_%d' % t] = np.asarray([lg, ld])
        for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):     # per step: the size of each variable's update
            out['delta_%d:%s' % (t, k)] = np.asarray(np.linalg.norm(v.numpy() - prev[k]))
            prev[k] = v.numpy().copy()
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        out['after:' + k] = v.numpy().copy()
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        out['state_after:' + k] = v.numpy().copy()
    out['global_step'] = np.asarray(m.global_step)
    return out


def mesh_codes_by_sine(z_support, mesh_num):
    """MeshCode.by_sine (math_func.py:257-291), restated with numpy in float64."""
    m0, m1 = mesh_num
    phi = np.float32(np.pi / 4.0 * np.linspace(0.0, 1.0, m0)).astype(np.float64)
    psi = np.float32(np.pi / 4.0 * np.linspace(0.0, 1.0, m1)).astype(np.float64)
    z = np.asarray(z_support, dtype=np.float64)
    out = np.zeros((m1, m0, z.shape[1]))
    for j in range(m1):
        for i in range(m0):
            out[j, i] = (np.cos(psi[j]) * z[0] + np.sin(psi[j]) * z[1]) * np.cos(phi[i]) \
                + (np.cos(psi[j]) * z[2] + np.sin(psi[j]) * z[3]) * np.sin(phi[i])
    return out.reshape(m0 * m1, -1)


def sprite_mosaic(images, mesh_num, if_invert):
    """write_sprite (graph_func.py:222-265) restated with loops: per-image min / range scaling, row-major mesh, uint8."""
    x = np.asarray(images, dtype=np.float32)
    if x.ndim == 3:
        x = x[..., None]
    if x.shape[3] == 1:
        x = np.concatenate([x, x, x], axis=3)
    n, h, w, c = x.shape
    rows, cols = mesh_num
    out = np.zeros((rows * h, cols * w, c), dtype=np.uint8)
    for i in range(n):
        t = x[i] - x[i].min()
        t = t / t.max()
        if if_invert:
            t = 1 - t
        r, q = divmod(i, cols)
        out[r * h:(r + 1) * h, q * w:(q + 1) * w] = (t * 255).astype(np.uint8)
    return out


def eval_inputs():
    """Model state for the eval_sampling fixtures: the tiny model after one training step (step_tiny_rep's `after` variables),
    with the batch-norm moving averages replaced by seeded non-trivial values so that inference-mode batch norm differs
    visibly from the training-mode one; codes = a (2, 3) sine mesh over four seeded supporting codes."""
    base = step_case('rep')
    rng = np.random.RandomState(21)
    out = {'mesh_num': np.asarray((2, 3))}
    for k, v in base.items():
        if k.startswith('after:'):
            out['var:' + k.split(':', 1)[1]] = v
        elif k.startswith('state_after:'):
            name = k.split(':', 1)[1]
            if name.endswith('moving_mean'):
                v = (rng.randn(*v.shape) * 0.3).astype(np.float32).astype(np.float64)
            elif name.endswith('moving_variance'):
                v = rng.uniform(0.4, 1.6, size=v.shape).astype(np.float32).astype(np.float64)
            out['state:' + name] = v
    out['z_support'] = rng.randn(4, base['code'].shape[1]).astype(np.float32)
    out['data'] = base['data'][:6]
    return out


def eval_case():
    out = eval_inputs()
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=3)
    for d in (m.gen_params, m.dis_params):
        for k in d:
            d[k] = torch.from_numpy(out['var:' + k]).double()
    for d in (m.gen_state, m.dis_state):
        for k in d:
            d[k] = torch.from_numpy(out['state:' + k]).double()
    code = mesh_codes_by_sine(out['z_support'], (2, 3))
    out['code'] = code
    res = m.eval_sampling(torch.from_numpy(code), torch.from_numpy(out['data']).double())
    for k, v in res.items():
        out[k] = v.numpy()
    out['sprite'] = sprite_mosaic(np.transpose(out['x_gen'], (0, 2, 3, 1)), (2, 3), False)
    out['sprite_inverted_gray'] = sprite_mosaic(out['x_gen'][:, 0], (3, 2), True)
    return out


def main():
    for lt in ('rep', 'rmb', 'mmd_g', 'mgb', 'mmd_t'):
        for b in (2, 3, 64, 128) if lt in ('rep', 'rmb') else (64,):
            np.savez_compressed(os.path.join(HERE, 'mmd_{}_{}.npz'.format(lt, b)), **mmd_case(lt, b))
    np.savez_compressed(os.path.join(HERE, 'mmd_rep_256.npz'), **mmd_case('rep', 256))
    np.savez_compressed(os.path.join(HERE, 'mmd_rmb_w_1_0.npz'), **mmd_case('rmb', 32, w=(1.0, 0.0)))
    cases = {'conv_k3s1_useu': ('c', 8, 16, 6, 3, 1), 'conv_k4s2_nouseu': ('c', 8, 16, 8, 4, 2), 'dense_nouseu': ('d', 64, 16, 1, 1, 1),
             'dense_useu': ('d', 16, 32, 1, 1, 1), 'conv_image': ('c', 3, 8, 8, 3, 1)}
    for i, (name, c) in enumerate(cases.items()):
        np.savez_compressed(os.path.join(HERE, 'sn_{}.npz'.format(name)), **sn_case(*c, seed=10 + i))
    for lt in ('rep', 'rmb'):
        np.savez_compressed(os.path.join(HERE, 'step_tiny_{}.npz'.format(lt)), **step_case(lt))
    np.savez_compressed(os.path.join(HERE, 'step_tiny_rep_pim.npz'), **step_case('rep', sn_mode='sn_paper'))
    np.savez_compressed(os.path.join(HERE, 'eval_tiny.npz'), **eval_case())
    for tag in IMBALANCED:
        np.savez_compressed(os.path.join(HERE, 'step_tiny_imbalanced_{}.npz'.format(tag)), **imbalanced_case(tag))


if __name__ == '__main__':
    main()
