"""CPU: the oracle reproduces the committed golden fixtures in tests/golden/.

Two families with identical keys: `ref_*.npz` were produced by EXECUTING THE REFERENCE'S OWN PYTHON (math_func.py,
layer_func.py, my_sngan.py, graph_func.py imported unmodified on top of oracle/tfshim; make_reference_fixtures.py) --
these pin the oracle to the reference; the un-prefixed twins were authored from the oracle itself (make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import architectures as oa
from oracle import mmd as omm
from oracle import net as onet

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _fixtures(pattern):
    return sorted(glob.glob(os.path.join(GOLD, pattern)) + glob.glob(os.path.join(GOLD, 'ref_' + pattern)))


def _stem(path):
    name = os.path.basename(path)
    return name[4:] if name.startswith('ref_') else name


@pytest.mark.parametrize('path', _fixtures('mmd_*.npz'), ids=os.path.basename)
def test_mmd_golden(path):
    z = np.load(path)
    parts = _stem(path).split('_')
    loss_type = 'mmd_' + parts[2] if parts[1] == 'mmd' else parts[1]       # mmd_<loss>_<B>.npz with loss in {rep, rmb, mgb, mmd_g, mmd_t}
    out = omm.gan_loss_with_grads(z['gen'], z['real'], loss_type, rep_weights=tuple(z['rep_weights']))
    for k in ['loss_gen', 'loss_dis', 'dLg_dgen', 'dLd_dgen', 'dLd_ddata', 'dLg_ddata']:
        assert np.allclose(out[k], z[k], rtol=0, atol=1e-13), k


@pytest.mark.parametrize('path', _fixtures('sn_*.npz'), ids=os.path.basename)
def test_sn_golden(path):
    z = np.load(path)
    op = str(z['op'])
    design = onet.update_layer_design({'name': 't', 'op': op, 'out': int(z['cout']), 'kernel': int(z['k']), 'strides': int(z['s']),
                                       'w_nm': 's', 'act_k': 1.5})
    sp = onet.LayerSpec(design, [int(z['cin'])] * 1 + [int(z['hin'])] * 2 if op != 'd' else [int(z['cin'])], 'n/t')
    assert bool(z['use_u']) == sp.use_u                       # routing decision: bit-exact
    w = torch.from_numpy(z['w']).requires_grad_(True)
    sigma, x_upd = onet.spectral_norm(sp, w, torch.from_numpy(z['x']))
    (ds,) = torch.autograd.grad(sigma, w)
    assert np.allclose(sigma.detach().numpy(), z['sigma'], atol=1e-13)
    assert np.allclose(x_upd.numpy(), z['x_update'], atol=1e-13)
    assert np.allclose(ds.numpy(), z['dsigma_dw'], atol=1e-13)


@pytest.mark.parametrize('family', ['', 'ref_'])
@pytest.mark.parametrize('loss_type', ['rep', 'rmb', 'rep_pim'])
def test_step_golden(loss_type, family):
    z = np.load(os.path.join(GOLD, '{}step_tiny_{}.npz'.format(family, loss_type)))
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    # *_pim: FLAGS.SPECTRAL_NORM_MODE = 'sn_paper' in the reference (power iteration on the reshaped kernel matrix)
    loss_type, sn_mode = (loss_type[:-4], 'sn_paper') if loss_type.endswith('_pim') else (loss_type, 'default')
    m = onet.OracleSNGan(arch, loss_type, dtype=torch.float64, seed=3, sn_mode=sn_mode)
    for store, pre in ((m.gen_params, 'before:'), (m.dis_params, 'before:'), (m.gen_state, 'state_before:'), (m.dis_state, 'state_before:')):
        for k in list(store.keys()):
            store[k] = torch.from_numpy(z[pre + k])
    data, code = torch.from_numpy(z['data']).double(), torch.from_numpy(z['code']).double()
    lg, ld, gg, gd, _, _ = m.grads(data, code)
    assert abs(float(lg) - float(z['loss_gen'])) < 1e-13 and abs(float(ld) - float(z['loss_dis'])) < 1e-13
    for k, v in list(gg.items()) + list(gd.items()):
        assert np.allclose(v.numpy(), z['grad:' + k], atol=1e-12), k
    m.step(data, code)
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        assert np.allclose(v.numpy(), z['after:' + k], atol=1e-12), k
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        assert np.allclose(v.numpy(), z['state_after:' + k], atol=1e-12), k


@pytest.mark.parametrize('family', ['', 'ref_'])
def test_eval_sampling_golden(family):
    """The eval_sampling graph (my_sngan.py:533-551; generator and discriminator with is_training=False on a sine mesh of codes)
    and write_sprite's mosaics (graph_func.py:222-266); `ref_` = the reference's own execution."""
    import sys
    sys.path.insert(0, GOLD)
    import make_golden as mg
    z = np.load(os.path.join(GOLD, family + 'eval_tiny.npz'))
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=3)
    for store, pre in ((m.gen_params, 'var:'), (m.dis_params, 'var:'), (m.gen_state, 'state:'), (m.dis_state, 'state:')):
        for k in list(store.keys()):
            store[k] = torch.from_numpy(z[pre + k]).double()
    code = mg.mesh_codes_by_sine(z['z_support'], tuple(z['mesh_num']))
    assert np.allclose(code, z['code'], atol=1e-14)
    state = {k: v.clone() for k, v in list(m.gen_state.items()) + list(m.dis_state.items())}
    res = m.eval_sampling(torch.from_numpy(code), torch.from_numpy(z['data']).double())
    for k in ('x_gen', 's_x', 's_gen'):
        assert np.allclose(res[k].numpy(), z[k], atol=1e-12), k
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        assert torch.equal(v, state[k]), k                       # no UPDATE_OPS in the eval graph
    assert np.array_equal(mg.sprite_mosaic(np.transpose(z['x_gen'], (0, 2, 3, 1)), (2, 3), False), z['sprite'])
    assert np.array_equal(mg.sprite_mosaic(z['x_gen'][:, 0], (3, 2), True), z['sprite_inverted_gray'])


@pytest.mark.parametrize('family', ['', 'ref_'])
@pytest.mark.parametrize('tag', ['d1g2', 'd3g1'])
def test_imbalanced_update_golden(tag, family):
    """Agent(imbalanced_update=(k_dis, k_gen)) (my_sngan.py:427-439; graph_func.py:876-908): an optimiser runs on the steps whose
    global step is a multiple of its k; the other variables, their Adam slots and beta powers stay put; UPDATE_OPS always run."""
    z = np.load(os.path.join(GOLD, '{}step_tiny_imbalanced_{}.npz'.format(family, tag)))
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    imb = tuple(int(k) for k in z['imbalanced_update'])
    m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=3)
    for store, pre in ((m.gen_params, 'before:'), (m.dis_params, 'before:'), (m.gen_state, 'state_before:'), (m.dis_state, 'state_before:')):
        for k in list(store.keys()):
            store[k] = torch.from_numpy(z[pre + k])
    for t in range(int(z['steps'])):
        prev = {k: v.clone() for k, v in list(m.gen_params.items()) + list(m.dis_params.items())}
        lg, ld = m.step(torch.from_numpy(z['data_%d' % t]).double(), torch.from_numpy(z['code_%d' % t]).double(), imbalanced_update=imb)
        assert np.allclose([lg, ld], z['losses_%d' % t], atol=1e-12)
        for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
            ref = float(z['delta_%d:%s' % (t, k)])
            runs = t % imb[0 if k.startswith('dis/') else 1] == 0
            assert (ref > 0) == runs, (t, k)
            if not k.endswith('_s/bias/bias'):   # the score-layer bias gradient is analytically zero: Adam amplifies its round-off
                assert abs(float((v - prev[k]).norm()) - ref) <= 1e-9 * ref + 1e-15, (t, k)
    for k, v in list(m.gen_params.items()) + list(m.dis_params.items()):
        assert np.allclose(v.numpy(), z['after:' + k], atol=4e-3 if k.endswith('_s/bias/bias') else 1e-12), k
    for k, v in list(m.gen_state.items()) + list(m.dis_state.items()):
        assert np.allclose(v.numpy(), z['state_after:' + k], atol=1e-12), k
    assert m.global_step == int(z['global_step']) == int(z['steps'])
    with pytest.raises(AttributeError):
        onet.imbalanced_schedule(0, (2, 3))                  # my_sngan.py:439


def test_reference_fixture_families_are_complete():
    """Every oracle-authored fixture has a reference-executed twin with the same keys."""
    for path in glob.glob(os.path.join(GOLD, '*.npz')):
        name = os.path.basename(path)
        if name.startswith('ref_'):
            continue
        twin = os.path.join(GOLD, 'ref_' + name)
        assert os.path.exists(twin), twin
        assert set(np.load(twin).files) == set(np.load(path).files), name


REF_STEPS = [('ref_step_cifar_rep.npz', 'cifar'), ('ref_step_cifar_rep_k27.npz', 'cifar'), ('ref_step_stl_rmb.npz', 'stl'),
             ('ref_step_celeba_rep.npz', 'celeba'), ('ref_step_lsun_rep.npz', 'lsun')]


@pytest.mark.parametrize('name,arch_name', REF_STEPS, ids=[r[0] for r in REF_STEPS])
def test_steps_against_reference_execution(name, arch_name):
    """Consecutive fused steps of the reference's own architecture dictionaries (parsed out of my_test_cifar.py / my_test_stl.py /
    my_test_celebA.py / my_test_lsun.py) executed by the reference's SNGan.__gpu_task__ / Net / SpectralNorm / GANLoss on
    oracle/tfshim (tests/golden/make_reference_fixtures.py): losses, every gradient and every variable after each update
    (norm + strided sample per variable).  The reference creates exactly the oracle's variable names (asserted at generation)."""
    z = np.load(os.path.join(GOLD, name))
    arch = oa.ARCHITECTURES[arch_name](act_k=float(z['act_k']))
    loss_type = str(z['loss_type']) if 'loss_type' in z.files else 'rep'
    lr_list = tuple(float(v) for v in z['lr_list']) if 'lr_list' in z.files else (5e-4, 2e-4)
    B, stride = int(z['batch']), int(z['sample_stride'])
    m = onet.OracleSNGan(arch, loss_type, lr_list=lr_list, dtype=torch.float64, seed=int(z['seed']))
    onet.warm_spectral_norm(m, int(z['warm']))
    for t in range(int(z['steps'])):
        data, code = onet.synthetic_batch(arch, B, seed=5 + 10 * t, dtype=torch.float32)
        data, code = data.double(), code.double()
        lg, ld, gg, gd, _, _ = m.grads(data, code)
        assert abs(float(lg) - float(z['loss_gen_%d' % t])) < 1e-12 and abs(float(ld) - float(z['loss_dis_%d' % t])) < 1e-12
        gmax = max(float(z['grad_norm_%d:%s' % (t, k)]) for k in gd)
        for k, v in list(gg.items()) + list(gd.items()):
            ref_norm = float(z['grad_norm_%d:%s' % (t, k)])
            assert abs(float(v.norm()) - ref_norm) <= 1e-9 * ref_norm + 1e-12 * gmax, k
            assert np.allclose(v.numpy().ravel()[::stride], z['grad_sample_%d:%s' % (t, k)], rtol=1e-8, atol=1e-12 * gmax), k
        m.step(data, code)
        for store in (m.gen_params, m.dis_params, m.gen_state, m.dis_state):
            for k, v in store.items():
                # the score-layer bias gradient is analytically zero (the loss is translation invariant); Adam amplifies its
                # round-off, so that one variable is compared with an absolute tolerance of one learning-rate step
                atol = 1e-3 if k.endswith('_s/bias/bias') else 1e-10
                assert np.allclose(v.detach().numpy().ravel()[::stride], z['var_sample_%d:%s' % (t, k)], rtol=1e-8, atol=atol), k
