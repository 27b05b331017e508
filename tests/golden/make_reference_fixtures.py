"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON (run here: python tests/golden/make_reference_fixtures.py).

/root/reference is TF-1.8 graph code and TensorFlow cannot be installed in this image, so the unmodified reference
modules (GeneralTools/math_func.py, GeneralTools/layer_func.py, GeneralTools/graph_func.py, DeepLearning/my_sngan.py) are
imported on top of `oracle/tfshim` -- an eager, PyTorch-CPU, float64 stand-in for the tf.* calls they make -- and their
own functions are run:

  ref_mmd_<loss>_<B>.npz   GANLoss.apply -> get_squared_dist, mmd_g, mmd_g_bounded, mixture_mmd_g   (math_func.py:767-858,
                           1048-1069, 1288-1473, 2088-2658); losses and their score gradients
  ref_sn_<case>.npz        SpectralNorm(sn_def).apply(kernel)  (math_func.py:397-749): sigma, the assign(in_rand, .) value in
                           UPDATE_OPS, d sigma / d kernel, the use_u routing flag
  ref_step_<net>_<loss>.npz  SNGan.init_net + SNGan.__gpu_task__ + multi_opt_config + apply_gradients + UPDATE_OPS
                           (my_sngan.py:85-108, 259-323, 412-426; graph_func.py:478-575, 848-854): Net / Routine /
                           ParametricOperation build the generator and discriminator from the architecture dictionary;
                           losses, scores, generated images, every gradient, and every variable after the update.
  ref_step_cifar_rep.npz, ref_step_cifar_rep_k27.npz, ref_step_stl_rmb.npz, ref_step_celeba_rep.npz, ref_step_lsun_rep.npz
                           the same through the architecture dictionaries parsed out of the reference's my_test_*.py scripts
                           (batch 4; gradients stored as norms plus a strided sample per variable to keep the file small;
                           the initial variables are the oracle's seeded initialisation, so only the seed is stored)

  ref_step_tiny_imbalanced_<tag>.npz   four steps under Agent(imbalanced_update=(k_dis, k_gen)): the op construction of
                           my_sngan.py:427-439 and the per-step op selection of graph_func.py:876-908
  ref_eval_tiny.npz        SNGan.eval_sampling's graph section (my_sngan.py:523-551): MeshCode.by_sine codes, the generator and the
                           discriminator with is_training=False, clipping; write_sprite's uint8 mosaics (graph_func.py:222-266)

Inputs (seeds, initial variables) are the ones tests/golden/make_golden.py uses for the oracle-authored twins, so each
ref_* file has the same keys as its twin and the same tests run against both.  Only the fixtures travel to the GPU box.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('MMDGAN_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tfshim'))
sys.path.insert(0, REFERENCE)

for _alias, _ty in (('int', int), ('float', float), ('bool', bool)):      # NumPy-1.x aliases the 2018 reference code uses
    if _alias not in np.__dict__:
        setattr(np, _alias, _ty)

import torch                                   # noqa: E402
import tensorflow as tf                        # noqa: E402  (oracle/tfshim)
from GeneralTools import math_func as rmf      # noqa: E402  (the reference)
from GeneralTools import graph_func as rgf     # noqa: E402
from GeneralTools.misc_fun import FLAGS        # noqa: E402
from DeepLearning.my_sngan import SNGan        # noqa: E402

from oracle import architectures as oa         # noqa: E402
from oracle import net as onet                 # noqa: E402
import make_golden as mg                       # noqa: E402  (input generators shared with the oracle-authored twins)

FLAGS.SILENT_MODE = True
assert tf.__version__.endswith('tfshim')


SAMPLE = 257   # stride of the per-variable sample in the CIFAR fixtures


# ------------------------------------------------------------------------------------------------ losses
def ref_mmd_case(loss_type, b, w=(0.0, -1.0)):
    base = mg.mmd_case(loss_type, b, w=w)                     # inputs (and the oracle's answers, discarded)
    g = torch.tensor(base['gen'], dtype=torch.float64, requires_grad=True)
    r = torch.tensor(base['real'], dtype=torch.float64, requires_grad=True)
    loss = rmf.GANLoss(do_summary=False)
    if loss_type in ('rep', 'rmb'):
        lg, ld = loss.apply(g, r, loss_type, batch_size=b, d=g.shape[1], rep_weights=list(w))   # my_sngan.py:284-287
    else:
        lg, ld = loss.apply(g, r, loss_type, batch_size=b, d=g.shape[1])                        # my_sngan.py:289-290
    dlg = torch.autograd.grad(lg, [g, r], retain_graph=True, allow_unused=True)
    dld = torch.autograd.grad(ld, [g, r], allow_unused=True)
    z = lambda t, ref: (torch.zeros_like(ref) if t is None else t).detach().numpy()
    return dict(gen=base['gen'], real=base['real'], rep_weights=np.asarray(w), loss_gen=lg.detach().numpy(),
                loss_dis=ld.detach().numpy(), dLg_dgen=z(dlg[0], g), dLg_ddata=z(dlg[1], r), dLd_dgen=z(dld[0], g),
                dLd_ddata=z(dld[1], r))


# ------------------------------------------------------------------------------------------------ spectral norm
def ref_sn_case(op, cin, cout, hin, k, s, seed):
    base = mg.sn_case(op, cin, cout, hin, k, s, seed)
    tf.reset_default_graph()
    w = torch.tensor(base['w'], dtype=torch.float64, requires_grad=True)
    if op == 'd':
        sn_def = {'op': 'd'}                                                         # layer_func.py:797-800
    else:
        hout = -(-hin // s)
        sn_def = {'op': op, 'strides': s, 'dilation': 1, 'padding': 'SAME', 'data_format': 'NCHW',
                  'input_shape': [64, cin, hin, hin], 'output_shape': [64, cout, hout, hout]}   # layer_func.py:804-809
    sn = rmf.SpectralNorm(sn_def, 'SN', num_iter=1)
    # the variable in_rand is created inside apply(); pre-seed the shim's variable store with the fixture's x
    with tf.variable_scope('SN'):
        x = tf.get_variable('in_rand', shape=list(base['x'].shape), initializer=lambda shape: torch.tensor(base['x'], dtype=torch.float64),
                            trainable=False)
    sigma = sn.apply(w)
    assert tuple(sn.x.shape) == tuple(base['x'].shape) and sn.x is x
    (upd,) = tf.get_collection(tf.GraphKeys.UPDATE_OPS)
    (dsdw,) = torch.autograd.grad(sigma, w)
    return dict(op=op, cin=cin, cout=cout, hin=hin, k=k, s=s, use_u=np.asarray(bool(sn.use_u)), w=base['w'], x=base['x'],
                sigma=sigma.detach().numpy(), x_update=upd.value.detach().numpy(), dsigma_dw=dsdw.numpy())


# ------------------------------------------------------------------------------------------------ the fused step
def reference_architecture(script):
    """The `architecture = {...}` literal of a my_test_*.py script, evaluated with the script's own act_k / w_nm."""
    src = open(os.path.join(REFERENCE, script)).read()
    tree = ast.parse(src)
    env = {'np': np}
    for node in tree.body:
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in ('act_k', 'w_nm', 'architecture'):
            exec(compile(ast.Module([node], []), script, 'exec'), env)
    return env['architecture']


class ReferenceRun(object):
    """SNGan.training's graph section (my_sngan.py:402-426) executed eagerly, one construction per step."""

    def __init__(self, architecture, loss_type, lr_list, rep_weights=(0.0, -1.0)):
        tf.reset_default_graph()
        self.mdl = SNGan(architecture, num_class=0, loss_type=loss_type, optimizer='adam', do_summary=False,
                         rep_weights=list(rep_weights))
        self.global_step = torch.zeros((), dtype=torch.int64)
        _, self.opt_ops = rgf.multi_opt_config(list(lr_list), end_lr=1e-7, optimizer='adam', global_step=self.global_step)
        self.code = None

    def set_variables(self, values):
        """Create the reference's variables under the reference's names with given initial values."""
        store = tf.shim_variables()
        assert not store
        self._init_values = {k: torch.as_tensor(np.asarray(v), dtype=torch.float64) for k, v in values.items()}

    def build(self, data, code, batch_size):
        """init_net + __gpu_task__ (is_training=True) with the code batch injected through tf.random_normal."""
        tf._S.collections.clear()
        self.mdl.init_net()                                                           # my_sngan.py:404
        orig_rn = tf.random_normal
        tf.random_normal = lambda shape, **kw: torch.as_tensor(code, dtype=torch.float64).reshape(tuple(shape))
        orig_gv = tf.get_variable
        init = getattr(self, '_init_values', None)

        def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
            full = '/'.join(tf._S.scope + [name])
            if init is not None and full not in tf.shim_variables():
                assert full in init, 'reference created a variable the oracle does not know: ' + full
                v0 = init[full]
                assert list(v0.shape) == ([shape] if isinstance(shape, int) else list(shape)), (full, list(v0.shape), shape)
                initializer = lambda s: v0.clone()                                    # noqa: E731
            return orig_gv(name, shape, dtype, initializer, trainable, **kw)
        tf.get_variable = get_variable
        try:
            data_batch = {'x': torch.as_tensor(data, dtype=torch.float64)}
            grads_list, loss_list = self.mdl.__gpu_task__(
                batch_size=batch_size, is_training=True, data_batch=data_batch, opt_op=self.opt_ops)   # my_sngan.py:419-421
        finally:
            tf.random_normal = orig_rn
            tf.get_variable = orig_gv
        if init is not None:
            missing = set(init) - set(tf.shim_variables())
            assert not missing, 'oracle variables the reference never created: {}'.format(sorted(missing))
        return grads_list, loss_list

    def run(self, grads_list, imbalanced_update=None):
        if imbalanced_update is None or imbalanced_update[0] == 1:
            dis_op = self.opt_ops[0].apply_gradients(grads_list[0], global_step=self.global_step)   # my_sngan.py:424, 431
            gen_op = self.opt_ops[1].apply_gradients(grads_list[1])                                 # my_sngan.py:425, 432
        elif imbalanced_update[1] == 1:
            dis_op = self.opt_ops[0].apply_gradients(grads_list[0])                                 # my_sngan.py:435
            gen_op = self.opt_ops[1].apply_gradients(grads_list[1], global_step=self.global_step)   # my_sngan.py:436
        else:
            raise AttributeError('One of the imbalanced_update must be 1.')                         # my_sngan.py:439
        op_list = [dis_op, gen_op]
        if imbalanced_update is not None:                                                           # graph_func.py:885-886
            global_step_value = int(self.global_step)
            op_list = [op_list[i] for i in range(2) if global_step_value % imbalanced_update[i] == 0]
        update_ops = tf.get_collection(tf.GraphKeys.UPDATE_OPS)                                     # graph_func.py:848
        for op in op_list + update_ops:          # one sess.run: every value was computed from pre-update variables
            op()


def _named_grads(grads_and_vars):
    names = {id(v): k for k, v in tf.shim_variables().items()}
    return {names[id(v)]: (torch.zeros_like(v) if g is None else g).detach().numpy() for g, v in grads_and_vars}


def ref_step_case(loss_type, sn_mode='default'):
    """Twin of make_golden.step_case: same architecture, initial variables, data and codes.  sn_mode 'sn_paper' runs the
    reference with FLAGS.SPECTRAL_NORM_MODE = 'sn_paper' (power iteration on the reshaped kernel matrix, layer_func.py:811-814)."""
    base = mg.step_case(loss_type, sn_mode)
    FLAGS.SPECTRAL_NORM_MODE = sn_mode
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    B = base['data'].shape[0]
    run = ReferenceRun(arch, loss_type, lr_list=(5e-4, 2e-4))
    init = {k.split(':', 1)[1]: v for k, v in base.items() if k.startswith('before:') or k.startswith('state_before:')}
    run.set_variables(init)
    grads_list, (lg, ld) = run.build(base['data'], base['code'], B)
    out = {k: v for k, v in base.items() if k in ('data', 'code') or k.startswith('before:') or k.startswith('state_before:')}
    out['loss_gen'], out['loss_dis'] = lg.detach().numpy(), ld.detach().numpy()
    for k, v in list(_named_grads(grads_list[0]).items()) + list(_named_grads(grads_list[1]).items()):
        out['grad:' + k] = v
    # scores / images are intermediate tensors of __gpu_task__; recompute them with the reference's own nets
    saved_updates = list(tf.get_collection(tf.GraphKeys.UPDATE_OPS))
    gen_batch = run.mdl.Gen({'x': torch.as_tensor(base['code'], dtype=torch.float64)}, is_training=True)
    dis_out = run.mdl.Dis(SNGan.concat_two_batches({'x': torch.as_tensor(base['data'], dtype=torch.float64)}, gen_batch), is_training=True)
    out['x_gen'] = gen_batch['x'].detach().numpy()
    out['scores'] = dis_out['x'].detach().numpy()
    tf._S.collections[tf.GraphKeys.UPDATE_OPS] = saved_updates      # drop the BN update ops of the recomputation
    run.run(grads_list)
    store = tf.shim_variables()
    for k in base:
        if k.startswith('after:') or k.startswith('state_after:'):
            out[k] = store[k.split(':', 1)[1]].detach().numpy()
    assert int(run.global_step) == 1
    FLAGS.SPECTRAL_NORM_MODE = 'default'
    return out


def ref_step_cifar(batch=4, seed=2, steps=2, act_k=None, script='my_test_cifar.py', loss_type='rep', lr_list=(5e-4, 2e-4), stride=SAMPLE):
    """The reference's own architecture dictionary (my_test_cifar.py:12-38 by default; my_test_stl.py, my_test_celebA.py and
    my_test_lsun.py for the other fixtures), `steps` consecutive fused steps.
    act_k: optional override of the script's 64^(1/8) -- a larger multiplier spreads the scores so that the kernel
    differences are O(1) and an fp32 implementation can be compared at 1e-3 (the *_k27 fixture)."""
    arch = reference_architecture(script)
    if act_k is not None:
        for layer in arch['discriminator']:
            layer['act_k'] = act_k
    m = onet.OracleSNGan(arch, loss_type, dtype=torch.float64, seed=seed)        # only for the seeded initial variables
    onet.warm_spectral_norm(m, 6)
    init = {}
    for d in (m.gen_params, m.dis_params, m.gen_state, m.dis_state):
        init.update({k: v.detach().clone() for k, v in d.items()})
    run = ReferenceRun(arch, loss_type, lr_list=lr_list)
    run.set_variables(init)
    out = {'seed': np.asarray(seed), 'batch': np.asarray(batch), 'steps': np.asarray(steps), 'warm': np.asarray(6),
           'act_k': np.asarray(arch['discriminator'][0]['act_k']), 'loss_type': np.asarray(loss_type), 'lr_list': np.asarray(lr_list),
           'sample_stride': np.asarray(stride)}
    for t in range(steps):
        data, code = onet.synthetic_batch(arch, batch, seed=5 + 10 * t, dtype=torch.float32)
        grads_list, (lg, ld) = run.build(data.double(), code.double(), batch)
        out['loss_gen_%d' % t], out['loss_dis_%d' % t] = lg.detach().numpy(), ld.detach().numpy()
        for k, v in list(_named_grads(grads_list[0]).items()) + list(_named_grads(grads_list[1]).items()):
            out['grad_norm_%d:%s' % (t, k)] = np.asarray(np.linalg.norm(v.ravel()))
            out['grad_sample_%d:%s' % (t, k)] = v.ravel()[::stride].copy()
        run.run(grads_list)
        for k, v in tf.shim_variables().items():
            a = v.detach().numpy().ravel()
            out['var_norm_%d:%s' % (t, k)] = np.asarray(np.linalg.norm(a))
            out['var_sample_%d:%s' % (t, k)] = a[::stride].copy()
    return out


def ref_imbalanced_case(tag):
    """Twin of make_golden.imbalanced_case: the reference's op construction for Agent(imbalanced_update=(k_dis, k_gen))
    (my_sngan.py:427-439) and MySession.full_run's op selection per step (graph_func.py:876-908)."""
    base = mg.imbalanced_case(tag)
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    imb = tuple(int(k) for k in base['imbalanced_update'])
    run = ReferenceRun(arch, 'rep', lr_list=(5e-4, 2e-4))
    run.set_variables({k.split(':', 1)[1]: v for k, v in base.items() if k.startswith('before:') or k.startswith('state_before:')})
    out = {k: v for k, v in base.items() if k.split(':')[0].split('_')[0] not in ('after', 'delta', 'state', 'losses') or k.startswith('state_before:')}
    prev = {k[7:]: v for k, v in base.items() if k.startswith('before:')}
    for t in range(int(base['steps'])):
        grads_list, (lg, ld) = run.build(base['data_%d' % t], base['code_%d' % t], base['data_%d' % t].shape[0])
        out['losses_%d' % t] = np.asarray([float(lg.detach()), float(ld.detach())])
        run.run(grads_list, imbalanced_update=imb)
        store = tf.shim_variables()
        for k in prev:
            now = store[k].detach().numpy().copy()
            out['delta_%d:%s' % (t, k)] = np.asarray(np.linalg.norm(now - prev[k]))
            prev[k] = now
    store = tf.shim_variables()
    for k in base:
        if k.startswith('state_after:') or k.startswith('after:'):
            out[k] = store[k.split(':', 1)[1]].detach().numpy().copy()
    out['global_step'] = np.asarray(int(run.global_step))
    return out


def ref_eval_case():
    """Twin of make_golden.eval_case: the reference's eval_sampling graph section (my_sngan.py:523-551) executed eagerly --
    MeshCode.by_sine for the codes (math_func.py:257-291), sample_codes, __gpu_task__(is_training=False), clip_by_value,
    Dis(concat_two_batches(...), is_training=False) -- and its write_sprite (graph_func.py:222-266), whose scipy.misc.imsave
    call (gone from SciPy) is replaced by a capture of the uint8 array it is handed."""
    import types
    base = mg.eval_inputs()
    arch = oa.tiny(channels=(16, 16), size=8, code=16, act_k=2.6)
    run = ReferenceRun(arch, 'rep', lr_list=(5e-4, 2e-4))
    init = {k.split(':', 1)[1]: v for k, v in base.items() if k.startswith('var:') or k.startswith('state:')}
    mdl = run.mdl
    tf._S.collections.clear()
    mdl.init_net()
    mesh_num = tuple(int(m) for m in base['mesh_num'])
    code_x = rmf.MeshCode(mdl.code_size, mesh_num=mesh_num).by_sine(base['z_support'].astype(np.float64), name='code_x')
    orig_gv = tf.get_variable

    def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
        full = '/'.join(tf._S.scope + [name])
        if full not in tf.shim_variables():
            v0 = torch.as_tensor(init[full], dtype=torch.float64)
            initializer = lambda s: v0.clone()                                    # noqa: E731
        return orig_gv(name, shape, dtype, initializer, trainable, **kw)
    tf.get_variable = get_variable
    try:
        code_batch = mdl.sample_codes(mesh_num[0] * mesh_num[1], code_x, None, name='code_te')
        gen_batch = mdl.__gpu_task__(code_batch=code_batch, is_training=False)
        gen_batch['x'] = tf.clip_by_value(gen_batch['x'], clip_value_min=-1, clip_value_max=1)
        data_batch = {'x': torch.as_tensor(base['data'], dtype=torch.float64)}
        dis_out = mdl.Dis(mdl.concat_two_batches(data_batch, gen_batch), is_training=False)
        s_x, s_gen = tf.split(dis_out['x'], num_or_size_splits=2, axis=0)
    finally:
        tf.get_variable = orig_gv
    assert set(init) == set(tf.shim_variables()), set(init) ^ set(tf.shim_variables())
    out = dict(base)
    out['code'] = code_x.detach().numpy()
    out['x_gen'], out['s_x'], out['s_gen'] = gen_batch['x'].detach().numpy(), s_x.detach().numpy(), s_gen.detach().numpy()
    captured = []
    import scipy
    fake = types.ModuleType('scipy.misc')
    fake.imsave = lambda path, arr: captured.append(np.array(arr))
    saved = sys.modules.get('scipy.misc'), getattr(scipy, 'misc', None)
    sys.modules['scipy.misc'], scipy.misc = fake, fake
    try:
        rgf.write_sprite('unused.png', np.transpose(out['x_gen'], (0, 2, 3, 1)), mesh_num=mesh_num, if_invert=False)
        rgf.write_sprite('unused.png', out['x_gen'][:, 0], mesh_num=[3, 2], if_invert=True)
    finally:
        if saved[0] is not None:
            sys.modules['scipy.misc'] = saved[0]
        else:
            del sys.modules['scipy.misc']
        if saved[1] is not None:
            scipy.misc = saved[1]
    out['sprite'], out['sprite_inverted_gray'] = captured
    return out


def main():
    for lt in ('rep', 'rmb', 'mmd_g', 'mgb', 'mmd_t'):
        for b in (2, 3, 64, 128) if lt in ('rep', 'rmb') else (64,):
            np.savez_compressed(os.path.join(HERE, 'ref_mmd_{}_{}.npz'.format(lt, b)), **ref_mmd_case(lt, b))
    np.savez_compressed(os.path.join(HERE, 'ref_mmd_rep_256.npz'), **ref_mmd_case('rep', 256))
    np.savez_compressed(os.path.join(HERE, 'ref_mmd_rmb_w_1_0.npz'), **ref_mmd_case('rmb', 32, w=(1.0, 0.0)))
    cases = {'conv_k3s1_useu': ('c', 8, 16, 6, 3, 1), 'conv_k4s2_nouseu': ('c', 8, 16, 8, 4, 2), 'dense_nouseu': ('d', 64, 16, 1, 1, 1),
             'dense_useu': ('d', 16, 32, 1, 1, 1), 'conv_image': ('c', 3, 8, 8, 3, 1)}
    for i, (name, c) in enumerate(cases.items()):
        np.savez_compressed(os.path.join(HERE, 'ref_sn_{}.npz'.format(name)), **ref_sn_case(*c, seed=10 + i))
    for lt in ('rep', 'rmb'):
        np.savez_compressed(os.path.join(HERE, 'ref_step_tiny_{}.npz'.format(lt)), **ref_step_case(lt))
    np.savez_compressed(os.path.join(HERE, 'ref_step_tiny_rep_pim.npz'), **ref_step_case('rep', sn_mode='sn_paper'))
    np.savez_compressed(os.path.join(HERE, 'ref_eval_tiny.npz'), **ref_eval_case())
    for tag in mg.IMBALANCED:
        np.savez_compressed(os.path.join(HERE, 'ref_step_tiny_imbalanced_{}.npz'.format(tag)), **ref_imbalanced_case(tag))
    np.savez_compressed(os.path.join(HERE, 'ref_step_cifar_rep.npz'), **ref_step_cifar())
    np.savez_compressed(os.path.join(HERE, 'ref_step_cifar_rep_k27.npz'), **ref_step_cifar(batch=8, act_k=2.7))
    # the other shipped architecture dictionaries, parsed from the reference's scripts (one step, batch 2, sparse samples)
    np.savez_compressed(os.path.join(HERE, 'ref_step_stl_rmb.npz'),
                        **ref_step_cifar(batch=2, steps=1, script='my_test_stl.py', loss_type='rmb', lr_list=(2e-4, 2e-4), stride=1031))
    np.savez_compressed(os.path.join(HERE, 'ref_step_celeba_rep.npz'),
                        **ref_step_cifar(batch=2, steps=1, script='my_test_celebA.py', lr_list=(1e-4, 2e-4), stride=4099))
    np.savez_compressed(os.path.join(HERE, 'ref_step_lsun_rep.npz'),
                        **ref_step_cifar(batch=2, steps=1, script='my_test_lsun.py', lr_list=(2e-4, 1e-4), stride=4099))


if __name__ == '__main__':
    main()
