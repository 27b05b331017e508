"""GPU: Agent / SNGan with FLAGS.CKPT_FORMAT = 'tf' -- the engine state goes through TensorFlow's checkpoint container
(GeneralTools/tf_bundle.py) and comes back bit for bit; training resumes from it and eval_sampling restores it.
(File name: runs after the parity suites.)"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tf_container_checkpoint_resume_and_eval(cuda, tmp_path, monkeypatch):
    from oracle import architectures as oa
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools import graph_func as gf
    from mmdgan_b200.GeneralTools import tf_bundle as tb
    from mmdgan_b200.DeepLearning.my_sngan import SNGan
    monkeypatch.setattr(FLAGS, 'DEFAULT_OUT', str(tmp_path) + '/')
    monkeypatch.setattr(FLAGS, 'SILENT_MODE', True)
    monkeypatch.setattr(FLAGS, 'CKPT_FORMAT', 'tf')
    arch = oa.tiny(act_k=2.6)
    images = (np.random.RandomState(0).rand(64, 3, 8, 8) * 255).astype(np.uint8)
    agent = gf.Agent('toy', 'sngan_rep', load_ckpt=True, do_save=True, query_step=100, imbalanced_update=(1, 2))
    mdl = SNGan(arch, num_class=0, loss_type='rep', optimizer='adam', rep_weights=[0.0, -1.0])
    torch.manual_seed(0)
    mdl.training(images, agent, 64, [5e-4, 2e-4], max_step=5, batch_size=16)
    prefix = gf.get_ckpt(agent.ckpt_folder)
    assert prefix == os.path.join(agent.ckpt_folder, 'toy.ckpt-5') and os.path.isfile(prefix + '.index')
    assert os.path.isfile(prefix + '.data-00000-of-00001') and os.path.isfile(os.path.join(agent.ckpt_folder, 'checkpoint'))
    z = tb.read_bundle(prefix)
    eng = mdl.engine
    for i, net in enumerate((eng.D, eng.G)):
        for name in net.var_offsets:
            assert np.array_equal(z[name], net.get_variable(name).cpu().numpy()), name
            off, shape = net.var_offsets[name]
            n = int(np.prod(shape))
            assert np.array_equal(z[name + '/Adam_{}'.format(i)].reshape(-1), net.m[off:off + n].cpu().numpy())
            assert np.array_equal(z[name + '/Adam_{}_1'.format(i)].reshape(-1), net.v[off:off + n].cpu().numpy())
        for name in net.state_names():
            assert np.array_equal(z[name], net.get_state(name).cpu().numpy()), name
    assert int(z['global_step']) == 5 and int(eng.D.step) == 5 and int(eng.G.step) == 3          # gen ran at steps 0, 2, 4
    assert np.isclose(z['beta2_power'], 0.999 ** 6) and np.isclose(z['beta2_power_1'], 0.999 ** 4)
    # resume in a fresh model: update counters recovered from the beta powers
    mdl2 = SNGan(arch, num_class=0, loss_type='rep', optimizer='adam', rep_weights=[0.0, -1.0])
    mdl2.training(images, agent, 64, [5e-4, 2e-4], max_step=2, batch_size=16)
    assert mdl2.global_step == 7 and int(mdl2.engine.D.step) == 7 and int(mdl2.engine.G.step) == 4
    assert sorted(f for f in os.listdir(agent.ckpt_folder) if f.endswith('.index')) == ['toy.ckpt-5.index', 'toy.ckpt-7.index']
    # eval_sampling restores the named bundle
    out = mdl2.eval_sampling('toy', 'sngan_rep', mesh_num=(3, 4), ckpt_file='toy.ckpt-5', do_sprite=False)
    assert out['global_step'] == 5 and out['x_gen'].shape == (12, 3, 8, 8) and np.isfinite(out['x_gen']).all()


def test_two_rank_step_with_fused_nvls_allreduce_adam(cuda):
    """The 2-rank equalities of tests/test_gpu_multi.py with the gradient all-reduce + Adam fused into one NVSwitch-multicast
    kernel (MMDGAN_NVLS_ADAM=1, csrc/nvls.cu).  The path is opt-in and was written without access to a multi-GPU box, so the
    test is opt-in too (MMDGAN_TEST_NVLS=1) until it has been run once."""
    import subprocess
    import sys
    if os.environ.get('MMDGAN_TEST_NVLS') != '1':
        pytest.skip('opt-in: set MMDGAN_TEST_NVLS=1 (needs 2 GPUs behind an NVSwitch)')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29900 + os.getpid() % 300), os.path.join(root, 'scripts', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, MMDGAN_NVLS_ADAM='1'))
    if 'NVLS_UNAVAILABLE' in out.stdout:
        pytest.skip('no NVSwitch multicast on this box')
    assert 'MULTI_GPU_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
