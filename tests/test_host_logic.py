"""CPU: host-side logic of the product -- layer DSL mirror, kernel launch planning, C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from oracle import architectures as oa
from oracle import net as onet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _routine(arch, key):
    from mmdgan_b200.GeneralTools.layer_func import Net, Routine
    net = Net(arch[key], key[:3], 'channels_first')
    r = Routine(net)
    r.add_input_layers([64, arch['code'][0][0]] if key == 'generator' else [64] + list(arch['input'][0]), [0])
    r.seq_links(list(range(net.num_layers)))
    r.add_output_layers([net.num_layers - 1])
    return r


@pytest.mark.parametrize('name', ['cifar', 'stl', 'celeba', 'lsun'])
def test_layer_dsl_matches_oracle_shape_inference_and_sn_routing(name):
    arch = oa.ARCHITECTURES[name]()
    for key, scope, in_shape in (('generator', 'gen', [arch['code'][0][0]]), ('discriminator', 'dis', list(arch['input'][0]))):
        layers = _routine(arch, key).ordered_layers()
        specs = onet.build_net(arch[key], scope, in_shape)
        assert len(layers) == len(specs)
        for ly, sp in zip(layers, specs):
            assert ly.kernel_shape == sp.kernel_shape
            assert ly.op_output_shape[1:] == sp.op_out_shape and ly.output_shape[1:] == sp.out_shape
            assert ly.kernel_name == sp.kernel_name
            assert ('bias' in ly.ops) == sp.has_bias and ('BN' in ly.ops) == sp.has_bn
            if sp.has_sn:
                assert ly.use_u is sp.use_u and ly.sn_x_shape == sp.x_shape      # integer routing decision: exact
                assert ly.sn_name == sp.sn_name


def test_product_experiment_definitions_match_the_oracle_twin():
    from mmdgan_b200 import experiments as ex
    for name in ['cifar', 'stl', 'celeba', 'lsun']:
        assert ex.ARCHITECTURES[name]() == oa.ARCHITECTURES[name]()
    assert ex.tiny() == oa.tiny()


def test_update_layer_design_defaults_and_errors():
    from mmdgan_b200.GeneralTools.layer_func import update_layer_design, Net, Routine
    d = update_layer_design({'name': 'l', 'out': 8})
    assert (d['op'], d['kernel'], d['strides'], d['padding'], d['bias'], d['act'], d['act_k']) == ('c', 3, 1, 'SAME', 'b', 'linear', False)
    assert update_layer_design({'name': 'l', 'out': 8, 'act_nm': 'bn'})['bias'] is None          # layer_func.py:1241-1242
    assert 'kernel' not in update_layer_design({'name': 'l', 'out': 8, 'op': 'd'})
    with pytest.raises(AttributeError):
        update_layer_design({'name': 'l', 'out': 8, 'op': 'zz'})                                # layer_func.py:1275
    r = Routine(Net([{'name': 'a', 'out': 8, 'type': 'res'}], 'n', 'channels_first'))
    with pytest.raises(NotImplementedError):
        r.add_input_layers([64, 3, 8, 8], [0])                                                  # layer_func.py:2067
    r = Routine(Net([{'name': 'a', 'out': 8}], 'n', 'channels_first'))
    r.add_input_layers([64, 3, 8, 8], [0])
    with pytest.raises(AttributeError):
        r.add_input_layers([64, 3, 8, 8], [0])                                                  # already added
    with pytest.raises(NotImplementedError):
        r({'x': torch.zeros(1)})                                                                # output layer not defined


def test_linear_op_launch_plans():
    from mmdgan_b200 import kernels as K
    # D l3 of the CIFAR net: 128 -> 128 k3 s1 on 16x16
    lop = K.LinearOp('c', [128, 16, 16], [128, 16, 16], 3, 1, device='cpu')
    assert (lop.f['kpad'], lop.f['bn'], lop.f['rows_pad'], lop.f['classes']) == (1152, 128, 128, 1)
    assert lop._fwd_geom()['cls'] == [(-1, -1, 0, 0)] and lop._dgrad_geom()['cls'] == [(-1, -1, 0, 0)]
    R, NC, bn, splits, P = lop.wgrad_plan(512)
    assert (R, NC, bn, P) == (128, 1152, 256, 512 * 256) and 1 <= splits <= P // 32
    # stride-2 conv: input gradient = four parity classes of 2x2 taps
    lop = K.LinearOp('c', [64, 32, 32], [128, 16, 16], 4, 2, device='cpu')
    assert lop.d['classes'] == 4 and lop.d['taps'] == 4 and lop.d['kpad'] == 4 * 128
    assert lop._dgrad_geom()['cls'] == [(-1, -1, 0, 0), (-1, 0, 0, 1), (0, -1, 1, 0), (0, 0, 1, 1)]
    # image layers: 3 channels padded to 8 (one 16-byte unit of bf16), weight gradient puts the small side on N
    lop = K.LinearOp('c', [64, 32, 32], [3, 32, 32], 3, 1, device='cpu')
    # round 2: both directions of the 3-channel layers run as dense products over 27 = 9 taps x 3 channels (im2col27 / tapsum27 + the
    # tensor-core GEMM): the weight gradient is the dense one, 8 * 1024 pixels contracted
    assert lop.Cs_out == 8 and lop.w_swapped and lop.f['bn'] == 16 and lop.wgrad_plan(8)[:3] == (64, 72, 128)
    # round 2: the many -> few direction (here the forward pass) runs as a dense [64 -> 27] product + the tap-sum kernel
    assert lop.img_op is not None and not lop.img_few_in and lop.img_op.f['ncols'] == 32
    lop = K.LinearOp('c', [3, 32, 32], [64, 32, 32], 3, 1, device='cpu')
    assert lop.Cs_in == 8 and not lop.w_swapped and lop.f['kpad'] == 128      # 9 taps x 8 channels = 72 -> whole 64-element K blocks
    assert lop.img_few_in and lop.img_op.d['ncols'] == 32       # its input gradient is the many -> few direction
    # operand planes: forward weights as two fp16 planes (three plane-pair products are fp32-grade), input-gradient weights as
    # three bf16 planes (three products for gradients, six when the adjoint acts as a forward operator in the spectral norm)
    assert lop.f['w'].dtype == torch.float16 and lop.f['w'].shape[0] == 2 and lop.d['w'].dtype == torch.bfloat16 and lop.d['w'].shape[0] == 3
    assert (lop.fwd_npass, lop.bwd_npass, lop.adj_npass) == (3, 3, 6)
    assert [K.pad_c(c) for c in (1, 3, 8, 9, 16, 17, 32, 33, 64, 65, 100)] == [8, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128]
    # dense with NCHW-flatten permutation folded into the packed weights
    lop = K.LinearOp('d', [8192], [16], in_flat=(512, 16), device='cpu')
    assert lop.w_swapped and lop.in_flat == (512, 16) and lop.d['kpad'] == 64
    with pytest.raises(NotImplementedError):
        K.LinearOp('c', [8, 8, 8], [8, 8, 8], 5, 1, device='cpu')
    with pytest.raises(AttributeError):
        K.LinearOp('sc', [8, 8, 8], [8, 8, 8], 3, 1, device='cpu')


def test_spectral_norm_routing_pico_and_pim(monkeypatch):
    """use_u / in_rand shapes: PICO compares the operator's input and output sizes (math_func.py:497-515), PIM
    ('sn_paper') reshapes the kernel to [k*k*Cin, Cout] and follows the dense rule (layer_func.py:811-814, math_func.py:477-486)."""
    from mmdgan_b200.GeneralTools.layer_func import Net, Routine
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200 import experiments as ex

    def shapes():
        r = Routine(Net(ex.cifar()['discriminator'], 'dis', 'channels_first'))
        r.add_input_layers([64, 3, 32, 32], [0])
        r.seq_links(list(range(8)))
        r.add_output_layers([7])
        return [(ly.use_u, ly.sn_x_shape, ly.sn_pim) for ly in r.ordered_layers()]
    pico = shapes()
    assert [p[0] for p in pico] == [True, False, True, False, True, False, True, False]
    assert pico[0][1] == [1, 3, 32, 32] and pico[1][1] == [1, 128, 16, 16] and pico[7][1] == [1, 16] and not any(p[2] for p in pico)
    monkeypatch.setattr(FLAGS, 'SPECTRAL_NORM_MODE', 'sn_paper')
    pim = shapes()
    assert [p[1] for p in pim] == [[1, 27], [1, 128], [1, 128], [1, 256], [1, 256], [1, 512], [1, 512], [1, 16]]
    assert [p[0] for p in pim] == [True] + [False] * 7 and [p[2] for p in pim] == [True] * 7 + [False]
    monkeypatch.setattr(FLAGS, 'SPECTRAL_NORM_MODE', 'no_such_mode')
    with pytest.raises(NotImplementedError):
        shapes()


def test_cabi_exports_every_declared_symbol():
    from mmdgan_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'mmdgan_b200.h')).read()
    declared = set(re.findall(r'\b(mmdgan_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    handle = _lib.load()
    assert handle.mmdgan_version() >= 100


def test_cabi_argument_validation_without_a_gpu():
    from mmdgan_b200 import _lib
    lib = _lib.load()
    d = _lib.MmdDesc()
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'rep', 0.0, -1.0) == 0
    assert list(d.cD) == [-1.0, 0.0, 1.0] and list(d.bmode) == [0, 0, 0] and d.n_sigma == 1
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'rmb', 0.0, -1.0) == 0
    assert list(d.bmode) == [1, 0, 2] and list(d.bval) == [0.25, 0.0, 4.0]
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'rmb', 1.0, 0.0) == 0 and list(d.bmode) == [1, 0, 2]
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'rmb', 2.0, 1.0) == 0 and list(d.bmode) == [1, 0, 1]
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'mmd_g', 0.0, 0.0) == 0 and d.n_sigma == 5 and d.family == 0
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'mmd_t', 0.0, 0.0) == 0 and d.n_sigma == 5 and d.family == 1 and d.beta == 2.0
    assert [round(v, 3) for v in d.sigma[:5]] == [0.2, 0.5, 1.0, 2.0, 5.0] and list(d.cD) == [-1.0, 2.0, -1.0]
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'rep', 0.0, 0.0) == _lib.MMDGAN_EINVAL
    assert b'w[0]-w[1] must be 1' in lib.mmdgan_last_error()
    assert lib.mmdgan_mmd_configure(ctypes.byref(d), b'hinge', 0.0, -1.0) == _lib.MMDGAN_EINVAL
    g = _lib.GemmDesc()
    assert lib.mmdgan_gather_gemm(ctypes.byref(g), None) == _lib.MMDGAN_EINVAL          # null pointers: rejected before any CUDA call
    assert lib.mmdgan_mmd_fwd_bwd(ctypes.byref(d), None) == _lib.MMDGAN_EINVAL
    assert lib.mmdgan_adam(None, None, None, None, 10, 1e-3, 0.5, 0.999, 1e-8, None, None) == _lib.MMDGAN_EINVAL
    # fused NVSwitch all-reduce + Adam (csrc/nvls.cu): argument validation happens before any CUDA call
    nv = lib.mmdgan_adam_allreduce_nvls
    assert nv(None, None, None, None, None, None, None, 0, 16, 1e-3, 0.5, 0.999, 1e-8, None, None) == _lib.MMDGAN_EINVAL
    fake16, step = ctypes.c_void_p(4096), ctypes.c_void_p(8192)
    assert nv(fake16, fake16, fake16, fake16, fake16, fake16, fake16, 2, 16, 1e-3, 0.5, 0.999, 1e-8, step, None) == _lib.MMDGAN_ESHAPE
    assert b'whole float4 groups' in lib.mmdgan_last_error()
    assert nv(fake16, fake16, fake16, ctypes.c_void_p(4100), fake16, fake16, fake16, 0, 16, 1e-3, 0.5, 0.999, 1e-8, step, None) == _lib.MMDGAN_ESHAPE
    assert nv(fake16, fake16, fake16, fake16, fake16, fake16, fake16, 16, 16, 1e-3, 0.5, 0.999, 1e-8, step, None) == _lib.MMDGAN_OK   # empty shard
    sc = lib.mmdgan_scatter_scores_nvls
    assert sc(None, 8, 16, 0, fake16, fake16, None) == _lib.MMDGAN_EINVAL
    assert sc(fake16, 8, 6, 0, fake16, fake16, None) == _lib.MMDGAN_ESHAPE and sc(fake16, 0, 16, 0, fake16, fake16, None) == _lib.MMDGAN_ESHAPE
    assert sc(fake16, 8, 16, 1, ctypes.c_void_p(4104), fake16, None) == _lib.MMDGAN_ESHAPE
    ar = lib.mmdgan_allreduce_small_nvls
    assert ar(None, fake16, 8, None) == _lib.MMDGAN_EINVAL and ar(fake16, fake16, 6, None) == _lib.MMDGAN_ESHAPE
    assert ar(fake16, fake16, 0, None) == _lib.MMDGAN_OK
    assert lib.mmdgan_mmd_workspace(256) >= 64 * 6 * 4


def test_cabi_plane_format_validation_without_a_gpu():
    """One MMA cannot mix an fp16 operand with a bf16 one (measured on the B200: illegal instruction), and the six-product mode
    is defined for three bf16 planes only: the C ABI rejects such descriptors before any CUDA call."""
    from mmdgan_b200 import _lib
    lib = _lib.load()
    g = _lib.GemmDesc()
    fake = 0x7f0000001000          # never dereferenced: validation fails first
    g.src, g.w, g.dst = fake, fake, fake
    g.src_plane = g.w_plane = g.dst_plane = 1 << 20
    g.Nimg, g.Hs, g.Ws, g.Cs, g.Hg, g.Wg, g.sy, g.sx, g.TH, g.TW = 2, 8, 8, 64, 8, 8, 1, 1, 3, 3
    g.kpad, g.classes, g.w_rows = 576, 1, 128
    g.Hd, g.Wd, g.Cd, g.osy, g.osx, g.Ncols = 8, 8, 128, 1, 1, 128
    g.out_mode, g.dst_npl, g.bn, g.npass = 0, 2, 128, 3
    g.src_fmt, g.w_fmt = _lib_fmt('F16A'), _lib_fmt('BF16')
    assert lib.mmdgan_gather_gemm(ctypes.byref(g), None) == _lib.MMDGAN_EINVAL
    assert b'both operands must be' in lib.mmdgan_last_error()
    g.src_fmt, g.w_fmt, g.npass = _lib_fmt('F16A'), _lib_fmt('F16W'), 6
    assert lib.mmdgan_gather_gemm(ctypes.byref(g), None) == _lib.MMDGAN_EINVAL
    assert b'npass 6 is the bf16 three-plane mode' in lib.mmdgan_last_error()
    w = _lib.WgradDesc()
    w.plain, w.g, w.out = fake, fake, fake
    w.plain_plane = w.g_plane = 1 << 20
    w.P, w.Cp, w.Nimg, w.Hs, w.Ws, w.Cs, w.Hg, w.Wg, w.sy, w.sx, w.TH, w.TW, w.splits = 128, 64, 2, 8, 8, 64, 8, 8, 1, 1, 3, 3, 1
    w.bn, w.npass, w.p_fmt, w.g_fmt = 128, 3, _lib_fmt('BF16'), _lib_fmt('F16A')
    assert lib.mmdgan_wgrad_gemm(ctypes.byref(w), None) == _lib.MMDGAN_EINVAL
    assert b'both operands must be' in lib.mmdgan_last_error()
    # fp16 formats carry at most two planes
    assert lib.mmdgan_to_planes(fake, fake, 1 << 20, 3, _lib_fmt('F16A'), 1024, None) == _lib.MMDGAN_ESHAPE


def _lib_fmt(name):
    return {'BF16': 0, 'F16A': 1, 'F16W': 2}[name]


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: a CPU tensor reaching a kernel wrapper raises instead of silently computing."""
    from mmdgan_b200 import kernels as K
    from mmdgan_b200._lib import MmdganError
    with pytest.raises(MmdganError):
        K.planes_value(torch.zeros(2, 4, 8, dtype=torch.bfloat16))
    from mmdgan_b200.GeneralTools.math_func import GANLoss
    with pytest.raises(RuntimeError):
        GANLoss().apply(torch.zeros(4, 16), torch.zeros(4, 16), 'rep', batch_size=4, d=16)
    with pytest.raises(NotImplementedError):
        GANLoss().apply(torch.zeros(4, 16), torch.zeros(4, 16), 'hinge', batch_size=4)


def test_reference_style_matrix_functions_match_oracle():
    from mmdgan_b200.GeneralTools import math_func as mf
    from oracle import mmd as omm
    g, r = torch.randn(6, 5, dtype=torch.float64), torch.randn(6, 5, dtype=torch.float64)
    a, b = mf.get_squared_dist(g, r), omm.get_squared_dist(g, r)
    for x, y in zip(a, b):
        assert torch.allclose(x, y)
    assert torch.allclose(torch.stack(mf.mmd_g(*a, 6, custom_weights=[0.0, -1.0])), torch.stack(omm.mmd_g(*b, 6, custom_weights=[0.0, -1.0])))
    assert torch.allclose(torch.stack(mf.mmd_g_bounded(*a, 6, lower_bound=0.25, upper_bound=4.0, custom_weights=[0.0, -1.0])),
                          torch.stack(omm.mmd_g_bounded(*b, 6, lower_bound=0.25, upper_bound=4.0, custom_weights=[0.0, -1.0])))
    assert mf.spatial_shape_after_conv([32, 32], 4, 2, 1, 'SAME') == [16, 16]
    assert mf.spatial_shape_after_transpose_conv(6, 4, 2, 1, 'SAME') == 12
    with pytest.raises(AttributeError):
        mf.get_squared_dist(g, r, mode='zz')


def test_mesh_codes_and_sprites_against_reference_fixture(tmp_path):
    """Host side of eval_sampling: MeshCode (math_func.py:220-335) and write_sprite / write_sprite_wrapper
    (graph_func.py:222-298) against the reference-executed fixture; no device work."""
    import numpy as np
    from PIL import Image
    from mmdgan_b200.GeneralTools.math_func import MeshCode
    from mmdgan_b200.GeneralTools.graph_func import sprite_array, write_sprite_wrapper
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_eval_tiny.npz'))
    code = MeshCode(16, mesh_num=(2, 3)).by_sine(z['z_support'])
    assert code.dtype == torch.float32 and np.allclose(code.numpy(), z['code'], atol=1e-6)
    nhwc = np.transpose(z['x_gen'], (0, 2, 3, 1))
    assert np.array_equal(sprite_array(nhwc, (2, 3)), z['sprite'])
    assert np.array_equal(sprite_array(z['x_gen'][:, 0], [3, 2], if_invert=True), z['sprite_inverted_gray'])
    # mesh_num None: smallest square, padded with black (graph_func.py:249-256)
    sq = sprite_array(nhwc[:5])
    assert sq.shape == (24, 24, 3) and not sq[16:, 8:].any() and np.array_equal(sq[:8, :24], z['sprite'][:8])
    # the wrapper takes NCHW, names the file <filename><index>.png and keeps an existing file
    path = write_sprite_wrapper(z['x_gen'], (2, 3), ['toy', 'other'], file_folder=str(tmp_path), file_index='_g_x_1_0', image_format='NCHW')
    assert path == os.path.join(str(tmp_path), 'toy_g_x_1_0.png')
    assert np.array_equal(np.asarray(Image.open(path)), z['sprite'])
    with pytest.warns(UserWarning, match='already exists'):
        write_sprite_wrapper(1.0 - z['x_gen'], (2, 3), 'toy', file_folder=str(tmp_path), file_index='_g_x_1_0', image_format='NCHW')
    assert np.array_equal(np.asarray(Image.open(path)), z['sprite'])
    # the other mesh modes
    mc = MeshCode(16, mesh_num=(4, 5))
    assert tuple(mc.get_batch('random').shape) == (20, 16) and tuple(mc.get_batch(1).shape) == (20, 16)
    f = mc.get_batch(2).reshape(4, 5, 16)
    cols = [int(f[i].abs().sum(0).argmax()) for i in range(4)]
    assert len(set(cols)) == 4
    for i, c in enumerate(cols):
        assert np.allclose(f[i, :, c].numpy(), np.linspace(-2.0, 2.0, 5)) and int((f[i] != 0).sum()) == 4    # 0 is on the grid
    with pytest.raises(AttributeError):
        mc.get_batch(3)
    g = MeshCode(2, mesh_num=(2, 3)).simple_grid()
    assert np.allclose(g, [[-1, -1], [-1, 0], [-1, 1], [1, -1], [1, 0], [1, 1]])
    with pytest.raises(AttributeError):
        mc.simple_grid()


def test_checkpoint_folder_helpers(tmp_path, monkeypatch):
    """prepare_folder (graph_func.py:161-180), get_ckpt (399-416: latest by global step, or the named file) and rollback's
    error (633) -- host logic only, no engine."""
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools.graph_func import prepare_folder, get_ckpt, rollback
    monkeypatch.setattr(FLAGS, 'DEFAULT_OUT', str(tmp_path) + '/')
    ckpt_folder, summary_folder, save_path = prepare_folder('cifar', sub_folder='sngan_rep')
    assert ckpt_folder == os.path.join(str(tmp_path), 'cifar_ckpt', 'sngan_rep') and os.path.isdir(ckpt_folder)
    assert summary_folder == os.path.join(str(tmp_path), 'cifar_log', 'sngan_rep') and os.path.isdir(summary_folder)
    assert save_path == os.path.join(ckpt_folder, 'cifar.ckpt')
    assert not os.path.isdir(prepare_folder('other', set_folder=False)[0])
    assert get_ckpt(ckpt_folder) is None and get_ckpt(os.path.join(str(tmp_path), 'missing')) is None
    for step in (5, 12, 100, 9):                     # numeric, not lexicographic, order
        open('{}-{}.npz'.format(save_path, step), 'wb').close()
    open(os.path.join(ckpt_folder, 'notes.txt'), 'w').close()
    assert get_ckpt(ckpt_folder) == save_path + '-100.npz'
    assert get_ckpt(ckpt_folder, 'cifar.ckpt-9.npz') == save_path + '-9.npz'
    assert get_ckpt(ckpt_folder, 'cifar.ckpt-7.npz') is None
    with pytest.raises(FileNotFoundError, match='No ckpt Model found'):
        rollback(None, os.path.join(str(tmp_path), 'missing'))


def test_nvls_shard_ranges_cover_the_flat_buffer():
    """parallel.shard_range: the per-rank shards of the fused all-reduce + Adam kernel are disjoint whole-float4 ranges that
    tile [0, n_flat) for every world size, also when n_flat / 4 is not a multiple of it."""
    from mmdgan_b200.parallel import shard_range
    for n_flat in (64, 64 * 3, 64 * 61, 6000000 // 64 * 64):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for rank in range(world):
                b, e = shard_range(n_flat, rank, world)
                assert b == pos and e >= b and b % 4 == 0 and e % 4 == 0
                pos = e
            assert pos == n_flat
    with pytest.raises(AssertionError):
        shard_range(66, 0, 2)


def test_refresh_job_table_partitions_the_flat_grid():
    """mmdgan_refresh: job j owns the blocks [block_start_j, block_start_{j+1}) of one flat grid (the kernel looks its job up from
    this table); the table is built on the host once per network.  Vector jobs and the split-K reduction tables carry device
    pointers and are exercised by the GPU step tests."""
    from mmdgan_b200 import _lib
    from mmdgan_b200 import kernels as K
    lib = _lib.load()
    packs = []
    for rows_pad, kpad, classes in ((128, 1152, 1), (64, 512, 4), (16, 64, 1), (4096, 8192, 1)):
        d = _lib.PackDesc()
        d.rows_pad, d.kpad, d.classes = rows_pad, kpad, classes
        packs.append(d)
    blob, n, total = K.build_refresh_jobs(packs, [], 'cpu')
    arr = (_lib.RefreshJob * n).from_buffer_copy(blob.numpy().tobytes())
    starts = [arr[i].block_start for i in range(n)]
    # one block per 32 x 64 tile of a packed operand, at most 2048 per job (larger operands stride)
    sizes = [4 * 18, 8 * 8, 1 * 1, 2048]
    assert n == 4 and starts == [sum(sizes[:i]) for i in range(4)] and total == sum(sizes)
    assert [arr[i].kind for i in range(n)] == [0, 0, 0, 0] and arr[1].pack.classes == 4 and arr[3].pack.kpad == 8192
    with pytest.raises(K._lib.MmdganError):          # a vector job needs device memory: the builder refuses a host tensor
        K.build_refresh_jobs(packs, [(torch.zeros(8), torch.zeros(8), 8, 0, 0)], 'cpu')
    # a table the kernel cannot index (more jobs than threads of a block) or an empty grid is refused without touching the device
    assert lib.mmdgan_refresh(ctypes.c_void_p(blob.data_ptr()), 257, 10, None) == _lib.MMDGAN_ESHAPE
    assert lib.mmdgan_refresh(ctypes.c_void_p(blob.data_ptr()), n, 0, None) == _lib.MMDGAN_ESHAPE
    assert lib.mmdgan_refresh(None, n, total, None) == _lib.MMDGAN_EINVAL
