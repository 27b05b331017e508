"""End-to-end parity of the fused SNGan step on the GPU against the CPU oracle (float64) on identical
inputs / weights / state.  Tolerance: 1e-3 normwise relative (the north-star bar) for losses, scores, every gradient
tensor, the spectral-norm and batch-norm state; observed errors of the parity mode (two 16-bit planes per operand, fp16
forward / bf16 gradients, three plane-pair products) are ~1e-5."""
import numpy as np
import pytest
import torch

from oracle import architectures as oa
from oracle import net as onet

pytestmark = pytest.mark.gpu

TOL = 1e-3


def rel(a, b):
    a = torch.as_tensor(a).double().cpu().reshape(-1)
    b = torch.as_tensor(b).double().cpu().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def make_pair(arch, B, loss_type, rep_weights=(0.0, -1.0), seed=3, warm=6, npass=3, use_graph=False, sn_mode='default'):
    """Oracle (float64) and engine holding the same variables and state."""
    from mmdgan_b200.engine import SNGanEngine
    orc = onet.OracleSNGan(arch, loss_type, rep_weights=rep_weights, dtype=torch.float64, seed=seed, sn_mode=sn_mode)
    if warm:
        onet.warm_spectral_norm(orc, warm)
    eng = SNGanEngine(arch, B, loss_type=loss_type, rep_weights=rep_weights, npass=npass, use_graph=use_graph)
    for net, params, state in ((eng.G, orc.gen_params, orc.gen_state), (eng.D, orc.dis_params, orc.dis_state)):
        assert list(net.var_offsets.keys()) == list(params.keys())
        for k, v in params.items():
            net.set_variable(k, v)
        for k, v in state.items():
            net.set_state(k, v)
        net.refresh()
    return orc, eng


def engine_activation(eng, net, L, nimg):
    """Activation of one layer in the oracle's layout ([n, C, H, W], or [n, F] in NCHW-flatten feature order)."""
    from mmdgan_b200 import kernels as K
    if L.op == 'd':
        c, hw = L.lop.out_flat
        return K.planes_value(L.a).reshape(nimg, hw, L.Cs_out // hw if hw > 1 else L.Cs_out)[:, :, :c].permute(0, 2, 1).reshape(nimg, -1).cpu()
    return K.planes_to_nchw(L.a, nimg, L.Cout, L.Hout, L.Wout).cpu()


def kink_masks(eng, B):
    """Sign pattern of every relu / lrelu activation of the engine's forward pass (see oracle.net._KinkWithMask)."""
    masks = {}
    for net, nimg in ((eng.G, B), (eng.D, 2 * B)):
        for L in net.layers:
            if L.act in ('relu', 'lrelu'):
                masks[L.ly.layer_scope] = engine_activation(eng, net, L, nimg) > 0
    return masks


def check_step(orc, eng, arch, B, seed, tol=TOL, degenerate=False, loss_atol=None, check_state=False):
    """One forward + loss + backward of the engine against the float64 oracle on the same batch.  check_state=True also runs
    the update phase and compares the UPDATE_OPS results (batch-norm moving statistics, every spectral-norm in_rand)."""
    data, code = onet.synthetic_batch(arch, B, seed=seed, dtype=torch.float64)
    eng.stage(data.float().cuda(), code.float().cuda())
    eng._phase_forward()
    eng._phase_loss()
    eng._phase_backward()
    torch.cuda.synchronize()
    col = {}
    lg, ld, gg, gd, ug, ud = orc.grads(data, code, col)
    # activations within fp32 rounding noise of a relu / lrelu kink may fall on either side: count them and let the
    # oracle differentiate on the engine's side of the tie (the forward values themselves are compared strictly)
    masks = kink_masks(eng, B)
    flips = 0
    for scope, m in masks.items():
        ref = col[scope + '/out'].detach().reshape(m.shape)
        differ = (ref > 0) != m
        flips += int(differ.sum())
        assert float(ref[differ].abs().max()) < 0.1 * tol * float(ref.abs().max()) if differ.any() else True, scope
    assert flips <= (tol / TOL) * 1e-5 * sum(m.numel() for m in masks.values()) + 2
    if flips:
        col = {}
        lg, ld, gg, gd, ug, ud = orc.grads(data, code, col, act_masks=masks)
    losses = eng.losses().cpu()
    scores = eng.D.layers[-1].a[0].cpu()
    s_ref = torch.cat([col['s_x'], col['s_gen']], 0).detach()
    assert rel(scores, s_ref) < tol
    # the kernel means are O(1): 2e-6 is the fp32 resolution of their differences
    atol = loss_atol if loss_atol is not None else (2e-6 if degenerate else 1e-7)
    assert abs(float(losses[0]) - float(lg)) <= tol * abs(float(lg)) + atol
    assert abs(float(losses[1]) - float(ld)) <= tol * abs(float(ld)) + atol
    x_gen = eng.generate(code.float().cuda()).cpu()
    assert rel(x_gen, col['x_gen'].detach()) < tol
    gmax = max(float(v.norm()) for v in gd.values())
    for name, ref in ([] if degenerate else list(gd.items()) + list(gg.items())):
        net = eng.D if name.startswith('dis/') else eng.G
        got = net.get_grad(name).cpu()
        if float(ref.norm()) < 1e-6 * gmax:       # e.g. the last bias of D: exactly zero by translation invariance
            assert float(got.double().norm()) < 1e-4 * gmax, name
        else:
            assert rel(got, ref) < tol, (name, rel(got, ref))
    # UPDATE_OPS: spectral-norm in_rand and sigma
    for L in eng.D.layers:
        if L.has_sn:
            assert abs(float(L.sigma) - float(col[L.ly.layer_scope + '/sigma'].detach())) < tol * float(col[L.ly.layer_scope + '/sigma'].detach())
    if check_state:
        eng._phase_update()
        torch.cuda.synchronize()
        for net, upd in ((eng.G, ug), (eng.D, ud)):
            assert set(upd.keys()) == set(net.state_names())
            for name, ref in upd.items():
                assert rel(net.get_state(name), ref.detach()) < tol, (name, rel(net.get_state(name), ref.detach()))
    return lg, ld, ug, ud


@pytest.mark.parametrize('loss_type', ['rep', 'rmb'])
def test_step_parity_tiny(cuda, loss_type):
    arch = oa.tiny(act_k=2.6)
    B = 16
    orc, eng = make_pair(arch, B, loss_type)
    check_step(orc, eng, arch, B, seed=5)


def test_step_parity_pim_spectral_norm_mode(cuda, monkeypatch):
    """FLAGS.SPECTRAL_NORM_MODE = 'sn_paper' (PIM, layer_func.py:811-814): the power iteration runs on the conv kernel reshaped
    to the [k*k*Cin, Cout] matrix instead of the conv operator; routing (use_u) and in_rand shapes follow the dense rule."""
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    monkeypatch.setattr(FLAGS, 'SPECTRAL_NORM_MODE', 'sn_paper')
    arch = oa.tiny(act_k=2.6)
    B = 16
    orc, eng = make_pair(arch, B, 'rep', sn_mode='sn_paper')
    assert all(L.ly.sn_pim for L in eng.D.layers if L.op != 'd')
    check_step(orc, eng, arch, B, seed=5)
    arch = oa.cifar(act_k=2.7)
    orc, eng = make_pair(arch, 4, 'rmb', sn_mode='sn_paper')
    check_step(orc, eng, arch, 4, seed=6)


def test_step_parity_tiny_first_step_unnormalised_in_rand(cuda):
    """The reference's very first step runs with the un-normalised in_rand (math_func.py:565-567): sigma is ~||x0||
    times too large, every score collapses onto the bias and the losses are differences of ~1.0 kernel means at the
    fp32 resolution limit.  Scores, sigma and losses (absolute 2e-6) are checked; the gradients of that step are
    pure cancellation noise in fp32 (in the reference too) and are not compared."""
    arch = oa.tiny(act_k=2.6)
    orc, eng = make_pair(arch, 8, 'rep', warm=0)
    check_step(orc, eng, arch, 8, seed=9, degenerate=True)


def test_step_parity_cifar_architecture(cuda):
    """my_test_cifar.py network (8 SN layers, 8192-feature flatten), small batch; act_k raised so that the pairwise
    distances are O(1) and the comparison is well conditioned."""
    arch = oa.cifar(act_k=2.7)
    B = 8
    orc, eng = make_pair(arch, B, 'rep')
    check_step(orc, eng, arch, B, seed=11)


def test_step_parity_stl_architecture_rmb(cuda):
    """my_test_stl.py network: 48x48 (non power-of-two rows), dense + batch norm first generator layer, rmb loss."""
    arch = oa.stl(act_k=2.7)
    B = 4
    orc, eng = make_pair(arch, B, 'rmb')
    check_step(orc, eng, arch, B, seed=13)


def test_step_parity_celeba_lsun_architecture(cuda):
    """my_test_celebA.py / my_test_lsun.py network: 64x64, 10 SN layers, 1024 channels, 16384-feature flatten."""
    arch = oa.celeba(act_k=2.3)
    B = 2
    orc, eng = make_pair(arch, B, 'rep')
    check_step(orc, eng, arch, B, seed=17)


def test_three_full_steps_and_state(cuda):
    """Three simultaneous G/D updates: loss trajectory, Adam-updated variables, BN moving statistics, in_rand."""
    arch = oa.tiny(act_k=2.6)
    B = 16
    orc, eng = make_pair(arch, B, 'rep')
    init = {k: v.clone() for k, v in list(orc.gen_params.items()) + list(orc.dis_params.items())}
    for it in range(3):
        data, code = onet.synthetic_batch(arch, B, seed=20 + it, dtype=torch.float64)
        lg_o, ld_o = orc.step(data, code)
        lg, ld = eng.step(data.float(), code.float())
        assert abs(lg - lg_o) <= 2e-3 * abs(lg_o) + 1e-6, (it, lg, lg_o)
        assert abs(ld - ld_o) <= 2e-3 * abs(ld_o) + 1e-6, (it, ld, ld_o)
    assert eng.global_step == orc.global_step == 3
    num = den = 0.0
    for name, ref in list(orc.gen_params.items()) + list(orc.dis_params.items()):
        net = eng.D if name.startswith('dis/') else eng.G
        got = net.get_variable(name).cpu().double()
        num += float((got - ref).norm() ** 2)
        den += float((ref - init[name]).norm() ** 2)
    assert (num / den) ** 0.5 < 2e-2      # Adam's first steps are sign-like: tiny gradient entries may flip
    for name, ref in list(orc.gen_state.items()) + list(orc.dis_state.items()):
        net = eng.D if name.startswith('dis/') else eng.G
        assert rel(net.get_state(name), ref) < 5e-3, name


def test_cuda_graph_replay_matches_eager(cuda):
    arch = oa.tiny(act_k=2.6)
    B = 16
    _, eng_e = make_pair(arch, B, 'rep', use_graph=False)
    _, eng_g = make_pair(arch, B, 'rep', use_graph=True)
    for it in range(4):
        data, code = onet.synthetic_batch(arch, B, seed=40 + it)
        le = eng_e.step(data, code)
        lg = eng_g.step(data, code)
        assert le == lg, (it, le, lg)       # same kernels, same order: bit-identical
    assert eng_g._graphs is not None and eng_g.kernel_launches_per_step > 0


def test_single_bf16_pass_mode_is_close_but_not_parity_grade(cuda):
    """npass=1 (opt-in speed mode: one bf16 plane, one product) follows the oracle only loosely (bf16 carries 8 significand
    bits and the MMD loss amplifies score errors); it is NOT the parity configuration."""
    arch = oa.tiny(act_k=2.6)
    B = 16
    orc, eng = make_pair(arch, B, 'rep', npass=1)
    check_step(orc, eng, arch, B, seed=5, tol=5e-1)


def test_fp16_plane_saturation_is_reported(cuda):
    """The forward operands are fp16 planes of 16 x value: an activation beyond +-4094 cannot be represented.  The producers
    saturate (never inf) and raise a device flag; step() turns it into a FloatingPointError instead of training on silently
    clipped activations."""
    arch = oa.tiny(act_k=2.6)
    B = 8
    orc, eng = make_pair(arch, B, 'rep')
    data, code = onet.synthetic_batch(arch, B, seed=3, dtype=torch.float32)
    eng.step(data, code)                                        # an ordinary step: no flag
    name = eng.G.layers[0].ly.bias_name                          # push the generator's first activation far out of range
    eng.G.set_variable(name, torch.full(eng.G.var_offsets[name][1], 1.0e4))
    eng.G.refresh()
    with pytest.raises(FloatingPointError):
        eng.step(data, code, check_nan=False)
