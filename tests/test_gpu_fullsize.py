"""GPU, BASELINE.json's FULL sizes (CIFAR-10 32x32 SNGAN + rep MMD at batch 256; the 1024-row global MMD of the LSUN config):
the float64 oracle cannot run these in seconds, so the CUDA path is checked through size-independent properties --
bit-reproducibility (eager == CUDA-graph replay), translation invariance of the MMD losses, linearity of the whole
backward pass in the score gradients, row-block == global for the data-parallel MMD form -- and, for one full-size
layer, against cuDNN's fp32 convolution (a torch fp32 reference of the same floating-point op, TF32 disabled)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).double().reshape(-1)
    b = torch.as_tensor(b).double().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def cifar_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, 32, 32, generator=g) * 2 - 1, torch.randn(B, 128, generator=g)


FULL_CONFIGS = [('cifar', 256, 'rep'), ('stl', 128, 'rmb'), ('celeba', 128, 'rep')]     # BASELINE.json configs[1..3], per-GPU batch


@pytest.mark.parametrize('name,B,loss_type', FULL_CONFIGS, ids=[c[0] for c in FULL_CONFIGS])
def test_fullsize_steps_are_bit_reproducible(cuda, name, B, loss_type):
    """The shipped architectures at their benchmark batch sizes: fused steps eager vs CUDA-graph replay give identical losses
    and identical parameters (no floating-point atomics anywhere; side streams only reorder independent work)."""
    from mmdgan_b200 import experiments as ex
    from mmdgan_b200.engine import SNGanEngine
    arch = ex.ARCHITECTURES[name]()
    c, h, w = arch['input'][0]
    e1 = SNGanEngine(arch, B, loss_type=loss_type, seed=7, use_graph=False)
    e2 = SNGanEngine(arch, B, loss_type=loss_type, seed=7, use_graph=True)
    for it in range(3):
        g = torch.Generator().manual_seed(100 + it)
        data, code = torch.rand(B, c, h, w, generator=g) * 2 - 1, torch.randn(B, arch['code'][0][0], generator=g)
        l1, l2 = e1.step(data, code), e2.step(data, code)
        assert l1 == l2, (it, l1, l2)
        assert np.isfinite(l1).all()
    assert e2._graphs is not None
    assert torch.equal(e1.D.w, e2.D.w) and torch.equal(e1.G.w, e2.G.w)
    assert torch.equal(e1.D.m, e2.D.m) and torch.equal(e1.G.v, e2.G.v)


def _warm_engine(B, loss_type='rep', steps=3):
    from mmdgan_b200 import experiments as ex
    from mmdgan_b200.engine import SNGanEngine
    eng = SNGanEngine(ex.cifar(), B, loss_type=loss_type, seed=11, use_graph=False)
    eng.grad_fork = False          # single stream: _phase_backward neither forks nor applies the early discriminator update
    for it in range(steps):        # the first steps normalise every spectral-norm in_rand
        eng.step(*cifar_batch(B, 200 + it))
    return eng


def test_fullsize_translation_invariance_and_backward_linearity(cuda):
    from mmdgan_b200 import kernels as K
    B = 256
    eng = _warm_engine(B)
    data, code = cifar_batch(B, 300)
    eng.stage(data.cuda(), code.cuda())
    eng._phase_forward()
    eng._phase_loss()
    last = eng.D.layers[-1]
    seed0 = last.dz_f32.clone()                                  # [3B, 16]: dL_D/ds_real, dL_D/ds_gen, dL_G/ds_gen
    # ---- the MMD losses are invariant under a common shift of all scores: the score gradients of each loss sum to zero
    #      (only rows [0, 2B) of loss_dis are complete: dL_G/ds_real is never materialised)
    col = seed0[:2 * B].double().sum(0)
    assert float(col.abs().max()) <= 1e-4 * float(seed0[:2 * B].double().abs().sum(0).max())

    def backward(seed):
        last.dz_f32.copy_(seed)
        K.to_planes(last.dz_f32, last.dz)
        eng._phase_backward()
        torch.cuda.synchronize()
        return eng.D.g.clone().double(), eng.G.g.clone().double()

    gd0, gg0 = backward(seed0)
    # ... hence the gradient of the score layer's bias vanishes
    name = last.ly.bias_name
    gmax = max(float(eng.D.view(gd0, n).norm()) for n in eng.D.var_offsets)
    assert float(eng.D.view(gd0, name).norm()) <= 1e-4 * gmax
    # ---- the whole backward pass (input gradients, weight gradients, batch-norm backward, the spectral-norm term) is linear
    #      in the score gradients for fixed activations: g(a s1 + b s2) = a g(s1) + b g(s2)
    g = torch.Generator().manual_seed(5)
    s1 = (torch.randn(seed0.shape, generator=g) * 1e-3).cuda()
    s2 = (torch.randn(seed0.shape, generator=g) * 1e-3).cuda()
    a, b = 0.7, -1.3
    gd1, gg1 = backward(s1)
    gd2, gg2 = backward(s2)
    gd3, gg3 = backward(a * s1 + b * s2)
    assert rel(a * gd1 + b * gd2, gd3) < 1e-3
    assert rel(a * gg1 + b * gg2, gg3) < 1e-3
    # and it is deterministic: the same seed gives the same bits
    gd0b, gg0b = backward(seed0)
    assert torch.equal(gd0, gd0b) and torch.equal(gg0, gg0b)


@pytest.mark.parametrize('loss_type', ['rep', 'rmb'])
def test_fullsize_mmd_row_blocks_match_global_and_gradients_sum_to_zero(cuda, loss_type):
    """Global batch 1024 (the 8 x 128 LSUN configuration): eight row blocks reproduce the single-launch kernel sums and
    gradients; every loss is shift invariant, so its gradients over all 2 x 1024 rows sum to zero per score column."""
    from mmdgan_b200 import kernels as K
    rng = np.random.RandomState(9)
    Bg, d, R = 1024, 16, 8
    gen = torch.from_numpy((rng.randn(Bg, d) * 0.4).astype(np.float32)).cuda()
    real = torch.from_numpy((rng.randn(Bg, d) * 0.4 + 0.1).astype(np.float32)).cuda()
    full = K.MmdKernel(loss_type, b=Bg)
    of = [torch.zeros(Bg, d, device=cuda) for _ in range(4)]
    full(gen, real, of[0], of[1], of[2], dLg_dreal=of[3])
    for ga, ra in ((of[0], of[3]), (of[1], of[2])):                 # (dLg_dgen, dLg_dreal), (dLd_dgen, dLd_dreal)
        tot = (ga.double().sum(0) + ra.double().sum(0)).abs().max()
        assert float(tot) <= 1e-4 * float((ga.double().abs().sum(0) + ra.double().abs().sum(0)).max())
    b = Bg // R
    sums = torch.zeros(6, device=cuda)
    oh = [torch.zeros(Bg, d, device=cuda) for _ in range(3)]
    for r in range(R):
        blk = K.MmdKernel(loss_type, b=b)
        sl = slice(r * b, (r + 1) * b)
        blk(gen[sl].contiguous(), real[sl].contiguous(), oh[0][sl], oh[1][sl], oh[2][sl], gen_all=gen, real_all=real, row0=r * b)
        sums += blk.sums
    assert torch.allclose(sums, full.sums, rtol=1e-4, atol=1e-7)
    for x, y in zip(oh, of[:3]):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-9)


def test_fullsize_layer_against_cudnn_fp32(cuda):
    """D's fifth layer at the benchmark size (256 -> 256, 3x3, 8x8, 512 images; M = 32768, K = 2304): forward (two fp16 planes,
    three products), input gradient and weight gradient (two bf16 planes) against cuDNN fp32 with TF32 disabled."""
    from mmdgan_b200 import kernels as K
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator().manual_seed(3)
        n, c, h = 512, 256, 8
        x = torch.randn(n, c, h, h, generator=g).cuda().requires_grad_(True)
        w = (torch.randn(3, 3, c, c, generator=g) * 0.03).cuda()            # canonical [k, k, Cin, Cout]
        dy = torch.randn(n, c, h, h, generator=g).cuda()
        wt = w.permute(3, 2, 0, 1).contiguous().requires_grad_(True)
        y_ref = F.conv2d(x, wt, padding=1)
        dx_ref, dwt_ref = torch.autograd.grad(y_ref, [x, wt], dy)
        dw_ref = dwt_ref.permute(2, 3, 1, 0)
        lop = K.LinearOp('c', [c, h, h], [c, h, h], 3, 1)
        lop.pack(w.contiguous())
        xs = K.new_value_planes(n * h * h, c)
        K.nchw_to_planes(x.detach().contiguous(), xs)
        yr = torch.zeros((1, n * h * h, c), device=cuda)
        lop.forward(xs, n, yr, out_mode=2)
        y = yr[0].reshape(n, h, h, c).permute(0, 3, 1, 2)
        assert rel(y, y_ref.detach()) < 2e-5
        # linearity at full size: f(x) + f(2x) == f(3x) to rounding
        xs3 = K.new_value_planes(n * h * h, c)
        K.nchw_to_planes((3.0 * x.detach()).contiguous(), xs3)
        y3 = torch.zeros_like(yr)
        lop.forward(xs3, n, y3, out_mode=2)
        assert rel(y3, 3.0 * yr) < 2e-5
        dys = K.new_planes(n * h * h, c, 2)
        K.nchw_to_planes(dy.contiguous(), dys)
        dxs = K.new_planes(n * h * h, c, 2)
        lop.dgrad(dys, n, dxs, out_mode=0)
        assert rel(K.planes_to_nchw(dxs, n, c, h, h), dx_ref) < 1e-4
        R, NC, bn, splits, P = lop.wgrad_plan(n)
        parts = torch.zeros(splits * R * NC, device=cuda)
        lop.wgrad(xs, dys, n, parts, splits)
        gw = torch.zeros(lop.canon_numel, device=cuda)
        lop.wgrad_reduce(parts, splits, n, gw)
        assert rel(gw.reshape(3, 3, c, c), dw_ref) < 1e-4
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
