"""Per-kernel parity on the GPU: every CUDA kernel, called through the C ABI, against the CPU oracle
(float64) on the same seeded inputs.  Tolerances: 5e-5 normwise for the parity mode of the tensor-core kernels (two 16-bit
planes per operand -- fp16 forward, bf16 gradients -- three plane-pair products) and the CUDA-core kernels, 1e-2 for the opt-in single
bf16 pass (north-star parity bar is 1e-3 on the parity-mode path)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mmd as omm
from oracle import net as onet

pytestmark = pytest.mark.gpu

TOL3 = 5e-5
TOL1 = 1e-2


def rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def raw_to_nchw(t, n, c, h, w):
    """fp32 NHWC rows [1, n*h*w, Cpad] -> [n, c, h, w]."""
    return t[0].reshape(n, h, w, -1)[..., :c].permute(0, 3, 1, 2).contiguous()


def _spec(op, cin, cout, hin, k, s):
    design = onet.update_layer_design({'name': 't', 'op': op, 'out': cout, 'kernel': k, 'strides': s})
    return onet.LayerSpec(design, [cin, hin, hin] if op != 'd' else [cin], 'n/t')


CASES = [
    # op, Cin, Cout, Hin, k, s, N
    ('c', 64, 128, 16, 3, 1, 4),
    ('c', 64, 128, 16, 4, 2, 4),
    ('c', 3, 64, 32, 3, 1, 3),
    ('c', 64, 3, 16, 3, 1, 5),
    ('c', 32, 32, 12, 3, 1, 2),      # 48-px family (non power-of-two rows), ragged M tile
    ('c', 256, 512, 8, 4, 2, 2),
    ('tc', 128, 64, 8, 4, 2, 3),
    ('tc', 512, 256, 4, 4, 2, 2),
    ('d', 128, 512, 1, 1, 1, 7),
    ('d', 2048, 16, 1, 1, 1, 9),
    # weight-gradient kernel variants: TMA-staged gather + CTA pair over several k-steps and images; a 4x4 grid with an odd
    # image count (the last 32-pixel k-step is half out of bounds); stride 2 on a 16-wide output grid; a 64-wide grid (cp.async path)
    ('c', 256, 256, 8, 3, 1, 6),
    ('tc', 512, 256, 4, 4, 2, 5),
    ('c', 128, 256, 32, 4, 2, 2),
    ('c', 64, 64, 64, 3, 1, 1),
]


@pytest.mark.parametrize('npass', [3, 1])
@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
def test_linear_op_fwd_dgrad_wgrad(cuda, case, npass):
    from mmdgan_b200 import kernels as K
    op, cin, cout, hin, k, s, n = case
    tol = TOL3 if npass == 3 else TOL1
    g = torch.Generator().manual_seed(cin * 7 + cout * 3 + hin + k + s + n)
    sp = _spec(op, cin, cout, hin, k, s)
    w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.1
    in_shape, out_shape = sp.in_shape, sp.op_out_shape
    x = torch.randn([n] + in_shape, generator=g, dtype=torch.float64)
    dy = torch.randn([n] + out_shape, generator=g, dtype=torch.float64)
    bias = torch.randn(cout, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    y_ref = onet._op_forward(sp, xr, wr)
    dx_ref, dw_ref = torch.autograd.grad(y_ref, [xr, wr], dy)
    alpha = 0.7
    shp = (1, -1) if op == 'd' else (1, -1, 1, 1)
    y_act_ref = F.leaky_relu(alpha * y_ref.detach() + bias.view(shp), 0.1)

    lop = K.LinearOp(op, in_shape, out_shape, k, s, npass=npass)
    wd = w.float().to(cuda).contiguous()
    lop.pack(wd)
    hin_, hout_ = (1, 1) if op == 'd' else (in_shape[1], out_shape[1])
    nv, ng = K.mode_planes(npass, 'value'), K.mode_planes(npass, 'grad')
    xs = K.new_value_planes(n * hin_ * hin_, lop.Cs_in, npass)      # forward operand: two fp16 planes in the parity mode
    K.nchw_to_planes(x.float().to(cuda).contiguous(), xs)
    # ---- forward with fused alpha, bias, lrelu
    ys = K.new_planes(n * hout_ * hout_, lop.Cs_out, 3)
    bias_d = torch.zeros(lop.Cs_out, device=cuda)
    bias_d[:cout] = bias.float().to(cuda)
    lop.forward(xs, n, ys, alpha_k=alpha, bias=bias_d, act=1, out_mode=0)
    y = K.planes_to_nchw(ys, n, cout, hout_, hout_).reshape(y_act_ref.shape)
    assert rel(y, y_act_ref) < tol
    # the planes are a proper split: p1 + p2 is the bf16 rounding residual of p0
    lo = K.planes_to_nchw(ys[1:], n, cout, hout_, hout_).reshape(y_act_ref.shape)
    assert float(lo.abs().max()) <= float(y.abs().max()) * 2.0 ** -8
    # ---- raw forward (out_mode 2) with column sums
    T = lop.fwd_tiles(n)
    cs = torch.zeros(T, lop.Cs_out, device=cuda)
    cq = torch.zeros(T, lop.Cs_out, device=cuda)
    yr = torch.zeros((1, n * hout_ * hout_, lop.Cs_out), device=cuda)
    lop.forward(xs, n, yr, out_mode=2, colsum=cs, colsumsq=cq)
    y2 = raw_to_nchw(yr, n, cout, hout_, hout_).reshape(y_ref.shape)
    assert rel(y2, y_ref.detach()) < tol
    red = [0] if op == 'd' else [0, 2, 3]
    assert rel(cs.sum(0)[:cout], y_ref.detach().sum(red)) < max(tol, 1e-4)
    assert rel(cq.sum(0)[:cout], (y_ref.detach() ** 2).sum(red)) < max(tol, 1e-4)
    # ---- input gradient
    dys = K.new_planes(n * hout_ * hout_, lop.Cs_out, ng)
    K.nchw_to_planes(dy.float().to(cuda).contiguous(), dys)
    dxs = K.new_planes(n * hin_ * hin_, lop.Cs_in, 2)
    lop.dgrad(dys, n, dxs, out_mode=0)
    dx = K.planes_to_nchw(dxs, n, cin, hin_, hin_).reshape(dx_ref.shape)
    assert rel(dx, dx_ref) < tol
    # ---- weight gradient (split-K partials -> canonical layout)
    R, NC, bn, splits, P = lop.wgrad_plan(n)
    parts = torch.zeros(splits * R * NC, device=cuda)
    lop.wgrad(xs, dys, n, parts, splits)
    gw = torch.zeros(lop.canon_numel, device=cuda)
    nb = K.lib().mmdgan_wgrad_reduce_blocks(R * NC)
    dots = torch.zeros(nb, dtype=torch.float64, device=cuda)
    lop.wgrad_reduce(parts, splits, n, gw, w_canon=wd, dots=dots)
    assert rel(gw.reshape(dw_ref.shape), dw_ref) < tol
    assert abs(float(dots.sum()) - float((dw_ref * w).sum())) <= max(tol, 1e-4) * float(dw_ref.norm() * w.norm())


@pytest.mark.parametrize('case', [c for c in CASES if c[2] % 128 == 0 or (c[0] != 'd' and c[1] % 128 == 0)], ids=str)
def test_linear_op_cta_pair_kernel(cuda, case, monkeypatch):
    """The tcgen05 cta_group::2 variant (normally selected only for large grids) forced on small problems: a 2-CTA
    cluster shares one 256 x bn tile; the peer CTA's operands are credited to the leader's mbarrier."""
    from mmdgan_b200 import kernels as K
    monkeypatch.setattr(K, 'GEMM_PAIR_MIN_TILES', 1)
    test_linear_op_fwd_dgrad_wgrad(cuda, case, 3)


def test_dgrad_fused_activation_derivative_and_wrap(cuda):
    """dgrad epilogue: multiply by lrelu'(a) read from the layer input activation, with the 3B virtual-batch row wrap
    and column sums restricted to the first 2B images (bias gradient of the previous layer)."""
    from mmdgan_b200 import kernels as K
    g = torch.Generator().manual_seed(5)
    b, cin, cout, h = 2, 32, 64, 8
    sp = _spec('c', cin, cout, h, 3, 1)
    w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.1
    a_prev = torch.randn(2 * b, cin, h, h, generator=g, dtype=torch.float64)      # activation of the 2B real+fake images
    dy = torch.randn(3 * b, cout, h, h, generator=g, dtype=torch.float64)         # 3B virtual rows
    dx_lin = F.conv_transpose2d(dy, w.permute(3, 2, 0, 1), padding=1)
    a_virtual = torch.cat([a_prev, a_prev[b:]], 0)
    dx_ref = dx_lin * torch.where(a_virtual > 0, 1.0, 0.1)
    lop = K.LinearOp('c', [cin, h, h], [cout, h, h], 3, 1)
    lop.pack(w.float().to(cuda).contiguous())
    aps = K.new_planes(2 * b * h * h, cin)
    K.nchw_to_planes(a_prev.float().to(cuda).contiguous(), aps)
    dys = K.new_planes(3 * b * h * h, cout, 2)
    K.nchw_to_planes(dy.float().to(cuda).contiguous(), dys)
    dxs = K.new_planes(3 * b * h * h, cin, 2)
    T = lop.dgrad_tiles(3 * b)
    cs = torch.zeros(T, cin, device=cuda)
    lop.dgrad(dys, 3 * b, dxs, aux=aps, aux_mode=1, aux_wrap=(2 * b * h * h, b * h * h), colsum=cs, colsum_rows=2 * b * h * h)
    dx = K.planes_to_nchw(dxs, 3 * b, cin, h, h)
    assert rel(dx, dx_ref) < TOL3
    assert rel(cs.sum(0), dx_ref[:2 * b].sum([0, 2, 3])) < 1e-4


def test_direct_conv_image_layers(cuda):
    """The image-channel layers (3 -> 64, 64 -> 3) run as direct CUDA-core convolutions; here the input gradient of D's
    first layer with everything the engine fuses into it: act_k / sigma, tanh'(x_gen) read from the planes, per-block
    column sums (bias gradient of the generator's last layer) -- and agreement with the tensor-core path."""
    from mmdgan_b200 import kernels as K
    g = torch.Generator().manual_seed(21)
    b, h = 3, 16
    sp = _spec('c', 3, 64, h, 3, 1)
    w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.1
    dy = torch.randn(b, 64, h, h, generator=g, dtype=torch.float64)
    xg = torch.tanh(torch.randn(b, 3, h, h, generator=g, dtype=torch.float64))
    alpha_k, sigma = 0.9, 1.7
    ref = F.conv_transpose2d(dy, w.permute(3, 2, 0, 1), padding=1) * (alpha_k / sigma) * (1.0 - xg * xg)
    outs = []
    for direct in (True, False):
        K.DIRECT_CONV = direct
        try:
            lop = K.LinearOp('c', [3, h, h], [64, h, h], 3, 1)
        finally:
            K.DIRECT_CONV = True
        assert (lop.direct_d == 'ls') == direct and (lop.direct_f == 'sl') == direct
        lop.pack(w.float().to(cuda).contiguous())
        dys = K.new_planes(b * h * h, 64, 2)
        K.nchw_to_planes(dy.float().to(cuda).contiguous(), dys)
        auxp = K.new_planes(b * h * h, 8, 3)
        K.nchw_to_planes(xg.float().to(cuda).contiguous(), auxp)
        dxs = K.new_planes(b * h * h, 8, 2)
        cs = torch.zeros(lop.dgrad_tiles(b), 8, device=cuda)
        lop.dgrad(dys, b, dxs, sigma=torch.tensor([sigma], device=cuda), alpha_k=alpha_k, aux=auxp, aux_mode=3, colsum=cs)
        dx = K.planes_to_nchw(dxs, b, 3, h, h)
        assert rel(dx, ref) < TOL3
        assert rel(cs.sum(0)[:3], ref.sum([0, 2, 3])) < 1e-4 and float(cs[:, 3:].abs().max()) == 0.0
        outs.append(dx)
    assert rel(outs[0], outs[1]) < TOL3
    # forward of G's last layer 64 -> 3 with bias + tanh, odd image size (partial 16 x 16 tiles)
    h = 24
    sp = _spec('c', 64, 3, h, 3, 1)
    w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.05
    x = torch.randn(2, 64, h, h, generator=g, dtype=torch.float64)
    bias = torch.randn(3, generator=g, dtype=torch.float64) * 0.1
    y_ref = torch.tanh(onet._op_forward(sp, x, w) + bias.view(1, -1, 1, 1))
    lop = K.LinearOp('c', [64, h, h], [3, h, h], 3, 1)
    assert lop.direct_f == 'ls'
    lop.pack(w.float().to(cuda).contiguous())
    xs = K.new_value_planes(2 * h * h, 64)
    K.nchw_to_planes(x.float().to(cuda).contiguous(), xs)
    ys = K.new_value_planes(2 * h * h, 8)
    bias_d = torch.zeros(8, device=cuda)
    bias_d[:3] = bias.float().to(cuda)
    lop.forward(xs, 2, ys, bias=bias_d, act=3, out_mode=0)
    assert rel(K.planes_to_nchw(ys, 2, 3, h, h), y_ref) < TOL3
    assert float(K.planes_value(ys)[:, 3:].abs().max()) == 0.0


MMD_CASES = [(2, 16), (3, 16), (64, 16), (200, 16), (256, 16), (96, 8), (40, 32), (130, 64), (33, 4)]


@pytest.mark.parametrize('loss_type', ['rep', 'rmb', 'mmd_g', 'mgb', 'mmd_t'])
@pytest.mark.parametrize('bd', MMD_CASES, ids=[str(c) for c in MMD_CASES])
def test_mmd_fused_parity(cuda, bd, loss_type):
    from mmdgan_b200 import kernels as K
    b, d = bd
    rng = np.random.RandomState(b * 7 + d)
    # scale so that distances straddle the rmb bounds 0.25 / 4.0
    gen = (rng.randn(b, d) * 0.35).astype(np.float32)
    real = (rng.randn(b, d) * 0.35 + 0.1).astype(np.float32)
    if b >= 3:
        gen[1] = gen[0]            # duplicate rows: exact-zero distance at the max(., 0) clamp
    ref = omm.gan_loss_with_grads(gen, real, loss_type, rep_weights=(0.0, -1.0))
    mk = K.MmdKernel(loss_type, (0.0, -1.0), b=b)
    gd, rd = torch.from_numpy(gen).to(cuda), torch.from_numpy(real).to(cuda)
    out = [torch.zeros(b, d, device=cuda) for _ in range(4)]
    for _ in range(2):             # twice: the workspace counter must reset itself
        losses = mk(gd, rd, out[0], out[1], out[2], dLg_dreal=out[3])
    torch.cuda.synchronize()
    lg, ld = float(losses[0]), float(losses[1])
    scale = 1.0
    assert abs(lg - float(ref['loss_gen'])) < 2e-6 * scale + 1e-4 * abs(float(ref['loss_gen']))
    assert abs(ld - float(ref['loss_dis'])) < 2e-6 * scale + 1e-4 * abs(float(ref['loss_dis']))
    gscale = max(np.abs(ref[k]).max() for k in ['dLg_dgen', 'dLd_dgen', 'dLd_ddata', 'dLg_ddata'])
    for t, key in zip(out, ['dLg_dgen', 'dLd_dgen', 'dLd_ddata', 'dLg_ddata']):
        assert np.abs(t.cpu().numpy() - ref[key]).max() < 2e-5 * gscale + 1e-9, key


def test_mmd_rep_weights_and_errors(cuda):
    from mmdgan_b200 import kernels as K
    from mmdgan_b200._lib import MmdganError
    rng = np.random.RandomState(0)
    gen = (rng.randn(48, 16) * 0.4).astype(np.float32)
    real = (rng.randn(48, 16) * 0.4).astype(np.float32)
    for w in [(1.0, 0.0), (-1.0, -2.0), (0.5, -0.5)]:
        for lt in ['rep', 'rmb']:
            ref = omm.gan_loss_with_grads(gen, real, lt, rep_weights=w)
            mk = K.MmdKernel(lt, w, b=48)
            o = [torch.zeros(48, 16, device=cuda) for _ in range(3)]
            losses = mk(torch.from_numpy(gen).to(cuda), torch.from_numpy(real).to(cuda), o[0], o[1], o[2])
            assert abs(float(losses[1]) - float(ref['loss_dis'])) < 1e-5
            assert np.abs(o[1].cpu().numpy() - ref['dLd_dgen']).max() < 2e-5 * np.abs(ref['dLd_dgen']).max() + 1e-9
            assert np.abs(o[2].cpu().numpy() - ref['dLd_ddata']).max() < 2e-5 * np.abs(ref['dLd_ddata']).max() + 1e-9
    with pytest.raises(MmdganError):
        K.MmdKernel('rep', (0.0, 0.0))          # w[0]-w[1] must be 1 (math_func.py:1340)
    with pytest.raises(MmdganError):
        K.MmdKernel('no_such_loss')             # NotImplementedError in the reference (math_func.py:2651)


def test_mmd_row_block_form_matches_global(cuda):
    """Multi-GPU row-block form: two halves with row0 offsets reproduce the single-launch sums and gradients."""
    from mmdgan_b200 import kernels as K
    rng = np.random.RandomState(3)
    B, d = 64, 16
    gen = torch.from_numpy((rng.randn(B, d) * 0.4).astype(np.float32)).to(cuda)
    real = torch.from_numpy((rng.randn(B, d) * 0.4).astype(np.float32)).to(cuda)
    full = K.MmdKernel('rmb', b=B)
    of = [torch.zeros(B, d, device=cuda) for _ in range(3)]
    full(gen, real, *of)
    sums = torch.zeros(6, device=cuda)
    oh = [torch.zeros(B, d, device=cuda) for _ in range(3)]
    for r in range(2):
        half = K.MmdKernel('rmb', b=B // 2)
        sl = slice(r * B // 2, (r + 1) * B // 2)
        half(gen[sl].contiguous(), real[sl].contiguous(), oh[0][sl], oh[1][sl], oh[2][sl], gen_all=gen, real_all=real, row0=r * B // 2)
        sums += half.sums
    assert torch.allclose(sums, full.sums, rtol=1e-5, atol=1e-7)
    for a, b_ in zip(oh, of):
        assert torch.allclose(a, b_, rtol=1e-5, atol=1e-8)


def test_bn_adam_sn_elementwise(cuda):
    from mmdgan_b200 import kernels as K
    g = torch.Generator().manual_seed(11)
    rows, Cc = 2 * 8 * 8, 64
    z = torch.randn(rows, Cc, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(Cc, generator=g) + 0.5, torch.randn(Cc, generator=g) * 0.1
    da = torch.randn(rows, Cc, generator=g)
    zr = z.double().clone().requires_grad_(True)
    gr, br = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    mean, var = zr.mean(0), zr.var(0, unbiased=False)
    a_ref = F.relu((zr - mean) / torch.sqrt(var + 1e-3) * gr + br)
    dz_ref, dg_ref, db_ref = torch.autograd.grad(a_ref, [zr, gr, br], da.double())
    zd, gd, bd, dad = z.to(cuda), gamma.to(cuda), beta.to(cuda), da.to(cuda)
    # statistics from per-"tile" partial sums
    T = 4
    ps = torch.stack([zd[i::T].sum(0) for i in range(T)]).contiguous()
    pq = torch.stack([(zd[i::T] ** 2).sum(0) for i in range(T)]).contiguous()
    mean_d, inv_d = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    mm, mv = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)
    K.bn_finalize(ps, pq, T, Cc, rows, mean_d, inv_d, mm, mv)
    assert rel(mean_d, mean.detach()) < 1e-5
    assert rel(mm, 0.01 * mean.detach()) < 1e-5
    assert rel(mv, 0.99 + 0.01 * var.detach() * rows / (rows - 1)) < 1e-5
    mm2, mv2 = torch.zeros(Cc, device=cuda), torch.ones(Cc, device=cuda)      # rank-2 batch norm: biased variance (nn.moments path)
    K.bn_finalize(ps, pq, T, Cc, rows, mean_d, inv_d, mm2, mv2, bessel=False)
    assert rel(mv2, 0.99 + 0.01 * var.detach()) < 1e-5
    a = K.new_planes(rows, Cc, 3)
    K.bn_apply(zd, mean_d, inv_d, gd, bd, Cc, rows * Cc, 2, a)
    assert rel(K.planes_value(a), a_ref.detach()) < 1e-5
    assert rel(K.planes_value(a[:2]), a_ref.detach()) < 2e-5      # two planes carry 16 significand bits
    a16 = K.new_value_planes(rows, Cc)                            # two fp16 planes of 16 x value: 22 significand bits
    K.bn_apply(zd, mean_d, inv_d, gd, bd, Cc, rows * Cc, 2, a16)
    assert a16.dtype == torch.float16 and rel(K.planes_value(a16), a_ref.detach()) < 1e-6
    rpb = 32
    nb = (rows + rpb - 1) // rpb
    p1, p2 = torch.zeros(nb, Cc, device=cuda), torch.zeros(nb, Cc, device=cuda)
    K.bn_bwd_reduce(dad, zd, mean_d, inv_d, gd, bd, Cc, rows, rpb, 2, p1, p2)
    dbeta, dgamma = torch.zeros(Cc, device=cuda), torch.zeros(Cc, device=cuda)
    K.reduce_tiles(p1, nb, Cc, dbeta)
    K.reduce_tiles(p2, nb, Cc, dgamma)
    assert rel(dbeta, db_ref) < 1e-5 and rel(dgamma, dg_ref) < 1e-5
    dz = K.new_planes(rows, Cc, 2)
    K.bn_bwd_apply(dad, zd, mean_d, inv_d, gd, bd, dbeta, dgamma, Cc, rows, 2, dz)
    assert rel(K.planes_value(dz), dz_ref) < 2e-5
    # ---- TF-Adam against the oracle's restatement
    n = 1000
    p0, gr0 = torch.randn(n, generator=g), torch.randn(n, generator=g)
    opt = onet.TFAdam({'p': p0.clone()}, 5e-4)
    params = {'p': p0.clone()}
    w, m, v = p0.clone().to(cuda), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    step = torch.zeros(1, dtype=torch.int32, device=cuda)
    for it in range(3):
        opt.apply(params, {'p': gr0 * (it + 1)})
        K.incr_step(step)
        K.adam(w, m, v, (gr0 * (it + 1)).to(cuda), n, 5e-4, step)
    assert torch.allclose(w.cpu(), params['p'], rtol=1e-5, atol=1e-7)
    # ---- l2 normalise
    vec = torch.randn(5000, generator=g).to(cuda)
    out, sig = K.new_planes(1, 5000, 3), torch.zeros(1, device=cuda)
    K.sn_normalize(vec, 5000, out, sigma_out=sig)
    assert abs(float(sig) - float(vec.double().norm())) < 1e-5 * float(vec.norm())
    assert rel(K.planes_value(out).flatten(), vec / (vec.norm() + 1e-10)) < 1e-6
    # ---- fp32 <-> bf16 planes round trip: three planes reproduce the fp32 value to 1 ulp, two planes to 2^-16
    xf = (torch.randn(4096, generator=g) * torch.logspace(-20, 20, 4096)).to(cuda)
    pl = K.new_planes(1, 4096, 3)
    K.to_planes(xf, pl)
    assert float(((K.planes_value(pl).flatten() - xf).abs() / xf.abs()).max()) <= 2.0 ** -23
    assert float(((K.planes_value(pl[:2]).flatten() - xf).abs() / xf.abs()).max()) <= 2.0 ** -16
    # fp16 planes: 22 bits for ordinary magnitudes, absolute accuracy 2^-29 below, saturation (never inf) beyond +-4094
    xa = torch.cat([torch.randn(4000, generator=g) * 3.0, torch.tensor([1e-6, -3e-9, 4000.0, -4090.0, 1e9, -1e9, 0.0, 1e-30])]).to(cuda)
    ph = K.new_value_planes(1, xa.numel())
    K.to_planes(xa, ph)
    back = K.planes_value(ph).flatten()
    assert bool(torch.isfinite(back).all())
    ok = xa.abs() < 4094
    assert float(((back - xa).abs()[ok] - (xa.abs()[ok] * 2.0 ** -21 + 2.0 ** -29)).max()) <= 0.0
    assert float(back[-4]) > 4000.0 and float(back[-3]) < -4000.0          # saturated, finite
