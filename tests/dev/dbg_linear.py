import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from oracle import net as onet
from mmdgan_b200 import kernels as K
from tests.test_gpu_kernels import _spec, rel, CASES, raw_to_nchw
cuda = torch.device('cuda')
for case in CASES:
  for npass in (3, 1):
    op, cin, cout, hin, k, s, n = case
    g = torch.Generator().manual_seed(1)
    sp = _spec(op, cin, cout, hin, k, s)
    w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.1
    in_shape, out_shape = sp.in_shape, sp.op_out_shape
    x = torch.randn([n] + in_shape, generator=g, dtype=torch.float64)
    dy = torch.randn([n] + out_shape, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    y_ref = onet._op_forward(sp, xr, wr)
    dx_ref, dw_ref = torch.autograd.grad(y_ref, [xr, wr], dy)
    lop = K.LinearOp(op, in_shape, out_shape, k, s, npass=npass)
    wd = w.float().to(cuda).contiguous(); lop.pack(wd)
    hin_, hout_ = (1, 1) if op == 'd' else (in_shape[1], out_shape[1])
    nv, ng = K.mode_planes(npass, 'value'), K.mode_planes(npass, 'grad')
    xs = K.new_value_planes(n * hin_ * hin_, lop.Cs_in, npass); K.nchw_to_planes(x.float().to(cuda).contiguous(), xs)
    yr = torch.zeros((1, n * hout_ * hout_, lop.Cs_out), device=cuda)
    lop.forward(xs, n, yr, out_mode=2)
    y2 = raw_to_nchw(yr, n, cout, hout_, hout_).reshape(y_ref.shape)
    dys = K.new_planes(n * hout_ * hout_, lop.Cs_out, ng); K.nchw_to_planes(dy.float().to(cuda).contiguous(), dys)
    dxs = K.new_planes(n * hin_ * hin_, lop.Cs_in, 2)
    lop.dgrad(dys, n, dxs, out_mode=0)
    dx = K.planes_to_nchw(dxs, n, cin, hin_, hin_).reshape(dx_ref.shape)
    R, NC, bn, splits, P = lop.wgrad_plan(n)
    parts = torch.zeros(splits * R * NC, device=cuda)
    lop.wgrad(xs, dys, n, parts, splits)
    gw = torch.zeros(lop.canon_numel, device=cuda)
    lop.wgrad_reduce(parts, splits, n, gw, w_canon=wd, dots=None)
    torch.cuda.synchronize()
    print(case, npass, 'fwd %.2e dgrad %.2e wgrad %.2e' % (rel(y2, y_ref.detach()), rel(dx, dx_ref), rel(gw.reshape(dw_ref.shape), dw_ref)),
          'plan', (R, NC, bn, splits, P), 'parts nz', int((parts != 0).sum()), 'absmax', float(parts.abs().max()), 'swapped', lop.w_swapped, flush=True)
