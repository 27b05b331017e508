import sys, torch
sys.path.insert(0, '.')
from oracle import net as onet
from mmdgan_b200 import kernels as K
from tests.test_gpu_kernels import _spec, rel, raw_to_nchw
stage = sys.argv[1]
cuda = torch.device('cuda')
op, cin, cout, hin, k, s, n = ('c', 64, 128, 16, 3, 1, 4)
g = torch.Generator().manual_seed(1)
sp = _spec(op, cin, cout, hin, k, s)
w = torch.randn(sp.kernel_shape, generator=g, dtype=torch.float64) * 0.1
x = torch.randn([n] + sp.in_shape, generator=g, dtype=torch.float64)
dy = torch.randn([n] + sp.op_out_shape, generator=g, dtype=torch.float64)
xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
y_ref = onet._op_forward(sp, xr, wr)
dx_ref, dw_ref = torch.autograd.grad(y_ref, [xr, wr], dy)
lop = K.LinearOp(op, sp.in_shape, sp.op_out_shape, k, s, npass=3)
wd = w.float().to(cuda).contiguous(); lop.pack(wd)
torch.cuda.synchronize(); print('pack ok', lop.f['w'].dtype, lop.d['w'].dtype, flush=True)
xs16 = K.new_value_planes(n * hin * hin, lop.Cs_in, 3); K.nchw_to_planes(x.float().to(cuda).contiguous(), xs16)
xsb = K.new_planes(n * hin * hin, lop.Cs_in, 3); K.nchw_to_planes(x.float().to(cuda).contiguous(), xsb)
dys = K.new_planes(n * hin * hin, lop.Cs_out, 2); K.nchw_to_planes(dy.float().to(cuda).contiguous(), dys)
torch.cuda.synchronize(); print('planes ok', rel(K.planes_value(xs16), K.planes_value(xsb)), flush=True)
if stage == 'fwd':
    yr = torch.zeros((1, n * hin * hin, lop.Cs_out), device=cuda)
    lop.forward(xs16, n, yr, out_mode=2)
    torch.cuda.synchronize(); print('fwd f16xf16', rel(raw_to_nchw(yr, n, cout, hin, hin), y_ref.detach()), flush=True)
if stage in ('wgrad_mixed', 'wgrad_bf16'):
    xs = xs16 if stage == 'wgrad_mixed' else xsb
    R, NC, bn, splits, P = lop.wgrad_plan(n)
    parts = torch.zeros(splits * R * NC, device=cuda)
    lop.wgrad(xs, dys, n, parts, splits)
    gw = torch.zeros(lop.canon_numel, device=cuda)
    lop.wgrad_reduce(parts, splits, n, gw)
    torch.cuda.synchronize(); print(stage, rel(gw.reshape(dw_ref.shape), dw_ref), flush=True)
