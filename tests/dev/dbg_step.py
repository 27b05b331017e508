import sys, torch
sys.path.insert(0, '.')
from oracle import architectures as oa, net as onet
from tests.test_gpu_step import make_pair, rel
which = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
arch, B, lt, seed = {'cifar': (oa.cifar(act_k=2.7), 8, 'rep', 11), 'stl': (oa.stl(act_k=2.7), 4, 'rmb', 13), 'tiny': (oa.tiny(act_k=2.6), 16, 'rep', 5)}[which]
orc, eng = make_pair(arch, B, lt)
data, code = onet.synthetic_batch(arch, B, seed=seed, dtype=torch.float64)
col = {}
lg, ld, gg, gd, ug, ud = orc.grads(data, code, col)
eng.stage(data.float().cuda(), code.float().cuda())
eng._phase_forward(); eng._phase_loss(); eng._phase_backward(); torch.cuda.synchronize()
print('losses', eng.losses().cpu().tolist(), float(lg), float(ld))
s_ref = torch.cat([col['s_x'], col['s_gen']], 0).detach()
print('scores rel', rel(eng.D.layers[-1].a[0].cpu(), s_ref))
from mmdgan_b200 import kernels as K
for net, nimg in ((eng.G, B), (eng.D, 2 * B)):
    for L in net.layers:
        ref = col[L.ly.layer_scope + '/out'].detach()
        if L.a.shape[0] == 1 and getattr(L, 'raw_out', False):
            got = L.a[0].cpu()
        else:
            if ref.dim() == 2 and L.op != 'd':
                ref = ref.reshape(nimg, L.Cout, L.Hout, L.Wout)
            if L.op == 'd':
                c, hw = L.lop.out_flat
                got = L.a[0].reshape(nimg, hw, c).permute(0, 2, 1).reshape(nimg, -1).cpu(); ref = ref.reshape(nimg, -1)
            else:
                got = K.planes_to_nchw(L.a, nimg, L.Cout, L.Hout, L.Wout).cpu()
        print('  act', L.ly.layer_scope, 'rel %.2e' % rel(got, ref), ('sigma %.6f ref %.6f' % (float(L.sigma), float(col[L.ly.layer_scope + '/sigma'].detach()))) if L.has_sn else '')
for name, ref in list(gd.items()) + list(gg.items()):
    net = eng.D if name.startswith('dis/') else eng.G
    print('  grad', name, 'rel %.2e' % rel(net.get_grad(name).cpu(), ref), '|ref| %.2e' % float(ref.norm()))
# ---- gradient w.r.t. every D pre-activation: oracle (autograd) vs engine dz (first 2B rows: loss_dis path)
lg2, ld2, gp, dp, _, _ = orc.forward_losses(data, code, col)
outs = [col[L.ly.layer_scope + '/out'] for L in eng.D.layers]
das = torch.autograd.grad(ld2, outs, retain_graph=True)
for L, out, da in zip(eng.D.layers, outs, das):
    dz_ref = da * (torch.where(out > 0, 1.0, 0.1) if L.act == 'lrelu' else 1.0)
    nimg = 2 * B
    if L.op == 'd':
        got = eng.D.layers[-1].dz[0][:nimg].cpu() if L is eng.D.layers[-1] else None
        ref = dz_ref.detach()
    else:
        got = K.planes_to_nchw(L.dz, nimg, L.Cout, L.Hout, L.Wout).cpu()
        ref = dz_ref.detach().reshape(nimg, L.Cout, L.Hout, L.Wout)
    if got is not None:
        e = (got.double() - ref)
        print('  dz', L.ly.layer_scope, 'rel %.2e' % rel(got, ref), 'per-image rel', ['%.1e' % rel(got[i], ref[i]) for i in range(min(nimg, 6))], 'max abs err %.2e at' % float(e.abs().max()), int(e.abs().argmax()))
