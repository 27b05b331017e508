"""Per-launch CUDA-event timing of every tcgen05 GEMM launch of one eager step (which layer, which pass, TFLOP/s)."""
import sys, json, torch
sys.path.insert(0, '.')
from mmdgan_b200 import experiments as oa
from mmdgan_b200 import kernels as K
from mmdgan_b200.engine import SNGanEngine
name = sys.argv[1] if len(sys.argv) > 1 else 'cifar'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
npass = int(sys.argv[3]) if len(sys.argv) > 3 else 3
arch = oa.ARCHITECTURES[name]()
eng = SNGanEngine(arch, B, loss_type="rep", npass=npass, use_graph=False)
eng.sn_fork = eng.grad_fork = False   # single stream: every launch timed alone
g = torch.Generator().manual_seed(0)
data = (torch.rand(B, *arch['input'][0], generator=g) * 2 - 1).cuda(); code = torch.randn(B, 128, generator=g).cuda()
recs = []
og, ow, od = K.LinearOp._gemm, K.LinearOp.wgrad, K.LinearOp._direct
DEPTH = [0]      # nested calls (an image layer = im2col / tap-sum + an inner dense GEMM) are timed at the outermost level only
def tg(self, g_, src, nimg, dst, geom, *a, **kw):
    if DEPTH[0] > 0: return og(self, g_, src, nimg, dst, geom, *a, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = og(self, g_, src, nimg, dst, geom, *a, **kw); e.record()
    d = geom['dims']; M = nimg * d[2] * d[3] * g_['classes']
    kind = 'fwd' if g_ is self.f else 'dgrad'
    recs.append((kind, self.op, self.Cin, self.Cout, self.Hin, nimg, M, g_['ncols'], g_['kpad'], 2.0 * M * g_['ncols'] * g_['taps'] * g_['Cs'], s, e, g_['bn']))
    return r
def tw(self, x_in, dy, nimg, partials, splits=None, **kw):
    if DEPTH[0] > 0: return ow(self, x_in, dy, nimg, partials, splits, **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    DEPTH[0] += 1
    s.record(); r = ow(self, x_in, dy, nimg, partials, splits, **kw); e.record()
    DEPTH[0] -= 1
    R, NC, bn, sp, P = self.wgrad_plan(nimg)
    recs.append(('wgrad', self.op, self.Cin, self.Cout, self.Hin, nimg, R, NC, P, 2.0 * R * NC * P, s, e, (bn, r)))
    return r
def td(self, fwd, src, nimg, dst, *a, **kw):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    DEPTH[0] += 1
    s.record(); r = od(self, fwd, src, nimg, dst, *a, **kw); e.record()
    DEPTH[0] -= 1
    M = nimg * self.Hin * self.Win
    recs.append(('fwd*' if fwd else 'dgrad*', self.op, self.Cin, self.Cout, self.Hin, nimg, M, self.Cout if fwd else self.Cin, 9 * (self.Cin if fwd else self.Cout),
                 2.0 * M * 9 * self.Cin * self.Cout, s, e, 'direct'))
    return r
for it in range(3):
    eng.stage(data, code); eng.step_device()
K.LinearOp._gemm, K.LinearOp.wgrad, K.LinearOp._direct = tg, tw, td
recs.clear()
eng.stage(data, code)
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); eng._run_phases(); t1.record(); torch.cuda.synchronize()
tot = 0
print('%-6s %-3s %5s %5s %4s %5s %8s %6s %7s %8s %8s %s' % ('kind', 'op', 'Cin', 'Cout', 'Hin', 'nimg', 'M/R', 'N', 'K/P', 'ms', 'TFLOP/s', 'bn'))
for r in recs:
    ms = r[10].elapsed_time(r[11]); tot += ms
    if ms > 0.02:
        print('%-6s %-3s %5d %5d %4d %5d %8d %6d %7d %8.3f %8.1f %s' % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], ms, r[9] / ms / 1e9, r[12]))
print('gemm total ms %.3f, eager step ms %.3f, launches %d' % (tot, t0.elapsed_time(t1), len(recs)))
