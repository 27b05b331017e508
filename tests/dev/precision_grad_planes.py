"""Round-2 study (CPU, drives the oracle): can the GRADIENT operand of the input-/weight-gradient GEMMs be ONE 16-bit plane
(2 plane-pair products per launch instead of 3)?  The forward stays as shipped (two fp16 planes per operand); the incoming
gradient of every layer is rounded to
  bf16x2   two bf16 planes (shipped)                                  16 bits
  f16x1s   one fp16 plane after a per-tensor power-of-two scale        11 bits, no underflow
  f16x1    one fp16 plane, unscaled                                    11 bits, underflow below 6e-8
  bf16x1   one bf16 plane                                              8 bits
and every gradient tensor is compared with the float64 oracle (the parity bar is 1e-3).
Result (CIFAR net, act_k 2.7, batch 8; median over the gradient tensors): bf16x2 7e-6, f16x1s 5.6e-4, f16x1 5.5e-4 (3e-2 at the
script's act_k 1.68: underflow), bf16x1 4e-3 -- a single 11-bit plane sits at half the 1e-3 bar with individual tensors at
8e-4, so the two-product gradient launches are NOT parity grade; three products stay.
Usage: python tests/dev/precision_grad_planes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import net as onet, architectures as oa
torch.set_num_threads(8)


def fwd_f16x2(x):
    if x.dtype != torch.float32:
        return x
    xs = (x * 16.0).clamp(-65504.0, 65504.0)
    h0 = xs.half().float()
    return (h0 + (xs - h0).half().float()) / 16.0


def g_bf16x2(g):
    hi = g.to(torch.bfloat16).float()
    return hi + (g - hi).to(torch.bfloat16).float()


def g_f16x1s(g):
    m = float(g.abs().max())
    if m == 0.0:
        return g
    s = 2.0 ** (14 - torch.frexp(torch.tensor(m))[1].item())      # max |g| * s in [8192, 16384)
    return (g * s).half().float() / s


GRAD = {'bf16x2': g_bf16x2, 'f16x1s': g_f16x1s, 'f16x1': lambda g: g.half().float(), 'bf16x1': lambda g: g.to(torch.bfloat16).float()}


class Round(torch.autograd.Function):
    mode = 'bf16x2'

    @staticmethod
    def forward(ctx, x):
        return fwd_f16x2(x)

    @staticmethod
    def backward(ctx, g):
        return GRAD[Round.mode](g)


def run(arch, B, seed, mode):
    if mode == 'f64':
        m = onet.OracleSNGan(arch, 'rep', dtype=torch.float64, seed=3)
        onet._maybe_tf32 = lambda x, on: x
    else:
        m = onet.OracleSNGan(arch, 'rep', dtype=torch.float32, seed=3)
        Round.mode = mode
        onet._maybe_tf32 = lambda x, on: Round.apply(x)
    onet.warm_spectral_norm(m, 6)
    data, code = onet.synthetic_batch(arch, B, seed=seed, dtype=m.dtype)
    lg, ld, gg, gd, _, _ = m.grads(data, code)
    return {**gg, **gd}


for name, arch, B in (('tiny', oa.tiny(act_k=2.6), 16), ('cifar k2.7', oa.cifar(act_k=2.7), 8), ('cifar k1.68', oa.cifar(), 8)):
    ref = run(arch, B, 11, 'f64')
    gmax = max(float(v.norm()) for k, v in ref.items() if k.startswith('dis/'))
    for mode in GRAD:
        got = run(arch, B, 11, mode)
        errs = sorted(((float((got[k].double() - v).norm() / v.norm()), k) for k, v in ref.items() if float(v.norm()) >= 1e-6 * gmax), reverse=True)
        print('%-12s %-7s worst %.2e %-28s median %.2e' % (name, mode, errs[0][0], errs[0][1], errs[len(errs) // 2][0]), flush=True)
