"""Torch-tensor level wrappers over the C ABI (include/mmdgan_b200.h): device memory and streams come from PyTorch,
every computation is a kernel of libmmdgan_b200.so.  No fallback: a missing library or a non-CUDA tensor raises.

GEMM operands ("planes") are [npl, rows, C] 16-bit CUDA tensors, NHWC row order (include/mmdgan_b200.h):
  torch.bfloat16  the sum of the planes is the value (3 planes = fp32, 2 planes = 16 significand bits): gradients, and
                  every operand when F16_FORWARD is off;
  torch.float16   two planes whose sum is 16 x value (activations) or 64 x value (packed forward weights): the operands of
                  forward launches -- 22 significand bits, so three plane-pair products are fp32-grade.
Gradient launches multiply three bf16 plane pairs; the weight-gradient GEMM mixes fp16 activations with bf16 gradients.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import GemmDesc, WgradDesc, WredDesc, PackDesc, MmdDesc, RefreshJob, DirectDesc

ACT = {'linear': 0, None: 0, 'lrelu': 1, 'relu': 2, 'tanh': 3}
PACK_CONV_FWD, PACK_CONV_DGRAD_S1, PACK_CONV_DGRAD_S2, PACK_TC_FWD, PACK_TC_DGRAD, PACK_DENSE_FWD, PACK_DENSE_DGRAD = range(7)


GEMM_BK = 64            # K elements per pipeline stage of the gather-GEMM = channel-chunk width of its K order (csrc/conv_gemm.cuh kGemmBK)
GEMM_PAIR = True        # use the CTA-pair (cta_group::2) gather-GEMM where the grid is large enough
GEMM_PAIR_MIN_TILES = int(os.environ.get('MMDGAN_PAIR_MIN_TILES', '256'))   # 128 x 128 output units; below that the single-CTA kernel fills the machine better
GEMM_BN_MAX = 256    # widest N tile of the gather-GEMM (256 halves the A re-reads of wide layers)
WGRAD_BN_MAX = int(os.environ.get('MMDGAN_WGRAD_BN', '256'))   # widest N tile of the weight-gradient GEMM (64 / 128 / 256)
FMT_BF16, FMT_F16A, FMT_F16W = 0, 1, 2
FMT_SCALE = {FMT_BF16: 1.0, FMT_F16A: 16.0, FMT_F16W: 64.0}
F16_FORWARD = int(os.environ.get('MMDGAN_F16_FORWARD', '1'))   # parity mode: forward operands as two fp16 planes (3 products) instead of three bf16 planes (6)
WGRAD_CTAS = int(os.environ.get('MMDGAN_WGRAD_CTAS', '222'))   # CTAs a weight-gradient launch aims for (tiles x split-K slices)
PAIR_BN256_AUX = int(os.environ.get('MMDGAN_BN256_AUX', '1'))   # 256-wide pair tiles for input gradients with N = 256 too (the persistent kernel hides their epilogue: 4.19 -> 4.15 ms)
PAIR_N64 = int(os.environ.get('MMDGAN_PAIR_N64', '0'))   # CTA-pair tiles for N = 64 layers: measured slower (0.36 vs 0.33 ms), off
IMG_GEMM = os.environ.get('MMDGAN_IMG_GEMM', '1') == '1'     # the image-channel layers as 1x1 GEMMs over 27 = 9 taps x 3 channels (im2col27 / tapsum27)
DIRECT_CONV = os.environ.get('MMDGAN_DIRECT_CONV', '1') == '1'     # image-channel 3x3 layers (<= 4 channels on one side): direct CUDA-core convolution instead of the GEMM
LAUNCHES = [0]   # kernels launched through the C ABI since import (bench.py reports the per-step count)


def lib():
    return _lib.load()


def check(rc):
    LAUNCHES[0] += 1
    return _lib.check(rc)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda or t.dtype not in (torch.float32, torch.float64, torch.int32, torch.int64, torch.bfloat16, torch.float16):
        raise _lib.MmdganError(_lib.MMDGAN_EINVAL, 'expected a CUDA float32/float64/int32/int64/bfloat16/float16 tensor, got {} on {}'.format(t.dtype, t.device))
    if t.dim() == 3:
        # planes [npl, rows, C]: every plane must be a dense [rows, C] block; the plane stride is free (row views)
        if t.stride(2) != 1 or t.stride(1) != t.shape[2]:
            raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'plane tensor must have dense [rows, C] planes')
    elif not t.is_contiguous():
        raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'tensor must be contiguous')
    return C.c_void_p(t.data_ptr())


def plane_stride(t):
    return t.stride(0) if t.shape[0] > 1 else 0


def _planes(t):
    if t.dtype not in (torch.bfloat16, torch.float16) or t.dim() != 3 or (t.dtype == torch.float16 and t.shape[0] > 2):
        raise _lib.MmdganError(_lib.MMDGAN_EINVAL, 'expected bfloat16 / float16 planes [npl, rows, C], got {} {}'.format(t.dtype, tuple(t.shape)))
    return t


def fmt_of(t, role='act'):
    """Plane format of a tensor: bf16 planes, or fp16 planes scaled for activations ('act') / packed weights ('w')."""
    return FMT_BF16 if t.dtype == torch.bfloat16 else (FMT_F16W if role == 'w' else FMT_F16A)


def fmt_need(fmt, npass):
    """Planes a launch with `npass` plane-pair products reads from an operand in format `fmt`."""
    return (3 if npass == 6 else (2 if npass == 3 else 1)) if fmt == FMT_BF16 else (2 if npass >= 3 else 1)


def pad_c(c):
    """Channel padding of a GEMM operand: 8, 16 or 32 (whole 16-byte units of one 128-byte K row), else whole 64-channel chunks."""
    return 8 if c <= 8 else (16 if c <= 16 else (32 if c <= 32 else (c + 63) // 64 * 64))


def fwd_passes(npass):
    """Engine precision mode (3 = parity, 1 = single bf16 pass) -> plane-pair products of a forward launch."""
    return (3 if F16_FORWARD else 6) if npass == 3 else 1


def mode_planes(npass, kind='value'):
    """Planes allocated for a tensor: values (activations, weights) carry 3, gradients 2; the single-pass mode 1."""
    return 1 if npass == 1 else (3 if kind == 'value' else 2)


def round_up(a, b):
    return (a + b - 1) // b * b


def pick_bn(ncols, lo=16, hi=128):
    bn = lo
    while bn < hi and bn < ncols:
        bn *= 2
    return bn


def new_planes(rows, c, npl=3, device='cuda', fmt=FMT_BF16):
    return torch.zeros((npl, rows, c), dtype=torch.bfloat16 if fmt == FMT_BF16 else torch.float16, device=device)


def new_value_planes(rows, c, npass=3, device='cuda'):
    """Planes of a VALUE that feeds forward launches: two fp16 planes in the parity mode (three bf16 planes with
    F16_FORWARD off, one bf16 plane in the single-pass mode)."""
    if npass == 3 and F16_FORWARD:
        return new_planes(rows, c, 2, device, FMT_F16A)
    return new_planes(rows, c, mode_planes(npass, 'value'), device)


# ------------------------------------------------------------------------------------------------ layout
def nchw_to_planes(x, dst):
    """x fp32 [N,C,H,W] (or [N,F]) -> dst bf16 planes [npl, N*H*W, Cpad]."""
    if x.dim() == 2:
        n, c = x.shape
        h = w = 1
    else:
        n, c, h, w = x.shape
    _planes(dst)
    check(lib().mmdgan_nchw_to_nhwc(_ptr(x), _ptr(dst), plane_stride(dst), dst.shape[0], fmt_of(dst), n, c, h, w, dst.shape[2], stream()))
    return dst


def planes_to_nchw(src, n, c, h, w):
    _planes(src)
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=src.device)
    check(lib().mmdgan_nhwc_to_nchw(_ptr(src), plane_stride(src), src.shape[0], fmt_of(src), _ptr(out), n, c, h, w, src.shape[2], stream()))
    return out


def to_planes(x, dst):
    """fp32 [rows, C] -> bf16 planes [npl, rows, C] (same element order)."""
    _planes(dst)
    n = dst.shape[1] * dst.shape[2]
    assert x.numel() == n and x.dtype == torch.float32
    check(lib().mmdgan_to_planes(_ptr(x), _ptr(dst), plane_stride(dst), dst.shape[0], fmt_of(dst), n, stream()))
    return dst


def convert_planes(src, dst, role='act'):
    """Planes in one format -> planes in another, same [rows, C] (bf16 re-split of fp16 activations for the weight gradients)."""
    _planes(src)
    _planes(dst)
    n = dst.shape[1] * dst.shape[2]
    assert src.shape[1] * src.shape[2] == n
    check(lib().mmdgan_convert_planes(_ptr(src), plane_stride(src), src.shape[0], fmt_of(src, role), _ptr(dst), plane_stride(dst),
                                      dst.shape[0], fmt_of(dst, role), n, stream()))
    return dst


def planes_value(src, role='act'):
    """planes [npl, rows, C] -> the fp32 values [rows, C] they carry."""
    _planes(src)
    out = torch.empty(src.shape[1:], dtype=torch.float32, device=src.device)
    check(lib().mmdgan_from_planes(_ptr(src), plane_stride(src), src.shape[0], fmt_of(src, role), _ptr(out), out.numel(), stream()))
    return out


# ------------------------------------------------------------------------------------------------ linear ops
class LinearOp(object):
    """One parametric op of the layer DSL ('d' / 'c' / 'tc', GeneralTools/layer_func.py:909-928) lowered to the
    gather-GEMM (forward, input gradient) and weight-gradient GEMM launches.

    in_shape / out_shape are the reference's per-sample NCHW shapes ([F] or [C,H,W]).  in_flat / out_flat = (C, HW)
    say that a dense layer's features are a flattened NCHW tensor whose internal storage is NHWC (the reshape at
    my_test_cifar.py:15 / :36), so the feature permutation is folded into the packed weights.
    """

    def __init__(self, op, in_shape, out_shape, kernel=3, strides=1, npass=3, in_flat=None, out_flat=None, device='cuda'):
        self.op, self.k, self.s, self.npass, self.device = op, kernel, strides, npass, device
        self.fwd_npass = fwd_passes(npass)          # forward-type launches (errors amplified by the loss): 6 plane pairs
        self.bwd_npass = 3 if npass == 3 else 1     # gradient launches (linear in the operands): 3 plane pairs
        self.adj_npass = 6 if npass == 3 else 1     # the adjoint used as a FORWARD operator (spectral norm): bf16 x 6
        self.npl = mode_planes(npass)               # planes of the bf16 packed operands
        self._wg_scale = 1.0
        self.sat_flag = None                        # optional int32 device flag: a value written as fp16 planes saturated
        self._ds_ws = None                          # workspace of the split-K small-N dense kernel
        if op == 'd':
            self.Cin, self.Cout = in_shape[0], out_shape[0]
            self.Hin = self.Win = self.Hout = self.Wout = 1
            self.in_flat = in_flat or (self.Cin, 1)
            self.out_flat = out_flat or (self.Cout, 1)
        elif op in ('c', 'tc'):
            self.Cin, self.Hin, self.Win = in_shape
            self.Cout, self.Hout, self.Wout = out_shape
            if not ((kernel == 3 and strides == 1) or (kernel == 4 and strides == 2)) or (op == 'tc' and strides != 2):
                raise NotImplementedError('{}: kernel {} / strides {} is not on the hot path'.format(op, kernel, strides))
            if op == 'c':
                assert self.Hout * strides == self.Hin and self.Wout * strides == self.Win
            else:
                assert self.Hin * strides == self.Hout and self.Win * strides == self.Wout
        else:
            raise AttributeError('layer op {} not supported.'.format(op))
        self.Cs_in, self.Cs_out = pad_c(self.Cin), pad_c(self.Cout)
        self.pad = 1 if op != 'd' else 0
        # direct CUDA-core kernels of the forward / input-gradient pass: 'sl' (<= 4 -> many channels) or 'ls' (many -> <= 4)
        self.direct_f = self.direct_d = None
        self.w_canon = None
        if op == 'c' and kernel == 3 and strides == 1 and DIRECT_CONV:
            if self.Cin <= 4 and self.Cout % 16 == 0 and self.Cout <= 128:
                self.direct_f, self.direct_d = 'sl', 'ls'
            elif self.Cout <= 4 and self.Cin % 16 == 0 and self.Cin <= 128:
                self.direct_f, self.direct_d = 'ls', 'sl'
        # round 2: the many -> few direction of those two layers as a dense product over 27 = 9 taps x 3 channels on the tensor-core
        # GEMM + the tap-sum kernel (csrc/direct_conv.cu, second half)
        self.img_op = None
        if self.direct_f and IMG_GEMM and min(self.Cin, self.Cout) == 3 and npass == 3:
            self.img_few_in = self.Cin <= 3
            cm = self.Cout if self.img_few_in else self.Cin
            self.img_op = LinearOp('d', [27], [cm], npass=npass, device=device) if self.img_few_in else \
                LinearOp('d', [cm], [27], npass=npass, device=device)
            # many -> few: the dense layer's canonical matrix [C][(kh, kw, co)] is a permuted copy of the conv kernel [kh, kw, C, co]
            self.img_wp = None if self.img_few_in else torch.zeros(cm * 27, dtype=torch.float32, device=device)
            self._img_bufs = {}
        k = kernel
        # ---- forward / dgrad operand geometry
        if op == 'd':
            self.f = dict(mode=PACK_DENSE_FWD, classes=1, taps=1, Cs=self.Cs_in, ncols=self.Cs_out)
            self.d = dict(mode=PACK_DENSE_DGRAD, classes=1, taps=1, Cs=self.Cs_out, ncols=self.Cs_in)
        elif op == 'c':
            self.f = dict(mode=PACK_CONV_FWD, classes=1, taps=k * k, Cs=self.Cs_in, ncols=self.Cs_out)
            if strides == 1:
                self.d = dict(mode=PACK_CONV_DGRAD_S1, classes=1, taps=k * k, Cs=self.Cs_out, ncols=self.Cs_in)
            else:
                self.d = dict(mode=PACK_CONV_DGRAD_S2, classes=4, taps=4, Cs=self.Cs_out, ncols=self.Cs_in)
        else:
            self.f = dict(mode=PACK_TC_FWD, classes=4, taps=4, Cs=self.Cs_in, ncols=self.Cs_out)
            self.d = dict(mode=PACK_TC_DGRAD, classes=1, taps=k * k, Cs=self.Cs_out, ncols=self.Cs_in)
        for g in (self.f, self.d):
            g['bn'] = pick_bn(g['ncols'])
            g['rows_pad'] = round_up(g['ncols'], g['bn'])
            g['kpad'] = round_up(g['taps'] * g['Cs'], GEMM_BK)
            if g is self.f and npass == 3 and F16_FORWARD:
                g['w'] = new_planes(g['classes'] * g['rows_pad'], g['kpad'], 2, device, FMT_F16W)
            else:
                g['w'] = new_planes(g['classes'] * g['rows_pad'], g['kpad'], self.npl, device)
        # ---- weight-gradient orientation: the small channel count goes to the N side
        if op == 'd':
            self.w_swapped = self.Cs_out < 32
        elif op == 'c':
            self.w_swapped = (self.Cs_out < 32 and strides == 1)
        else:
            self.w_swapped = False
        self.canon_shape = ([self.Cin, self.Cout] if op == 'd' else
                            [k, k, self.Cin, self.Cout] if op == 'c' else [k, k, self.Cout, self.Cin])
        self.canon_numel = int(math.prod(self.canon_shape))

    # -------------------------------------------------------------------------------------------- packing
    def pack_descs(self, w_canon):
        """The two packing jobs (forward, input-gradient operand) of this op as PackDesc structures."""
        self.w_canon = w_canon
        out = []
        for g in (self.f, self.d):
            d = PackDesc()
            d.w, d.out = _ptr(w_canon), _ptr(g['w'])
            d.plane, d.npl, d.fmt = plane_stride(g['w']), g['w'].shape[0], fmt_of(g['w'], 'w')
            d.mode, d.k, d.Cin, d.Cout, d.Cs = g['mode'], self.k, self.Cin, self.Cout, g['Cs']
            d.rows_pad, d.kpad, d.classes = g['rows_pad'], g['kpad'], g['classes']
            if self.op == 'd':
                d.in_C, d.in_HW = self.in_flat
                d.out_C, d.out_HW = self.out_flat
            else:
                d.in_C = d.in_HW = d.out_C = d.out_HW = 1
            out.append(d)
        if self.img_op is not None:
            out += self.img_op.pack_descs(self.img_canon())
        return out

    # -------------------------------------------------------------------------------------------- image layers as 1x1 GEMMs
    def img_canon(self):
        """Canonical [in][out] matrix of the dense stand-in: the conv kernel itself ([3,3,3,C] = [27][C]) or its permuted copy."""
        return self.w_canon if self.img_few_in else self.img_wp

    def pre_refresh(self):
        """many -> few: [kh, kw, C, co] -> [C][(kh, kw, co)] before the packing launch (one small copy kernel, capturable)."""
        if self.img_op is not None and not self.img_few_in and self.w_canon is not None:
            cm, cf = self.Cin, self.Cout
            self.img_wp.view(cm, 3, 3, 3)[:, :, :, :cf].copy_(self.w_canon.view(3, 3, cm, cf).permute(2, 0, 1, 3))

    def _img_buf(self, key, maker):
        if key not in self._img_bufs:
            self._img_bufs[key] = maker()
        return self._img_bufs[key]

    def _img_tapsum(self, T, nimg, flip, dst, sigma, alpha_k, bias, act, aux, aux_mode, colsum, out_mode):
        if (out_mode == 0) != (dst.dtype in (torch.bfloat16, torch.float16)):
            raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'out_mode 0 writes planes, out_mode 2 one fp32 plane')
        a = (_ptr(aux), plane_stride(aux), aux.shape[0], fmt_of(aux)) if aux is not None else (None, 0, 0, 0)
        check(lib().mmdgan_tapsum3x3_small(_ptr(T), nimg, self.Hin, self.Win, 1 if flip else 0, float(alpha_k), _ptr(sigma), _ptr(bias), act,
                                           a[0], a[1], a[2], a[3], aux_mode, _ptr(dst), plane_stride(dst), dst.shape[0],
                                           fmt_of(dst) if out_mode == 0 else 0, dst.shape[2], out_mode, _ptr(colsum), _ptr(self.sat_flag), stream()))

    def _img(self, fwd, src, nimg, dst, sigma, alpha_k, bias, act, aux, aux_mode, colsum, out_mode):
        px = nimg * self.Hin * self.Win
        op = self.img_op
        op.sat_flag = self.sat_flag
        rows = lambda t, c: t.as_strided((t.shape[0], px, c), (t.stride(0), c, 1), t.storage_offset())
        assert fwd != self.img_few_in       # many -> few only: forward of C -> 3, or input gradient of 3 -> C
        T = self._img_buf(('T', nimg, fwd), lambda: torch.zeros((1, px, 32), dtype=torch.float32, device=src.device))
        if fwd:
            op.forward(rows(src, src.shape[2]), px, T, out_mode=2)
        else:
            op.dgrad(rows(src, src.shape[2]), px, T, out_mode=2)
        self._img_tapsum(T, nimg, not fwd, dst, sigma, alpha_k, bias, act, aux, aux_mode, colsum, out_mode)

    def pack(self, w_canon):
        """canonical weights -> forward and input-gradient GEMM operands (unscaled; act_k / sigma is an epilogue alpha)."""
        self.w_canon = w_canon
        if self.img_op is not None:
            self.pre_refresh()
            self.img_op.pack(self.img_canon())
        for g in (self.f, self.d):
            d = PackDesc()
            d.w, d.out = _ptr(w_canon), _ptr(g['w'])
            d.plane, d.npl, d.fmt = plane_stride(g['w']), g['w'].shape[0], fmt_of(g['w'], 'w')
            d.mode, d.k, d.Cin, d.Cout, d.Cs = g['mode'], self.k, self.Cin, self.Cout, g['Cs']
            d.rows_pad, d.kpad, d.classes = g['rows_pad'], g['kpad'], g['classes']
            if self.op == 'd':
                d.in_C, d.in_HW = self.in_flat
                d.out_C, d.out_HW = self.out_flat
            else:
                d.in_C = d.in_HW = d.out_C = d.out_HW = 1
            check(lib().mmdgan_pack_weights(C.byref(d), stream()))

    # -------------------------------------------------------------------------------------------- GEMM launches
    def _gemm(self, g, src, nimg, dst, geom, sigma, alpha_k, bias, act, aux, aux_mode, aux_wrap, colsum, colsumsq,
              colsum_rows, out_mode, npass):
        d = GemmDesc()
        _planes(src)
        d.src, d.src_plane = _ptr(src), plane_stride(src)
        d.src_fmt, d.w_fmt = fmt_of(src), fmt_of(g['w'], 'w')
        d.Nimg = nimg
        (d.Hs, d.Ws, d.Hg, d.Wg, d.sy, d.sx, d.TH, d.TW, d.Hd, d.Wd, d.osy, d.osx) = geom['dims']
        d.Cs = g['Cs']
        assert src.shape[2] == g['Cs'], 'source channels {} != {}'.format(src.shape[2], g['Cs'])
        assert src.shape[1] >= nimg * d.Hs * d.Ws
        d.w, d.w_plane, d.w_rows = _ptr(g['w']), plane_stride(g['w']), g['w'].shape[1]
        d.kpad, d.classes = g['kpad'], g['classes']
        d.dst, d.dst_plane, d.dst_npl = _ptr(dst), plane_stride(dst), dst.shape[0]
        d.dst_fmt = fmt_of(dst) if out_mode == 0 else 0
        if (out_mode == 0) != (dst.dtype in (torch.bfloat16, torch.float16)) or (out_mode == 2 and (dst.dtype != torch.float32 or dst.shape[0] != 1)):
            raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'out_mode 0 writes bf16 planes, out_mode 2 one fp32 plane')
        d.Cd, d.Ncols = dst.shape[2], g['ncols']
        assert dst.shape[1] >= nimg * d.Hd * d.Wd and dst.shape[2] >= g['ncols']
        d.alpha_k = float(alpha_k) / (FMT_SCALE[d.src_fmt] * FMT_SCALE[d.w_fmt])     # the products carry the operands' plane scales
        d.sigma, d.bias, d.act = _ptr(sigma), _ptr(bias), act
        if aux is not None:
            _planes(aux)
            d.aux, d.aux_plane, d.aux_npl, d.aux_fmt = _ptr(aux), plane_stride(aux), aux.shape[0], fmt_of(aux)
        d.aux_mode = aux_mode
        d.aux_wrap_at, d.aux_wrap_len = aux_wrap if aux_wrap else (0, 0)
        d.colsum, d.colsumsq, d.colsum_rows = _ptr(colsum), _ptr(colsumsq), colsum_rows
        bn, pair = g['bn'], 0
        m_tiles = (nimg * d.Hg * d.Wg + 127) // 128
        if GEMM_PAIR and g['ncols'] % 128 == 0 and m_tiles * (g['ncols'] // 128) * g['classes'] >= GEMM_PAIR_MIN_TILES:
            # tcgen05 cta_group::2: a 2-CTA cluster shares a 256 x bn tile (half the weight bytes per CTA).  bn = 256 doubles
            # the epilogue per CTA: measured to pay off for plain forward epilogues and for very wide layers only
            pair = 1
            bn = 256 if (g['ncols'] % 256 == 0 and (g['ncols'] >= 512 or aux is None or PAIR_BN256_AUX)) else 128
        elif GEMM_PAIR and PAIR_N64 and g['ncols'] == 64 and m_tiles * g['classes'] >= GEMM_PAIR_MIN_TILES:
            # N = 64 layers are bound by the traffic of the gathered operand: a CTA pair halves the weight bytes per CTA
            pair, bn = 1, 64
        elif GEMM_BN_MAX >= 256 and g['ncols'] % 256 == 0 and npass != 6:
            # 128 x 256 tiles halve the re-reads of the gathered operand, but leave only two pipeline stages per CTA and
            # double the epilogue: measured to pay off only while the grid still fills the 2 x 148 CTA slots
            tiles = m_tiles * (g['ncols'] // 256) * g['classes']
            if tiles >= 190 and (g['ncols'] >= 512 or aux is None):
                bn = 256
        d.cta_pair = pair
        d.out_mode, d.bn, d.npass = out_mode, bn, npass
        d.sat_flag = _ptr(self.sat_flag)
        if src.shape[0] < fmt_need(d.src_fmt, npass) or g['w'].shape[0] < fmt_need(d.w_fmt, npass) or (npass == 6 and (d.src_fmt or d.w_fmt)):
            raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'npass {} does not fit operands {} x {} / {} x {}'.format(
                npass, src.shape[0], src.dtype, g['w'].shape[0], g['w'].dtype))
        for i, (oy, ox, ooy, oox) in enumerate(geom['cls']):
            d.cls[i].oy, d.cls[i].ox, d.cls[i].ooy, d.cls[i].oox = oy, ox, ooy, oox
            d.cls[i].wrow = i * g['rows_pad']
        check(lib().mmdgan_gather_gemm(C.byref(d), stream()))

    def _fwd_geom(self):
        if self.op == 'd':
            return dict(dims=(1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1), cls=[(0, 0, 0, 0)])
        if self.op == 'c':
            return dict(dims=(self.Hin, self.Win, self.Hout, self.Wout, self.s, self.s, self.k, self.k, self.Hout, self.Wout, 1, 1),
                        cls=[(-self.pad, -self.pad, 0, 0)])
        return dict(dims=(self.Hin, self.Win, self.Hin, self.Win, 1, 1, 2, 2, self.Hout, self.Wout, 2, 2),
                    cls=[(ph - 1, pw - 1, ph, pw) for ph in (0, 1) for pw in (0, 1)])

    def _dgrad_geom(self):
        if self.op == 'd':
            return dict(dims=(1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1), cls=[(0, 0, 0, 0)])
        if self.op == 'c' and self.s == 1:
            o = -(self.k - 1 - self.pad)
            return dict(dims=(self.Hout, self.Wout, self.Hin, self.Win, 1, 1, self.k, self.k, self.Hin, self.Win, 1, 1),
                        cls=[(o, o, 0, 0)])
        if self.op == 'c':
            return dict(dims=(self.Hout, self.Wout, self.Hout, self.Wout, 1, 1, 2, 2, self.Hin, self.Win, 2, 2),
                        cls=[(ph - 1, pw - 1, ph, pw) for ph in (0, 1) for pw in (0, 1)])
        return dict(dims=(self.Hout, self.Wout, self.Hin, self.Win, 2, 2, 4, 4, self.Hin, self.Win, 1, 1), cls=[(-1, -1, 0, 0)])

    def _direct(self, fwd, src, nimg, dst, sigma, alpha_k, bias, act, aux, aux_mode, colsum, out_mode):
        """3x3 / stride-1 image-channel layer.  many -> few channels (forward of C -> 3, input gradient of 3 -> C): a dense [C -> 27]
        product on the tensor-core GEMM + the tap-sum kernel (measured 0.057 / 0.058 ms against 0.091 / 0.087 for the direct kernel);
        few -> many and the weight gradients stay on the direct CUDA-core kernel / the general weight-gradient GEMM (the im2col27 +
        dense [27 -> C] form was measured SLOWER there: 0.173 vs 0.150 ms forward, 0.110 vs 0.077 ms weight gradient)."""
        if self.img_op is not None and fwd != self.img_few_in:
            return self._img(fwd, src, nimg, dst, sigma, alpha_k, bias, act, aux, aux_mode, colsum, out_mode)
        d = DirectDesc()
        _planes(src)
        d.src, d.src_plane, d.src_npl, d.Cs = _ptr(src), plane_stride(src), src.shape[0], src.shape[2]
        d.src_fmt = fmt_of(src)
        d.N, d.H, d.W = nimg, self.Hin, self.Win
        ci, co = self.Cin, self.Cout
        d.w, d.w_tap = _ptr(self.w_canon), ci * co
        if fwd:
            d.Cin, d.Cout, d.w_in, d.w_out, d.flip = ci, co, co, 1, 0
        else:
            d.Cin, d.Cout, d.w_in, d.w_out, d.flip = co, ci, 1, co, 1
        assert src.shape[1] >= nimg * self.Hin * self.Win and dst.shape[1] >= nimg * self.Hin * self.Win
        if (out_mode == 0) != (dst.dtype in (torch.bfloat16, torch.float16)):
            raise _lib.MmdganError(_lib.MMDGAN_ESHAPE, 'out_mode 0 writes planes, out_mode 2 one fp32 plane')
        d.dst, d.dst_plane, d.dst_npl, d.Cd, d.out_mode = _ptr(dst), plane_stride(dst), dst.shape[0], dst.shape[2], out_mode
        d.dst_fmt = fmt_of(dst) if out_mode == 0 else 0
        d.alpha_k, d.sigma, d.bias, d.act = float(alpha_k), _ptr(sigma), _ptr(bias), act
        if aux is not None:
            _planes(aux)
            d.aux, d.aux_plane, d.aux_npl, d.aux_fmt = _ptr(aux), plane_stride(aux), aux.shape[0], fmt_of(aux)
        d.aux_mode = aux_mode
        d.colsum = _ptr(colsum)
        d.sat_flag = _ptr(self.sat_flag)
        check(lib().mmdgan_direct_conv(C.byref(d), stream()))

    def fwd_tiles(self, nimg):
        g = self._fwd_geom()['dims']
        return lib().mmdgan_gather_gemm_tiles(nimg, g[2], g[3]) * self.f['classes']

    def dgrad_tiles(self, nimg):
        if self.direct_d == 'ls':       # rows of the per-block column-sum workspace of the direct / tap-sum kernel
            if self.img_op is not None and self.img_few_in:
                return lib().mmdgan_tapsum_blocks(nimg, self.Hin, self.Win)
            return lib().mmdgan_direct_conv_blocks(nimg, self.Hin, self.Win)
        g = self._dgrad_geom()['dims']
        return lib().mmdgan_gather_gemm_tiles(nimg, g[2], g[3]) * self.d['classes']

    def forward(self, src, nimg, dst, sigma=None, alpha_k=1.0, bias=None, act=0, colsum=None, colsumsq=None, out_mode=0):
        if (self.op == 'd' and self.Cs_out in (8, 16, 32) and out_mode == 2 and act == 0 and colsum is None
                and self.npass == 3 and nimg <= 4096):      # (many rows x few columns, e.g. the 27-column image stand-in: tensor cores)
            # a handful of output columns (the critic scores): fp32 CUDA-core kernel instead of a 94 %-padded MMA tile
            _planes(src)
            npl = min(src.shape[0], self.f['w'].shape[0])
            need = int(lib().mmdgan_dense_small_workspace(nimg, self.Cs_in, self.Cs_out)) // 4
            if self._ds_ws is None or self._ds_ws.numel() < need:        # K-slice partials; sized on the first (eager) call
                self._ds_ws = torch.empty(need, dtype=torch.float32, device=src.device)
            check(lib().mmdgan_dense_small_fwd(_ptr(src), plane_stride(src), npl, fmt_of(src), nimg, self.Cs_in, _ptr(self.f['w']),
                                               plane_stride(self.f['w']), fmt_of(self.f['w'], 'w'), self.f['kpad'], self.Cs_out,
                                               float(alpha_k), _ptr(sigma), _ptr(bias), _ptr(dst), dst.shape[2], _ptr(self._ds_ws), stream()))
            return
        if self.direct_f and colsum is None and colsumsq is None and self.w_canon is not None:
            return self._direct(True, src, nimg, dst, sigma, alpha_k, bias, act, None, 0, None, out_mode)
        self._gemm(self.f, src, nimg, dst, self._fwd_geom(), sigma, alpha_k, bias, act, None, 0, None, colsum, colsumsq, 0, out_mode,
                   self.fwd_npass)

    def dgrad(self, dy, nimg, dst, sigma=None, alpha_k=1.0, aux=None, aux_mode=0, aux_wrap=None, colsum=None, colsum_rows=0,
              out_mode=0, npass=None):
        """Input gradient (npass defaults to the 3-pair gradient mode; the spectral-norm power iteration, which uses the
        adjoint as a FORWARD operator, passes the 6-pair mode)."""
        if (self.direct_d and aux_wrap is None and not colsum_rows and self.w_canon is not None
                and (self.direct_d == 'ls' or (aux is None and colsum is None))):
            return self._direct(False, dy, nimg, dst, sigma, alpha_k, None, 0, aux, aux_mode, colsum, out_mode)
        self._gemm(self.d, dy, nimg, dst, self._dgrad_geom(), sigma, alpha_k, None, 0, aux, aux_mode, aux_wrap, colsum, None,
                   colsum_rows, out_mode, self.bwd_npass if npass is None else npass)

    # -------------------------------------------------------------------------------------------- weight gradient
    def wgrad_plan(self, nimg):
        """(R, NC, bn, splits, P) of the weight-gradient launch for nimg samples."""
        if self.op == 'd':
            P = nimg
            R, NC = (self.Cs_in, self.Cs_out) if self.w_swapped else (self.Cs_out, self.Cs_in)
        elif self.op == 'c':
            if self.w_swapped:
                P, R, NC = nimg * self.Hin * self.Win, self.Cs_in, self.k * self.k * self.Cs_out
            else:
                P, R, NC = nimg * self.Hout * self.Wout, self.Cs_out, self.k * self.k * self.Cs_in
        else:
            P, R, NC = nimg * self.Hin * self.Win, self.Cs_in, self.k * self.k * self.Cs_out
        bn = pick_bn(NC, lo=64, hi=WGRAD_BN_MAX)
        tiles = ((R + 127) // 128) * ((NC + bn - 1) // bn)
        ksteps = (P + 31) // 32
        splits = max(1, min((WGRAD_CTAS + tiles - 1) // tiles, max(1, ksteps // 8)))
        return R, NC, bn, splits, P

    def wgrad(self, x_in, dy, nimg, partials, splits=None, scratch=None):
        """partials [splits, R, NC] <- weight-gradient GEMM of (layer input x_in, output gradient dy).  An fp16-plane operand
        (forward activations) is first re-split into two bf16 planes -- into `scratch` (a flat bf16 buffer) when given."""
        ops = []
        for i, t in enumerate((x_in, dy)):
            if t.dtype == torch.float16:
                n = t.shape[1] * t.shape[2]
                if scratch is not None and i == 0:
                    assert scratch.numel() >= 2 * n
                    buf = scratch[:2 * n].view(2, t.shape[1], t.shape[2])
                else:
                    buf = new_planes(t.shape[1], t.shape[2], 2, t.device)
                t = convert_planes(t, buf)
            ops.append(t)
        x_in, dy = ops
        R, NC, bn, sp, P = self.wgrad_plan(nimg)
        splits = splits or sp
        d = WgradDesc()
        if self.op == 'd':
            plain, gath = (x_in, dy) if self.w_swapped else (dy, x_in)
            geo = (1, 1, 1, 1, 1, 1, 1, 1, 0, 0)
        elif self.op == 'c':
            if self.w_swapped:
                plain, gath = x_in, dy
                o = -(self.k - 1 - self.pad)
                geo = (self.Hout, self.Wout, self.Hin, self.Win, 1, 1, self.k, self.k, o, o)
            else:
                plain, gath = dy, x_in
                geo = (self.Hin, self.Win, self.Hout, self.Wout, self.s, self.s, self.k, self.k, -self.pad, -self.pad)
        else:
            plain, gath = x_in, dy
            geo = (self.Hout, self.Wout, self.Hin, self.Win, 2, 2, 4, 4, -1, -1)
        _planes(plain)
        _planes(gath)
        d.plain, d.plain_plane, d.P, d.Cp = _ptr(plain), plane_stride(plain), P, R
        assert plain.shape[2] == R and plain.shape[1] >= P
        d.g, d.g_plane = _ptr(gath), plane_stride(gath)
        d.Nimg = nimg
        d.Hs, d.Ws, d.Hg, d.Wg, d.sy, d.sx, d.TH, d.TW, d.oy, d.ox = geo
        d.Cs = gath.shape[2]
        assert d.TH * d.TW * d.Cs == NC
        d.splits, d.out, d.bn, d.npass = splits, _ptr(partials), bn, self.bwd_npass
        d.p_fmt, d.g_fmt = fmt_of(plain), fmt_of(gath)
        self._wg_scale = 1.0 / (FMT_SCALE[d.p_fmt] * FMT_SCALE[d.g_fmt])      # removed by the wgrad_reduce that follows
        assert partials.numel() >= splits * R * NC
        check(lib().mmdgan_wgrad_gemm(C.byref(d), stream()))
        return splits

    def wgrad_reduce(self, partials, splits, nimg, out_canon, w_canon=None, dots=None):
        """split partials -> gradient in the canonical (reference) weight layout; optional per-block <G, W>."""
        d, nblocks = self.wgrad_reduce_desc(partials, splits, nimg, out_canon, w_canon, dots)
        check(lib().mmdgan_wgrad_reduce(C.byref(d), stream()))
        return nblocks

    def wgrad_reduce_desc(self, partials, splits, nimg, out_canon, w_canon=None, dots=None):
        """(descriptor, number of blocks = length of `dots`) of that reduction, for a single launch or a batched job table."""
        R, NC, _, _, _ = self.wgrad_plan(nimg)
        d = WredDesc()
        d.partials, d.splits, d.R, d.NC = _ptr(partials), splits, R, NC
        d.scale = self._wg_scale
        d.r_perm_C = d.c_perm_C = 1
        d.r_perm_HW = d.c_perm_HW = 1
        ci, co = self.Cin, self.Cout
        if self.op == 'd':
            if self.w_swapped:      # r = in', c = out'
                d.Cg, d.Cvalid, d.Rvalid = self.Cs_out, co, ci
                d.base, d.sr, d.st, d.sc = 0, co, 0, 1
                (d.r_perm_C, d.r_perm_HW), (d.c_perm_C, d.c_perm_HW) = self.in_flat, self.out_flat
            else:                   # r = out', c = in'
                d.Cg, d.Cvalid, d.Rvalid = self.Cs_in, ci, co
                d.base, d.sr, d.st, d.sc = 0, 1, 0, co
                (d.r_perm_C, d.r_perm_HW), (d.c_perm_C, d.c_perm_HW) = self.out_flat, self.in_flat
        elif self.op == 'c':
            if self.w_swapped:      # r = ci, t mirrored, c = co
                d.Cg, d.Cvalid, d.Rvalid = self.Cs_out, co, ci
                d.base, d.sr, d.st, d.sc = (self.k * self.k - 1) * ci * co, co, -ci * co, 1
            else:                   # r = co, t = (kh,kw), c = ci
                d.Cg, d.Cvalid, d.Rvalid = self.Cs_in, ci, co
                d.base, d.sr, d.st, d.sc = 0, 1, ci * co, co
        else:                       # tc: r = ci, t = (kh,kw), c = co ; canon [k,k,Cout,Cin]
            d.Cg, d.Cvalid, d.Rvalid = self.Cs_out, co, ci
            d.base, d.sr, d.st, d.sc = 0, 1, co * ci, ci
        d.w, d.out, d.dots = (w_canon.data_ptr() if w_canon is not None else None), out_canon.data_ptr(), (dots.data_ptr() if dots is not None else None)
        for t in (w_canon, out_canon, dots):
            _ptr(t)             # validation only
        return d, lib().mmdgan_wgrad_reduce_blocks(R * NC)


# ------------------------------------------------------------------------------------------------ small wrappers
def build_wred_jobs(descs_blocks, device):
    """Device job table of mmdgan_wgrad_reduce_batched: [(WredDesc, nblocks)] -> (blob, start, njobs, total blocks)."""
    n = len(descs_blocks)
    arr = (WredDesc * n)()
    start = [0]
    for i, (d, nb) in enumerate(descs_blocks):
        arr[i] = d
        start.append(start[-1] + int(nb))
    blob = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    st = torch.tensor(start, dtype=torch.int32, device=device)
    return blob, st, n, start[-1]


def wgrad_reduce_batched(blob, start, njobs, total):
    check(lib().mmdgan_wgrad_reduce_batched(C.c_void_p(blob.data_ptr()), _ptr(start), njobs, total, stream()))


def build_sn_combine_jobs(items, device):
    """items: [(g, s, dots, ndots, sigma, act_k, n)] -> (blob, njobs, blocks)."""
    n = len(items)
    arr = (_lib.SnCombineJob * n)()
    blocks = 1
    for i, (g, s, dots, ndots, sigma, act_k, cnt) in enumerate(items):
        for t in (g, s, dots, sigma):
            _ptr(t)
        arr[i].g, arr[i].s, arr[i].dots, arr[i].sigma = g.data_ptr(), s.data_ptr(), dots.data_ptr(), sigma.data_ptr()
        arr[i].n, arr[i].ndots, arr[i].act_k = int(cnt), int(ndots), float(act_k)
        blocks = max(blocks, min(1184, (int(cnt) + 1023) // 1024))
    blob = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return blob, n, blocks


def sn_grad_combine_batched(blob, njobs, blocks):
    check(lib().mmdgan_sn_grad_combine_batched(C.c_void_p(blob.data_ptr()), njobs, blocks, stream()))


def reduce_tiles(partials, T, Cc, out, scale=1.0):
    check(lib().mmdgan_reduce_tiles(_ptr(partials), T, Cc, float(scale), _ptr(out), stream()))


def colsum_small(x, rows, Cc, out):
    # column sums of a [rows, C] matrix = the tile reduction with one "tile" per row
    check(lib().mmdgan_reduce_tiles(_ptr(x), rows, Cc, 1.0, _ptr(out), stream()))


def colsum_planes(x, rows, Cc, out):
    # column sums of the values carried by bf16 planes [npl, rows, C]
    _planes(x)
    assert x.dtype == torch.bfloat16, 'gradients are bf16 planes'
    check(lib().mmdgan_colsum_planes(_ptr(x), plane_stride(x), x.shape[0], rows, Cc, _ptr(out), stream()))


def build_refresh_jobs(pack_descs, permutes, device):
    """Device-resident job table for mmdgan_refresh: pack_descs = [PackDesc], permutes = [(src, dst, n, C, HW)].
    Returns (table, njobs, total_blocks): job j owns blocks [block_start_j, block_start_{j+1}) of the flat grid -- one block per
    32 x 64 tile of a packed operand (at most 2048 per job: larger operands stride), one per 256 features of a vector."""
    n = len(pack_descs) + len(permutes)
    assert 0 < n <= 256
    arr = (RefreshJob * n)()
    start = 0
    for i, d in enumerate(pack_descs):
        arr[i].kind = 0
        arr[i].pack = d
        arr[i].block_start = start
        start += max(1, min(2048, ((d.rows_pad * d.classes + 31) // 32) * ((d.kpad + 63) // 64)))
    for j, (src, dst, cnt, Cc, HW) in enumerate(permutes):
        job = arr[len(pack_descs) + j]
        job.kind = 1
        job.block_start = start
        job.src, job.dst = _ptr(src), _ptr(dst)
        job.n, job.C, job.HW, job.inverse = cnt, Cc, HW, 0
        start += max(1, min(64, (int(cnt) + 255) // 256))
    blob = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    return blob, n, start


def refresh(blob, njobs, total_blocks):
    check(lib().mmdgan_refresh(C.c_void_p(blob.data_ptr()), njobs, total_blocks, stream()))


def sn_normalize(v, n, out, sigma_out=None, eps=1e-10):
    _planes(out)
    check(lib().mmdgan_sn_normalize(_ptr(v), n, float(eps), _ptr(sigma_out), _ptr(out), plane_stride(out), out.shape[0], fmt_of(out), stream()))


def sn_grad_combine(g, s, dots, ndots, sigma, act_k, n):
    check(lib().mmdgan_sn_grad_combine(_ptr(g), _ptr(s), _ptr(dots), ndots, _ptr(sigma), float(act_k), n, stream()))


def scale_by_sigma(g, sigma, act_k, n):
    check(lib().mmdgan_scale_by_sigma(_ptr(g), _ptr(sigma), float(act_k), n, stream()))


def permute_features(src, dst, n, Cc, HW, inverse=False):
    check(lib().mmdgan_permute_features(_ptr(src), _ptr(dst), n, Cc, HW, 1 if inverse else 0, stream()))


def bn_finalize(psum, psq, T, Cc, rows, mean, invstd, moving_mean=None, moving_var=None, eps=1e-3, momentum=0.99, bessel=True):
    """bessel: Bessel-corrected variance into the moving average (TF's fused kernel, rank-4 inputs); False for the rank-2 inputs
    of a dense layer, where TF 1.8 falls back to nn.moments (biased)."""
    check(lib().mmdgan_bn_finalize(_ptr(psum), _ptr(psq), T, Cc, rows, float(eps), float(momentum), _ptr(mean), _ptr(invstd),
                                   _ptr(moving_mean), _ptr(moving_var), 1 if bessel else 0, stream()))


def bn_inference_stats(moving_mean, moving_var, Cc, mean, invstd, eps=1e-3):
    """tf.layers.batch_normalization(training=False): the moving averages normalise (layer_func.py:953-966)."""
    check(lib().mmdgan_bn_inference_stats(_ptr(moving_mean), _ptr(moving_var), Cc, float(eps), _ptr(mean), _ptr(invstd), stream()))


def bn_apply(z, mean, invstd, gamma, beta, Cc, total, act, out, sat_flag=None):
    _planes(out)
    check(lib().mmdgan_bn_apply(_ptr(z), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), Cc, total, act, _ptr(out),
                                plane_stride(out), out.shape[0], fmt_of(out), _ptr(sat_flag), stream()))


def bn_bwd_reduce(da, z, mean, invstd, gamma, beta, Cc, rows, rows_per_block, act, psum, psumx):
    check(lib().mmdgan_bn_bwd_reduce(_ptr(da), _ptr(z), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), Cc, rows,
                                     rows_per_block, act, _ptr(psum), _ptr(psumx), stream()))


def bn_bwd_apply(da, z, mean, invstd, gamma, beta, dbeta, dgamma, Cc, rows, act, out):
    _planes(out)
    assert out.dtype == torch.bfloat16, 'gradients are bf16 planes'
    check(lib().mmdgan_bn_bwd_apply(_ptr(da), _ptr(z), _ptr(mean), _ptr(invstd), _ptr(gamma), _ptr(beta), _ptr(dbeta),
                                    _ptr(dgamma), Cc, rows, act, _ptr(out), plane_stride(out), out.shape[0], stream()))


def adam(w, m, v, g, n, lr, step, beta1=0.5, beta2=0.999, eps=1e-8):
    check(lib().mmdgan_adam(_ptr(w), _ptr(m), _ptr(v), _ptr(g), n, float(lr), float(beta1), float(beta2), float(eps), _ptr(step),
                            stream()))


def adam_allreduce_nvls(w, m, v, g_mc, w_mc, m_mc, v_mc, begin, end, lr, step, beta1=0.5, beta2=0.999, eps=1e-8):
    """Gradient all-reduce fused with Adam through NVSwitch multicast (csrc/nvls.cu).  w / m / v: this rank's replicas;
    *_mc: integer multicast addresses of the four buffers; [begin, end): this rank's shard of the flat buffer."""
    check(lib().mmdgan_adam_allreduce_nvls(_ptr(w), _ptr(m), _ptr(v), C.c_void_p(g_mc), C.c_void_p(w_mc), C.c_void_p(m_mc),
                                           C.c_void_p(v_mc), int(begin), int(end), float(lr), float(beta1), float(beta2),
                                           float(eps), _ptr(step), stream()))


def scatter_scores_nvls(s_local, b, rank, gen_all_mc, real_all_mc):
    """This rank's [2b, d] scores -> rows [rank*b, (rank+1)*b) of real_all / gen_all on every rank (csrc/nvls.cu)."""
    check(lib().mmdgan_scatter_scores_nvls(_ptr(s_local), int(b), int(s_local.shape[1]), int(rank), C.c_void_p(gen_all_mc),
                                           C.c_void_p(real_all_mc), stream()))


def allreduce_small_nvls(out, in_mc, n):
    check(lib().mmdgan_allreduce_small_nvls(_ptr(out), C.c_void_p(in_mc), int(n), stream()))


def sample_normal(out, seed, draw_counter=None, raw=None):
    """out (float32, any shape) <- N(0, 1) samples drawn on the device (Philox-4x32-10 + Box-Muller; csrc/elementwise.cu)."""
    check(lib().mmdgan_sample_normal(_ptr(out), out.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(draw_counter), _ptr(raw), stream()))


def incr_counter(counter):
    check(lib().mmdgan_incr_counter(_ptr(counter), stream()))


def losses_from_sums(sums, cD, losses):
    check(lib().mmdgan_losses_from_sums(_ptr(sums), float(cD[0]), float(cD[1]), float(cD[2]), _ptr(losses), stream()))


def incr_step(step):
    check(lib().mmdgan_incr_step(_ptr(step), stream()))


def nan_flag(x, n, flag):
    check(lib().mmdgan_nan_flag(_ptr(x), n, _ptr(flag), stream()))


# ------------------------------------------------------------------------------------------------ MMD
class MmdKernel(object):
    """Configured fused MMD loss kernel (mmdgan_mmd_fwd_bwd); owns its small workspace."""

    def __init__(self, loss_type, rep_weights=(0.0, -1.0), b=64, device='cuda', sigma=None, alpha=None, beta=None):
        """sigma: bandwidth list of the Gaussian mixture ('mmd_g' / 'fixed_g', GANLoss.sigma, math_func.py:2108); alpha / beta:
        scales and exponent offset of the t-distribution mixture ('mmd_t' / 'fixed_t', math_func.py:2109-2110).  None keeps the
        reference defaults; the single-bandwidth losses (rep, rmb, mgb) are fixed at sigma = 1 as in the reference."""
        self.desc = MmdDesc()
        _lib.check(lib().mmdgan_mmd_configure(C.byref(self.desc), loss_type.encode(), float(rep_weights[0]), float(rep_weights[1])))
        scales = sigma if self.desc.family == 0 else alpha
        if scales is not None and self.desc.n_sigma > 1:
            scales = [float(s) for s in scales]
            if not 1 <= len(scales) <= 8 or min(scales) <= 0.0:
                raise _lib.MmdganError(_lib.MMDGAN_EINVAL, 'the fused MMD kernel takes 1..8 positive kernel scales, got {}'.format(scales))
            self.desc.n_sigma = len(scales)
            for i in range(8):
                self.desc.sigma[i] = scales[i] if i < len(scales) else 0.0
        if beta is not None and self.desc.family == 1:
            self.desc.beta = float(beta)
        self.b = b
        self.ws = torch.zeros(int(lib().mmdgan_mmd_workspace(b)) // 4 + 4, dtype=torch.float32, device=device)
        self.sums = torch.zeros(6, dtype=torch.float32, device=device)
        self.losses = torch.zeros(2, dtype=torch.float32, device=device)

    def __call__(self, gen_loc, real_loc, dLg_dgen, dLd_dgen, dLd_dreal, dLg_dreal=None, gen_all=None, real_all=None, row0=0):
        d = self.desc
        b, dim = gen_loc.shape
        assert b <= self.b
        gen_all = gen_loc if gen_all is None else gen_all
        real_all = real_loc if real_all is None else real_all
        d.gen_loc, d.real_loc, d.gen_all, d.real_all = _ptr(gen_loc), _ptr(real_loc), _ptr(gen_all), _ptr(real_all)
        d.b, d.Bg, d.row0, d.d = b, gen_all.shape[0], row0, dim
        d.sums, d.losses = _ptr(self.sums), _ptr(self.losses)
        d.dLg_dgen, d.dLg_dreal, d.dLd_dgen, d.dLd_dreal = _ptr(dLg_dgen), _ptr(dLg_dreal), _ptr(dLd_dgen), _ptr(dLd_dreal)
        d.workspace = _ptr(self.ws)
        check(lib().mmdgan_mmd_fwd_bwd(C.byref(d), stream()))
        return self.losses
