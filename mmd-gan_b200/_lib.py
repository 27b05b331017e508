"""ctypes binding of libmmdgan_b200.so (the C ABI declared in include/mmdgan_b200.h).

The library is built in-tree by `make -C mmd-gan_b200/csrc` (or __graft_entry__.build()).  There is NO CPU or
PyTorch fallback: if the shared object is missing, or the device is not sm_100, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmmdgan_b200.so')

MMDGAN_OK, MMDGAN_EINVAL, MMDGAN_ESHAPE, MMDGAN_EARCH, MMDGAN_ECUDA, MMDGAN_ENCCL = 0, -1, -2, -3, -4, -5


class MmdganError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('mmdgan_b200 error {}: {}'.format(code, msg))
        self.code = code


class GemmClass(C.Structure):
    _fields_ = [('oy', C.c_int), ('ox', C.c_int), ('ooy', C.c_int), ('oox', C.c_int), ('wrow', C.c_int),
                ('pad0', C.c_int), ('pad1', C.c_int), ('pad2', C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [
        ('src', C.c_void_p), ('src_plane', C.c_longlong),
        ('Nimg', C.c_int), ('Hs', C.c_int), ('Ws', C.c_int), ('Cs', C.c_int),
        ('Hg', C.c_int), ('Wg', C.c_int), ('sy', C.c_int), ('sx', C.c_int), ('TH', C.c_int), ('TW', C.c_int),
        ('w', C.c_void_p), ('w_plane', C.c_longlong), ('w_rows', C.c_longlong),
        ('kpad', C.c_int), ('classes', C.c_int), ('src_fmt', C.c_int), ('w_fmt', C.c_int),
        ('dst', C.c_void_p), ('dst_plane', C.c_longlong), ('dst_npl', C.c_int), ('dst_fmt', C.c_int),
        ('Hd', C.c_int), ('Wd', C.c_int), ('Cd', C.c_int), ('osy', C.c_int), ('osx', C.c_int), ('Ncols', C.c_int),
        ('alpha_k', C.c_float), ('sigma', C.c_void_p), ('bias', C.c_void_p), ('act', C.c_int),
        ('aux', C.c_void_p), ('aux_plane', C.c_longlong), ('aux_npl', C.c_int), ('aux_fmt', C.c_int), ('aux_mode', C.c_int), ('aux_wrap_at', C.c_longlong), ('aux_wrap_len', C.c_longlong),
        ('colsum', C.c_void_p), ('colsumsq', C.c_void_p), ('colsum_rows', C.c_longlong),
        ('out_mode', C.c_int), ('bn', C.c_int), ('npass', C.c_int), ('sat_flag', C.c_void_p), ('cta_pair', C.c_int),
        ('cls', GemmClass * 4)]


class WgradDesc(C.Structure):
    _fields_ = [
        ('plain', C.c_void_p), ('plain_plane', C.c_longlong), ('P', C.c_longlong), ('Cp', C.c_int),
        ('g', C.c_void_p), ('g_plane', C.c_longlong),
        ('Nimg', C.c_int), ('Hs', C.c_int), ('Ws', C.c_int), ('Cs', C.c_int),
        ('Hg', C.c_int), ('Wg', C.c_int), ('sy', C.c_int), ('sx', C.c_int), ('TH', C.c_int), ('TW', C.c_int),
        ('oy', C.c_int), ('ox', C.c_int), ('splits', C.c_int),
        ('out', C.c_void_p), ('bn', C.c_int), ('npass', C.c_int), ('p_fmt', C.c_int), ('g_fmt', C.c_int)]


class DirectDesc(C.Structure):
    _fields_ = [
        ('src', C.c_void_p), ('src_plane', C.c_longlong), ('src_npl', C.c_int), ('Cs', C.c_int),
        ('src_fmt', C.c_int), ('dst_fmt', C.c_int), ('aux_fmt', C.c_int), ('pad0', C.c_int),
        ('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int), ('Cout', C.c_int),
        ('w', C.c_void_p), ('w_tap', C.c_longlong), ('w_in', C.c_longlong), ('w_out', C.c_longlong), ('flip', C.c_int),
        ('dst', C.c_void_p), ('dst_plane', C.c_longlong), ('dst_npl', C.c_int), ('Cd', C.c_int), ('out_mode', C.c_int),
        ('alpha_k', C.c_float), ('sigma', C.c_void_p), ('bias', C.c_void_p), ('act', C.c_int),
        ('aux', C.c_void_p), ('aux_plane', C.c_longlong), ('aux_npl', C.c_int), ('aux_mode', C.c_int),
        ('colsum', C.c_void_p), ('sat_flag', C.c_void_p)]


class WredDesc(C.Structure):
    _fields_ = [
        ('partials', C.c_void_p), ('scale', C.c_float), ('splits', C.c_int), ('R', C.c_int), ('NC', C.c_int), ('Cg', C.c_int),
        ('Cvalid', C.c_int), ('Rvalid', C.c_int),
        ('r_perm_C', C.c_int), ('r_perm_HW', C.c_int), ('c_perm_C', C.c_int), ('c_perm_HW', C.c_int),
        ('base', C.c_longlong), ('sr', C.c_longlong), ('st', C.c_longlong), ('sc', C.c_longlong),
        ('w', C.c_void_p), ('out', C.c_void_p), ('dots', C.c_void_p)]


class SnCombineJob(C.Structure):
    _fields_ = [('g', C.c_void_p), ('s', C.c_void_p), ('dots', C.c_void_p), ('sigma', C.c_void_p), ('n', C.c_longlong),
                ('ndots', C.c_int), ('act_k', C.c_float)]


class PackDesc(C.Structure):
    _fields_ = [
        ('w', C.c_void_p), ('out', C.c_void_p), ('plane', C.c_longlong), ('npl', C.c_int), ('fmt', C.c_int),
        ('mode', C.c_int), ('k', C.c_int), ('Cin', C.c_int), ('Cout', C.c_int), ('Cs', C.c_int),
        ('rows_pad', C.c_int), ('kpad', C.c_int), ('classes', C.c_int),
        ('in_C', C.c_int), ('in_HW', C.c_int), ('out_C', C.c_int), ('out_HW', C.c_int)]


class RefreshJob(C.Structure):
    _fields_ = [('kind', C.c_int), ('block_start', C.c_int), ('pack', PackDesc), ('src', C.c_void_p), ('dst', C.c_void_p),
                ('n', C.c_int), ('C', C.c_int), ('HW', C.c_int), ('inverse', C.c_int)]


class MmdDesc(C.Structure):
    _fields_ = [
        ('gen_loc', C.c_void_p), ('real_loc', C.c_void_p), ('gen_all', C.c_void_p), ('real_all', C.c_void_p),
        ('b', C.c_int), ('Bg', C.c_int), ('row0', C.c_int), ('d', C.c_int),
        ('n_sigma', C.c_int), ('sigma', C.c_float * 8), ('cD', C.c_float * 3), ('bmode', C.c_int * 3),
        ('bval', C.c_float * 3), ('family', C.c_int), ('beta', C.c_float),
        ('sums', C.c_void_p), ('losses', C.c_void_p),
        ('dLg_dgen', C.c_void_p), ('dLg_dreal', C.c_void_p), ('dLd_dgen', C.c_void_p), ('dLd_dreal', C.c_void_p),
        ('workspace', C.c_void_p)]


# name -> (restype, argtypes); every symbol include/mmdgan_b200.h declares
_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
SYMBOLS = {
    'mmdgan_last_error': (C.c_char_p, []),
    'mmdgan_version': (_I, []),
    'mmdgan_check_device': (_I, []),
    'mmdgan_nchw_to_nhwc': (_I, [_P, _P, _LL, _I, _I, _I, _I, _I, _I, _I, _P]),
    'mmdgan_nhwc_to_nchw': (_I, [_P, _LL, _I, _I, _P, _I, _I, _I, _I, _I, _P]),
    'mmdgan_to_planes': (_I, [_P, _P, _LL, _I, _I, _LL, _P]),
    'mmdgan_from_planes': (_I, [_P, _LL, _I, _I, _P, _LL, _P]),
    'mmdgan_convert_planes': (_I, [_P, _LL, _I, _I, _P, _LL, _I, _I, _LL, _P]),
    'mmdgan_pack_weights': (_I, [C.POINTER(PackDesc), _P]),
    'mmdgan_permute_features': (_I, [_P, _P, _I, _I, _I, _I, _P]),
    'mmdgan_refresh': (_I, [_P, _I, _LL, _P]),
    'mmdgan_dense_small_fwd': (_I, [_P, _LL, _I, _I, _I, _I, _P, _LL, _I, _I, _I, _F, _P, _P, _P, _I, _P, _P]),
    'mmdgan_dense_small_workspace': (C.c_size_t, [_I, _I, _I]),
    'mmdgan_losses_from_sums': (_I, [_P, _F, _F, _F, _P, _P]),
    'mmdgan_wgrad_reduce_batched': (_I, [_P, _P, _I, _I, _P]),
    'mmdgan_sn_grad_combine_batched': (_I, [_P, _I, _I, _P]),
    'mmdgan_tapsum3x3_small': (_I, [_P, _I, _I, _I, _I, _F, _P, _P, _I, _P, _LL, _I, _I, _I, _P, _LL, _I, _I, _I, _I, _P, _P, _P]),
    'mmdgan_tapsum_blocks': (_I, [_I, _I, _I]),
    'mmdgan_sample_normal': (_I, [_P, _LL, C.c_ulonglong, _P, _P, _P]),
    'mmdgan_incr_counter': (_I, [_P, _P]),
    'mmdgan_direct_conv': (_I, [C.POINTER(DirectDesc), _P]),
    'mmdgan_direct_conv_blocks': (_I, [_I, _I, _I]),
    'mmdgan_gather_gemm': (_I, [C.POINTER(GemmDesc), _P]),
    'mmdgan_gather_gemm_tiles': (_I, [_I, _I, _I]),
    'mmdgan_wgrad_gemm': (_I, [C.POINTER(WgradDesc), _P]),
    'mmdgan_wgrad_reduce': (_I, [C.POINTER(WredDesc), _P]),
    'mmdgan_wgrad_reduce_blocks': (_I, [_LL]),
    'mmdgan_sn_grad_combine': (_I, [_P, _P, _P, _I, _P, _F, _LL, _P]),
    'mmdgan_scale_by_sigma': (_I, [_P, _P, _F, _LL, _P]),
    'mmdgan_sn_normalize': (_I, [_P, _LL, _F, _P, _P, _LL, _I, _I, _P]),
    'mmdgan_reduce_tiles': (_I, [_P, _I, _I, _F, _P, _P]),
    'mmdgan_colsum_small': (_I, [_P, _I, _I, _P, _P]),
    'mmdgan_colsum_planes': (_I, [_P, _LL, _I, _I, _I, _P, _P]),
    'mmdgan_bn_finalize': (_I, [_P, _P, _I, _I, _LL, _F, _F, _P, _P, _P, _P, _I, _P]),
    'mmdgan_bn_inference_stats': (_I, [_P, _P, _I, _F, _P, _P, _P]),
    'mmdgan_bn_apply': (_I, [_P, _P, _P, _P, _P, _I, _LL, _I, _P, _LL, _I, _I, _P, _P]),
    'mmdgan_bn_bwd_reduce': (_I, [_P, _P, _P, _P, _P, _P, _I, _LL, _I, _I, _P, _P, _P]),
    'mmdgan_bn_bwd_apply': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _LL, _I, _P, _LL, _I, _P]),
    'mmdgan_mmd_configure': (_I, [C.POINTER(MmdDesc), C.c_char_p, _F, _F]),
    'mmdgan_mmd_workspace': (C.c_size_t, [_I]),
    'mmdgan_mmd_fwd_bwd': (_I, [C.POINTER(MmdDesc), _P]),
    'mmdgan_adam': (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _P, _P]),
    'mmdgan_adam_allreduce_nvls': (_I, [_P, _P, _P, _P, _P, _P, _P, _LL, _LL, _F, _F, _F, _F, _P, _P]),
    'mmdgan_scatter_scores_nvls': (_I, [_P, _I, _I, _I, _P, _P, _P]),
    'mmdgan_allreduce_small_nvls': (_I, [_P, _P, _I, _P]),
    'mmdgan_incr_step': (_I, [_P, _P]),
    'mmdgan_nan_flag': (_I, [_P, _I, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmdganError(MMDGAN_EINVAL, '{} not found: run `make -C mmd-gan_b200/csrc` or __graft_entry__.build(); '
                                          'this framework has no CPU fallback'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != MMDGAN_OK:
        msg = load().mmdgan_last_error()
        raise MmdganError(rc, msg.decode() if msg else '')
    return rc
