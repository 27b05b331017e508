"""Eager stand-in for the TensorFlow-1.8 API surface used by the hot path of richardwth/MMD-GAN, backed by PyTorch-CPU.

TEST INFRASTRUCTURE ONLY (see ../README.md).  It exists so that the UNMODIFIED reference modules can be imported and
executed in this image to generate golden vectors (tests/golden/make_reference_fixtures.py).  Tensors are torch tensors
(autograd on); graph-mode constructs (variable scopes, collections, while_loop, control_dependencies, assign) are
executed eagerly.  `tf.float32` maps to the dtype in TFSHIM_DTYPE (default float64, so the fixtures are not limited by
fp32 round-off).  Library semantics restated from TF-1.8: SAME padding, conv2d_transpose == conv2d backprop-input,
Maximum/Minimum tie gradients (first argument wins), fused batch norm, AdamOptimizer.
"""
import builtins
import contextlib
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import contrib  # noqa: F401

__version__ = '1.8.0-tfshim'

_DT = {'float64': torch.float64, 'float32': torch.float32}[os.environ.get('TFSHIM_DTYPE', 'float64')]
float32 = _DT
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001
Tensor = torch.Tensor
AUTO_REUSE = 'auto_reuse'


# ----------------------------------------------------------------------------------------------- tensor cosmetics
class _Shape(list):
    def as_list(self):
        return list(self)


def _get_shape(self):
    return _Shape(int(s) for s in self.shape)


torch.Tensor.get_shape = _get_shape


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x), dtype=dtype if dtype is not None else (_DT if np.asarray(x).dtype.kind == 'f' else None))


# ----------------------------------------------------------------------------------------------- graph state
class _State(object):
    def __init__(self):
        self.reset()

    def reset(self, seed=0):
        self.scope = []
        self.variables = {}          # full name -> tensor
        self.trainable = []          # full names, creation order
        self.collections = {}
        self.gen = torch.Generator().manual_seed(seed)


_S = _State()


def reset_default_graph(seed=0):
    _S.reset(seed)


def shim_variables():
    """name -> tensor of every variable created so far (shim-only helper)."""
    return _S.variables


def shim_generator():
    return _S.gen


class GraphKeys(object):
    UPDATE_OPS = 'update_ops'
    TRAINABLE_VARIABLES = 'trainable_variables'
    GLOBAL_VARIABLES = 'variables'


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, **kwargs):
    name = name_or_scope if name_or_scope is not None else default_name
    _S.scope.append(str(name))
    try:
        yield name
    finally:
        _S.scope.pop()


@contextlib.contextmanager
def name_scope(name, default_name=None, values=None):
    yield name


@contextlib.contextmanager
def control_dependencies(control_inputs):
    yield


@contextlib.contextmanager
def device(name):
    yield


def add_to_collection(name, value):
    _S.collections.setdefault(name, []).append(value)


def get_collection(name, scope=None):
    if name == GraphKeys.TRAINABLE_VARIABLES:
        names = [n for n in _S.trainable if scope is None or n.startswith(scope)]
        return [_S.variables[n] for n in names]
    return list(_S.collections.get(name, []))


def trainable_variable_names(scope=None):
    return [n for n in _S.trainable if scope is None or n.startswith(scope)]


# ----------------------------------------------------------------------------------------------- initialisers
def _fans(shape):
    shape = list(shape)
    if len(shape) < 1:
        return 1.0, 1.0
    if len(shape) == 1:
        return float(shape[0]), float(shape[0])
    if len(shape) == 2:
        return float(shape[0]), float(shape[1])
    rf = float(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def _trunc_normal(shape, mean, std):
    out = torch.randn(tuple(shape), generator=_S.gen, dtype=torch.float64)
    for _ in range(64):
        bad = out.abs() > 2.0
        if not bad.any():
            break
        out[bad] = torch.randn(int(bad.sum()), generator=_S.gen, dtype=torch.float64)
    return (out * std + mean).to(_DT)


def zeros_initializer(*a, **k):
    return lambda shape: torch.zeros(tuple(shape), dtype=_DT)


def ones_initializer(*a, **k):
    return lambda shape: torch.ones(tuple(shape), dtype=_DT)


def constant_initializer(value=0.0, **k):
    return lambda shape: torch.full(tuple(shape), float(value), dtype=_DT)


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    return lambda shape: _trunc_normal(shape, mean, stddev)


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    return lambda shape: (torch.randn(tuple(shape), generator=_S.gen, dtype=torch.float64) * stddev + mean).to(_DT)


def variance_scaling_initializer(scale=1.0, mode='fan_in', distribution='normal', seed=None, dtype=None):
    def init(shape):
        fi, fo = _fans(shape)
        n = {'fan_in': fi, 'fan_out': fo, 'fan_avg': (fi + fo) / 2.0}[mode.lower()]
        s = scale / max(1.0, n)
        if distribution == 'normal':
            return _trunc_normal(shape, 0.0, math.sqrt(s) / .87962566103423978)
        lim = math.sqrt(3.0 * s)
        return ((torch.rand(tuple(shape), generator=_S.gen, dtype=torch.float64) * 2 - 1) * lim).to(_DT)
    return init


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
    full = '/'.join(_S.scope + [name])
    if full in _S.variables:
        return _S.variables[full]
    if shape is None:
        shape = []
    if isinstance(shape, int):
        shape = [shape]
    if initializer is None:
        initializer = variance_scaling_initializer(1.0, 'fan_avg', 'uniform')
    if isinstance(initializer, type) or (callable(initializer) and getattr(initializer, '__name__', '') in ('zeros_initializer', 'ones_initializer')):
        initializer = initializer()   # the reference passes `tf.ones_initializer` (the class) in one place
    v = initializer(list(shape)).clone()
    v.requires_grad_(builtins.bool(trainable))
    _S.variables[full] = v
    if trainable:
        _S.trainable.append(full)
    return v


def assign(ref, value, name=None):
    def op():
        with torch.no_grad():
            ref.copy_(value.detach() if isinstance(value, torch.Tensor) else _t(value))
    op.ref, op.value = ref, value
    return op


def assign_add(ref, value, name=None):
    def op():
        with torch.no_grad():
            ref.add_(value)
    return op


# ----------------------------------------------------------------------------------------------- TF tie semantics
class _TFMaximum(torch.autograd.Function):
    """MaximumGrad of TF: xmask = x >= y; dx = where(xmask, g, 0); dy = where(xmask, 0, g) (math_grad.py)."""
    @staticmethod
    def forward(ctx, x, y):
        mask = x >= y
        ctx.save_for_backward(mask)
        ctx.xs, ctx.ys = x.shape, y.shape
        return torch.where(mask, x, y)

    @staticmethod
    def backward(ctx, g):
        mask, = ctx.saved_tensors
        gx = torch.where(mask, g, torch.zeros_like(g))
        gy = g - gx
        return _unbroadcast(gx, ctx.xs), _unbroadcast(gy, ctx.ys)


class _TFMinimum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        mask = x <= y
        ctx.save_for_backward(mask)
        ctx.xs, ctx.ys = x.shape, y.shape
        return torch.where(mask, x, y)

    @staticmethod
    def backward(ctx, g):
        mask, = ctx.saved_tensors
        gx = torch.where(mask, g, torch.zeros_like(g))
        gy = g - gx
        return _unbroadcast(gx, ctx.xs), _unbroadcast(gy, ctx.ys)


def _unbroadcast(g, shape):
    if tuple(g.shape) == tuple(shape):
        return g
    while g.dim() > len(shape):
        g = g.sum(0)
    for i, s in enumerate(shape):
        if s == 1 and g.shape[i] != 1:
            g = g.sum(i, keepdim=True)
    return g


def _pair(x, y):
    x = _t(x)
    y = _t(y)
    dt = x.dtype if x.dtype.is_floating_point else y.dtype
    x, y = x.to(dt), y.to(dt)
    x, y = torch.broadcast_tensors(x, y)
    return x, y


def maximum(x, y, name=None):
    if not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
        return max(x, y)
    return _TFMaximum.apply(*_pair(x, y))


def minimum(x, y, name=None):
    if not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
        return min(x, y)
    return _TFMinimum.apply(*_pair(x, y))


# ----------------------------------------------------------------------------------------------- maths
def constant(value, dtype=None, shape=None, name=None):
    t = torch.as_tensor(np.asarray(value))
    if dtype is None:
        dtype = _DT if t.dtype.is_floating_point else (torch.int32 if t.dtype in (torch.int64, torch.int32) else t.dtype)
    t = t.to(dtype)
    if shape is not None:
        t = t.expand(tuple(shape)).clone() if t.numel() == 1 else t.reshape(tuple(shape))
    return t


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


def identity(x, name=None):
    return x


def stop_gradient(x, name=None):
    return x.detach()


def zeros(shape, dtype=None, name=None):
    return torch.zeros(tuple(shape), dtype=dtype or _DT)


def ones(shape, dtype=None, name=None):
    return torch.ones(tuple(shape), dtype=dtype or _DT)


def eye(n, dtype=None, name=None):
    return torch.eye(n, dtype=dtype or _DT)


def shape(x, name=None):
    return list(x.shape)


def _ax(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
    kd = keepdims if keep_dims is None else keep_dims
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_ax(axis), keepdim=kd)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
    kd = keepdims if keep_dims is None else keep_dims
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_ax(axis), keepdim=kd)


def reduce_max(x, axis=None, keepdims=False, name=None):
    return x.max() if axis is None else x.amax(dim=_ax(axis), keepdim=keepdims)


def reduce_min(x, axis=None, keepdims=False, name=None):
    return x.min() if axis is None else x.amin(dim=_ax(axis), keepdim=keepdims)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def multiply(x, y, name=None):
    return x * y


def add(x, y, name=None):
    return x + y


def add_n(inputs, name=None):
    out = inputs[0]
    for t in inputs[1:]:
        out = out + t
    return out


def subtract(x, y, name=None):
    return x - y


def divide(x, y, name=None):
    return x / y


def square(x, name=None):
    return x * x


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def exp(x, name=None):
    return torch.exp(_t(x))


def log(x, name=None):
    return torch.log(_t(x))


def pow(x, y, name=None):  # noqa: A001
    return torch.pow(_t(x), y)


def abs(x, name=None):  # noqa: A001
    return torch.abs(x)


def sin(x, name=None):
    return torch.sin(_t(x))


def cos(x, name=None):
    return torch.cos(_t(x))


def less(x, y, name=None):
    return x < y


def greater(x, y, name=None):
    return x > y


def logical_not(x, name=None):
    return ~x


def where(cond, x=None, y=None, name=None):
    return torch.where(cond, x, y)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return torch.clamp(t, clip_value_min, clip_value_max)


def expand_dims(x, axis=None, name=None, dim=None):
    return x.unsqueeze(axis if axis is not None else dim)


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    axis = axis if axis is not None else squeeze_dims
    if axis is None:
        return x.squeeze()
    if isinstance(axis, (list, tuple)):
        for a in sorted((a % x.dim() for a in axis), reverse=True):
            x = x.squeeze(a)
        return x
    return x.squeeze(axis)


def reshape(x, shape, name=None):
    return x.reshape(tuple(int(s) for s in shape))


def transpose(x, perm=None, name=None):
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*perm)


def concat(values, axis, name=None):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0, name=None):
    return torch.stack(list(values), dim=axis)


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    if isinstance(num_or_size_splits, int):
        return list(torch.chunk(value, num_or_size_splits, dim=axis))
    return list(torch.split(value, list(num_or_size_splits), dim=axis))


def gather(params, indices, axis=0, name=None):
    return torch.index_select(params, axis, indices.reshape(-1).long()).reshape(
        tuple(indices.shape) + tuple(params.shape[1:])) if axis == 0 else torch.index_select(params, axis, indices.long())


def diag_part(x, name=None):
    return torch.diagonal(x)


def matrix_diag_part(x, name=None):
    return torch.diagonal(x, dim1=-2, dim2=-1)


def diag(x, name=None):
    return torch.diag(x)


def trace(x, name=None):
    return torch.diagonal(x).sum()


def norm(tensor, ord='euclidean', axis=None, keepdims=None, name=None, keep_dims=None):  # noqa: A002
    kd = builtins.bool(keepdims if keepdims is not None else keep_dims)
    assert ord in ('euclidean', 2, 'fro'), ord
    sq = tensor * tensor
    s = sq.sum() if axis is None else sq.sum(dim=_ax(axis), keepdim=kd)
    return torch.sqrt(s)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return (torch.randn(tuple(shape), generator=_S.gen, dtype=torch.float64) * stddev + mean).to(dtype or _DT)


def random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    dtype = dtype or _DT
    if not dtype.is_floating_point:
        return torch.randint(int(minval), int(maxval), tuple(shape), generator=_S.gen).to(dtype)
    maxval = 1.0 if maxval is None else maxval
    return (torch.rand(tuple(shape), generator=_S.gen, dtype=torch.float64) * (maxval - minval) + minval).to(dtype)


def gradients(ys, xs, grad_ys=None, name=None):
    single = not isinstance(xs, (list, tuple))
    xs_l = [xs] if single else list(xs)
    ys_l = ys if isinstance(ys, (list, tuple)) else [ys]
    total = None
    for y in ys_l:
        total = y.sum() if total is None else total + y.sum()
    gs = torch.autograd.grad(total, xs_l, retain_graph=True, allow_unused=True, create_graph=True)
    return list(gs)


def while_loop(cond, body, loop_vars, **kwargs):
    loop_vars = tuple(loop_vars)
    while builtins_bool(cond(*loop_vars)):
        loop_vars = tuple(body(*loop_vars))
    return loop_vars


def builtins_bool(x):
    return x.item() != 0 if isinstance(x, torch.Tensor) else (True if x else False)


# ----------------------------------------------------------------------------------------------- tf.nn
def _same_pad(size, k, s, d=1):
    out = -(-size // s)
    keff = (k - 1) * d + 1
    total = max((out - 1) * s + keff - size, 0)
    return total // 2, total - total // 2


def _hw(v, data_format):
    if isinstance(v, int):
        return v, v
    v = list(v)
    if len(v) == 2:
        return v[0], v[1]
    return (v[2], v[3]) if data_format in ('NCHW', 'channels_first') else (v[1], v[2])


def _conv2d_nchw(x, w_hwio, strides, padding, dilations):
    sh, sw = strides
    dh, dw = dilations
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    w = w_hwio.permute(3, 2, 0, 1)
    if padding == 'SAME':
        pt, pb = _same_pad(x.shape[2], kh, sh, dh)
        pl, pr = _same_pad(x.shape[3], kw, sw, dw)
        x = F.pad(x, (pl, pr, pt, pb))
    elif padding != 'VALID':
        raise ValueError(padding)
    return F.conv2d(x, w, stride=(sh, sw), dilation=(dh, dw))


class _NN(object):
    @staticmethod
    def conv2d(input, filter, strides, padding, use_cudnn_on_gpu=True, data_format='NHWC', dilations=(1, 1, 1, 1), name=None):  # noqa: A002
        nchw = data_format in ('NCHW', 'channels_first')
        x = input if nchw else input.permute(0, 3, 1, 2)
        y = _conv2d_nchw(x, filter, _hw(strides, data_format), padding, _hw(dilations, data_format))
        return y if nchw else y.permute(0, 2, 3, 1)

    @staticmethod
    def conv2d_transpose(value, filter, output_shape, strides, padding='SAME', data_format='NHWC', name=None):  # noqa: A002
        """TF defines conv2d_transpose as conv2d_backprop_input(input_sizes=output_shape, filter, out_backprop=value):
        the vector-Jacobian product of conv2d(., filter) at any point, taken here literally with autograd."""
        nchw = data_format in ('NCHW', 'channels_first')
        dy = value if nchw else value.permute(0, 3, 1, 2)
        oshape = list(output_shape) if nchw else [output_shape[0], output_shape[3], output_shape[1], output_shape[2]]
        x0 = torch.zeros(tuple(int(s) for s in oshape), dtype=dy.dtype, requires_grad=True)
        y0 = _conv2d_nchw(x0, filter, _hw(strides, data_format), padding, (1, 1))
        assert tuple(y0.shape) == tuple(dy.shape), (y0.shape, dy.shape)
        dx, = torch.autograd.grad(y0, x0, dy, create_graph=True)
        return dx if nchw else dx.permute(0, 2, 3, 1)

    @staticmethod
    def bias_add(value, bias, data_format='NHWC', name=None):
        if data_format in ('NCHW', 'channels_first') and value.dim() == 4:
            return value + bias.reshape(1, -1, 1, 1)
        return value + bias

    @staticmethod
    def relu(x, name=None):
        return torch.relu(x)

    @staticmethod
    def leaky_relu(features, alpha=0.2, name=None):
        # TF-1.8: math_ops.maximum(alpha * features, features)
        return maximum(alpha * features, features)

    @staticmethod
    def tanh(x, name=None):
        return torch.tanh(x)

    @staticmethod
    def sigmoid(x, name=None):
        return torch.sigmoid(x)

    @staticmethod
    def softplus(x, name=None):
        return F.softplus(x)

    @staticmethod
    def softsign(x, name=None):
        return F.softsign(x)

    @staticmethod
    def softmax(x, axis=-1, name=None):
        return torch.softmax(x, dim=axis)

    @staticmethod
    def elu(x, name=None):
        return F.elu(x)

    @staticmethod
    def selu(x, name=None):
        return F.selu(x)


nn = _NN()
tanh = torch.tanh
sigmoid = torch.sigmoid


# ----------------------------------------------------------------------------------------------- tf.layers
class _Layers(object):
    @staticmethod
    def batch_normalization(inputs, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, beta_initializer=None,
                            gamma_initializer=None, gamma_constraint=None, training=False, renorm=False, fused=None,
                            name=None, **kwargs):
        """tf.layers.batch_normalization: normalise with the biased batch variance; moving = moving*momentum + batch*(1-momentum)
        via UPDATE_OPS.  TF 1.8 keeps `fused=True` only for rank-4 inputs (normalization.py build(): fused = ... and ndims == 4
        ...): the fused kernel feeds the Bessel-corrected variance into the moving average, the nn.moments fallback taken by
        rank-2 inputs (a dense layer followed by batch norm) the biased one."""
        x = inputs
        axis = axis % x.dim()
        C = x.shape[axis]
        with variable_scope(name or 'batch_normalization'):
            gamma = get_variable('gamma', [C], initializer=gamma_initializer or ones_initializer(), trainable=True) if scale else None
            beta = get_variable('beta', [C], initializer=beta_initializer or zeros_initializer(), trainable=True) if center else None
            mm = get_variable('moving_mean', [C], initializer=zeros_initializer(), trainable=False)
            mv = get_variable('moving_variance', [C], initializer=ones_initializer(), trainable=False)
        red = [d for d in range(x.dim()) if d != axis]
        bshape = [1] * x.dim()
        bshape[axis] = C
        if builtins_bool(training):
            mean = x.mean(dim=red)
            var = ((x - mean.reshape(bshape)) ** 2).mean(dim=red)
            n = x.numel() // C
            unbiased = var * (n / max(n - 1.0, 1.0)) if (x.dim() == 4 and fused is not False) else var
            add_to_collection(GraphKeys.UPDATE_OPS, assign(mm, mm.detach() * momentum + mean.detach() * (1 - momentum)))
            add_to_collection(GraphKeys.UPDATE_OPS, assign(mv, mv.detach() * momentum + unbiased.detach() * (1 - momentum)))
        else:
            mean, var = mm, mv
        y = (x - mean.reshape(bshape)) / torch.sqrt(var.reshape(bshape) + epsilon)
        if gamma is not None:
            y = y * gamma.reshape(bshape)
        if beta is not None:
            y = y + beta.reshape(bshape)
        return y


layers = _Layers()


# ----------------------------------------------------------------------------------------------- tf.summary (no-ops)
class _Summary(object):
    def __getattr__(self, item):
        return lambda *a, **k: None


summary = _Summary()


# ----------------------------------------------------------------------------------------------- tf.train
class _AdamOptimizer(object):
    """tf.train.AdamOptimizer (TF-1.8 training/adam.py): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m, v EMA; var -= lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False, name='Adam'):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.slots = {}

    def compute_gradients(self, loss, var_list=None, **kwargs):
        var_list = list(var_list)
        grads = torch.autograd.grad(loss, var_list, retain_graph=True, allow_unused=True)
        return list(zip(grads, var_list))

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = list(grads_and_vars)

        def op():
            self.t += 1
            lr = self.lr() if callable(self.lr) else self.lr
            lr = float(lr)
            lr_t = lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
            with torch.no_grad():
                for g, v in gv:
                    if g is None:
                        continue
                    m, s = self.slots.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                    m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
                    s.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
                    v.sub_(lr_t * m / (s.sqrt() + self.eps))
                if global_step is not None:
                    global_step.add_(1)
        return op

    def slot(self, var):
        return self.slots.get(id(var))


class _Train(object):
    AdamOptimizer = _AdamOptimizer


train = _Train()


# dtype names referenced in default arguments of modules the hot path never calls (input_func.py)
uint8 = torch.uint8
int8 = torch.int8
int16 = torch.int16
float16 = torch.float16
string = 'string'


def __getattr__(name):
    raise AttributeError("tfshim: tf.{} is not on the MMD-GAN hot path and is not provided".format(name))
