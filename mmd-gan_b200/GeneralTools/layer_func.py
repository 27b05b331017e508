"""Layer DSL of the reference, rebuilt over the B200 kernels: architecture dictionary -> Net -> Layer -> ops.

Mirrors the public surface of GeneralTools/layer_func.py for the hot-path subset (type 'default'; ops 'd' / 'c' / 'tc';
bias; batch norm; activations linear / relu / lrelu / tanh; spectral-norm weight normalisation with act_k):
  update_layer_design   layer_func.py:1189-1275   (same keys, same defaults, same bias/BN rule)
  Layer                 layer_func.py:1278-1685, 2043-2108  (static shape inference, op order kernel->bias->BN->act)
  Net                   layer_func.py:2111-2150   (dense layers get data_format None)
  Routine               layer_func.py:2207-2494   (add_input_layers / seq_links / add_output_layers / __call__)
Error behaviour follows the reference: AttributeError for unsupported ops, NotImplementedError for unimplemented
layer types / norms, assertion messages prefixed with the layer scope.

Unlike the TF graph, nothing here owns device memory: a Net is a static description; parameters, buffers and the
kernel launches of the fused training step live in mmdgan_b200.engine.SNGanEngine, and Routine.__call__ (inference
forward) delegates to an engine bound with Routine.bind().
"""
import numpy as np

from .misc_fun import FLAGS
from .math_func import spatial_shape_after_conv, spatial_shape_after_transpose_conv

HOT_PATH_OPS = {'d', 'c', 'tc'}
KNOWN_OPS = {'d', 'dcd', 'dck', 'sc', 'c', 'tc', 'avg', 'max', 'sum', 'cck', 'tcck', 'i'}


def update_layer_design(layer_design):
    """Fill a layer-design dictionary with the reference's defaults (layer_func.py:1230-1275)."""
    template = {'name': None, 'type': 'default', 'op': 'c', 'out': None, 'bias': 'b',
                'act': 'linear', 'act_nm': None, 'act_k': False,
                'w_nm': None, 'w_p': None,
                'kernel': 3, 'strides': 1, 'dilation': 1, 'padding': 'SAME', 'scale': None,
                'in_reshape': None, 'out_reshape': None, 'aux': None}
    for key in layer_design:
        template[key] = layer_design[key]
    if template['act_nm'] in {'bn', 'BN'} and template['bias'] in {'b', 'bias'}:
        template['bias'] = None  # batch normalization is not used with common bias
    if template['act_nm'] in {'cbn', 'CBN'}:
        template['bias'] = None
    if template['op'] in {'tc'}:
        template['scale'] = None
    if template['scale'] is not None:
        assert isinstance(template['scale'], (list, tuple)), 'Value for key "scale" must be list or tuple.'
    if template['w_nm'] is not None:
        assert not isinstance(template['w_nm'], (list, tuple)), 'Value for key "w_nm" must not be list or tuple.'
    if template['op'] not in KNOWN_OPS:
        raise AttributeError('layer op {} not supported.'.format(template['op']))
    if template['op'] in {'d', 'dcd', 'dck'}:
        keys = ['name', 'op', 'type', 'out', 'bias', 'act', 'act_nm', 'act_k', 'w_nm', 'w_p', 'in_reshape', 'out_reshape', 'aux']
    elif template['op'] in {'i'}:
        keys = ['name', 'op', 'act', 'act_nm', 'type', 'in_reshape', 'out_reshape']
    else:
        keys = ['name', 'op', 'type', 'out', 'bias', 'act', 'act_nm', 'act_k', 'w_nm', 'w_p',
                'kernel', 'strides', 'dilation', 'padding', 'scale', 'in_reshape', 'out_reshape', 'aux']
    return {key: template[key] for key in keys}


class Layer(object):
    """Static description of one layer (layer_func.py:1278).  build_layer() infers shapes and registers the ops."""

    def __init__(self, design, input_shape=None, name_prefix='', data_format=None, num_class=0):
        self.design = design
        self.layer_scope = name_prefix + self.design['name']
        self.input_shape = None if input_shape is None else list(input_shape)
        self.output_shape = None
        if data_format in {'channels_first', 'NCHW'}:
            self.data_format, self.data_format_alias = 'channels_first', 'NCHW'
        elif data_format in {'channels_last', 'NHWC'}:
            raise NotImplementedError('{}: channels_last is not on the hot path (FLAGS.IMAGE_FORMAT, misc_fun.py:50)'.format(
                self.layer_scope))
        else:
            self.data_format_alias = self.data_format = data_format
        self.num_class = num_class
        if self.num_class < 2:
            assert not self.design['type'] in {'project'}, '{}: cannot use {} for one class'.format(
                self.layer_scope, self.design['type'])
            assert not self.design['act_nm'] in {'cbn', 'CBN'}, '{}: cannot use {} for one class'.format(
                self.layer_scope, self.design['act_nm'])
        self.is_layer_built = False
        self.ops = {}
        # filled by build_layer
        self.op_input_shape = None      # after in_reshape
        self.op_output_shape = None     # before out_reshape
        self.kernel_shape = None
        self.use_u = None
        self.sn_x_shape = None
        self.sn_pim = False             # SPECTRAL_NORM_MODE 'sn_paper': the conv kernel as a [k*k*C, C'] matrix (layer_func.py:811-814)

    # ---- spectral-norm routing (math_func.py:470-528): integer compares, must match the reference bit for bit
    def _sn_routing(self):
        op = self.design['op']
        if FLAGS.SPECTRAL_NORM_MODE not in {'default', 'PICO', 'pico', 'sn_paper', 'PIM', 'pim'}:
            raise NotImplementedError('{}: SPECTRAL_NORM_MODE {} is not implemented.'.format(self.layer_scope, FLAGS.SPECTRAL_NORM_MODE))
        self.sn_pim = FLAGS.SPECTRAL_NORM_MODE in {'sn_paper', 'PIM', 'pim'} and op in {'c', 'tc'}
        if op == 'd' or self.sn_pim:
            # PIM: tf.reshape(kernel, (-1, kernel_shape[3])) then the dense routine (layer_func.py:811-814; math_func.py:477-486)
            num_in, num_out = self.kernel_shape if op == 'd' else (int(np.prod(self.kernel_shape[:3])), self.kernel_shape[3])
            self.use_u = True if num_in <= num_out else False
            self.sn_x_shape = [1, num_in] if self.use_u else [1, num_out]
        else:
            self.use_u = True if np.prod(self.op_input_shape[1:]) <= np.prod(self.op_output_shape[1:]) else False
            if op == 'c':
                self.sn_x_shape = [1] + list(self.op_input_shape[1:] if self.use_u else self.op_output_shape[1:])
            else:
                self.sn_x_shape = [1] + list(self.op_output_shape[1:] if self.use_u else self.op_input_shape[1:])

    def build_layer(self):
        if self.is_layer_built:
            return
        d = self.design
        if d['type'] not in {'default'}:
            raise NotImplementedError('{}: {} is not implemented.'.format(self.layer_scope, d['type']))
        if d['op'] not in HOT_PATH_OPS:
            raise AttributeError('{}: type {} not supported'.format(self.layer_scope, d['op']))
        if d.get('scale') is not None:
            raise NotImplementedError('{}: image scaling is not on the hot path'.format(self.layer_scope))
        if d['act'] not in {'linear', 'relu', 'lrelu', 'tanh'}:
            raise NotImplementedError('Function {} is not implemented.'.format(d['act']))
        if d['act_nm'] not in {None, 'bn', 'BN'}:
            raise NotImplementedError('{}: activation normalisation {} is not on the hot path'.format(self.layer_scope, d['act_nm']))
        if d.get('w_nm') not in {None, 's'}:
            raise NotImplementedError('{}: {} method not implemented'.format(self.layer_scope, d['w_nm']))
        batch = self.input_shape[0]
        in_shape = list(self.input_shape) if d['in_reshape'] is None else [batch] + list(d['in_reshape'])
        self.op_input_shape = in_shape
        if d['op'] == 'd':
            assert len(in_shape) == 2, '{}: the input shape {} is not 2-D for a dense op.'.format(self.layer_scope, in_shape)
            self.kernel_shape = [in_shape[1], d['out']]
            out_shape = [batch, d['out']]
        else:
            assert len(in_shape) == 4, '{}: the input shape {} is not 4-D for a conv op.'.format(self.layer_scope, in_shape)
            fan_in, h, w = in_shape[1:]
            if d['dilation'] != 1 or d['padding'] not in {'SAME', 'same'}:
                raise NotImplementedError('{}: dilation / VALID padding are not on the hot path'.format(self.layer_scope))
            if d['op'] == 'c':
                self.kernel_shape = [d['kernel'], d['kernel'], fan_in, d['out']]
                h, w = spatial_shape_after_conv([h, w], d['kernel'], d['strides'], d['dilation'], d['padding'])
            else:
                self.kernel_shape = [d['kernel'], d['kernel'], d['out'], fan_in]
                h, w = spatial_shape_after_transpose_conv([h, w], d['kernel'], d['strides'], d['dilation'], d['padding'])
            out_shape = [batch, d['out'], h, w]
        self.op_output_shape = out_shape
        self.ops['kernel'] = {'op': d['op'], 'kernel_shape': self.kernel_shape, 'w_nm': d.get('w_nm'), 'act_k': d.get('act_k')}
        if d.get('bias') is not None:
            self.ops['bias'] = {'op': 'bias', 'kernel_shape': out_shape[1]}
        if d['act_nm'] in {'bn', 'BN'}:
            self.ops['BN'] = {'op': 'bn', 'kernel_shape': out_shape[1]}
        if d.get('w_nm') == 's':
            if not isinstance(d['act_k'], (float, int)):
                raise NotImplementedError('{}: act_k must be a number with spectral normalisation'.format(self.layer_scope))
            self._sn_routing()
        self.output_shape = out_shape if d['out_reshape'] is None else [batch] + list(d['out_reshape'])
        assert int(np.prod(self.output_shape[1:])) == int(np.prod(out_shape[1:])), \
            '{}: the output shape {} does not match existed shape {}.'.format(self.layer_scope, out_shape[1:], self.output_shape[1:])
        self.is_layer_built = True

    # ---- variable names as the reference's checkpoint holds them
    @property
    def kernel_name(self):
        return self.layer_scope + '/kernel/kernel'

    @property
    def bias_name(self):
        return self.layer_scope + '/bias/bias'

    @property
    def sn_name(self):
        return self.layer_scope + '/kernel/SN/in_rand'

    def bn_name(self, what):
        return self.layer_scope + '/BN/BN/' + what

    def __call__(self, layer_input, is_training=True):
        raise NotImplementedError('{}: layers run inside a bound Routine (Routine.bind(engine))'.format(self.layer_scope))

    apply = __call__


class Net(object):
    """layer_func.py:2111-2150."""

    def __init__(self, net_design, net_name='net', data_format=None, num_class=0):
        self.net_def = net_design
        self.num_layers = len(net_design)
        self.net_name = net_name
        self.layers = []
        for i in range(self.num_layers):
            layer_design = update_layer_design(self.net_def[i])
            if layer_design['op'] in {'d', 'dcd', 'dck'}:
                layer_data_format = None
            elif layer_design['op'] in {'i'} and self.layers[i - 1].design['op'] in {'d', 'dcd', 'dck'}:
                layer_data_format = None
            else:
                layer_data_format = data_format
            self.layers.append(Layer(layer_design, name_prefix=self.net_name + '/', data_format=layer_data_format,
                                     num_class=num_class))


class Routine(object):
    """layer_func.py:2207-2494, sequential subset: one input layer, seq_links, one output layer."""

    def __init__(self, net_object):
        self.net = net_object
        self.operations = []
        self.layer_indices = []
        self.output_layer_indices = []
        self.output_added = False
        self._runner = None

    def add_input_layers(self, input_shape, out_layer_indices):
        for out_index in out_layer_indices:
            if out_index in self.layer_indices:
                raise AttributeError('Layer {} has already been added.'.format(out_index))
            self.layer_indices.append(out_index)
            layer = self.net.layers[out_index]
            layer.input_shape = list(input_shape)
            layer.build_layer()
            self.operations.append([None, None, layer, [out_index]])

    def seq_links(self, in_layer_indices):
        if self.net.layers[in_layer_indices[0]].output_shape is None:
            raise NotImplementedError('Input layer {} has not been defined yet.'.format(in_layer_indices[0]))
        for out_index in in_layer_indices[1:]:
            if out_index in self.layer_indices:
                raise AttributeError('Layer {} has already been linked.'.format(out_index))
            self.layer_indices.append(out_index)
        for index in range(len(in_layer_indices) - 1):
            in_shape = self.net.layers[in_layer_indices[index]].output_shape[:]
            layer = self.net.layers[in_layer_indices[index + 1]]
            layer.input_shape = in_shape
            layer.build_layer()
            self.operations.append([[in_layer_indices[index]], None, layer, [in_layer_indices[index + 1]]])

    def link(self, in_layer_indices, out_layer_indices, input_fun=None):
        if len(in_layer_indices) == 1 and len(out_layer_indices) == 1:
            return self.seq_links([in_layer_indices[0], out_layer_indices[0]])
        raise NotImplementedError('{}: only sequential links are on the hot path'.format(in_layer_indices))

    def add_output_layers(self, in_layer_indices):
        for out_index in in_layer_indices:
            if out_index in self.output_layer_indices:
                raise AttributeError('Layer {} has already been added as output layer.'.format(out_index))
            self.output_layer_indices.append(out_index)
            if self.net.layers[out_index].output_shape is None:
                raise NotImplementedError('Output layer {} has not been linked yet.'.format(out_index))
        self.operations.append([in_layer_indices, None, None, None])
        self.output_added = True

    def ordered_layers(self):
        if not self.output_added:
            raise NotImplementedError('Output layer has not been defined.')
        return [self.net.layers[i] for i in self.layer_indices]

    def bind(self, runner):
        """runner(routine_inputs: dict, is_training) -> dict; set by the engine that owns the parameters."""
        self._runner = runner

    def __call__(self, routine_inputs, is_training=True):
        if not self.output_added:
            raise NotImplementedError('Output layer has not been defined.')
        if self._runner is None:
            raise NotImplementedError('{}: routine is not bound to an engine'.format(self.net.net_name))
        if not isinstance(routine_inputs, dict):
            routine_inputs = {'x': routine_inputs}
        return self._runner(routine_inputs, is_training)

    apply = __call__
