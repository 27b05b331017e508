"""Step loop, optimiser configuration and checkpointing, mirroring the hot-path part of GeneralTools/graph_func.py:
  opt_config / multi_opt_config (graph_func.py:478-575): Adam(beta1 .5, beta2 .999, eps 1e-8), constant learning rate
  prepare_folder (graph_func.py:161-180), Agent (1144-1219), MySession.full_run (820-946): the per-step loop with the
  NaN assert (856), loss print every query_step (860-866) and one checkpoint at the last step (869-871).

The TF session / graph / summary machinery has no equivalent here: a "session" is an SNGanEngine whose step is a
replayed CUDA graph.  Checkpoints are keyed by the reference's variable names, as .npz files or -- with
FLAGS.CKPT_FORMAT = 'tf' -- in TensorFlow's own checkpoint container (tf_bundle.py), which is also what load_ckpt reads
when the folder holds a model trained with the reference.
"""
import os
import time

import numpy as np

from .misc_fun import FLAGS


def opt_config(initial_lr, lr_decay_steps=None, end_lr=1e-7, optimizer='adam', name_suffix='', global_step=None,
               target_step=1e5):
    """graph_func.py:478-527.  Returns (learning_rate, optimiser hyper-parameters)."""
    if optimizer in ['Adam', 'adam']:
        return initial_lr, dict(kind='adam', lr=initial_lr, beta1=0.5, beta2=0.999, epsilon=1e-8, name='Adam' + name_suffix)
    if optimizer in ['SGD', 'sgd', 'Momentum', 'momentum', 'RMSProp', 'rmsprop']:
        raise NotImplementedError('Optimizer {} is not on the hot path (only adam is built).'.format(optimizer))
    raise AttributeError('Optimizer {} not supported.'.format(optimizer))


def multi_opt_config(lr_list, lr_decay_steps=None, end_lr=1e-7, optimizer='adam', global_step=None, target_step=1e5):
    """graph_func.py:540-575."""
    if isinstance(optimizer, str):
        optimizer = [optimizer]
    if len(lr_list) == 1:
        return opt_config(lr_list[0], lr_decay_steps, end_lr, optimizer[0], '', global_step, target_step)
    if len(optimizer) == 1:
        optimizer = optimizer * len(lr_list)
    combo = [opt_config(lr_list[i], lr_decay_steps, end_lr, optimizer[i], '_' + str(i), global_step, target_step)
             for i in range(len(lr_list))]
    return [c[0] for c in combo], [c[1] for c in combo]


def prepare_folder(filename, sub_folder='', set_folder=True):
    """graph_func.py:161-180: <DEFAULT_OUT>/<file>_ckpt/<sub_folder>/ and .../<file>_log/<sub_folder>/."""
    ckpt_folder = os.path.join(FLAGS.DEFAULT_OUT, filename + '_ckpt', sub_folder)
    summary_folder = os.path.join(FLAGS.DEFAULT_OUT, filename + '_log', sub_folder)
    save_path = os.path.join(ckpt_folder, filename + '.ckpt')
    if set_folder:
        os.makedirs(ckpt_folder, exist_ok=True)
        os.makedirs(summary_folder, exist_ok=True)
    return ckpt_folder, summary_folder, save_path


_BETA_NAMES = (('beta1_power', 'beta2_power'), ('beta1_power_1', 'beta2_power_1'))   # optimiser 0 = dis, 1 = gen (my_sngan.py:424-426)


def collect_variables(engine, global_step, tf_names=False):
    """All global variables under the reference's names: weights, biases, BN gamma/beta/moving stats, SN in_rand, Adam
    slots (<var>/Adam_i, <var>/Adam_i_1) and global_step (graph_func.py:708-717) as host arrays.  The number of updates each
    optimiser has applied is kept as `<net>/adam_step`; with tf_names=True it is stored the way TF-1.8 `AdamOptimizer` keeps
    it -- the non-slot variables beta1_power / beta2_power = beta ** (updates + 1) (suffix `_1` for the second optimiser)."""
    out = {'global_step': np.asarray(global_step, dtype=np.int32)}
    for i, net in enumerate((engine.D, engine.G)):        # Adam_0 = discriminator, Adam_1 = generator (my_sngan.py:414)
        for name in net.var_offsets:
            out[name] = net.get_variable(name).cpu().numpy()
            off, shape = net.var_offsets[name]
            n = int(np.prod(shape))
            out[name + '/Adam_{}'.format(i)] = net.m[off:off + n].reshape(shape).cpu().numpy()
            out[name + '/Adam_{}_1'.format(i)] = net.v[off:off + n].reshape(shape).cpu().numpy()
        for name in net.state_names():
            out[name] = net.get_state(name).cpu().numpy()
        steps = int(net.step.cpu().numpy().reshape(-1)[0])
        if tf_names:
            out[_BETA_NAMES[i][0]] = np.asarray(np.float32(0.5) ** np.float32(steps + 1), dtype=np.float32)
            out[_BETA_NAMES[i][1]] = np.asarray(np.float32(0.999) ** np.float32(steps + 1), dtype=np.float32)
        else:
            out[net.name + '/adam_step'] = net.step.cpu().numpy()
    return out


def _adam_steps(z, i, net_name, global_step):
    """Updates applied by optimiser i: `<net>/adam_step` (.npz), else recovered from TF's beta2_power = .999 ** (t + 1) while
    that is a normal float32 (t < ~87 000; afterwards the bias correction is 1 to float32 precision and global_step serves)."""
    key = net_name + '/adam_step'
    if key in z:
        return int(np.asarray(z[key]).reshape(-1)[0])
    b2 = z.get(_BETA_NAMES[i][1]) if hasattr(z, 'get') else None
    if b2 is not None and float(b2) > 1e-37:
        return max(int(round(np.log(float(b2)) / np.log(0.999))) - 1, 0)
    return int(global_step)


def apply_variables(engine, z):
    """Inverse of collect_variables.  Adam slots are optional (a checkpoint saved from an inference graph has none: the
    moments then stay as they are)."""
    import torch
    gs = int(np.asarray(z['global_step']).reshape(-1)[0])
    for i, net in enumerate((engine.D, engine.G)):
        for name in net.var_offsets:
            net.set_variable(name, torch.from_numpy(np.asarray(z[name])))
            off, shape = net.var_offsets[name]
            n = int(np.prod(shape))
            km, kv = name + '/Adam_{}'.format(i), name + '/Adam_{}_1'.format(i)
            if km in z and kv in z:
                net.m[off:off + n].copy_(torch.from_numpy(np.asarray(z[km])).reshape(-1))
                net.v[off:off + n].copy_(torch.from_numpy(np.asarray(z[kv])).reshape(-1))
        for name in net.state_names():
            net.set_state(name, torch.from_numpy(np.asarray(z[name])))
        net.step.fill_(_adam_steps(z, i, net.name, gs))
        net.refresh()
    engine.global_step = gs
    return gs


def _is_tf_prefix(path):
    return os.path.isfile(path + '.index')


def save_checkpoint(engine, save_path, global_step, ckpt_format=None, max_to_keep=2):
    """`<save_path>-<global_step>`: a .npz file (default) or, with ckpt_format='tf' (FLAGS.CKPT_FORMAT), the files
    `tf.train.Saver(max_to_keep=2).save(sess, save_path, global_step)` writes (graph_func.py:708-717, 869-871): tensor
    bundle + the folder's `checkpoint` state file, older bundles beyond max_to_keep removed."""
    fmt = ckpt_format if ckpt_format is not None else getattr(FLAGS, 'CKPT_FORMAT', 'npz')
    if fmt == 'npz':
        path = '{}-{}.npz'.format(save_path, global_step)
        np.savez(path, **collect_variables(engine, global_step))
        return path
    if fmt != 'tf':
        raise AttributeError('Checkpoint format {} not supported.'.format(fmt))
    from . import tf_bundle
    prefix = '{}-{}'.format(save_path, global_step)
    tf_bundle.write_bundle(prefix, collect_variables(engine, global_step, tf_names=True))
    folder = os.path.dirname(prefix)
    state = tf_bundle.read_checkpoint_state(folder)
    kept = [p for p in (state[1] if state else []) if p != prefix and _is_tf_prefix(p)] + [prefix]
    for old in kept[:-max_to_keep]:
        for f in (old + '.index', old + '.data-00000-of-00001', old + '.meta'):
            if os.path.isfile(f):
                os.remove(f)
    tf_bundle.write_checkpoint_state(folder, prefix, kept[-max_to_keep:])
    return prefix


def get_ckpt(ckpt_folder, ckpt_file=None):
    """graph_func.py:399-416: latest checkpoint in the folder (or the named one).  Knows both containers: `<file>.ckpt-N.npz`
    and TF bundles `<file>.ckpt-N{.index, .data-00000-of-00001}` (returned as the prefix, as TF does); the highest global step
    wins, and a folder holding only TF files is resolved through its `checkpoint` state file first, like
    tf.train.get_checkpoint_state."""
    if ckpt_file is not None:
        path = os.path.join(ckpt_folder, ckpt_file)
        return path if (os.path.exists(path) or _is_tf_prefix(path)) else None
    if not os.path.isdir(ckpt_folder):
        return None
    cands = {}
    for f in os.listdir(ckpt_folder):
        if '.ckpt-' not in f:
            continue
        if f.endswith('.npz'):
            stem = f[:-4]
        elif f.endswith('.index'):
            stem = f[:-6]
        else:
            continue
        tail = stem.rsplit('-', 1)[1]
        if tail.isdigit():
            cands.setdefault(int(tail), []).append(f)
    if not cands:
        from . import tf_bundle
        state = tf_bundle.read_checkpoint_state(ckpt_folder)
        return state[0] if state is not None and _is_tf_prefix(state[0]) else None
    best = sorted(cands[max(cands)])               # same step in both containers: '.index' sorts first, the bundle wins
    f = best[0]
    return os.path.join(ckpt_folder, f[:-6] if f.endswith('.index') else f)


def load_checkpoint(engine, path):
    """path: a .npz checkpoint or a TF bundle prefix (`.../cifar.ckpt-6284`)."""
    if path.endswith('.npz') and os.path.isfile(path):
        return apply_variables(engine, np.load(path))
    if _is_tf_prefix(path):
        from . import tf_bundle
        return apply_variables(engine, tf_bundle.read_bundle(path))
    raise FileNotFoundError('No ckpt Model found at {}'.format(path))


def print_tensor_in_ckpt(ckpt_folder, all_tensor_values=False, all_tensor_names=False):
    """graph_func.py:419-443: list the variables of the latest checkpoint under FLAGS.DEFAULT_OUT/<ckpt_folder> (either
    container), one `name (dtype) shape` line each as TF's inspect_checkpoint prints them; values on request.  Returns the
    {name: (dtype, shape)} listing."""
    if not isinstance(ckpt_folder, str):  # if list, use the name of the first file
        ckpt_folder = ckpt_folder[0]
    output_folder = os.path.join(FLAGS.DEFAULT_OUT, ckpt_folder)
    print(output_folder)
    path = get_ckpt(output_folder)
    print(path)
    if path is None:
        raise FileNotFoundError('No ckpt Model found at {}'.format(output_folder))
    if path.endswith('.npz') and os.path.isfile(path):
        z = dict(np.load(path))
        listing = {k: (v.dtype, tuple(v.shape)) for k, v in z.items()}
    else:
        from . import tf_bundle
        listing = tf_bundle.list_bundle(path)
        z = tf_bundle.read_bundle(path) if all_tensor_values else {}
    for name in sorted(listing):
        dtype, shape = listing[name]
        print('{} ({}) {}'.format(name, dtype, list(shape)))
        if all_tensor_values:
            print(z[name])
    return listing


def rollback(engine, ckpt_folder, ckpt_file=None):
    """graph_func.py:606-636 without the session: restore the engine's variables from the latest (or the named) checkpoint of
    the folder and return its global step."""
    path = get_ckpt(ckpt_folder, ckpt_file)
    if path is None:
        raise FileNotFoundError('No ckpt Model found at {}'.format(ckpt_folder))        # graph_func.py:633
    step = load_checkpoint(engine, path)
    FLAGS.print('Model reloaded from {}.'.format(path))
    return step


def sprite_array(images, mesh_num=None, if_invert=False):
    """The uint8 mosaic the reference writes (graph_func.py:222-265): channels-last images [n, h, w(, c)], EACH image shifted
    by its own minimum and divided by its own range (no guard: a constant image divides by zero exactly as there), optionally
    inverted, laid out row-major on a mesh_num = (rows, columns) grid -- or, with mesh_num None, on the smallest square that
    holds them, padded with black."""
    x = np.asarray(images)
    if x.ndim == 3:
        x = x[..., None]
    if x.shape[3] == 1:
        x = np.repeat(x, 3, axis=3)
    x = x.astype(np.float32)
    flat = x.reshape(x.shape[0], -1)
    flat = flat - flat.min(axis=1, keepdims=True)
    with np.errstate(divide='ignore', invalid='ignore'):
        flat = flat / flat.max(axis=1, keepdims=True)
    x = flat.reshape(x.shape)
    if if_invert:
        x = 1 - x
    if mesh_num is None:
        side = int(np.ceil(np.sqrt(x.shape[0])))
        mesh_num = (side, side)
        x = np.concatenate([x, np.zeros((side * side - x.shape[0],) + x.shape[1:], np.float32)], axis=0)
    rows, cols = (int(m) for m in mesh_num)
    n, h, w, c = x.shape
    assert n == rows * cols, '{} images do not fill a {} x {} mesh'.format(n, rows, cols)
    mosaic = x.reshape(rows, cols, h, w, c).transpose(0, 2, 1, 3, 4).reshape(rows * h, cols * w, c)
    with np.errstate(invalid='ignore'):
        return (mosaic * 255).astype(np.uint8)


def write_sprite(sprite_path, images, mesh_num=None, if_invert=False):
    """graph_func.py:222-266; the PNG is written with PIL (scipy.misc.imsave, which the reference calls, no longer exists)."""
    from PIL import Image
    Image.fromarray(sprite_array(images, mesh_num, if_invert)).save(sprite_path)


def write_sprite_wrapper(images, mesh_num, filename, file_folder=None, file_index='', if_invert=False,
                         image_format='channels_last'):
    """graph_func.py:269-298: <file_folder>/<filename><file_index>.png; an existing file is kept (with a warning)."""
    import warnings
    if not isinstance(filename, str):
        filename = filename[0]
    if file_folder is None:
        file_folder = FLAGS.DEFAULT_OUT
    images = np.asarray(images)
    if image_format in {'channels_first', 'NCHW'}:
        images = np.transpose(images, axes=(0, 2, 3, 1))
    sprite_path = os.path.join(file_folder, filename + file_index + '.png')
    if os.path.isfile(sprite_path):
        warnings.warn('This file already exists: ' + sprite_path)
    else:
        write_sprite(sprite_path, images, mesh_num=mesh_num, if_invert=if_invert)
    return sprite_path


class Agent(object):
    """graph_func.py:1144-1219.  train() is MySession.full_run (graph_func.py:820-908): every step one fused engine step;
    imbalanced_update = (k_dis, k_gen) selects per step which of the two optimisers is applied."""

    def __init__(self, filename, sub_folder, load_ckpt=False, do_trace=False, do_save=True, debug_mode=False, debug_step=800,
                 query_step=500, log_device=False, imbalanced_update=None, print_loss=True):
        self.ckpt_folder, self.summary_folder, self.save_path = prepare_folder(filename, sub_folder=sub_folder)
        self.load_ckpt = load_ckpt
        self.do_trace = do_trace
        self.do_save = do_save
        self.debug = debug_mode
        self.debug_step = debug_step
        self.log_device = log_device
        self.query_step = query_step
        self.imbalanced_update = imbalanced_update
        self.print_loss = print_loss
        if isinstance(imbalanced_update, str):
            raise NotImplementedError("imbalanced_update='dynamic' serves sngan_mmd_rand_g only (graph_func.py:910-950): not on the hot path.")
        if imbalanced_update is not None:
            if not isinstance(imbalanced_update, (list, tuple)):
                raise AttributeError('Imbalanced_update not identified.')                     # my_sngan.py:445
            assert len(imbalanced_update) == 2, 'Imbalanced_update length does not match that of op_list. Expected 2 got {}.'.format(
                len(imbalanced_update))                                                       # graph_func.py:878-880
            if 1 not in tuple(imbalanced_update):
                raise AttributeError('One of the imbalanced_update must be 1.')               # my_sngan.py:439

    def train(self, engine, batch_fn, max_step, step_per_epoch, loss_names='<loss_gen>, <loss_dis>', force_print=False):
        """engine: SNGanEngine; batch_fn(step) -> (data NCHW float32 in [-1,1], codes [B, code_size]) host tensors."""
        if self.load_ckpt:
            path = get_ckpt(self.ckpt_folder)
            if path is not None:
                step = load_checkpoint(engine, path)
                FLAGS.print('Model reloaded from {} (global step {}).'.format(path, step), force_print)
            else:
                FLAGS.print('No ckpt found; variables initialised.', force_print)
        start_time = time.time()
        loss_value = None
        nxt = batch_fn(0) if max_step > 0 else None

        def finish(pending, step):
            # losses of an enqueued step; check if model produces nan outcome (graph_func.py:856)
            value = engine.result(pending, check_nan=False)
            assert not any(np.isnan(value)), 'Model diverged with loss = {} at step {}'.format(value, step)
            gs = pending[2]
            if gs % self.query_step == (self.query_step - 1) and self.print_loss:
                epoch = step // max(step_per_epoch, 1)
                FLAGS.print('Epoch {}, global steps {}, loss_list {}'.format(
                    epoch, gs, ['{}'.format(['<{:.2f}>'.format(l) for l in value])]))
            return value

        pending = None
        for step in range(max_step):
            data_x, code_x = nxt
            # the next batch is produced now and its host -> device copy overlaps this step (engine.step_async(prefetch=...))
            nxt = batch_fn(step + 1) if step + 1 < max_step else None
            # imbalanced update (graph_func.py:885-886): optimiser i runs when the global step is a multiple of imbalanced_update[i]
            update = (True, True) if self.imbalanced_update is None else \
                tuple(engine.global_step % int(k) == 0 for k in self.imbalanced_update)
            # step i + 1 is enqueued before the losses of step i are read: the device never waits for the host's per-step work
            enqueued = engine.step_async(data_x, code_x, update=update, prefetch=nxt)
            if pending is not None:
                loss_value = finish(pending, step - 1)
            pending = enqueued
        if pending is not None:
            loss_value = finish(pending, max_step - 1)
            if self.do_save:
                save_checkpoint(engine, self.save_path, engine.global_step)
        duration = time.time() - start_time
        FLAGS.print('Training for {} steps took {:.3f} sec.'.format(max_step, duration))     # graph_func.py:945-946
        return loss_value
