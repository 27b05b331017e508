"""Experiment definitions of the reference (my_test_cifar.py:12-38, my_test_stl.py:10-32, my_test_celebA.py:11-38,
my_test_lsun.py:11-38) as builders of the architecture dictionary SNGan consumes: layer names, channel counts, kernels,
strides, activations, act_k and the spectral-norm flags are the reference's; `tiny` is a small network with the same op
mix for smoke runs.  (oracle/architectures.py is the test-side twin; tests/test_host_logic.py checks they agree.)
"""
import numpy as np


def _dis_block(name, out, act_k, **kw):
    d = {'name': name, 'out': out, 'act': 'lrelu', 'act_k': act_k, 'w_nm': 's'}
    d.update(kw)
    return d


def cifar(act_k=None):
    """my_test_cifar.py:10-38."""
    act_k = float(np.power(64.0, 0.125)) if act_k is None else act_k
    return {
        'input': [(3, 32, 32)],
        'code': [(128, 'linear')],
        'generator': [
            {'name': 'l1', 'out': 512 * 4 * 4, 'op': 'd', 'act': 'linear', 'act_nm': None, 'out_reshape': [512, 4, 4]},
            {'name': 'l2_up', 'out': 256, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l3_up', 'out': 128, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l4_up', 'out': 64, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l5_t32', 'out': 3, 'act': 'tanh'}],
        'discriminator': [
            _dis_block('l1_f32', 64, act_k),
            _dis_block('l2_ds', 128, act_k, kernel=4, strides=2),
            _dis_block('l3', 128, act_k),
            _dis_block('l4_ds', 256, act_k, kernel=4, strides=2),
            _dis_block('l5', 256, act_k),
            _dis_block('l6_ds', 512, act_k, kernel=4, strides=2),
            _dis_block('l7', 512, act_k, op='c', out_reshape=[4 * 4 * 512]),
            {'name': 'l8_s', 'out': 16, 'op': 'd', 'act_k': act_k, 'bias': 'b', 'w_nm': 's'}]}


def stl(act_k=None):
    """my_test_stl.py:8-32."""
    act_k = float(np.power(64.0, 0.125)) if act_k is None else act_k
    return {
        'input': [(3, 48, 48)],
        'code': [(128, 'linear')],
        'generator': [
            {'name': 'l1', 'out': 512 * 6 * 6, 'op': 'd', 'act': 'relu', 'act_nm': 'bn', 'out_reshape': [512, 6, 6]},
            {'name': 'l2_up', 'out': 256, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l3_up', 'out': 128, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l4_up', 'out': 64, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l5_t32', 'out': 3, 'act': 'tanh'}],
        'discriminator': [
            _dis_block('l1_f32', 64, act_k),
            _dis_block('l2_ds', 128, act_k, kernel=4, strides=2),
            _dis_block('l3', 128, act_k),
            _dis_block('l4_ds', 256, act_k, kernel=4, strides=2),
            _dis_block('l5', 256, act_k),
            _dis_block('l6_ds', 512, act_k, kernel=4, strides=2),
            _dis_block('l7', 512, act_k, op='c', out_reshape=[6 * 6 * 512]),
            {'name': 'l8_s', 'out': 16, 'op': 'd', 'act_k': act_k, 'bias': 'b', 'w_nm': 's'}]}


def celeba(act_k=None):
    """my_test_celebA.py:9-38 (my_test_lsun.py:9-38 is the same network)."""
    act_k = float(np.power(64.0, 0.1)) if act_k is None else act_k
    return {
        'input': [(3, 64, 64)],
        'code': [(128, 'linear')],
        'generator': [
            {'name': 'l1', 'out': 1024 * 4 * 4, 'op': 'd', 'act': 'linear', 'act_nm': None,
             'out_reshape': [1024, 4, 4]},
            {'name': 'l2_up', 'out': 512, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l3_up', 'out': 256, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l4_up', 'out': 128, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l5_up', 'out': 64, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l6_t32', 'out': 3, 'act': 'tanh'}],
        'discriminator': [
            _dis_block('l1_f32', 64, act_k),
            _dis_block('l2_ds', 128, act_k, kernel=4, strides=2),
            _dis_block('l3', 128, act_k),
            _dis_block('l4_ds', 256, act_k, kernel=4, strides=2),
            _dis_block('l5', 256, act_k),
            _dis_block('l6_ds', 512, act_k, kernel=4, strides=2),
            _dis_block('l7', 512, act_k),
            _dis_block('l8_ds', 1024, act_k, kernel=4, strides=2),
            _dis_block('l9', 1024, act_k, op='c', out_reshape=[4 * 4 * 1024]),
            {'name': 'l10_s', 'out': 16, 'op': 'd', 'act_k': act_k, 'bias': 'b', 'w_nm': 's'}]}


lsun = celeba

ARCHITECTURES = {'cifar': cifar, 'stl': stl, 'celeba': celeba, 'lsun': lsun}


def tiny(channels=(32, 32), size=8, code=32, score=16, act_k=1.3):
    """A small network with the same op mix (d / tc+bn / c, SN on every D layer) for fast tests."""
    c1, c2 = channels
    return {
        'input': [(3, size, size)],
        'code': [(code, 'linear')],
        'generator': [
            {'name': 'l1', 'out': c2 * (size // 4) ** 2, 'op': 'd', 'act': 'linear', 'act_nm': None,
             'out_reshape': [c2, size // 4, size // 4]},
            {'name': 'l2_up', 'out': c1, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l3_up', 'out': c1, 'op': 'tc', 'act': 'relu', 'act_nm': 'bn', 'kernel': 4, 'strides': 2},
            {'name': 'l4_t', 'out': 3, 'act': 'tanh'}],
        'discriminator': [
            _dis_block('l1_f', c1, act_k),
            _dis_block('l2_ds', c2, act_k, kernel=4, strides=2),
            _dis_block('l3', c2, act_k, op='c', out_reshape=[(size // 2) ** 2 * c2]),
            {'name': 'l4_s', 'out': score, 'op': 'd', 'act_k': act_k, 'bias': 'b', 'w_nm': 's'}]}
