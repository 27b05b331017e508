"""CPU: the TensorFlow checkpoint container (GeneralTools/tf_bundle.py) and its use by the checkpoint helpers of
graph_func.py.  The reference ships no checkpoint file, so these tests pin the codec to the FORMAT: crc32c check values from
RFC 3720, the LevelDB table constants (magic number, 48-byte footer, 5-byte block trailers, restart arrays), the
tensor_bundle.proto field numbers written out by hand -- and the round trip."""
import os
import struct

import numpy as np
import pytest
import torch

from mmdgan_b200.GeneralTools import tf_bundle as tb
from mmdgan_b200.GeneralTools.input_func import crc32c


def test_crc32c_check_values_and_fast_path():
    # RFC 3720 B.4 / the usual "123456789" check value
    assert crc32c(b'123456789') == 0xE3069283
    assert crc32c(bytes(32)) == 0x8A9136AA and crc32c(b'\xff' * 32) == 0x62A8AB43
    assert crc32c(bytes(range(32))) == 0x46DD794E
    rng = np.random.default_rng(0)
    for n in (0, 1, 4 * 8192 - 1, 4 * 8192, 4 * 8192 + 1, 50001, 131072):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb.crc32c_fast(data) == crc32c(data), n
    assert tb.crc32c_fast(bytes(40000)) == crc32c(bytes(40000))            # all-zero data exercises the preset term alone
    arr = rng.standard_normal((300, 200)).astype(np.float32)               # ndarray input == its bytes
    assert tb.crc32c_fast(arr) == crc32c(arr.tobytes())
    for c in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tb.unmask_crc(tb.mask_crc(c)) == c
    assert tb.mask_crc(0) == 0xA282EAD8


def test_table_file_structure(tmp_path):
    path = str(tmp_path / 't.index')
    items = [(b'', b'hdr')] + [('k{:04d}'.format(i).encode(), bytes([i % 251]) * 40) for i in range(400)]
    tb.write_table(path, items, block_size=1024)
    data = open(path, 'rb').read()
    # footer: 40 bytes of (padded) block handles, then the magic number as two little-endian fixed32
    assert data[-8:] == struct.pack('<Q', 0xdb4775248b80fb57)
    footer = data[-48:]
    pos = 0
    moff, pos = tb._read_varint(footer, pos)
    msize, pos = tb._read_varint(footer, pos)
    ioff, pos = tb._read_varint(footer, pos)
    isize, pos = tb._read_varint(footer, pos)
    assert footer[pos:40] == bytes(40 - pos)
    assert msize == 8 and data[moff:moff + 8] == struct.pack('<II', 0, 1)         # empty metaindex block: one restart at 0
    assert ioff == moff + msize + 5 and ioff + isize + 5 + 48 == len(data)       # blocks are followed by 5-byte trailers
    # every block trailer: type 0, masked crc32c over contents + type byte
    index = data[ioff:ioff + isize]
    assert data[ioff + isize] == 0
    assert struct.unpack('<I', data[ioff + isize + 1:ioff + isize + 5])[0] == tb.mask_crc(crc32c(index + b'\x00'))
    handles = list(tb._block_entries(index))
    assert len(handles) > 10                                                      # 1 KB blocks
    (nrest,) = struct.unpack('<I', index[-4:])
    assert nrest == len(handles)                                                  # index block: restart interval 1
    expect_off, seen = 0, []
    for sep, handle in handles:
        boff, p = tb._read_varint(handle, 0)
        bsize, p = tb._read_varint(handle, p)
        assert boff == expect_off and p == len(handle)
        expect_off = boff + bsize + 5
        entries = list(tb._block_entries(data[boff:boff + bsize]))
        assert entries[-1][0] <= sep                                              # separator >= last key of its block
        if seen:
            assert seen[-1][0] < entries[0][0] and (sep > seen[-1][0])
        (nr,) = struct.unpack('<I', data[boff + bsize - 4:boff + bsize])
        assert nr == -(-len(entries) // 16)                                       # data blocks: a restart every 16 entries
        seen.extend(entries)
    assert seen == items
    # first data block by hand: entry 0 is the empty key (shared 0, non_shared 0, value_len 3)
    assert data[:6] == b'\x00\x00\x03hdr'
    # entry 1 'k0000' shares nothing with ''; entry 2 'k0001' shares 4 bytes with it
    assert data[6:9] == b'\x00\x05\x28' and data[9:14] == b'k0000'
    assert data[54:57] == b'\x04\x01\x28' and data[57:58] == b'1'
    assert tb.read_table(path) == items
    # corruption is detected
    bad = bytearray(data)
    bad[20] ^= 1
    open(path, 'wb').write(bytes(bad))
    with pytest.raises(ValueError, match='crc32c mismatch'):
        tb.read_table(path)
    open(path, 'wb').write(data[:-1] + b'\x00')
    with pytest.raises(ValueError, match='bad magic'):
        tb.read_table(path)


def test_bundle_protos_by_hand():
    # BundleHeaderProto{num_shards: 1, version{producer: 1}}
    assert tb._encode_header(1) == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    # BundleEntryProto{dtype: DT_FLOAT, shape{dim{size: 3} dim{size: 16}}, offset: 300, size: 192, crc32c: fixed32}
    e = tb._encode_entry(1, (3, 16), 0, 300, 192, 0x01020304)
    assert e == bytes([0x08, 0x01, 0x12, 0x08, 0x12, 0x02, 0x08, 0x03, 0x12, 0x02, 0x08, 0x10,
                       0x20, 0xAC, 0x02, 0x28, 0xC0, 0x01, 0x35, 0x04, 0x03, 0x02, 0x01])
    d = tb._decode_entry(e)
    assert d['dtype'] == 1 and d['shape'] == [3, 16] and d['shard_id'] == 0 and d['offset'] == 300 and d['size'] == 192
    assert d['crc32c'] == 0x01020304 and not d['sliced']
    # scalar int32 at offset 0: the (empty) shape message is present, zero-valued scalars are omitted (proto3)
    assert tb._encode_entry(3, (), 0, 0, 4, 7) == bytes([0x08, 0x03, 0x12, 0x00, 0x28, 0x04, 0x35, 0x07, 0x00, 0x00, 0x00])
    assert tb._decode_entry(tb._encode_entry(1, (0, 4), 0, 0, 0, 0))['shape'] == [0, 4]


def test_bundle_round_trip_and_errors(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {
        'global_step': np.asarray(6284, np.int32),
        'beta1_power': np.asarray(0.5 ** 9, np.float32),
        'dis/l1_f/kernel/kernel': rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
        'dis/l1_f/kernel/kernel/Adam_0': rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
        'dis/l1_f/kernel/SN/in_rand': rng.standard_normal((1, 3, 32, 32)).astype(np.float32),
        'gen/l2_up/BN/BN/moving_variance': rng.random(256).astype(np.float32),
        'big': rng.standard_normal((512, 300)).astype(np.float32),         # > 4 chunks: the vectorised crc path
        'f64': rng.standard_normal(5), 'i64': np.arange(4, dtype=np.int64), 'empty': np.zeros((0, 4), np.float32),
    }
    prefix = str(tmp_path / 'cifar.ckpt-6284')
    assert tb.write_bundle(prefix, tensors) == prefix
    assert sorted(os.listdir(tmp_path)) == ['cifar.ckpt-6284.data-00000-of-00001', 'cifar.ckpt-6284.index']
    # the data shard is the tensors' bytes back to back in name order
    raw = open(prefix + '.data-00000-of-00001', 'rb').read()
    assert raw == b''.join(np.ascontiguousarray(tensors[k]).tobytes() for k in sorted(tensors, key=str.encode))
    listing = tb.list_bundle(prefix)
    assert listing['dis/l1_f/kernel/kernel'] == (np.dtype(np.float32), (3, 3, 3, 64)) and listing['global_step'] == (np.dtype(np.int32), ())
    out = tb.read_bundle(prefix)
    assert set(out) == set(tensors)
    for k, v in tensors.items():
        assert out[k].dtype == np.asarray(v).dtype and out[k].shape == np.asarray(v).shape and np.array_equal(out[k], v), k
    assert list(tb.read_bundle(prefix, names=['big'])) == ['big']
    with pytest.raises(KeyError, match='not found in checkpoint'):
        tb.read_bundle(prefix, names=['dis/l9/kernel/kernel'])
    with pytest.raises(TypeError):
        tb.write_bundle(str(tmp_path / 'x'), {'s': np.asarray(['a'])})
    # a flipped data byte is caught by the per-tensor crc32c
    bad = bytearray(raw)
    bad[len(bad) // 2] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(bad))
    with pytest.raises(ValueError, match='crc32c mismatch'):
        tb.read_bundle(prefix)
    assert set(tb.read_bundle(prefix, check_crc=False)) == set(tensors)


def test_checkpoint_state_file(tmp_path):
    folder = str(tmp_path)
    assert tb.read_checkpoint_state(folder) is None
    a, b = os.path.join(folder, 'cifar.ckpt-100'), os.path.join(folder, 'cifar.ckpt-200')
    tb.write_checkpoint_state(folder, b, [a, b])
    assert open(os.path.join(folder, 'checkpoint')).read() == (
        'model_checkpoint_path: "cifar.ckpt-200"\nall_model_checkpoint_paths: "cifar.ckpt-100"\n'
        'all_model_checkpoint_paths: "cifar.ckpt-200"\n')
    assert tb.read_checkpoint_state(folder) == (b, [a, b])


def test_checkpoint_state_with_relative_default_out(tmp_path, monkeypatch):
    """FLAGS.DEFAULT_OUT is relative by default ('MMD-GAN/Results/'): the state file must still hold folder-relative paths, so
    that Saver(max_to_keep=2) pruning finds the older bundles and get_ckpt's state-file fallback resolves."""
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools import graph_func as gf
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(FLAGS, 'DEFAULT_OUT', 'rel/sub/')
    monkeypatch.setattr(FLAGS, 'CKPT_FORMAT', 'tf')
    monkeypatch.setattr(FLAGS, 'SILENT_MODE', True)
    folder, _, save_path = gf.prepare_folder('cifar', 'sngan_rep')
    assert not os.path.isabs(save_path)
    src = _FakeEngine(5)
    for s in (1, 2, 3):
        gf.save_checkpoint(src, save_path, s)
    names = sorted(os.listdir(folder))
    assert names == ['checkpoint', 'cifar.ckpt-2.data-00000-of-00001', 'cifar.ckpt-2.index',
                     'cifar.ckpt-3.data-00000-of-00001', 'cifar.ckpt-3.index']
    assert open(os.path.join(folder, 'checkpoint')).read().splitlines()[0] == 'model_checkpoint_path: "cifar.ckpt-3"'
    latest, every = tb.read_checkpoint_state(folder)
    assert os.path.isfile(latest + '.index') and all(os.path.isfile(p + '.index') for p in every)
    assert os.path.samefile(latest + '.index', save_path + '-3.index')


class _FakeNet(object):
    """The slice of engine.NetState the checkpoint helpers touch, on the CPU."""

    def __init__(self, name, shapes, states, seed):
        g = torch.Generator().manual_seed(seed)
        self.name = name
        self.var_offsets, off = {}, 0
        for k, shape in shapes.items():
            self.var_offsets[k] = (off, shape)
            off += int(np.prod(shape))
        self.w = torch.randn(off, generator=g)
        self.m = torch.randn(off, generator=g)
        self.v = torch.rand(off, generator=g)
        self.state = {k: torch.randn(shape, generator=g) for k, shape in states.items()}
        self.step = torch.zeros(1, dtype=torch.int32)
        self.refreshed = 0

    def _view(self, name):
        off, shape = self.var_offsets[name]
        return self.w[off:off + int(np.prod(shape))].view(shape)

    def get_variable(self, name):
        return self._view(name).clone()

    def set_variable(self, name, value):
        self._view(name).copy_(value)

    def state_names(self):
        return list(self.state)

    def get_state(self, name):
        return self.state[name].clone()

    def set_state(self, name, value):
        self.state[name].copy_(value)

    def refresh(self):
        self.refreshed += 1


class _FakeEngine(object):
    def __init__(self, seed):
        self.D = _FakeNet('dis', {'dis/l1_f/kernel/kernel': (3, 3, 3, 16), 'dis/l1_f/bias/bias': (16,)},
                          {'dis/l1_f/kernel/SN/in_rand': (1, 3, 8, 8)}, seed)
        self.G = _FakeNet('gen', {'gen/l1/kernel/kernel': (16, 64), 'gen/l2_up/BN/BN/gamma': (16,)},
                          {'gen/l2_up/BN/BN/moving_mean': (16,)}, seed + 1)
        self.global_step = 0


def _same(a, b):
    for na, nb in ((a.D, b.D), (a.G, b.G)):
        assert torch.equal(na.w, nb.w) and torch.equal(na.m, nb.m) and torch.equal(na.v, nb.v)
        assert int(na.step) == int(nb.step)
        for k in na.state:
            assert torch.equal(na.state[k], nb.state[k])


@pytest.mark.parametrize('fmt', ['npz', 'tf'])
def test_save_load_checkpoint_both_containers(tmp_path, monkeypatch, fmt):
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools import graph_func as gf
    monkeypatch.setattr(FLAGS, 'DEFAULT_OUT', str(tmp_path) + '/')
    monkeypatch.setattr(FLAGS, 'CKPT_FORMAT', fmt)
    monkeypatch.setattr(FLAGS, 'SILENT_MODE', True)
    folder, _, save_path = gf.prepare_folder('cifar', 'sngan_rep')
    src = _FakeEngine(3)
    src.D.step.fill_(12)
    src.G.step.fill_(4)                                    # imbalanced update: the generator ran every third step
    paths = [gf.save_checkpoint(src, save_path, s) for s in (3, 9, 12)]
    if fmt == 'npz':
        assert paths[-1] == save_path + '-12.npz' and gf.get_ckpt(folder) == paths[-1]
    else:
        assert paths[-1] == save_path + '-12'
        names = sorted(os.listdir(folder))                 # Saver(max_to_keep=2): the oldest bundle is gone
        assert names == ['checkpoint', 'cifar.ckpt-12.data-00000-of-00001', 'cifar.ckpt-12.index',
                         'cifar.ckpt-9.data-00000-of-00001', 'cifar.ckpt-9.index']
        assert tb.read_checkpoint_state(folder) == (save_path + '-12', [save_path + '-9', save_path + '-12'])
        assert gf.get_ckpt(folder) == save_path + '-12'
        assert gf.get_ckpt(folder, 'cifar.ckpt-9') == save_path + '-9' and gf.get_ckpt(folder, 'cifar.ckpt-3') is None
        z = tb.read_bundle(paths[-1])
        # TF-1.8 AdamOptimizer non-slot variables: beta ** (updates + 1); first optimiser = dis, second (`_1`) = gen
        assert np.isclose(z['beta1_power'], 0.5 ** 13) and np.isclose(z['beta2_power'], 0.999 ** 13)
        assert np.isclose(z['beta1_power_1'], 0.5 ** 5) and np.isclose(z['beta2_power_1'], 0.999 ** 5)
        assert z['global_step'].dtype == np.int32 and z['global_step'].shape == () and int(z['global_step']) == 12
        assert 'dis/l1_f/kernel/kernel/Adam_0' in z and 'gen/l1/kernel/kernel/Adam_1_1' in z and 'dis/adam_step' not in z
    dst = _FakeEngine(50)
    assert gf.rollback(dst, folder) == 12 and dst.global_step == 12
    _same(src, dst)
    assert dst.D.refreshed == 1 and dst.G.refreshed == 1


def test_load_reference_style_bundle_without_adam_slots(tmp_path, monkeypatch):
    """A bundle holding only the model variables (what a Saver over the inference graph writes): weights and state are
    restored, the Adam moments stay, the update counters fall back to global_step."""
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools import graph_func as gf
    monkeypatch.setattr(FLAGS, 'SILENT_MODE', True)
    src, dst = _FakeEngine(7), _FakeEngine(8)
    z = {k: v for k, v in gf.collect_variables(src, 200000, tf_names=True).items() if '/Adam_' not in k and 'power' not in k}
    prefix = str(tmp_path / 'lsun.ckpt-200000')
    tb.write_bundle(prefix, z)
    m0 = dst.D.m.clone()
    assert gf.load_checkpoint(dst, prefix) == 200000
    assert torch.equal(dst.D.w, src.D.w) and torch.equal(dst.G.w, src.G.w) and torch.equal(dst.D.m, m0)
    assert int(dst.D.step) == 200000 and int(dst.G.step) == 200000
    # beta2_power below float32's normal range (> ~87 000 updates): same fallback
    z2 = gf.collect_variables(src, 150000, tf_names=True)
    z2['beta2_power'] = np.asarray(0.0, np.float32)
    assert gf._adam_steps(z2, 0, 'dis', 150000) == 150000
    with pytest.raises(FileNotFoundError, match='No ckpt Model found'):
        gf.load_checkpoint(dst, str(tmp_path / 'missing.ckpt-1'))
    # npz and bundle of the same step in one folder: the bundle is returned
    folder = str(tmp_path)
    open(os.path.join(folder, 'lsun.ckpt-200000.npz'), 'wb').close()
    assert gf.get_ckpt(folder) == prefix


def test_print_tensor_in_ckpt(tmp_path, monkeypatch, capsys):
    from mmdgan_b200.GeneralTools.misc_fun import FLAGS
    from mmdgan_b200.GeneralTools import graph_func as gf
    monkeypatch.setattr(FLAGS, 'DEFAULT_OUT', str(tmp_path) + '/')
    monkeypatch.setattr(FLAGS, 'SILENT_MODE', True)
    folder, _, save_path = gf.prepare_folder('cifar', 'sngan_rep')
    with pytest.raises(FileNotFoundError):
        gf.print_tensor_in_ckpt('cifar_ckpt/sngan_rep')
    gf.save_checkpoint(_FakeEngine(1), save_path, 3, ckpt_format='tf')
    listing = gf.print_tensor_in_ckpt(['cifar_ckpt/sngan_rep'])
    out = capsys.readouterr().out
    assert 'dis/l1_f/kernel/kernel (float32) [3, 3, 3, 16]' in out and 'global_step (int32) []' in out
    assert listing['gen/l1/kernel/kernel/Adam_1_1'] == (np.dtype(np.float32), (16, 64))
    gf.save_checkpoint(_FakeEngine(1), save_path, 4, ckpt_format='npz')
    assert gf.print_tensor_in_ckpt('cifar_ckpt/sngan_rep', all_tensor_values=True)['global_step'][1] == ()


def test_against_tensorboards_tensorflow_protos_and_crc(tmp_path):
    """Independent pins from Google's own code shipped with TensorBoard (tensorboard.compat): the protobuf-generated
    TensorShapeProto / VersionDef serialisers, the DataType enum, TF's masked crc32c, and its TFRecord reader reading a file
    written by input_func.write_tfrecords.  (tensor_bundle.proto itself is not part of TensorBoard.)"""
    pytest.importorskip('tensorboard')
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as pw
    from mmdgan_b200.GeneralTools.input_func import masked_crc32c, write_tfrecords, encode_example
    # the shape sub-message of BundleEntryProto (field 2) and the version sub-message of BundleHeaderProto (field 3)
    for shape in ((), (7,), (3, 3, 128, 256), (0, 4), (1, 3, 32, 32)):
        ref = tensor_shape_pb2.TensorShapeProto()
        for dsize in shape:
            ref.dim.add().size = dsize
        ours = tb._encode_entry(1, shape, 0, 0, 0, 0)
        fields = dict((f, v) for f, _, v in tb._parse_fields(ours))
        assert bytes(fields[2]) == ref.SerializeToString(), shape
        back = tensor_shape_pb2.TensorShapeProto.FromString(bytes(fields[2]))
        assert [d.size for d in back.dim] == list(shape)
    hdr = dict((f, v) for f, _, v in tb._parse_fields(tb._encode_header(1)))
    assert bytes(hdr[3]) == versions_pb2.VersionDef(producer=1).SerializeToString() and hdr[1] == 1
    enum = {np.float32: types_pb2.DT_FLOAT, np.float64: types_pb2.DT_DOUBLE, np.int32: types_pb2.DT_INT32,
            np.uint8: types_pb2.DT_UINT8, np.int16: types_pb2.DT_INT16, np.int8: types_pb2.DT_INT8, np.int64: types_pb2.DT_INT64,
            np.bool_: types_pb2.DT_BOOL, np.uint16: types_pb2.DT_UINT16, np.float16: types_pb2.DT_HALF,
            np.uint32: types_pb2.DT_UINT32, np.uint64: types_pb2.DT_UINT64}
    assert {np.dtype(k): v for k, v in enum.items()} == tb._DTYPE_ENUM
    rng = np.random.default_rng(5)
    for n in (0, 1, 9, 1000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert pw.masked_crc32c(data) == masked_crc32c(data) == tb.mask_crc(crc32c(data))
    big = rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    assert pw.crc32c(big) == tb.crc32c_fast(big)
    # TFRecord framing: TensorBoard's reader accepts what the writer of the input pipeline produces
    recs = [encode_example({'x': rng.integers(0, 256, 3 * 8 * 8, dtype=np.uint8).tobytes()}) for _ in range(5)]
    path = str(tmp_path / 'toy.tfrecords')
    write_tfrecords(path, recs)
    reader = pw.PyRecordReader_New(path)
    got = []
    while True:
        try:
            reader.GetNext()
        except Exception:
            break
        got.append(reader.record())
    assert got == recs
