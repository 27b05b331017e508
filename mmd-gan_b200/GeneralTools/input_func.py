"""Input pipeline of the reference without TensorFlow: TFRecord files of tf.train.Example{x: bytes(uint8 CHW)[, y: int64]}.

Mirrors GeneralTools/input_func.py for the path's data contract:
  my_np2tfrecord(filename, data, label)                       input_func.py:55-105   (writer, used to prepare datasets)
  ReadTFRecords(...).shape2image(C, H, W).next_batch()        input_func.py:721-965  (parse -> uint8 -> float32 ->
      x / 127.5 - 1 -> reshape CHW -> shuffle(buffer 10000) -> batch -> repeat)  ->  {'x': [B, C, H, W] float32 in [-1, 1]}

The TFRecord container (length, masked crc32c, payload, masked crc32c) and the protobuf wire format of
tf.train.Example are decoded here in pure Python / numpy (third-party formats: TensorFlow 1.8 `tf.python_io` and
protobuf 3, neither installable in this image).  Host-side code: it produces the NCHW batch that
SNGanEngine.step() copies to the GPU; no arithmetic of the training step happens here.
"""
import os
import struct

import numpy as np

from .misc_fun import FLAGS

# ---------------------------------------------------------------------------------------------- crc32c (Castagnoli)
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = tab
    return _CRC_TABLE


def crc32c(data):
    tab = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- protobuf wire format
def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _ld(field, payload):            # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _bytes_feature(value):          # Feature{bytes_list = 1: BytesList{value = 1}}
    return _ld(1, _ld(1, value))


def _int64_feature(value):          # Feature{int64_list = 3: Int64List{value = 1 (packed)}}
    return _ld(3, _ld(1, _varint(int(value))))


def _float_feature(values):         # Feature{float_list = 2: FloatList{value = 1 (packed)}}
    return _ld(2, _ld(1, np.asarray(values, dtype='<f4').tobytes()))


def encode_example(features):
    """features: {name: serialized Feature} -> serialized tf.train.Example (Example{features = 1: Features{feature = 1: map}})."""
    entries = b''
    for key in sorted(features):
        entries += _ld(1, _ld(1, key.encode()) + _ld(2, features[key]))
    return _ld(1, entries)


def _parse_fields(buf):
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 2:
            n, pos = _read_varint(buf, pos)
            out.append((field, wire, buf[pos:pos + n]))
            pos += n
        elif wire == 0:
            v, pos = _read_varint(buf, pos)
            out.append((field, wire, v))
        elif wire == 5:
            out.append((field, wire, buf[pos:pos + 4]))
            pos += 4
        elif wire == 1:
            out.append((field, wire, buf[pos:pos + 8]))
            pos += 8
        else:
            raise ValueError('unsupported protobuf wire type {}'.format(wire))
    return out


def decode_example(buf):
    """serialized tf.train.Example -> {name: bytes | np.int64 array | np.float32 array}."""
    out = {}
    for f, _, features in _parse_fields(buf):
        if f != 1:
            continue
        for f2, _, entry in _parse_fields(features):
            if f2 != 1:
                continue
            key, feat = None, None
            for f3, _, v in _parse_fields(entry):
                if f3 == 1:
                    key = bytes(v).decode()
                elif f3 == 2:
                    feat = v
            for kind, _, lst in _parse_fields(feat):
                vals = _parse_fields(lst)
                if kind == 1:                                    # bytes_list
                    out[key] = bytes(vals[0][2])
                elif kind == 2:                                  # float_list (packed or not)
                    raw = b''.join(bytes(v[2]) for v in vals)
                    out[key] = np.frombuffer(raw, dtype='<f4').copy()
                elif kind == 3:                                  # int64_list (packed or not)
                    ints = []
                    for _, wire, v in vals:
                        if wire == 0:
                            ints.append(v)
                        else:
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                ints.append(x)
                    out[key] = np.asarray(ints, dtype=np.uint64).astype(np.int64)
    return out


# ---------------------------------------------------------------------------------------------- TFRecord container
def write_tfrecords(path, records):
    with open(path, 'wb') as f:
        for rec in records:
            head = struct.pack('<Q', len(rec))
            f.write(head)
            f.write(struct.pack('<I', masked_crc32c(head)))
            f.write(rec)
            f.write(struct.pack('<I', masked_crc32c(rec)))


def read_tfrecords(path, check_crc=False):
    """Yields the payload of every record; the 12-byte header crc is always checked, the payload crc on request."""
    with open(path, 'rb') as f:
        data = f.read()
    pos, n = 0, len(data)
    while pos < n:
        head = data[pos:pos + 8]
        (length,) = struct.unpack('<Q', head)
        (hcrc,) = struct.unpack('<I', data[pos + 8:pos + 12])
        if hcrc != masked_crc32c(head):
            raise IOError('{}: corrupted record header at byte {}'.format(path, pos))
        rec = data[pos + 12:pos + 12 + length]
        if check_crc:
            (dcrc,) = struct.unpack('<I', data[pos + 12 + length:pos + 16 + length])
            if dcrc != masked_crc32c(rec):
                raise IOError('{}: corrupted record payload at byte {}'.format(path, pos))
        yield rec
        pos += 16 + length


def my_np2tfrecord(filename, data, label=None, file_folder=None):
    """input_func.py:55-105: one Example per row; uint8 rows as a bytes feature, float32 rows as a float list."""
    folder = FLAGS.DEFAULT_IN if file_folder is None else file_folder
    path = os.path.join(folder, filename + '.tfrecords')
    data = np.asarray(data)
    if data.dtype == np.int32:
        data = data.astype(np.float32)
    if data.dtype == np.uint8:
        feature_fun = lambda x: _bytes_feature(x.tobytes())
    elif data.dtype == np.float32:
        feature_fun = _float_feature
    else:
        raise AttributeError('Supported data type: uint8, float32, int32; got {}'.format(data.dtype))
    if label is not None and np.asarray(label).shape[0] != data.shape[0]:
        raise ValueError('Data size and label size do not match.')

    def gen():
        for i in range(data.shape[0]):
            feats = {'x': feature_fun(data[i].reshape(-1))}
            if label is not None:
                feats['y'] = _int64_feature(int(np.asarray(label)[i].reshape(-1)[0]))
            yield encode_example(feats)
    write_tfrecords(path, gen())
    return path


class ReadTFRecords(object):
    """input_func.py:721-965 for the unconditional image case (x_dtype string -> uint8 -> float32)."""

    def __init__(self, filename, num_features=None, num_labels=0, x_dtype='string', y_dtype='int64', batch_size=64,
                 skip_count=0, file_repeat=1, num_epoch=None, file_folder=None, num_threads=8, buffer_size=10000,
                 shuffle_file=False, seed=None):
        folder = FLAGS.DEFAULT_IN if file_folder is None else file_folder
        names = [filename] if isinstance(filename, str) else list(filename)
        files = [os.path.join(folder, n + '.tfrecords') for n in names]
        for file in files:
            assert os.path.isfile(file), 'File {} does not exist.'.format(file)
        if file_repeat > 1:
            files = files * int(file_repeat)
        self.rng = np.random.RandomState(seed)
        if shuffle_file:
            self.rng.shuffle(files)
        self.files = files
        self.num_features, self.num_labels = num_features, num_labels
        self.x_dtype, self.batch_size, self.buffer_size = x_dtype, batch_size, buffer_size
        self.image_shape = None
        self._cache = {}
        self._stream = None

    def shape2image(self, channels, height, width, resize=None):
        if resize is not None:
            raise NotImplementedError('resize is not on the hot path')
        if FLAGS.IMAGE_FORMAT != 'channels_first':
            raise NotImplementedError('channels_last is not on the hot path')
        self.image_shape = (channels, height, width)

    def _load(self, path):
        if path not in self._cache:
            xs, ys = [], []
            for rec in read_tfrecords(path):
                ex = decode_example(rec)
                x = ex['x']
                x = np.frombuffer(x, dtype=np.uint8) if isinstance(x, bytes) else x       # tf.decode_raw(..., tf.uint8)
                if self.num_features is not None:
                    assert x.size == self.num_features, 'record has {} features, expected {}'.format(x.size, self.num_features)
                xs.append(x)
                if self.num_labels:
                    ys.append(ex['y'][:self.num_labels].astype(np.int32))
            self._cache[path] = (np.stack(xs), np.stack(ys) if ys else None)
        return self._cache[path]

    def _examples(self):
        while True:                                            # dataset.repeat()
            for path in self.files:
                xs, ys = self._load(path)
                for i in range(xs.shape[0]):
                    yield xs[i], (None if ys is None else ys[i])

    def _shuffled(self):
        """tf.data shuffle(buffer_size): keep a buffer, emit a uniformly random element, refill from the stream."""
        src = self._examples()
        buf = [next(src) for _ in range(self.buffer_size)] if self.buffer_size > 1 else []
        if not buf:
            for item in src:
                yield item
        while True:
            k = self.rng.randint(len(buf))
            item = buf[k]
            buf[k] = next(src)
            yield item

    def next_batch(self, sample_same_class=False):
        """{'x': float32 [B, C, H, W] in [-1, 1] (x / 127.5 - 1, input_func.py:839)[, 'y': int32 [B, num_labels]]}."""
        if sample_same_class:
            raise NotImplementedError('sample_same_class is not on the hot path')
        if self._stream is None:
            total = sum(self._load(p)[0].shape[0] for p in set(self.files))
            self.buffer_size = max(1, min(self.buffer_size, total))
            self._stream = self._shuffled()
        items = [next(self._stream) for _ in range(self.batch_size)]
        x = np.stack([it[0] for it in items]).astype(np.float32) / 127.5 - 1.0
        if self.image_shape is not None:
            x = x.reshape((self.batch_size,) + self.image_shape)
        out = {'x': x}
        if self.num_labels:
            out['y'] = np.stack([it[1] for it in items])
        return out
