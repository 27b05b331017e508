"""TensorFlow checkpoint files ("tensor bundle", the V2 format `tf.train.Saver` writes) without TensorFlow.

The reference saves and restores its variables through `tf.train.Saver(var_list=tf.global_variables(), max_to_keep=2)`
(GeneralTools/graph_func.py:708-717, 719-760, 606-636) and finds the latest file through
`tf.train.get_checkpoint_state` (graph_func.py:399-416).  On disk that is

    <folder>/checkpoint                              text proto CheckpointState
    <folder>/<file>.ckpt-<step>.index                an SSTable (LevelDB table format) name -> BundleEntryProto
    <folder>/<file>.ckpt-<step>.data-00000-of-00001  the raw little-endian tensor bytes, back to back

This module reads and writes those three files in pure Python / numpy so that weights trained with the reference can be
loaded into the engine (and the other way round).  Third-party formats, restated from their published definitions
(TensorFlow 1.8 `tensorflow/core/util/tensor_bundle`, `tensorflow/core/lib/io/table*`, `tensor_bundle.proto`): the
reference repository ships no checkpoint, so there is NO TF-written file to pin this codec against -- the tests check the
round trip, the block / footer structure against the format's constants, crc32c against its published check values, and
the pieces TensorBoard's bundled TensorFlow stubs cover (protobuf-generated TensorShapeProto / VersionDef, the DataType enum,
TF's masked crc32c) against those.
Host-side code only; no arithmetic of the training step happens here.
"""
import os
import struct

import numpy as np

from .input_func import _ld, _parse_fields, _read_varint, _varint, crc32c

TABLE_MAGIC = 0xdb4775248b80fb57           # leveldb kTableMagicNumber, stored little-endian in the last 8 bytes
FOOTER_BYTES = 48                          # two padded block handles (2 x 20 bytes) + magic
BLOCK_TRAILER_BYTES = 5                    # 1 byte compression type + 4 bytes masked crc32c
RESTART_INTERVAL = 16
BLOCK_SIZE = 262144                        # tensorflow::table::Options::block_size

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_ENUM = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------- crc32c for megabytes
_CRC_BYTE_TABLE = None


def _byte_table():
    global _CRC_BYTE_TABLE
    if _CRC_BYTE_TABLE is None:
        tab = np.arange(256, dtype=np.uint32)
        for _ in range(8):
            tab = np.where(tab & 1, (tab >> 1) ^ np.uint32(0x82F63B78), tab >> 1).astype(np.uint32)
        _CRC_BYTE_TABLE = tab
    return _CRC_BYTE_TABLE


def _advance_tables(nbytes):
    """The linear map "feed `nbytes` zero bytes to the raw crc register" as four 256-entry lookup tables (one per register byte)."""
    tab = _byte_table()
    basis = (np.uint32(1) << np.arange(32, dtype=np.uint32)).astype(np.uint32)     # images of the 32 unit vectors
    n = nbytes
    # feeding one zero byte: reg -> tab[reg & 0xFF] ^ (reg >> 8); compose by repeated doubling on the basis images
    def step(v):
        return tab[v & np.uint32(0xFF)] ^ (v >> np.uint32(8))

    def apply(mat, v):                      # mat: images of the unit vectors; v: uint32 array
        out = np.zeros_like(v)
        for bit in range(32):
            out ^= np.where((v >> np.uint32(bit)) & np.uint32(1), mat[bit], np.uint32(0)).astype(np.uint32)
        return out

    result = basis.copy()                   # identity
    power = step(basis)                     # one zero byte
    while n:
        if n & 1:
            result = apply(power, result)
        n >>= 1
        if n:
            power = apply(power, power)
    tables = []
    for byte in range(4):
        vals = (np.arange(256, dtype=np.uint32) << np.uint32(8 * byte)).astype(np.uint32)
        tables.append(apply(result, vals))
    return tables


def crc32c_fast(data, chunk=8192):
    """crc32c (Castagnoli) of a bytes-like object.  The crc register is linear in the data, so the buffer is cut into equal
    chunks whose raw registers are computed side by side (one numpy operation per byte POSITION instead of per byte) and then
    folded together with the "append `chunk` zero bytes" operator.  Same value as input_func.crc32c, ~100x faster on
    megabyte tensors (a 6 M-parameter network with its Adam slots is 72 MB)."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    n = int(buf.size)
    if n < 4 * chunk:
        return crc32c(buf.tobytes())
    tab = _byte_table()
    nchunks = -(-n // chunk)
    padded = np.zeros(nchunks * chunk, dtype=np.uint8)
    padded[nchunks * chunk - n:] = buf           # leading zeros leave a zero register unchanged
    cols = np.ascontiguousarray(padded.reshape(nchunks, chunk).T)
    regs = np.zeros(nchunks, dtype=np.uint32)
    for j in range(chunk):
        regs = tab[(regs ^ cols[j]) & np.uint32(0xFF)] ^ (regs >> np.uint32(8))
    t0, t1, t2, t3 = (t.tolist() for t in _advance_tables(chunk))
    acc = 0
    for r in regs.tolist():                      # Horner: acc = advance(acc, chunk) ^ raw(chunk_i)
        acc = t0[acc & 0xFF] ^ t1[(acc >> 8) & 0xFF] ^ t2[(acc >> 16) & 0xFF] ^ t3[acc >> 24] ^ r
    a0, a1, a2, a3 = (t.tolist() for t in _advance_tables(n))
    init = a0[0xFF] ^ a1[0xFF] ^ a2[0xFF] ^ a3[0xFF]      # the 0xFFFFFFFF preset pushed through n bytes
    return (acc ^ init ^ 0xFFFFFFFF) & 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- SSTable (LevelDB table format)
def _block_handle(offset, size):
    return _varint(offset) + _varint(size)


class _BlockBuilder(object):
    """Entries `varint shared | varint non_shared | varint value_len | key suffix | value`, a restart point (full key) every
    `interval` entries, then the fixed32 restart offsets and their count."""

    def __init__(self, interval):
        self.interval = interval
        self.reset()

    def reset(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b''
        self.entries = 0

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            limit = min(len(key), len(self.last_key))
            while shared < limit and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1
        self.entries += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=BLOCK_SIZE):
    """items: (key bytes, value bytes) pairs in strictly increasing bytewise key order."""
    out = bytearray()

    def emit(contents):                     # block + trailer (type 0 = uncompressed, masked crc32c of contents + type)
        offset = len(out)
        out.extend(contents)
        out.append(0)
        out.extend(struct.pack('<I', mask_crc(crc32c_fast(contents + b'\x00'))))
        return offset, len(contents)

    data = _BlockBuilder(RESTART_INTERVAL)
    index = _BlockBuilder(1)
    prev = None
    for key, value in items:
        assert prev is None or key > prev, 'table keys must be strictly increasing'
        prev = key
        data.add(key, value)
        if data.size_estimate() >= block_size:
            last = data.last_key
            index.add(last, _block_handle(*emit(data.finish())))    # any key in [last key, next first key) separates the blocks
            data.reset()
    if data.entries:
        last = data.last_key
        index.add(last, _block_handle(*emit(data.finish())))
    meta = _BlockBuilder(RESTART_INTERVAL)
    meta_handle = _block_handle(*emit(meta.finish()))
    index_handle = _block_handle(*emit(index.finish()))
    footer = meta_handle + index_handle
    footer += b'\x00' * (40 - len(footer))
    footer += struct.pack('<II', TABLE_MAGIC & 0xFFFFFFFF, TABLE_MAGIC >> 32)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(bytes(out))


def _read_block(data, offset, size, check_crc):
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    if check_crc:
        (stored,) = struct.unpack('<I', data[offset + size + 1:offset + size + 5])
        if unmask_crc(stored) != crc32c_fast(data[offset:offset + size + 1]):
            raise ValueError('table block at offset {}: crc32c mismatch'.format(offset))
    if ctype != 0:
        raise NotImplementedError('compressed table block (type {}): TensorFlow writes checkpoints uncompressed'.format(ctype))
    return contents


def _block_entries(block):
    (nrestarts,) = struct.unpack('<I', block[-4:])
    end = len(block) - 4 - 4 * nrestarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, check_crc=True):
    """-> list of (key, value) in file order."""
    with open(path, 'rb') as f:
        data = f.read()
    if len(data) < FOOTER_BYTES:
        raise ValueError('{} is too short to be a table file'.format(path))
    footer = data[-FOOTER_BYTES:]
    lo, hi = struct.unpack('<II', footer[40:])
    if (hi << 32) | lo != TABLE_MAGIC:
        raise ValueError('{} is not a table file (bad magic number)'.format(path))
    pos = 0
    _, pos = _read_varint(footer, pos)      # metaindex handle (unused: no filter / meta blocks in a checkpoint index)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, check_crc)):
        boff, p = _read_varint(handle, 0)
        bsize, p = _read_varint(handle, p)
        out.extend(_block_entries(_read_block(data, boff, bsize, check_crc)))
    return out


# ---------------------------------------------------------------------------------------------- tensor_bundle.proto
def _encode_header(num_shards=1):
    # BundleHeaderProto{num_shards = 1; endianness = 2 (LITTLE = 0, default: omitted); version = 3: VersionDef{producer = 1}}
    return _varint((1 << 3) | 0) + _varint(num_shards) + _ld(3, _varint((1 << 3) | 0) + _varint(1))


def _encode_entry(dtype_enum, shape, shard_id, offset, size, crc_masked):
    # BundleEntryProto{dtype = 1; shape = 2: TensorShapeProto{dim = 2: Dim{size = 1}}; shard_id = 3; offset = 4; size = 5;
    #                  crc32c = 6 (fixed32)}; proto3: zero-valued scalars are omitted
    out = _varint((1 << 3) | 0) + _varint(dtype_enum)
    dims = b''.join(_ld(2, (_varint((1 << 3) | 0) + _varint(int(d))) if int(d) else b'') for d in shape)
    out += _ld(2, dims)
    if shard_id:
        out += _varint((3 << 3) | 0) + _varint(shard_id)
    if offset:
        out += _varint((4 << 3) | 0) + _varint(offset)
    if size:
        out += _varint((5 << 3) | 0) + _varint(size)
    out += _varint((6 << 3) | 5) + struct.pack('<I', crc_masked)
    return out


def _decode_entry(buf):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for field, wire, val in _parse_fields(buf):
        if field == 1:
            e['dtype'] = val
        elif field == 2:
            for f2, _, dim in _parse_fields(val):
                if f2 == 2:
                    size = 0
                    for f3, _, v in _parse_fields(dim):
                        if f3 == 1:
                            size = v
                    e['shape'].append(size)
        elif field == 3:
            e['shard_id'] = val
        elif field == 4:
            e['offset'] = val
        elif field == 5:
            e['size'] = val
        elif field == 6:
            (e['crc32c'],) = struct.unpack('<I', val)
        elif field == 7:
            e['sliced'] = True
    return e


def _decode_header(buf):
    h = dict(num_shards=0, endianness=0)
    for field, wire, val in _parse_fields(buf):
        if field == 1:
            h['num_shards'] = val
        elif field == 2:
            h['endianness'] = val
    return h


def _shard_name(prefix, shard, num_shards):
    return '{}.data-{:05d}-of-{:05d}'.format(prefix, shard, num_shards)


def write_bundle(prefix, tensors):
    """tensors: {variable name: numpy array}.  Writes <prefix>.index and <prefix>.data-00000-of-00001 (one shard, tensors in
    name order, as a Saver on one device does)."""
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b'', _encode_header(1))]
    offset = 0
    with open(_shard_name(prefix, 0, 1), 'wb') as f:
        for name in names:
            arr = np.asarray(tensors[name])
            shape = arr.shape                               # (ascontiguousarray turns a 0-d array into shape (1,))
            arr = np.ascontiguousarray(arr)
            if arr.dtype not in _DTYPE_ENUM:
                raise TypeError('variable {}: dtype {} has no checkpoint encoding here'.format(name, arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder('<'), copy=False).tobytes()
            f.write(raw)
            items.append((name.encode(), _encode_entry(_DTYPE_ENUM[arr.dtype], shape, 0, offset, len(raw),
                                                       mask_crc(crc32c_fast(raw)))))
            offset += len(raw)
    write_table(prefix + '.index', items)
    return prefix


def list_bundle(prefix):
    """-> {variable name: (numpy dtype, shape)} without touching the data shards (cf. print_tensor_in_ckpt, graph_func.py:419-437)."""
    out = {}
    for key, value in read_table(prefix + '.index'):
        if key == b'':
            continue
        e = _decode_entry(value)
        out[key.decode()] = (np.dtype(_DTYPES[e['dtype']]) if e['dtype'] in _DTYPES else None, tuple(e['shape']))
    return out


def read_bundle(prefix, names=None, check_crc=True):
    """-> {variable name: numpy array}; `names` restricts the read.  Raises on big-endian bundles, sliced (partitioned)
    variables and dtypes outside the numeric set."""
    entries, header = {}, None
    for key, value in read_table(prefix + '.index'):
        if key == b'':
            header = _decode_header(value)
        else:
            entries[key.decode()] = _decode_entry(value)
    if header is None:
        raise ValueError('{}.index has no bundle header'.format(prefix))
    if header['endianness'] != 0:
        raise NotImplementedError('big-endian tensor bundle')
    wanted = list(entries) if names is None else list(names)
    shards, out = {}, {}
    for name in wanted:
        if name not in entries:
            raise KeyError('variable {} not found in checkpoint {}'.format(name, prefix))
        e = entries[name]
        if e['sliced']:
            raise NotImplementedError('variable {} is stored as slices (partitioned variable)'.format(name))
        if e['dtype'] not in _DTYPES:
            raise TypeError('variable {}: checkpoint dtype enum {} not supported'.format(name, e['dtype']))
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap(_shard_name(prefix, sid, header['num_shards']), dtype=np.uint8, mode='r')
        raw = np.asarray(shards[sid][e['offset']:e['offset'] + e['size']])
        dt = np.dtype(_DTYPES[e['dtype']]).newbyteorder('<')
        count = int(np.prod(e['shape'])) if e['shape'] else 1
        if raw.size != count * dt.itemsize:
            raise ValueError('variable {}: {} bytes on disk, shape {} needs {}'.format(name, raw.size, e['shape'], count * dt.itemsize))
        if check_crc and e['crc32c'] is not None and unmask_crc(e['crc32c']) != crc32c_fast(raw):
            raise ValueError('variable {}: crc32c mismatch in {}'.format(name, prefix))
        out[name] = raw.view(dt).reshape(e['shape']).astype(dt.newbyteorder('='), copy=True)
    return out


# ---------------------------------------------------------------------------------------------- CheckpointState
def write_checkpoint_state(folder, latest, all_paths=None):
    """<folder>/checkpoint as tf.train.update_checkpoint_state writes it (paths relative to the folder)."""
    all_paths = list(all_paths) if all_paths is not None else [latest]
    # always relative to the state file's folder (TF's generate_checkpoint_state_proto rewrites relative paths against
    # save_dir): a cwd-relative prefix such as 'MMD-GAN/Results/x_ckpt/x.ckpt-3' must not be stored as is, because
    # read_checkpoint_state joins every relative entry onto the folder
    rel = lambda p: os.path.relpath(os.path.abspath(p), os.path.abspath(folder))      # noqa: E731
    lines = ['model_checkpoint_path: "{}"'.format(rel(latest))]
    lines += ['all_model_checkpoint_paths: "{}"'.format(rel(p)) for p in all_paths]
    with open(os.path.join(folder, 'checkpoint'), 'w') as f:
        f.write('\n'.join(lines) + '\n')


def read_checkpoint_state(folder):
    """-> (model_checkpoint_path, all_model_checkpoint_paths) with absolute paths, or None (tf.train.get_checkpoint_state)."""
    path = os.path.join(folder, 'checkpoint')
    if not os.path.isfile(path):
        return None
    latest, every = None, []
    with open(path) as f:
        for line in f:
            key, _, val = line.partition(':')
            val = val.strip()
            if len(val) >= 2 and val[0] == '"' and val[-1] == '"':
                val = val[1:-1]
            if not os.path.isabs(val):
                val = os.path.join(folder, val)
            if key.strip() == 'model_checkpoint_path':
                latest = val
            elif key.strip() == 'all_model_checkpoint_paths':
                every.append(val)
    return (latest, every) if latest is not None else None
