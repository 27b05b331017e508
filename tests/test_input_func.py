"""CPU: TFRecord / tf.train.Example codec and the ReadTFRecords pipeline (input_func.py:55-105, 721-965)."""
import struct

import numpy as np
import pytest


def test_crc32c_known_answers():
    from mmdgan_b200.GeneralTools import input_func as inp
    assert inp.crc32c(b'123456789') == 0xE3069283            # the standard CRC-32C check value
    assert inp.crc32c(b'') == 0
    assert inp.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros


def test_example_wire_format_is_protobuf():
    from mmdgan_b200.GeneralTools import input_func as inp
    ex = inp.encode_example({'x': inp._bytes_feature(b'\x01\x02\x03'), 'y': inp._int64_feature(7)})
    # Example{1: Features{1: {1:'x', 2: Feature{1: BytesList{1: 01 02 03}}}, 1: {1:'y', 2: Feature{3: Int64List{1: [7]}}}}}
    expect = bytes.fromhex('0a1a' '0a0c' '0a0178' '1207' '0a05' '0a03' '010203' '0a0a' '0a0179' '1205' '1a03' '0a01' '07')
    assert ex == expect                                      # hand-assembled nested length-delimited fields
    dec = inp.decode_example(ex)
    assert dec['x'] == b'\x01\x02\x03' and dec['y'].tolist() == [7]
    # an unpacked int64 list (older writers) decodes as well
    feat = inp._ld(3, inp._varint((1 << 3) | 0) + inp._varint(9))
    assert inp.decode_example(inp.encode_example({'y': feat}))['y'].tolist() == [9]


def test_tfrecord_round_trip_and_corruption(tmp_path):
    from mmdgan_b200.GeneralTools import input_func as inp
    rng = np.random.RandomState(0)
    data = rng.randint(0, 256, size=(37, 3 * 8 * 8)).astype(np.uint8)
    path = inp.my_np2tfrecord('toy', data, file_folder=str(tmp_path))
    recs = list(inp.read_tfrecords(path, check_crc=True))
    assert len(recs) == 37
    assert np.array_equal(np.frombuffer(inp.decode_example(recs[5])['x'], np.uint8), data[5])
    # container framing: uint64 length, masked crc of the length, payload, masked crc of the payload
    raw = open(path, 'rb').read()
    (n0,) = struct.unpack('<Q', raw[:8])
    assert raw[12:12 + n0] == recs[0]
    bad = bytearray(raw)
    bad[3] ^= 0xFF
    open(path, 'wb').write(bytes(bad))
    with pytest.raises(IOError):
        list(inp.read_tfrecords(path))
    with pytest.raises(AttributeError):
        inp.my_np2tfrecord('toy2', data.astype(np.float64), file_folder=str(tmp_path))


def test_read_tfrecords_pipeline(tmp_path):
    from mmdgan_b200.GeneralTools import input_func as inp
    data = np.random.RandomState(3).randint(0, 256, size=(50, 192)).astype(np.uint8)
    labels = (np.arange(50) % 10).reshape(50, 1)
    inp.my_np2tfrecord('lab', data, labels, file_folder=str(tmp_path))
    r = inp.ReadTFRecords('lab', 192, num_labels=1, batch_size=16, file_folder=str(tmp_path), buffer_size=20, seed=1)
    r.shape2image(3, 8, 8)
    seen = []
    for _ in range(10):                                         # 160 draws from 50 examples: dataset.repeat()
        b = r.next_batch()
        assert b['x'].shape == (16, 3, 8, 8) and b['x'].dtype == np.float32
        assert b['x'].min() >= -1.0 and b['x'].max() <= 1.0
        assert b['y'].shape == (16, 1) and b['y'].dtype == np.int32
        # the pixel scaling of input_func.py:839 and the CHW reshape
        back = np.rint((b['x'].reshape(16, -1) + 1.0) * 127.5).astype(np.uint8)
        for row, y in zip(back, b['y'][:, 0]):
            idx = int(np.where((data == row).all(1))[0][0])
            assert idx % 10 == y
            seen.append(idx)
    assert len(set(seen)) == 50 and seen[:16] != sorted(seen[:16])      # everything is visited; order is shuffled
    # without shuffling the file order is kept
    r2 = inp.ReadTFRecords('lab', 192, batch_size=8, file_folder=str(tmp_path), buffer_size=1)
    r2.shape2image(3, 8, 8)
    x = r2.next_batch()['x']
    assert np.allclose(x.reshape(8, -1), data[:8].astype(np.float32) / 127.5 - 1.0)
    with pytest.raises(AssertionError):
        inp.ReadTFRecords('missing', 192, file_folder=str(tmp_path))
