"""TFRecord / tf.Example reader of the reference's dataset files (Input_Pipeline/*.py) without TensorFlow: record framing
against a hand-assembled file (independent bit-wise CRC from oracle/bundle_format.py), the proto encoding byte for byte,
round trips, corruption detection, the file naming of get_filenames.  Host-only."""
import os
import struct

import numpy as np
import pytest

from oracle import bundle_format as bf
from tgan import tfrecord as tr


def _frame(data):
    head = struct.pack('<Q', len(data))
    return head + struct.pack('<I', bf.mask(bf.crc32c_bitwise(head))) + data + struct.pack('<I', bf.mask(bf.crc32c_bitwise(data)))


def test_record_framing_hand_assembled(tmp_path):
    p = str(tmp_path / 'a.tfrecords')
    payloads = [b'', b'x', bytes(range(256)) * 3]
    open(p, 'wb').write(b''.join(_frame(d) for d in payloads))
    assert list(tr.read_records(p)) == payloads
    q = str(tmp_path / 'b.tfrecords')
    tr.write_records(q, payloads)
    assert open(q, 'rb').read() == open(p, 'rb').read()          # the writer produces exactly the hand-assembled bytes


def test_example_bytes():
    # Example{features{feature{"label": Feature{int64_list{value:[7]}}}}}, packed
    want = bytes([0x0a, 0x10, 0x0a, 0x0e, 0x0a, 0x05]) + b'label' + bytes([0x12, 0x05, 0x1a, 0x03, 0x0a, 0x01, 0x07])
    assert tr.make_example({'label': 7}) == want
    assert tr.parse_example(want) == {'label': [7]}
    # the unpacked encoding of the same list (value as a varint field) parses to the same thing
    unpacked = bytes([0x0a, 0x0f, 0x0a, 0x0d, 0x0a, 0x05]) + b'label' + bytes([0x12, 0x04, 0x1a, 0x02, 0x08, 0x07])
    assert tr.parse_example(unpacked) == {'label': [7]}
    ex = tr.make_example({'image': b'\x00\x01\xff', 'label': 300, 'height': 1, 'width': 1, 'w': [0.5, -2.0], 'neg': -3})
    got = tr.parse_example(ex)
    assert got == {'image': [b'\x00\x01\xff'], 'label': [300], 'height': [1], 'width': [1], 'w': [0.5, -2.0], 'neg': [-3]}
    # unpacked float (fixed32 field)
    fl = bytes([0x0a, 0x0e, 0x0a, 0x0c, 0x0a, 0x01]) + b'w' + bytes([0x12, 0x07, 0x12, 0x05, 0x0d]) + struct.pack('<f', 1.5)
    assert tr.parse_example(fl) == {'w': [1.5]}


@pytest.mark.parametrize('name,shape', [('cifar10', (32, 32, 3)), ('mnist', (28, 28, 1))])
def test_dataset_roundtrip_and_names(tmp_path, name, shape):
    rng = np.random.default_rng(3)
    d = tmp_path / 'Tfrecord'
    d.mkdir()

    class Cfg:
        DATA_NAME = name
        NUM_CLASSES = 10
    nl = 40
    ds = tr.RecordDataset(str(tmp_path), Cfg, num_label=nl, subset='train')
    fl, fu = ds.get_filenames()
    assert os.path.basename(fl) == '%s_train_%s.tfrecords' % (name, str(nl).zfill(6))
    assert os.path.basename(fu) == '%s_train_%s.tfrecords' % (name, str(tr.TRAIN_SIZE[name] - nl).zfill(6))
    il, ll = rng.integers(0, 256, (nl,) + shape, dtype=np.uint8), rng.integers(0, 10, nl)
    iu, lu = rng.integers(0, 256, (70,) + shape, dtype=np.uint8), rng.integers(0, 10, 70)
    tr.write_image_records(fl, il, ll)
    tr.write_image_records(fu, iu, lu)
    (gl, gll), (gu, glu) = ds.load()
    assert gl.dtype == np.uint8 and gl.shape == (nl,) + shape and np.array_equal(gl, il) and np.array_equal(gll, ll)
    assert np.array_equal(gu, iu) and np.array_equal(glu, lu)
    te = tr.RecordDataset(str(tmp_path), Cfg, subset='test')
    assert [os.path.basename(f) for f in te.get_filenames()] == ['%s_test.tfrecords' % name]
    with pytest.raises(AssertionError, match='Invalid data subset'):
        tr.RecordDataset(str(tmp_path), Cfg, subset='val').get_filenames()
    with pytest.raises(ValueError):
        tr.RecordDataset(str(tmp_path), Cfg, data_name='imagenet')


def test_corruption_and_bad_files(tmp_path):
    p = str(tmp_path / 'c.tfrecords')
    img = np.arange(2 * 4 * 4 * 3, dtype=np.uint8).reshape(2, 4, 4, 3)
    tr.write_image_records(p, img, [1, 2])
    raw = bytearray(open(p, 'rb').read())
    bad = bytearray(raw)
    bad[bytes(raw).index(img[0].tobytes()) + 5] ^= 1          # one bit inside the first image's pixel bytes
    open(p, 'wb').write(bad)
    with pytest.raises(ValueError, match='corrupted record data'):
        tr.load_image_records(p, 3)
    got = tr.load_image_records(p, 3, verify=False)[0]                             # checksums ignored: the flipped pixel comes through
    assert got.shape == (2, 4, 4, 3) and (got != img).sum() == 1
    bad = bytearray(raw)
    bad[0] ^= 1
    open(p, 'wb').write(bad)
    with pytest.raises(ValueError, match='corrupted record length'):
        list(tr.read_records(p))
    open(p, 'wb').write(raw[:-3])
    with pytest.raises(ValueError, match='truncated'):
        list(tr.read_records(p))
    open(p, 'wb').write(raw)
    with pytest.raises(ValueError, match='is not 4x4x1'):
        tr.load_image_records(p, 1)
    tr.write_records(p, [tr.make_example({'image': b'abc', 'label': 1})])
    with pytest.raises(ValueError, match="feature 'height' missing"):
        tr.load_image_records(p, 3)
    open(p, 'wb').write(b'')
    with pytest.raises(ValueError, match='no records'):
        tr.load_image_records(p, 3)


def test_reader_on_committed_fixture(tmp_path):
    """tests/golden/fmt_records.tfrecords: written by the independent encoder of oracle/gen_format_fixtures.py; the product
    writer reproduces the file byte for byte"""
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    want = np.load(os.path.join(g, 'fmt_expected.npz'))
    im, lb = tr.load_image_records(os.path.join(g, 'fmt_records.tfrecords'), 3)
    assert np.array_equal(im, want['images']) and np.array_equal(lb, want['labels'])
    p = str(tmp_path / 'w.tfrecords')
    tr.write_image_records(p, want['images'], want['labels'])
    assert open(p, 'rb').read() == open(os.path.join(g, 'fmt_records.tfrecords'), 'rb').read()
