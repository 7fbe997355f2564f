"""TF tensor-bundle checkpoint files (Training/Saver.py): CRC-32C known answers, the table / bundle
writer against an independent parser (oracle/bundle_format.py), the reader against a hand-assembled index, corruption
detection, and the Saver's file-name logic.  Host-only: runs without a GPU."""
import os
import struct

import numpy as np
import pytest

from oracle import bundle_format as bf
from tgan import checkpoint as ck


# RFC 3720 B.4 / LevelDB crc32c_test known answers
KATS = [(b'\x00' * 32, 0x8a9136aa), (b'\xff' * 32, 0x62a8ab43), (bytes(range(32)), 0x46dd794e),
        (bytes(range(31, -1, -1)), 0x113fdb5c), (b'123456789', 0xe3069283), (b'a', 0xc1d04330), (b'', 0)]


@pytest.mark.parametrize('data,want', KATS)
def test_crc32c_known_answers(data, want):
    assert ck.crc32c(data) == want
    assert bf.crc32c_bitwise(data) == want


def test_crc32c_matches_bitwise_on_random_buffers_and_extends():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 64, 65, 1000, 4099):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ck.crc32c(b) == bf.crc32c_bitwise(b)
        k = n // 3
        assert ck.crc32c(b[k:], ck.crc32c(b[:k])) == ck.crc32c(b)          # streaming
    a = rng.standard_normal((33, 17)).astype(np.float32)
    assert ck.crc32c(a) == bf.crc32c_bitwise(a.tobytes())
    assert ck.crc32c(a[:, ::2]) == bf.crc32c_bitwise(np.ascontiguousarray(a[:, ::2]).tobytes())


def test_mask_roundtrip_and_value():
    # LevelDB crc32c_test "Mask": masking is not the identity, not idempotent, and invertible
    c = ck.crc32c(b'foo')
    assert ck.mask_crc(c) != c and ck.mask_crc(ck.mask_crc(c)) != c
    assert ck.unmask_crc(ck.mask_crc(c)) == c
    assert ck.unmask_crc(ck.unmask_crc(ck.mask_crc(ck.mask_crc(c)))) == c
    assert ck.mask_crc(c) == bf.mask(c)
    assert ck.mask_crc(0) == 0xa282ead8


def test_varint():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63 + 5):
        b = ck.put_varint(v)
        assert ck.get_varint(b, 0) == (v, len(b))
        assert bf.varint(b, 0) == (v, len(b))
    assert ck.put_varint(300) == b'\xac\x02'
    with pytest.raises(ValueError):
        ck.get_varint(b'\x80', 0)


def _tensors(rng, n):
    t = {}
    for i in range(n):
        shape = tuple(int(s) for s in rng.integers(1, 6, rng.integers(0, 4)))
        t['classifier/conv%d_%d/V' % (i // 7, i % 7)] = rng.standard_normal(shape).astype(np.float32)
    t['Train/beta1_power'] = np.float32(0.5).reshape(())
    t['global_step'] = np.int64(12345).reshape(())
    t['classifier/labels'] = rng.integers(0, 10, (4, 3)).astype(np.int32)
    return t


@pytest.mark.parametrize('n', [1, 40, 600])          # 600 names -> many 4 KB blocks and a multi-entry index block
def test_bundle_roundtrip_and_independent_parse(tmp_path, n):
    rng = np.random.default_rng(n)
    t = _tensors(rng, n)
    prefix = str(tmp_path / 'model_0003.ckpt')
    ck.write_bundle(prefix, t)
    # (1) our reader
    r = ck.BundleReader(prefix)
    assert sorted(r.keys()) == sorted(t)
    for k, v in t.items():
        got = r.get(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    # (2) the independent parser sees a well-formed table with the same entries
    items = bf.parse_index(open(prefix + '.index', 'rb').read())
    assert items[0] == (b'', b'\x08\x01\x1a\x02\x08\x01')
    assert [k.decode() for k, _ in items[1:]] == sorted(t, key=lambda s: s.encode())
    data = open(prefix + '.data-00000-of-00001', 'rb').read()
    assert len(data) == sum(v.nbytes for v in t.values())
    for k, v in items[1:]:
        e = ck.decode_entry(v)
        raw = data[e['offset']:e['offset'] + e['size']]
        assert raw == t[k.decode()].tobytes()
        assert e['crc32c'] == bf.mask(bf.crc32c_bitwise(raw))
        assert e['shape'] == t[k.decode()].shape and e['shard_id'] == 0


def test_entry_proto_bytes():
    # dtype DT_FLOAT(1), shape [3,3,128,256], offset 4096, size 1179648, crc 0x01020304 -- field by field
    e = ck.encode_entry(np.float32, (3, 3, 128, 256), 4096, 1179648, 0x01020304)
    want = (b'\x08\x01' + b'\x12\x12' + b'\x12\x02\x08\x03' * 2 + b'\x12\x03\x08\x80\x01' + b'\x12\x03\x08\x80\x02'
            + b'\x20\x80\x20' + b'\x28\x80\x80\x48' + b'\x35\x04\x03\x02\x01')
    assert e == want
    d = ck.decode_entry(e)
    assert (d['dtype'], d['shape'], d['offset'], d['size'], d['crc32c']) == (1, (3, 3, 128, 256), 4096, 1179648, 0x01020304)
    # scalar: empty (but present) shape, zero offset omitted
    assert ck.encode_entry(np.float32, (), 0, 4, 7) == b'\x08\x01\x12\x00\x28\x04\x35\x07\x00\x00\x00'
    assert ck.decode_entry(b'\x08\x09\x28\x08')['shape'] == ()        # shape message absent (proto3 default)


def _hand_index(entries):
    """assemble a one-block table byte by byte, without the writer (no prefix compression, restart at every entry)"""
    def block(kvs):
        body, restarts = b'', []
        for k, v in kvs:
            restarts.append(len(body))
            body += bytes([0, len(k), len(v)]) + k + v
        if not restarts:
            restarts = [0]
        body += b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))
        return body
    out = b''
    handles = []
    for kvs in entries:
        b = block(kvs)
        handles.append((len(out), len(b), kvs[-1][0] if kvs else b''))
        out += b + b'\x00' + struct.pack('<I', bf.mask(bf.crc32c_bitwise(b + b'\x00')))
    meta = block([])
    mh = ck.put_varint(len(out)) + ck.put_varint(len(meta))
    out += meta + b'\x00' + struct.pack('<I', bf.mask(bf.crc32c_bitwise(meta + b'\x00')))
    idx = block([(k + b'~', ck.put_varint(o) + ck.put_varint(s)) for o, s, k in handles])   # separator > last key
    ih = ck.put_varint(len(out)) + ck.put_varint(len(idx))
    out += idx + b'\x00' + struct.pack('<I', bf.mask(bf.crc32c_bitwise(idx + b'\x00')))
    foot = mh + ih
    return out + foot + b'\x00' * (40 - len(foot)) + struct.pack('<II', 0x8b80fb57, 0xdb477524)


def test_reader_on_hand_assembled_bundle(tmp_path):
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.array([7, 8], np.int64)
    data = a.tobytes() + b.tobytes()
    ea = b'\x08\x01\x12\x08\x12\x02\x08\x02\x12\x02\x08\x03\x28\x18\x35' + struct.pack('<I', bf.mask(bf.crc32c_bitwise(a.tobytes())))
    eb = b'\x08\x09\x12\x04\x12\x02\x08\x02\x20\x18\x28\x10\x35' + struct.pack('<I', bf.mask(bf.crc32c_bitwise(b.tobytes())))
    prefix = str(tmp_path / 'hand.ckpt')
    open(prefix + '.index', 'wb').write(_hand_index([[(b'', ck.HEADER_PROTO), (b'a/kernel', ea)], [(b'b', eb)]]))
    open(prefix + '.data-00000-of-00001', 'wb').write(data)
    r = ck.BundleReader(prefix)
    assert r.keys() == ['a/kernel', 'b']
    assert np.array_equal(r.get('a/kernel'), a) and r.get('a/kernel').dtype == np.float32
    assert np.array_equal(r.get('b'), b) and r.get('b').dtype == np.int64


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / 'm.ckpt')
    ck.write_bundle(prefix, {'w': np.arange(1000, dtype=np.float32), 'v': np.ones(3, np.float32)})
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[40] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(raw)
    r = ck.BundleReader(prefix)
    with pytest.raises(ValueError, match='checksum'):
        r.get('w') if r.entries['w']['offset'] <= 40 < r.entries['w']['offset'] + 4000 else r.get('v')
    idx = bytearray(open(prefix + '.index', 'rb').read())
    idx[3] ^= 0x10
    open(prefix + '.index', 'wb').write(idx)
    with pytest.raises(ValueError, match='checksum'):
        ck.BundleReader(prefix)
    idx[-1] ^= 0xff
    open(prefix + '.index', 'wb').write(idx)
    with pytest.raises(ValueError, match='magic'):
        ck.BundleReader(prefix)


def test_saver_paths(tmp_path):
    """Training/Saver.py:38-62: newest Run_* directory, newest model_* file, epoch parsed from the name."""
    s = ck.Saver(str(tmp_path))
    with pytest.raises(ValueError, match='Cannot find ckpt file'):
        s._findfilename()
    s.set_save_path(comments='hello')
    run = s.save_dir
    assert os.path.basename(run).startswith('Run_') and open(os.path.join(run, 'Comments.txt')).read() == 'hello'
    with pytest.raises(ValueError, match='Cannot find ckpt file'):
        ck.Saver(str(tmp_path))._findfilename()
    for ep in (10, 20):
        ck.write_bundle(os.path.join(run, 'model_%s.ckpt' % str(ep).zfill(4)), {'x': np.zeros(2, np.float32)})
    d, f, e = ck.Saver(str(tmp_path))._findfilename()
    assert (d, f, e) == (run, os.path.join(run, 'model_0020.ckpt'), 20)
    d, f, e = ck.Saver(str(tmp_path))._findfilename(dir_names=os.path.basename(run), epoch=10)
    assert (f, e) == (os.path.join(run, 'model_0010.ckpt'), 10)


def test_reader_on_committed_fixture():
    """tests/golden/fmt_ckpt.*: written by the independent encoder of oracle/gen_format_fixtures.py"""
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    want = np.load(os.path.join(g, 'fmt_expected.npz'))
    r = ck.BundleReader(os.path.join(g, 'fmt_ckpt'))
    assert r.keys() == ['Train/beta1_power', 'classifier/conv1_1/V', 'classifier/conv1_1/V/Adam_optimizer']
    for k in r.keys():
        a = r.get(k)
        assert a.dtype == np.float32 and np.array_equal(a, want[k.replace('/', '.')]) and a.shape == want[k.replace('/', '.')].shape
    assert r.get('Train/beta1_power').shape == () and r.num_shards == 1
