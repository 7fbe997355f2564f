"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Independent, deliberately naive restatement of the pieces of TensorFlow's tensor-bundle checkpoint format that
tgan/checkpoint.py reads and writes (tf.train.Saver behind Training/Saver.py:29-36): a bit-at-a-time CRC-32C, the LevelDB
masked-CRC, and a linear parser of the `.index` string table.  Restated from the published format descriptions
(tensorflow/core/lib/hash/crc32c.h, core/lib/io/format.{h,cc}, table_format of LevelDB, core/protobuf/tensor_bundle.proto);
TensorFlow is not installable in this image, so no TF-written file pins it: "parity unpinned".
"""
import struct


def crc32c_bitwise(data, crc=0):
    """Castagnoli CRC, reflected polynomial 0x82F63B78, one bit at a time."""
    c = crc ^ 0xffffffff
    for b in bytes(data):
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
    return c ^ 0xffffffff


def mask(crc):
    """crc32c::Mask: rotate right by 15 bits and add a constant"""
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def varint(buf, pos):
    v = s = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7f) << s
        s += 7
        if b < 0x80:
            return v, pos


def parse_block(buf, off, size):
    """-> [(key, value)] of one table block; checks the block trailer (type 0, masked crc)."""
    body, typ = buf[off:off + size], buf[off + size]
    stored = struct.unpack_from('<I', buf, off + size + 1)[0]
    assert typ == 0, 'compressed block'
    assert stored == mask(crc32c_bitwise(body + bytes([typ]))), 'block crc'
    nrestart = struct.unpack_from('<I', body, len(body) - 4)[0]
    limit = len(body) - 4 - 4 * nrestart
    restarts = struct.unpack_from('<%dI' % nrestart, body, limit)
    pos, key, out, starts = 0, b'', [], []
    while pos < limit:
        starts.append(pos)
        shared, pos = varint(body, pos)
        rest, pos = varint(body, pos)
        vlen, pos = varint(body, pos)
        key = key[:shared] + body[pos:pos + rest]
        pos += rest
        out.append((key, body[pos:pos + vlen]))
        pos += vlen
    for r in restarts:                       # every restart point is an entry boundary with a full key
        assert r in starts or (r == 0 and not starts)
        if starts:
            assert varint(body, r)[0] == 0
    return out


def parse_index(buf):
    """-> [(key, value)] of a whole `.index` file"""
    assert struct.unpack('<Q', buf[-8:])[0] == 0xdb4775248b80fb57, 'magic'
    foot = buf[-48:]
    p = 0
    moff, p = varint(foot, p)
    msz, p = varint(foot, p)
    ioff, p = varint(foot, p)
    isz, p = varint(foot, p)
    assert parse_block(buf, moff, msz) == []
    items = []
    for sep, h in parse_index_block(buf, ioff, isz):
        o, q = varint(h, 0)
        s, q = varint(h, q)
        blk = parse_block(buf, o, s)
        assert blk and blk[-1][0] <= sep
        items += blk
    assert all(a[0] < b[0] for a, b in zip(items, items[1:])), 'keys not strictly increasing'
    return items


def parse_index_block(buf, off, size):
    return parse_block(buf, off, size)
