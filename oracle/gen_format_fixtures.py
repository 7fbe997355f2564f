"""TEST INFRASTRUCTURE ONLY.  Writes the small file-format fixtures under tests/golden/ with a naive, independent encoder
(no prefix compression, a restart point at every entry, bit-wise CRC) -- NOT with the product writers in tgan/:

    tests/golden/fmt_ckpt.index, fmt_ckpt.data-00000-of-00001    a TF tensor-bundle checkpoint (3 tensors, 2 data blocks)
    tests/golden/fmt_records.tfrecords                            a TFRecord file of 3 tf.Example image records
    tests/golden/fmt_expected.npz                                 the arrays they hold

    python oracle/gen_format_fixtures.py        (re-run only if the fixtures are to change; they are committed)

TensorFlow is not available here, so these files pin the product reader against this independent restatement of the
published formats, not against TF itself ("parity unpinned").
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bundle_format import crc32c_bitwise, mask  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def vi(v):
    b = b''
    while True:
        c = v & 0x7f
        v >>= 7
        if v:
            b += bytes([c | 0x80])
        else:
            return b + bytes([c])


def ld(field, payload):
    return vi((field << 3) | 2) + vi(len(payload)) + payload


def block(kvs):
    body, restarts = b'', []
    for k, v in kvs:
        restarts.append(len(body))
        body += vi(0) + vi(len(k)) + vi(len(v)) + k + v
    restarts = restarts or [0]
    return body + b''.join(struct.pack('<I', r) for r in restarts) + struct.pack('<I', len(restarts))


def table(blocks):
    out, handles = b'', []

    def emit(b):
        nonlocal out
        h = vi(len(out)) + vi(len(b))
        out += b + b'\x00' + struct.pack('<I', mask(crc32c_bitwise(b + b'\x00')))
        return h
    for kvs in blocks:
        handles.append((kvs[-1][0], emit(block(kvs))))
    mh = emit(block([]))
    ih = emit(block(handles))
    foot = mh + ih
    return out + foot + b'\x00' * (40 - len(foot)) + struct.pack('<Q', 0xdb4775248b80fb57)


def entry(dtype_code, shape, off, arr):
    dims = b''.join(ld(2, b'\x08' + vi(s)) for s in shape)
    e = b'\x08' + vi(dtype_code) + ld(2, dims)
    if off:
        e += b'\x20' + vi(off)
    e += b'\x28' + vi(arr.nbytes) + b'\x35' + struct.pack('<I', mask(crc32c_bitwise(arr.tobytes())))
    return e


def main():
    rng = np.random.default_rng(20261018)
    t = {'classifier/conv1_1/V': rng.standard_normal((3, 3, 3, 8)).astype(np.float32),
         'classifier/conv1_1/V/Adam_optimizer': rng.standard_normal((3, 3, 3, 8)).astype(np.float32),
         'Train/beta1_power': np.float32(0.25).reshape(())}
    names = sorted(t, key=lambda s: s.encode())
    data, ents, off = b'', [], 0
    for n in names:
        a = t[n]
        ents.append((n.encode(), entry(1, a.shape, off, a)))
        data += a.tobytes()
        off += a.nbytes
    hdr = (b'', b'\x08\x01\x1a\x02\x08\x01')
    open(os.path.join(OUT, 'fmt_ckpt.index'), 'wb').write(table([[hdr, ents[0]], ents[1:]]))
    open(os.path.join(OUT, 'fmt_ckpt.data-00000-of-00001'), 'wb').write(data)

    imgs = rng.integers(0, 256, (3, 4, 4, 3), dtype=np.uint8)
    labels = np.array([3, 0, 9], np.int64)
    recs = b''
    for i in range(3):
        feats = b''
        for k, feat in sorted({'height': ld(3, ld(1, vi(4))), 'image': ld(1, ld(1, imgs[i].tobytes())),
                               'label': ld(3, ld(1, vi(int(labels[i])))), 'width': ld(3, ld(1, vi(4)))}.items()):
            feats += ld(1, ld(1, k.encode()) + ld(2, feat))
        ex = ld(1, feats)
        head = struct.pack('<Q', len(ex))
        recs += head + struct.pack('<I', mask(crc32c_bitwise(head))) + ex + struct.pack('<I', mask(crc32c_bitwise(ex)))
    open(os.path.join(OUT, 'fmt_records.tfrecords'), 'wb').write(recs)
    np.savez(os.path.join(OUT, 'fmt_expected.npz'), images=imgs, labels=labels, **{k.replace('/', '.'): v for k, v in t.items()})
    print('wrote fixtures to', OUT)


if __name__ == '__main__':
    main()
