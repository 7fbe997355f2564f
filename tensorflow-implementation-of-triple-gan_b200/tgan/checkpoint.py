"""Checkpoint import / export in TensorFlow's tensor-bundle ("V2 checkpoint") format, without TensorFlow
(SURVEY.md §8f rank 3).  Mirrors the reference's Training/Saver.py (class Saver :13-66), whose `save` / `restore`
are `tf.train.Saver().save(sess, name)` / `.restore(sess, name)` over ALL global variables: the model variables under
their variable-scope names, the non-trainable statistics (pop_mean, moving_mean / moving_variance), the
Adam slots `<var>/Adam_optimizer`, `<var>/Adam_optimizer_1` (the optimisers are built with name='Adam_optimizer',
train_base.py:91-97), each optimiser's `beta1_power` / `beta2_power` accumulators and the classifier's
`<var>/ExponentialMovingAverage` shadows (Train_goodGAN.py:101-103).

File format restated from TensorFlow's published sources (tensorflow/core/util/tensor_bundle, core/lib/io/table*,
core/lib/hash/crc32c.h); TensorFlow is not installable here, so the format is checked against its specification and by
round trips, not against files written by TF ("parity unpinned", like the oracle):

  <prefix>.data-00000-of-00001   raw little-endian tensor bytes, back to back, in key order
  <prefix>.index                 an immutable sorted string table (LevelDB table format):
        data blocks | metaindex block | index block | 48-byte footer
      block      = entries (varint32 shared, varint32 non_shared, varint32 value_len, key suffix, value),
                   restart offsets (fixed32 each), fixed32 restart count; then on file a 1-byte compression type
                   (0 = none) and a fixed32 masked CRC-32C of block + type
      footer     = metaindex handle, index handle (varint64 offset, varint64 size), zero padding to 40 bytes,
                   magic 0xdb4775248b80fb57 (fixed64)
      key ""     -> BundleHeaderProto {num_shards = 1, endianness = LITTLE, version {producer = 1}}
      key <name> -> BundleEntryProto {dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32,
                   masked CRC-32C of the tensor bytes)}
  masked crc = rotate_right(crc, 15) + 0xa282ead8 (mod 2^32)

The CRC itself is libtgan's host routine `tgan_crc32c` (csrc/pipeline.cu).
"""
import ctypes
import os
import struct
from datetime import datetime, timedelta, timezone

import numpy as np

from . import _lib

MAGIC = 0xdb4775248b80fb57
_DT = {1: np.dtype('<f4'), 2: np.dtype('<f8'), 3: np.dtype('<i4'), 4: np.dtype('u1'), 5: np.dtype('<i2'),
       6: np.dtype('i1'), 9: np.dtype('<i8'), 10: np.dtype('?')}
_DT_CODE = {v: k for k, v in _DT.items()}
BLOCK_SIZE = 4096          # TF's table builder default is 256 KB; readers accept any block size
RESTART_INTERVAL = 16


# ---------------------------------------------------------------- primitives ----------------------
def crc32c(data, crc=0):
    """CRC-32C of a bytes-like / contiguous ndarray (libtgan host routine)."""
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return int(_lib.load().tgan_crc32c(crc, a.ctypes.data, a.nbytes))
    b = bytes(data)
    return int(_lib.load().tgan_crc32c(crc, ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b)))


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xffffffff) + 0xa282ead8) & 0xffffffff


def unmask_crc(m):
    rot = (m - 0xa282ead8) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


def put_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(buf, pos):
    v, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError('checkpoint: truncated varint')
        b = buf[pos]
        pos += 1
        v |= (b & 0x7f) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise ValueError('checkpoint: varint too long')


def _pb_fields(buf):
    """minimal protobuf wire-format walk -> list of (field, wire_type, value)"""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        else:
            raise ValueError('checkpoint: unsupported protobuf wire type %d' % wt)
        out.append((f, wt, v))
    return out


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def encode_entry(dtype, shape, offset, size, crc_masked, shard_id=0):
    """BundleEntryProto (proto3: zero-valued scalars are omitted, the shape message is always present)"""
    dims = b''.join(b'\x12' + put_varint(len(d)) + d for d in (b'\x08' + put_varint(int(s)) for s in shape))
    out = b'\x08' + put_varint(_DT_CODE[np.dtype(dtype)])
    out += b'\x12' + put_varint(len(dims)) + dims
    if shard_id:
        out += b'\x18' + put_varint(shard_id)
    if offset:
        out += b'\x20' + put_varint(offset)
    if size:
        out += b'\x28' + put_varint(size)
    out += b'\x35' + struct.pack('<I', crc_masked)
    return out


def decode_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            dims = []
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    sz = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            sz = _signed64(v3)
                    dims.append(sz)
                elif f2 == 3 and v2:
                    raise ValueError('checkpoint: tensor of unknown rank')
            e['shape'] = tuple(dims)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = v
        elif f == 7:
            e['sliced'] = True
    return e


HEADER_PROTO = b'\x08\x01\x1a\x02\x08\x01'      # num_shards = 1, (endianness LITTLE = default), version {producer: 1}


# ---------------------------------------------------------------- table (.index) ------------------
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.ri = restart_interval
        self.reset()

    def reset(self):
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last = b''

    def add(self, key, value):
        shared = 0
        if self.count < self.ri:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def write_table(path, items, block_size=BLOCK_SIZE):
    """items: list of (key bytes, value bytes), strictly increasing keys."""
    out = bytearray()

    def emit(contents):
        off = len(out)
        out.extend(contents)
        out.append(0)                                                    # kNoCompression
        out.extend(struct.pack('<I', mask_crc(crc32c(contents + b'\x00'))))
        return put_varint(off) + put_varint(len(contents))

    data, index = _BlockBuilder(RESTART_INTERVAL), _BlockBuilder(1)
    prev = None
    for k, v in items:
        if prev is not None and not k > prev:
            raise ValueError('checkpoint: keys must be strictly increasing')
        data.add(k, v)
        prev = k
        if data.size() >= block_size:
            index.add(k, emit(data.finish()))       # separator = the block's last key (any key in [last, next) is legal)
            data.reset()
    if not data.empty():
        index.add(prev, emit(data.finish()))
    meta_h = emit(_BlockBuilder(RESTART_INTERVAL).finish())
    index_h = emit(index.finish())
    footer = meta_h + index_h
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', MAGIC)
    out.extend(footer)
    with open(path, 'wb') as f:
        f.write(out)


def _read_block(buf, handle_off, handle_size, verify=True):
    contents = buf[handle_off:handle_off + handle_size]
    trailer = buf[handle_off + handle_size:handle_off + handle_size + 5]
    if len(contents) != handle_size or len(trailer) != 5:
        raise ValueError('checkpoint: index block out of range')
    if verify and unmask_crc(struct.unpack('<I', trailer[1:])[0]) != crc32c(bytes(contents) + bytes(trailer[:1])):
        raise ValueError('checkpoint: index block checksum mismatch')
    if trailer[0] != 0:
        raise NotImplementedError('checkpoint: compressed index blocks (type %d) are not supported' % trailer[0])
    nres = struct.unpack_from('<I', contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * nres
    pos, key, out = 0, b'', []
    while pos < end:
        shared, pos = get_varint(contents, pos)
        non_shared, pos = get_varint(contents, pos)
        vlen, pos = get_varint(contents, pos)
        key = key[:shared] + bytes(contents[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(contents[pos:pos + vlen])))
        pos += vlen
    return out


def read_table(path, verify=True):
    buf = open(path, 'rb').read()
    if len(buf) < 48 or struct.unpack('<Q', buf[-8:])[0] != MAGIC:
        raise ValueError('checkpoint: %s is not a tensor-bundle index (bad magic)' % path)
    foot = buf[-48:]
    _, p = get_varint(foot, 0)
    _, p = get_varint(foot, p)
    ioff, p = get_varint(foot, p)
    isz, p = get_varint(foot, p)
    items = []
    for _, h in _read_block(buf, ioff, isz, verify):
        off, q = get_varint(h, 0)
        sz, q = get_varint(h, q)
        items.extend(_read_block(buf, off, sz, verify))
    return items


# ---------------------------------------------------------------- bundle ---------------------------
def write_bundle(prefix, tensors):
    """tensors: {name: ndarray}.  Writes <prefix>.index and <prefix>.data-00000-of-00001."""
    items = [(b'', HEADER_PROTO)]
    off = 0
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in sorted(tensors, key=lambda s: s.encode()):
            a = np.asarray(tensors[name])
            shape = a.shape                              # (np.ascontiguousarray would turn a scalar into [1])
            if a.dtype.byteorder == '>':
                a = a.astype(a.dtype.newbyteorder('<'))
            a = np.ascontiguousarray(a)
            f.write(a.tobytes())
            items.append((name.encode(), encode_entry(a.dtype, shape, off, a.nbytes, mask_crc(crc32c(a)))))
            off += a.nbytes
    write_table(prefix + '.index', items)


class BundleReader:
    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        items = read_table(prefix + '.index', verify)
        if not items or items[0][0] != b'':
            raise ValueError('checkpoint: missing bundle header')
        hdr = {f: v for f, _, v in _pb_fields(items[0][1])}
        self.num_shards = hdr.get(1, 0)
        if hdr.get(2, 0) != 0:
            raise NotImplementedError('checkpoint: big-endian bundles are not supported')
        self.entries = {k.decode(): decode_entry(v) for k, v in items[1:]}
        self._files = {}

    def keys(self):
        return list(self.entries)

    def has(self, name):
        return name in self.entries

    def shape(self, name):
        return self.entries[name]['shape']

    def get(self, name):
        e = self.entries[name]
        if e['sliced']:
            raise NotImplementedError('checkpoint: partitioned variable %s' % name)
        if e['dtype'] not in _DT:
            raise NotImplementedError('checkpoint: dtype %d of %s' % (e['dtype'], name))
        sid = e['shard_id']
        if sid not in self._files:
            self._files[sid] = np.memmap('%s.data-%05d-of-%05d' % (self.prefix, sid, self.num_shards), dtype=np.uint8, mode='r')
        raw = np.asarray(self._files[sid][e['offset']:e['offset'] + e['size']])
        dt = _DT[e['dtype']]
        n = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if raw.nbytes != e['size'] or n * dt.itemsize != e['size']:
            raise ValueError('checkpoint: %s: size mismatch' % name)
        if self.verify and e['crc32c'] is not None and unmask_crc(e['crc32c']) != crc32c(raw):
            raise ValueError('checkpoint: %s: tensor checksum mismatch' % name)
        return raw.view(dt).reshape(e['shape']).copy()


# ---------------------------------------------------------------- trainer state <-> TF names ------
ADAM_NAME = 'Adam_optimizer'       # train_base.py:91
EMA_NAME = 'ExponentialMovingAverage'
# the three optimisers are created in this order (Train_goodGAN.py:85-94); TF uniquifies their accumulators' names
_POWER_SUFFIX = {'discriminator': '', 'good_generator': '_1', 'classifier': '_2'}
_POWER_SCOPES = ('Train/', '')     # built under tf.name_scope('Train') (:78); bare names accepted on import


def _opt_of(trainer, grp):
    return {'discriminator': trainer.d_optimizer, 'good_generator': trainer.g_optimizer, 'classifier': trainer.c_optimizer}[grp]


def state_dict(trainer):
    """Everything `tf.train.Saver()` would save for the reference's graph, as {TF variable name: ndarray}."""
    st = trainer.store
    if getattr(trainer, 'fused_dp', None) is not None:      # Adam slots are sharded across the ranks: make them whole
        trainer.fused_dp.gather_slots(st)
    out = {n: st.vars[n].data.detach().cpu().numpy().copy() for n in st.order}
    for grp, fb in st.flat.items():
        if 'm' not in fb:
            continue
        m, v = fb['m'].detach().cpu().numpy(), fb['v'].detach().cpu().numpy()
        for p, o in zip(fb['params'], fb['offsets']):
            out['%s/%s' % (p.name, ADAM_NAME)] = m[o:o + p.size].reshape(p.shape).copy()
            out['%s/%s_1' % (p.name, ADAM_NAME)] = v[o:o + p.size].reshape(p.shape).copy()
        pw = _opt_of(trainer, grp)._state().detach().cpu().numpy()
        out['Train/beta1_power' + _POWER_SUFFIX[grp]] = np.float32(pw[1]).reshape(())
        out['Train/beta2_power' + _POWER_SUFFIX[grp]] = np.float32(pw[2]).reshape(())
    fb = st.flat['classifier']
    sh = trainer.ema.shadow.detach().cpu().numpy()
    for p, o in zip(fb['params'], fb['offsets']):
        out['%s/%s' % (p.name, EMA_NAME)] = sh[o:o + p.size].reshape(p.shape).copy()
    return out


def load_state_dict(trainer, get, has, strict=True):
    """Inverse of state_dict.  Model variables must all be present (`strict`), optimiser / EMA state is optional
    (the reference runs `initialize_uninitialized_vars` after a restore, Train_goodGAN.py:147)."""
    import torch
    st = trainer.store
    P = {}
    for n in st.order:
        if has(n):
            a = np.asarray(get(n), np.float32)
            if tuple(a.shape) != st.vars[n].shape:
                raise ValueError('checkpoint: shape of %s is %s, the model expects %s' % (n, a.shape, st.vars[n].shape))
            P[n] = a
        elif strict:
            raise KeyError('checkpoint: variable %s not found' % n)
    st.load_numpy(P)
    for grp, fb in st.flat.items():
        if 'm' not in fb:
            continue
        for slot, suffix in (('m', ADAM_NAME), ('v', ADAM_NAME + '_1')):
            for p, o in zip(fb['params'], fb['offsets']):
                k = '%s/%s' % (p.name, suffix)
                if has(k):
                    fb[slot][o:o + p.size].copy_(torch.from_numpy(np.asarray(get(k), np.float32).reshape(-1)))
        opt = _opt_of(trainer, grp)
        for i, nm in ((1, 'beta1_power'), (2, 'beta2_power')):
            for sc in _POWER_SCOPES:
                k = sc + nm + _POWER_SUFFIX[grp]
                if has(k):
                    opt._state()[i:i + 1].copy_(torch.from_numpy(np.asarray(get(k), np.float32).reshape(1)))
                    break
    fb = st.flat['classifier']
    for p, o in zip(fb['params'], fb['offsets']):
        k = '%s/%s' % (p.name, EMA_NAME)
        src = get(k) if has(k) else P.get(p.name)      # an EMA slot starts at the variable's value
        if src is not None:
            trainer.ema.shadow[o:o + p.size].copy_(torch.from_numpy(np.asarray(src, np.float32).reshape(-1)))
    st.bump()


# ---------------------------------------------------------------- Training/Saver.py mirror --------
def _eastern_now():
    try:
        from zoneinfo import ZoneInfo
        return datetime.now(ZoneInfo('US/Eastern'))
    except Exception:                                     # no tz database in the image: EST
        return datetime.now(timezone(timedelta(hours=-5)))


class Saver(object):
    """Training/Saver.py:13-66 with the trainer in the place of the tf.Session."""

    def __init__(self, save_dir, **kwargs):
        self.save_dir = save_dir

    def set_save_path(self, **kwargs):
        self.save_dir = os.path.join(self.save_dir, 'Run_' + _eastern_now().strftime("%Y-%m-%d_%H_%M_%S"))
        if not os.path.exists(self.save_dir):
            os.makedirs(self.save_dir)
        if 'comments' in kwargs:
            self.comments = kwargs.get('comments')
            self._write_comments()

    def save(self, sess, save_name):
        save_name = os.path.join(self.save_dir, save_name)
        write_bundle(save_name, state_dict(sess))
        base = os.path.basename(save_name)
        with open(os.path.join(self.save_dir, 'checkpoint'), 'w') as f:      # tf.train.Saver's CheckpointState file
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))

    def restore(self, sess, dir_names=None, epoch=None):
        self.save_dir, filename, start_epoch = self._findfilename(dir_names, epoch)
        r = BundleReader(filename)
        load_state_dict(sess, r.get, r.has)
        return start_epoch

    def _findfilename(self, dir_names=None, epoch=None):
        if dir_names is None:
            dir_names = next(os.walk(self.save_dir))[1]
            dir_names = sorted(f for f in dir_names if f.startswith('Run'))
            if not dir_names:
                raise ValueError('Cannot find ckpt file!')
            save_dir = os.path.join(self.save_dir, dir_names[-1])
        else:
            save_dir = os.path.join(self.save_dir, dir_names)
        checkpoints = sorted(f for f in next(os.walk(save_dir))[2] if f.startswith("model"))
        if not checkpoints:
            raise ValueError('Cannot find ckpt file!')
        name, suffix, _ = checkpoints[-1].split('.')
        if epoch is None:
            start_epoch = name.split('_')[1]
            checkpoints = os.path.join(save_dir, name + '.' + suffix)
        else:
            start_epoch = epoch
            name_prefix = name.split('_')[0]
            checkpoints = os.path.join(save_dir, name_prefix + '_' + str(epoch).zfill(4) + '.' + suffix)
        return save_dir, checkpoints, int(start_epoch)

    def _write_comments(self):
        with open(os.path.join(self.save_dir, 'Comments.txt'), 'w') as txt_file:
            txt_file.write(self.comments)
