"""TFRecord files of the reference's datasets, read (and written) without TensorFlow, feeding the device-side pipeline
(SURVEY.md §8f rank 4).

The reference stores CIFAR-10 / SVHN / MNIST as TFRecord files of tf.Example protos with the features
`image` (raw uint8 bytes), `label`, `height`, `width` (int64) -- parser() of Input_Pipeline/cifar10Dataset.py:42-62,
svhnDataset.py:41-65, mnistDataset.py:42-67 -- under `<DATA_DIR>/Tfrecord/<name>_train_<num_label>.tfrecords` (labelled),
`<name>_train_<train_size - num_label>.tfrecords` (unlabelled) and `<name>_test.tfrecords` (get_filenames, :23-33).
Here a file is decoded once on the host into one uint8 array + one label vector and handed to
`pipeline.DeviceDataset`, which keeps it in HBM; the per-example float conversion and one-hot encoding of `parser` run
on the device at batch-formation time (`tgan_gather_images_u8`, `tgan_gather_onehot`).

Formats, restated from their published descriptions (tensorflow/core/lib/io/record_writer.h, core/example/*.proto):

  record  = uint64 length | uint32 masked_crc32c(length bytes) | data[length] | uint32 masked_crc32c(data)
  Example = { features(1): Features { feature(1): map<string, Feature> } },  map entry = { key(1), value(2) }
  Feature = oneof { bytes_list(1): {value(1) bytes...}, float_list(2): {value(1) packed float}, int64_list(3): {value(1)
            packed varint} }      (unpacked encodings of the repeated scalars are accepted too)

No TF-written file is available in this image, so like the checkpoint format this is checked against the specification
(hand-assembled records, CRC known answers) and by round trips: "parity unpinned".
"""
import os
import struct

import numpy as np

from .checkpoint import _pb_fields, _signed64, crc32c, mask_crc, put_varint, unmask_crc

TRAIN_SIZE = {'cifar10': 50000, 'svhn': 73257, 'mnist': 60000}      # cifar10Dataset.py:21, svhnDataset.py:21, mnistDataset.py:20
CHANNELS = {'cifar10': 3, 'svhn': 3, 'mnist': 1}                      # reshape([height, width, C]) of the three parsers


# ---------------------------------------------------------------- record framing -------------------
def read_records(path, verify=True):
    """Yield the payload of every record of a TFRecord file."""
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) != 12:
                raise ValueError('tfrecord: truncated record header in %s' % path)
            n, lcrc = struct.unpack('<QI', head)
            if verify and unmask_crc(lcrc) != crc32c(head[:8]):
                raise ValueError('tfrecord: corrupted record length in %s' % path)
            body = f.read(n + 4)
            if len(body) != n + 4:
                raise ValueError('tfrecord: truncated record in %s' % path)
            data = body[:n]
            if verify and unmask_crc(struct.unpack('<I', body[n:])[0]) != crc32c(data):
                raise ValueError('tfrecord: corrupted record data in %s' % path)
            yield data


def write_records(path, payloads):
    with open(path, 'wb') as f:
        for data in payloads:
            head = struct.pack('<Q', len(data))
            f.write(head + struct.pack('<I', mask_crc(crc32c(head))) + data + struct.pack('<I', mask_crc(crc32c(data))))


# ---------------------------------------------------------------- tf.Example ------------------------
def parse_example(buf):
    """-> {feature name: list of bytes | list of int | list of float}"""
    out = {}
    for f, wt, feats in _pb_fields(buf):
        if f != 1 or wt != 2:
            continue
        for f2, wt2, entry in _pb_fields(feats):
            if f2 != 1 or wt2 != 2:
                continue
            key, val = None, []
            for f3, _, v3 in _pb_fields(entry):
                if f3 == 1:
                    key = v3.decode()
                elif f3 == 2:
                    for kind, _, lst in _pb_fields(v3):
                        for f5, wt5, v5 in _pb_fields(lst):
                            if f5 != 1:
                                continue
                            if kind == 1:
                                val.append(v5)
                            elif kind == 2:
                                if wt5 == 2:
                                    val.extend(struct.unpack('<%df' % (len(v5) // 4), v5))
                                else:
                                    val.append(struct.unpack('<f', struct.pack('<I', v5))[0])
                            elif kind == 3:
                                if wt5 == 2:
                                    p = 0
                                    while p < len(v5):
                                        x, p = _varint_at(v5, p)
                                        val.append(_signed64(x))
                                else:
                                    val.append(_signed64(v5))
            if key is not None:
                out[key] = val
    return out


def _varint_at(b, p):
    v = s = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7f) << s
        if c < 0x80:
            return v, p
        s += 7


def _ld(field, payload):
    return bytes([(field << 3) | 2]) + put_varint(len(payload)) + payload


def make_example(features):
    """features: {name: bytes | int | float | list of those} -> serialized tf.Example (packed repeated scalars)"""
    entries = b''
    for k in sorted(features):
        v = features[k]
        v = v if isinstance(v, (list, tuple)) else [v]
        if isinstance(v[0], (bytes, bytearray)):
            feat = _ld(1, b''.join(_ld(1, bytes(x)) for x in v))
        elif isinstance(v[0], float):
            feat = _ld(2, _ld(1, struct.pack('<%df' % len(v), *v)))
        else:
            feat = _ld(3, _ld(1, b''.join(put_varint(int(x) & 0xffffffffffffffff) for x in v)))
        entries += _ld(1, _ld(1, k.encode()) + _ld(2, feat))
    return _ld(1, entries)


# ---------------------------------------------------------------- datasets --------------------------
def load_image_records(path, channels, verify=True):
    """Decode a dataset file of the reference's layout -> (images uint8 [M, H, W, C], labels int64 [M]).
    Mirrors parser(): decode_raw(image, uint8) reshaped to [height, width, C]; label cast to int."""
    imgs, labels, shape = [], [], None
    for rec in read_records(path, verify):
        ex = parse_example(rec)
        for k in ('image', 'label', 'height', 'width'):
            if k not in ex or not ex[k]:
                raise ValueError('tfrecord: feature %r missing in %s' % (k, path))
        h, w = int(ex['height'][0]), int(ex['width'][0])
        raw = np.frombuffer(ex['image'][0], np.uint8)
        if raw.size != h * w * channels:
            raise ValueError('tfrecord: image of %d bytes is not %dx%dx%d in %s' % (raw.size, h, w, channels, path))
        if shape is None:
            shape = (h, w, channels)
        elif shape != (h, w, channels):
            raise ValueError('tfrecord: mixed image sizes in %s' % path)
        imgs.append(raw.reshape(shape))
        labels.append(int(ex['label'][0]))
    if not imgs:
        raise ValueError('tfrecord: %s holds no records' % path)
    return np.stack(imgs), np.asarray(labels, np.int64)


def write_image_records(path, images, labels):
    """The inverse (used by the tests and to export synthetic datasets in the reference's format)."""
    images = np.asarray(images, np.uint8)
    write_records(path, (make_example({'image': images[i].tobytes(), 'label': int(labels[i]), 'height': int(images.shape[1]),
                                       'width': int(images.shape[2])}) for i in range(images.shape[0])))


class RecordDataset(object):
    """cifar10Dataset / svhnDataset / mnistDataset of Input_Pipeline/*.py: same constructor and `get_filenames`;
    `load()` replaces input_from_tfrecord_filename + parser, `to_device()` hands the arrays to the device pipeline."""

    def __init__(self, data_dir, config, num_label=None, subset='train', use_augmentation=False, data_name=None):
        self.data_dir = os.path.join(data_dir, "Tfrecord")
        self.subset = subset
        self.use_augmentation = use_augmentation
        self.config = config
        self.num_label = num_label
        self.name = data_name or config.DATA_NAME
        if self.name not in TRAIN_SIZE:
            raise ValueError('unknown dataset %r' % (self.name,))
        self.train_size = TRAIN_SIZE[self.name]

    def get_filenames(self):
        assert self.subset in ['train', 'test'], 'Invalid data subset "%s"' % self.subset
        if self.subset == 'train':
            return [os.path.join(self.data_dir, '%s_%s_%s.tfrecords' % (self.name, self.subset, str(self.num_label).zfill(6))),
                    os.path.join(self.data_dir, '%s_%s_%s.tfrecords' % (self.name, self.subset,
                                                                        str(self.train_size - self.num_label).zfill(6)))]
        return [os.path.join(self.data_dir, '%s_%s.tfrecords' % (self.name, self.subset))]

    def load(self, verify=True):
        """-> [(images, labels)] per file: [labelled, unlabelled] for 'train', [test] for 'test'."""
        return [load_image_records(f, CHANNELS[self.name], verify) for f in self.get_filenames()]

    def to_device(self, verify=True):
        from . import pipeline
        return [pipeline.DeviceDataset(im, lb, self.config.NUM_CLASSES, self.name) for im, lb in self.load(verify)]
