"""TFRecord files of the reference's block-diagonal ("sparse") path -- SURVEY section 8(f) row 2.

Reader side: ``tf.data.TFRecordDataset`` + ``tf.io.parse_single_example`` with the feature_spec of
``task_sparse_gcn.py:93-101,153-166``, then ``dataset.batch`` (``:112``) -- here one native pass per
feature key over the records of a batch (``kgcn_tfrecord_scan`` / ``kgcn_tfexample_gather`` of
``include/kgcn_b200.h``, implemented in ``csrc/recordio.cu``), which yields directly the concatenated
arrays ``construct_batched_adjacency_and_feature_matrices`` (``kgcn/data_util.py:698-845``) takes.

Writer side: ``convert_to_example`` / ``save_tfrecords`` of ``kgcn/preprocessing/utils.py:178-226`` (host numpy
+ the protobuf wire format, written out by hand: the ``tensorflow.Example`` schema is five tiny messages).

No TensorFlow, no ``protobuf`` runtime on the product path; the tests cross-check both directions against the
``google.protobuf`` runtime with a dynamically built copy of the Example schema.
"""
import collections
import ctypes
import glob as _glob
import os
import struct

import numpy as np

from . import _lib
from .data_util import DataLoadError, construct_batched_adjacency_and_feature_matrices

FLOAT, INT64 = 1, 2
_KINDS = {"float32": FLOAT, "int64": INT64, np.float32: FLOAT, np.int64: INT64, FLOAT: FLOAT, INT64: INT64}
_DTYPES = {FLOAT: np.float32, INT64: np.int64}

FixedLenFeature = collections.namedtuple("FixedLenFeature", ["shape", "dtype"])
VarLenFeature = collections.namedtuple("VarLenFeature", ["dtype"])
# tf.io.VarLenFeature parses to a SparseTensor; the reference only ever reads `.values` of the batched one
# (example_model/sparse.py:49-57), i.e. the concatenation over the batch -- `row_splits` keeps the per-record extent.
VarLenBatch = collections.namedtuple("VarLenBatch", ["values", "row_splits"])


def sparse_feature_spec(task_num):
    """The feature_spec of ``task_sparse_gcn.py:153-166`` (identical at ``:282-295``)."""
    var, fixed = VarLenFeature, FixedLenFeature
    return {
        "adj_column": var(INT64), "adj_degrees": var(INT64), "adj_elem_len": fixed([1], INT64),
        "adj_row": var(INT64), "adj_values": var(FLOAT), "feature_column": var(INT64),
        "feature_elem_len": fixed([1], INT64), "feature_row": var(INT64), "feature_values": var(FLOAT),
        "label": fixed([task_num], INT64), "mask_label": fixed([task_num], INT64), "size": fixed([2], INT64),
    }


class TFRecordFile:
    """One ``.tfrecords`` file held in host memory with its record index.

    ``verify_crc=True`` checks the two CRC-32C words of every record like TensorFlow's reader
    (a mismatch raises ``DataLoadError``; TF: ``DataLossError``)."""

    def __init__(self, path, verify_crc=True):
        self.path = path
        self.buf = np.fromfile(path, dtype=np.uint8)
        n = ctypes.c_int64(0)
        base = self.buf.ctypes.data if self.buf.size else None
        rc = _lib.lib.kgcn_tfrecord_scan(base, self.buf.size, int(verify_crc), None, None, 0, ctypes.byref(n))
        if rc != 0:
            raise DataLoadError("%s: %s" % (path, _lib.last_error()))
        self.offsets = np.empty(n.value, np.int64)
        self.lengths = np.empty(n.value, np.int64)
        if n.value:
            _lib.check(_lib.lib.kgcn_tfrecord_scan(base, self.buf.size, 0, _lib.ptr(self.offsets), _lib.ptr(self.lengths),
                                                   n.value, ctypes.byref(n)))

    def __len__(self):
        return int(self.offsets.shape[0])

    def record(self, i):
        """The serialized Example of record ``i`` (bytes)."""
        o, n = int(self.offsets[i]), int(self.lengths[i])
        return self.buf[o:o + n].tobytes()

    def gather(self, key, kind, records=None):
        """Values stored under ``key`` in the given records, concatenated, plus per-record counts."""
        kind = _KINDS[kind]
        if records is None:
            off, length = self.offsets, self.lengths
        else:
            records = np.asarray(records, np.int64).reshape(-1)
            off, length = np.ascontiguousarray(self.offsets[records]), np.ascontiguousarray(self.lengths[records])
        n_rec = int(off.shape[0])
        counts = np.zeros(n_rec, np.int64)
        if n_rec == 0:
            return np.empty(0, _DTYPES[kind]), counts
        # an int64 varint takes >= 1 byte and a float 4, so the record bytes bound the value count
        capacity = int(length.sum()) if kind == INT64 else int(length.sum()) // 4
        values = np.empty(max(capacity, 1), _DTYPES[kind])
        total = ctypes.c_int64(0)
        rc = _lib.lib.kgcn_tfexample_gather(self.buf.ctypes.data, _lib.ptr(off), _lib.ptr(length), n_rec,
                                            key.encode("utf-8"), kind, _lib.ptr(values), capacity, _lib.ptr(counts),
                                            ctypes.byref(total))
        if rc != 0:
            raise DataLoadError("%s: %s" % (self.path, _lib.last_error()))
        return values[:total.value].copy(), counts


def parse_examples(tfr, feature_spec, records=None):
    """``dataset.map(parse_single_example).batch(len(records))`` for one file (``task_sparse_gcn.py:99,112,120``).

    ``FixedLenFeature(shape)`` -> array ``[n_records, *shape]``; a record that lacks the key or holds a different
    number of values raises (the reference gives no ``default_value``, so TensorFlow raises InvalidArgumentError).
    ``VarLenFeature`` -> ``VarLenBatch(values, row_splits)``; a missing key contributes no values."""
    out = {}
    for key, spec in feature_spec.items():
        values, counts = tfr.gather(key, spec.dtype, records)
        if isinstance(spec, FixedLenFeature):
            want = int(np.prod(spec.shape)) if len(spec.shape) else 1
            bad = np.nonzero(counts != want)[0]
            if bad.size:
                raise DataLoadError("%s: feature '%s' of record %d holds %d values, the spec needs %d"
                                    % (tfr.path, key, int(bad[0]), int(counts[bad[0]]), want))
            out[key] = values.reshape([counts.shape[0]] + list(spec.shape))
        else:
            out[key] = VarLenBatch(values, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
    return out


def block_diagonal_batch(parsed, max_degree=5, normalize=True, split_adj=False):
    """Parsed batch -> ``(channels, features)`` exactly as ``example_model/sparse.py:47-62`` wires
    ``construct_batched_adjacency_and_feature_matrices`` (``input_dim = size[0, 1]``)."""
    return construct_batched_adjacency_and_feature_matrices(
        parsed["size"][:, 0], parsed["adj_row"].values, parsed["adj_column"].values, parsed["adj_values"].values,
        parsed["adj_elem_len"][:, 0], parsed["adj_degrees"].values, parsed["feature_row"].values,
        parsed["feature_column"].values, parsed["feature_values"].values, parsed["feature_elem_len"][:, 0],
        int(parsed["size"][0, 1]), max_degree=max_degree, normalize=normalize, split_adj=split_adj)


class SparseDataset:
    """The records of a glob of ``.tfrecords`` files, batched in file order (``task_sparse_gcn.py:104-133``
    without the shuffle; ``info`` mirrors ``:135-142``: ``num_elements`` and ``input_dim = size[1]`` of the
    last record)."""

    def __init__(self, files, task_num, verify_crc=True):
        paths = sorted(_glob.glob(files)) if isinstance(files, str) else list(files)
        if not paths:
            raise DataLoadError("no tfrecords file matches %r" % (files,))
        self.files = [TFRecordFile(p, verify_crc) for p in paths]
        self.spec = sparse_feature_spec(task_num)
        self.index = [(f, r) for f, t in enumerate(self.files) for r in range(len(t))]
        num = len(self.index)
        input_dim = None
        if num:
            f, r = self.index[-1]
            input_dim = int(self.files[f].gather("size", INT64, [r])[0][1])
        self.info = {"num_elements": num, "input_dim": input_dim}

    def __len__(self):
        return len(self.index)

    def batches(self, batch_size, order=None, rank=0, world_size=1):
        """Yields parsed batches (dict as :func:`parse_examples`).  Records of one batch may span files.

        Data-parallel runs: every rank walks the same global batches of ``batch_size`` records and
        parses only its contiguous shard of each (molecules are independent, the block-diagonal matrix of a
        shard is the corresponding diagonal block of the global one) -- no data-path collective."""
        order = range(len(self.index)) if order is None else order
        order = list(order)
        for lo in range(0, len(order), batch_size):
            picked = order[lo:lo + batch_size]
            if world_size > 1:
                from .trainer import shard_range
                a, b = shard_range(len(picked), rank, world_size)
                picked = picked[a:b]
            chunk = [self.index[i] for i in picked]
            if not chunk:
                yield parse_examples(self.files[0], self.spec, [])
                continue
            parts = []
            start = 0
            while start < len(chunk):          # runs of consecutive records from the same file
                end = start
                while end < len(chunk) and chunk[end][0] == chunk[start][0]:
                    end += 1
                parts.append(parse_examples(self.files[chunk[start][0]], self.spec, [r for _, r in chunk[start:end]]))
                start = end
            yield parts[0] if len(parts) == 1 else _concat_parsed(parts)


def _concat_parsed(parts):
    out = {}
    for key, first in parts[0].items():
        if isinstance(first, VarLenBatch):
            values = np.concatenate([p[key].values for p in parts])
            counts = np.concatenate([np.diff(p[key].row_splits) for p in parts])
            out[key] = VarLenBatch(values, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64))
        else:
            out[key] = np.concatenate([p[key] for p in parts], 0)
    return out


# ---------------------------------------------------------------------------------------------------
# writer: kgcn/preprocessing/utils.py:178-226
def _varint(v):
    v &= (1 << 64) - 1                     # int64 two's complement, ten bytes when negative
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_field(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _int64_feature(values):
    packed = b"".join(_varint(int(v)) for v in values)
    lst = _len_field(1, packed) if len(values) else b""       # Int64List.value, packed
    return _len_field(3, lst)                                  # Feature.int64_list


def _float_feature(values):
    arr = np.asarray(values, dtype="<f4").reshape(-1)
    lst = _len_field(1, arr.tobytes()) if arr.size else b""   # FloatList.value, packed
    return _len_field(2, lst)                                  # Feature.float_list


def serialize_example(features):
    """``Example(features=Features(feature={key: Feature}))`` -> bytes.  ``features`` maps a key to
    ``("int64", values)`` or ``("float", values)``.  Map entries are written in sorted key order (what the
    protobuf runtimes do for deterministic output)."""
    body = b""
    for key in sorted(features):
        kind, values = features[key]
        feat = _int64_feature(values) if kind == "int64" else _float_feature(values)
        entry = _len_field(1, key.encode("utf-8")) + _len_field(2, feat)
        body += _len_field(1, entry)                           # Features.feature map entry
    return _len_field(1, body)                                 # Example.features


def convert_to_example(adj, feature, label_data=None, label_mask=None):
    """``kgcn/preprocessing/utils.py:178-214``: dense adjacency + dense feature matrix -> serialized Example.

    ``adj_degrees[e]`` is 0 for a diagonal entry, else the column sum at the entry's ROW index (``:185-191``;
    the reference indexes ``np.sum(adj, 0)`` with the row -- kept).  ``size`` is the feature matrix shape.
    (The reference's ``mask_label`` line ends in a stray comma that makes the value a tuple, ``:211``, so its
    own labelled path cannot serialise; the intended int64 list is written here.)"""
    adj = np.asarray(adj)
    adj_row, adj_col = np.nonzero(adj)
    degrees = np.sum(adj, 0)
    adj_degrees = [0 if r == c else int(degrees[r]) for r, c in zip(adj_row, adj_col)]
    feature = np.asarray(feature)
    feature_row, feature_col = np.nonzero(feature)
    feats = {
        "adj_row": ("int64", adj_row), "adj_column": ("int64", adj_col),
        "adj_values": ("float", adj[adj_row, adj_col]), "adj_elem_len": ("int64", [len(adj_row)]),
        "adj_degrees": ("int64", adj_degrees),
        "feature_row": ("int64", feature_row), "feature_column": ("int64", feature_col),
        "feature_values": ("float", feature[feature_row, feature_col]),
        "feature_elem_len": ("int64", [len(feature_row)]), "size": ("int64", list(feature.shape)),
    }
    if label_data is not None:
        feats["label"] = ("int64", np.nan_to_num(np.asarray(label_data, dtype=np.float64)).astype(int))
        feats["mask_label"] = ("int64", np.asarray(label_mask).astype(int))
    return serialize_example(feats)


def write_tfrecords(path, examples):
    """``TFRecordWriter(path).write(e) for e in examples`` (``utils.py:217-226``)."""
    crc = _lib.lib.kgcn_crc32c_masked
    with open(path, "wb") as fh:
        for e in examples:
            head = struct.pack("<Q", len(e))
            fh.write(head + struct.pack("<I", crc(head, 8)) + e + struct.pack("<I", crc(e, len(e))))


def save_tfrecords(save_dir, train_list, eval_list, test_list, idx):
    """``kgcn/preprocessing/utils.py:217-226``: the three split files ``{idx}_{train,test,eval}_.tfrecords``."""
    for name, lst in (("train", train_list), ("test", test_list), ("eval", eval_list)):
        write_tfrecords(os.path.join(save_dir, "%s_%s_.tfrecords" % (idx, name)), lst)
