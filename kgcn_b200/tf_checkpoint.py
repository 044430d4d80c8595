"""Reader for TensorFlow V2 checkpoints ("tensor bundles") -- SURVEY section 8(f) row 4.

The reference saves and restores weights with ``tf.train.Saver`` (``kgcn/core.py``: ``saver.save`` /
``saver.restore``) and ships one trained checkpoint, ``model/reaction/model.best.ckpt.{index,data-00000-of-00001}``.
TensorFlow is not installable here, so the on-disk format is restated from its published definition
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table = the LevelDB table format):

* ``<prefix>.index`` is an immutable sorted string table:
  ``[data block]* [metaindex block] [index block] footer``; footer = 48 bytes: two block handles
  (varint64 offset, varint64 size each), zero padding to 40 bytes, magic ``0xdb4775248b80fb57`` little-endian.
  Every block is followed by a 5-byte trailer: compression type (0 = none, 1 = snappy) and the masked CRC-32C of
  block + type byte.  A block holds prefix-compressed entries ``varint shared | varint non_shared | varint
  value_len | key suffix | value`` and ends with its restart array (``u32 * n``, then ``u32 n``).  The index block
  maps a separator key to the handle of a data block.
* key ``""`` -> ``BundleHeaderProto`` (num_shards = 1, endianness = 2, version = 3); every other key is a variable
  name -> ``BundleEntryProto``: dtype = 1, shape = 2 (``TensorShapeProto``: repeated dim = 2 {size = 1}), shard_id = 3,
  offset = 4, size = 5, crc32c = 6 (fixed32, masked CRC-32C of the tensor bytes), slices = 7.
* ``<prefix>.data-SSSSS-of-NNNNN`` holds the raw little-endian tensor bytes at ``offset``.

Both checksums (block trailers and per-tensor) are verified with the library's ``kgcn_crc32c``.
"""
import struct

import numpy as np

from . import _lib
from .data_util import DataLoadError

TABLE_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


class CheckpointError(DataLoadError):
    pass


def _masked_crc(data):
    return _lib.lib.kgcn_crc32c_masked(data, len(data))


def _varint(buf, pos):
    value, shift = 0, 0
    while True:
        if pos >= len(buf) or shift > 63:
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one protobuf message; length-delimited values as bytes."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            value, pos = _varint(buf, pos)
        elif wire == 1:
            value, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        elif wire == 2:
            n, pos = _varint(buf, pos)
            if pos + n > len(buf):
                raise CheckpointError("truncated protobuf field")
            value, pos = bytes(buf[pos:pos + n]), pos + n
        elif wire == 5:
            value, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        else:
            raise CheckpointError("unsupported protobuf wire type %d" % wire)
        yield field, wire, value


def _read_block(table, offset, size, what):
    if offset + size + 5 > len(table):
        raise CheckpointError("%s block runs past the end of the index file" % what)
    block, trailer = table[offset:offset + size], table[offset + size:offset + size + 5]
    if _masked_crc(table[offset:offset + size + 1]) != struct.unpack("<I", trailer[1:])[0]:
        raise CheckpointError("%s block at byte %d fails its CRC-32C" % (what, offset))
    if trailer[0] != 0:
        raise CheckpointError("%s block is compressed (type %d); tensor bundles are written uncompressed" % (what, trailer[0]))
    return block


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("table block too short")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise CheckpointError("table block restart array is corrupt")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        value_len, pos = _varint(block, pos)
        if shared > len(key) or pos + non_shared + value_len > end:
            raise CheckpointError("table block entry is corrupt")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + value_len])
        pos += value_len


def _handle(buf, pos=0):
    offset, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return offset, size, pos


def read_table(path):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    table = open(path, "rb").read()
    if len(table) < 48 or struct.unpack("<Q", table[-8:])[0] != TABLE_MAGIC:
        raise CheckpointError("%s is not a TensorFlow checkpoint index (bad table magic)" % path)
    footer = table[-48:]
    _, _, pos = _handle(footer)                      # metaindex handle (unused: no filter policy)
    idx_off, idx_size, _ = _handle(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(table, idx_off, idx_size, "index")):
        off, size, _ = _handle(handle)
        out.extend(_block_entries(_read_block(table, off, size, "data")))
    return out


class BundleEntry:
    __slots__ = ("name", "dtype", "shape", "shard_id", "offset", "size", "crc32c", "sliced")

    def __repr__(self):
        return "BundleEntry(%r, %s, %r)" % (self.name, np.dtype(self.dtype).name if self.dtype else "?", self.shape)


def _parse_entry(name, blob):
    e = BundleEntry()
    e.name, e.dtype, e.shape, e.shard_id, e.offset, e.size, e.crc32c, e.sliced = name, None, [], 0, 0, 0, None, False
    for field, _, value in _fields(blob):
        if field == 1:
            e.dtype = DTYPES.get(value, value)
        elif field == 2:
            for f2, _, dim in _fields(value):
                if f2 == 2:
                    size = 0
                    for f3, _, v in _fields(dim):
                        if f3 == 1:
                            size = v - (1 << 64) if v >> 63 else v
                    e.shape.append(size)
        elif field == 3:
            e.shard_id = value
        elif field == 4:
            e.offset = value
        elif field == 5:
            e.size = value
        elif field == 6:
            e.crc32c = value
        elif field == 7:
            e.sliced = True
    return e


class CheckpointReader:
    """``tf.train.load_checkpoint(prefix)`` / ``tf.train.NewCheckpointReader``: ``get_variable_to_shape_map()``,
    ``get_variable_to_dtype_map()``, ``has_tensor(name)``, ``get_tensor(name)``."""

    def __init__(self, prefix, verify_crc=True):
        self.prefix, self.verify_crc = prefix, verify_crc
        pairs = read_table(prefix + ".index")
        if not pairs or pairs[0][0] != b"":
            raise CheckpointError("%s.index has no bundle header" % prefix)
        self.num_shards, self.endianness, self.version = 1, 0, None
        for field, _, value in _fields(pairs[0][1]):
            if field == 1:
                self.num_shards = value
            elif field == 2:
                self.endianness = value
            elif field == 3:
                self.version = dict((f, v) for f, _, v in _fields(value))
        if self.endianness != 0:
            raise CheckpointError("big-endian tensor bundles are not supported")
        self.entries = {}
        for key, blob in pairs[1:]:
            name = key.decode("utf-8")
            self.entries[name] = _parse_entry(name, blob)
        self._shards = {}

    def _shard(self, shard_id):
        if shard_id not in self._shards:
            path = "%s.data-%05d-of-%05d" % (self.prefix, shard_id, self.num_shards)
            self._shards[shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
        return self._shards[shard_id]

    def get_variable_to_shape_map(self):
        return {n: list(e.shape) for n, e in self.entries.items()}

    def get_variable_to_dtype_map(self):
        return {n: e.dtype for n, e in self.entries.items()}

    def has_tensor(self, name):
        return name in self.entries

    def get_tensor(self, name):
        if name not in self.entries:
            raise KeyError("%s: no variable named %r in the checkpoint" % (self.prefix, name))   # TF: NotFoundError
        e = self.entries[name]
        if e.sliced:
            raise CheckpointError("%s: partitioned variables (slices) are not supported" % name)
        if not isinstance(e.dtype, type):
            raise CheckpointError("%s: unsupported dtype enum %r" % (name, e.dtype))
        count = int(np.prod(e.shape)) if e.shape else 1
        if count * np.dtype(e.dtype).itemsize != e.size:
            raise CheckpointError("%s: %d bytes stored for shape %r of %s" % (name, e.size, e.shape, np.dtype(e.dtype).name))
        shard = self._shard(e.shard_id)
        if e.offset + e.size > shard.shape[0]:
            raise CheckpointError("%s: data shard is truncated" % name)
        raw = np.array(shard[e.offset:e.offset + e.size])
        if self.verify_crc and e.crc32c is not None and e.size and \
                _lib.lib.kgcn_crc32c_masked(raw.ctypes.data, e.size) != e.crc32c:
            raise CheckpointError("%s: tensor bytes fail their CRC-32C" % name)
        return raw.view(e.dtype).reshape(e.shape)

    def tensors(self, skip_slots=True):
        """name -> array; ``skip_slots`` drops optimizer state (``.../Adam``, ``.../Adam_1``, ``beta*_power``)."""
        out = {}
        for name in self.entries:
            leaf = name.rsplit("/", 1)[-1]
            if skip_slots and (leaf in ("Adam", "Adam_1") or name in ("beta1_power", "beta2_power")):
                continue
            out[name] = self.get_tensor(name)
        return out


def load_checkpoint(prefix, verify_crc=True):
    return CheckpointReader(prefix, verify_crc)


# ---------------------------------------------------------------------------------------------------
# writer: tf.train.Saver.save for a single shard
_DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items()}
BLOCK_SIZE = 262144          # tensorflow/core/lib/io/table_options.h
RESTART_INTERVAL = 16


def _enc_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _enc_field(field, wire, payload):
    if wire == 0:
        return _enc_varint((field << 3) | 0) + _enc_varint(payload)
    if wire == 5:
        return _enc_varint((field << 3) | 5) + struct.pack("<I", payload)
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval, self.buf, self.restarts, self.count, self.last_key = restart_interval, bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            limit = min(len(key), len(self.last_key))
            while shared < limit and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _enc_varint(shared) + _enc_varint(len(key) - shared) + _enc_varint(len(value)) + key[shared:] + value
        self.last_key, self.count = key, self.count + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    """LevelDB BytewiseComparator::FindShortestSeparator: a short key k with start <= k < limit."""
    n = min(len(start), len(limit))
    i = 0
    while i < n and start[i] == limit[i]:
        i += 1
    if i < n and start[i] < 0xFF and start[i] + 1 < limit[i]:
        return start[:i] + bytes([start[i] + 1])
    return start


def _short_successor(key):
    """LevelDB BytewiseComparator::FindShortSuccessor: a short key >= key."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def write_table(path, pairs):
    """Writes sorted (key, value) byte pairs as a LevelDB-format table (no compression, no filter block)."""
    out = bytearray()

    def emit(block):
        handle = _enc_varint(len(out)) + _enc_varint(len(block))
        out.extend(block + b"\x00" + struct.pack("<I", _masked_crc(block + b"\x00")))
        return handle

    index, data, last, pending = _BlockBuilder(1), _BlockBuilder(RESTART_INTERVAL), None, None
    for key, value in pairs:
        if last is not None and key <= last:
            raise CheckpointError("table keys must be strictly increasing")
        if pending is not None:                      # index key of a finished block: a short key in [last, key)
            index.add(_shortest_separator(last, key), pending)
            pending = None
        data.add(key, value)
        last = key
        if data.size() >= BLOCK_SIZE:
            pending, data = emit(data.finish()), _BlockBuilder(RESTART_INTERVAL)
    if data.buf or last is None:
        pending = emit(data.finish())
    if pending is not None:
        index.add(_short_successor(last or b""), pending)
    meta_handle = emit(_BlockBuilder(1).finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as fh:
        fh.write(out)


def save_checkpoint(prefix, tensors):
    """``{name: array}`` -> ``prefix.index`` + ``prefix.data-00000-of-00001`` in the layout described at the top of
    this file (one shard, little-endian, tensors in name order, per-tensor and per-block CRC-32C)."""
    header = _enc_field(1, 0, 1) + _enc_field(3, 2, _enc_field(1, 0, 1))          # num_shards = 1, version.producer = 1
    pairs, offset = [(b"", header)], 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(tensors, key=lambda n: n.encode("utf-8")):
            if not name:
                raise CheckpointError("a variable needs a non-empty name")
            arr = np.asarray(tensors[name])           # (ascontiguousarray would turn a scalar into shape [1])
            if arr.dtype not in _DTYPE_ENUM:
                raise CheckpointError("%s: dtype %s cannot be stored" % (name, arr.dtype))
            raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes(order="C")
            shape = b"".join(_enc_field(2, 2, _enc_field(1, 0, int(d))) for d in arr.shape)
            entry = _enc_field(1, 0, _DTYPE_ENUM[arr.dtype]) + _enc_field(2, 2, shape)
            if offset:
                entry += _enc_field(4, 0, offset)
            entry += _enc_field(5, 0, len(raw)) + _enc_field(6, 5, _masked_crc(raw) if raw else _lib.lib.kgcn_crc32c_masked(None, 0))
            pairs.append((name.encode("utf-8"), entry))
            fh.write(raw)
            offset += len(raw)
    write_table(prefix + ".index", pairs)
