"""Minimal, independent reader of a TensorFlow V2 checkpoint ("tensor bundle": ``<prefix>.index`` +
``<prefix>.data-0000k-of-0000n``).

TEST INFRASTRUCTURE ONLY.  Nothing under ``kgcn_b200/`` imports this module, and this module imports nothing from
``kgcn_b200/``: it exists so that the real-weights golden (``oracle/make_ckpt_golden.py``) and the checkpoint tests read
the reference's shipped ``model/reaction/model.best.ckpt`` WITHOUT the product's own parser
(``kgcn_b200/tf_checkpoint.py``) -- the product reader is then checked against this one instead of against itself.

Format restated from TensorFlow's public sources (no TensorFlow here):
* ``.index`` is an immutable sorted string table in LevelDB's block format (tensorflow/core/lib/io/table*.cc, format.cc):
  a 48-byte footer {metaindex handle, index handle, padding, magic 0xdb4775248b80fb57}; every block is
  ``contents | 1 byte compression | 4 bytes masked CRC-32C``; block contents are prefix-compressed entries
  {varint shared, varint non_shared, varint value_len, key suffix, value} followed by the restart array;
* key ``""`` holds ``BundleHeaderProto`` {1: num_shards, 2: endianness, 3: version}; every other key is a tensor name whose
  value is ``BundleEntryProto`` {1: dtype, 2: TensorShapeProto{2: Dim{1: size}}, 3: shard_id, 4: offset, 5: size,
  6: fixed32 masked CRC-32C of the tensor bytes, 7: slices} (tensorflow/core/protobuf/tensor_bundle.proto);
* CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), masked as ``rotr(crc, 15) + 0xa282ead8`` (lib/hash/crc32c.h).
Only what those checkpoints contain is supported: no compression, little endian, no partitioned variables.
"""
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_}

_TABLE = []
for _n in range(256):
    _c = _n
    for _ in range(8):
        _c = (_c >> 1) ^ (0x82F63B78 if _c & 1 else 0)
    _TABLE.append(_c)


def crc32c(data):
    crc = 0xFFFFFFFF
    for b in bytes(data):
        crc = _TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block(buf, offset, size):
    body, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if trailer[0] != 0:
        raise ValueError("compressed table block (type %d)" % trailer[0])
    want = struct.unpack("<I", trailer[1:5])[0]
    if masked(crc32c(body + trailer[:1])) != want:
        raise ValueError("table block at %d: CRC mismatch" % offset)
    n_restarts = struct.unpack("<I", body[-4:])[0]
    end = len(body) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(body, pos)
        non_shared, pos = _varint(body, pos)
        vlen, pos = _varint(body, pos)
        key = key[:shared] + body[pos:pos + non_shared]
        pos += non_shared
        out.append((key, body[pos:pos + vlen]))
        pos += vlen
    return out


def _proto_fields(buf):
    """-> list of (field number, wire type, value); value is an int (varint / fixed) or bytes (length-delimited)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 1:
            v = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        elif wire == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            v = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        else:
            raise ValueError("wire type %d" % wire)
        out.append((field, wire, v))
    return out


def read_index(prefix):
    """-> (header dict, {tensor name: dict(dtype, shape, shard_id, offset, size, crc32c)}) in table (sorted) order."""
    buf = open(prefix + ".index", "rb").read()
    footer = buf[-48:]
    if struct.unpack("<Q", footer[40:])[0] != _MAGIC:
        raise ValueError("not a table file: bad magic")
    _, pos = _varint(footer, 0)          # metaindex handle (unused)
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    header, entries = None, {}
    for _, handle in _block(buf, ioff, isize):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block(buf, off, size):
            f = _proto_fields(value)
            if key == b"":
                header = {"num_shards": 0, "endianness": 0}
                for field, _, v in f:
                    if field == 1:
                        header["num_shards"] = v
                    elif field == 2:
                        header["endianness"] = v
                continue
            e = {"dtype": None, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
            for field, _, v in f:
                if field == 1:
                    e["dtype"] = _DTYPES[v]
                elif field == 2:
                    for f2, _, dim in _proto_fields(v):
                        if f2 == 2:
                            sz = [x for g, _, x in _proto_fields(dim) if g == 1]
                            e["shape"].append(sz[0] if sz else 0)
                elif field == 3:
                    e["shard_id"] = v
                elif field == 4:
                    e["offset"] = v
                elif field == 5:
                    e["size"] = v
                elif field == 6:
                    e["crc32c"] = v
                elif field == 7:
                    raise ValueError("partitioned variable %r" % key)
            entries[key.decode()] = e
    if header is None or header["endianness"] != 0:
        raise ValueError("missing header or big-endian bundle")
    return header, entries


def read_tensors(prefix, verify=True):
    """-> {name: numpy array} of every tensor in the bundle, each checked against its stored CRC-32C."""
    header, entries = read_index(prefix)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e["shard_id"] not in shards:
            shards[e["shard_id"]] = open("%s.data-%05d-of-%05d" % (prefix, e["shard_id"], header["num_shards"]), "rb").read()
        raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
        if verify and masked(crc32c(raw)) != e["crc32c"]:
            raise ValueError("tensor %r: CRC mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=e["dtype"]).reshape(e["shape"]).copy()
    return out
