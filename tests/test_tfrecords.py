"""SURVEY section 8(f) row 2, file half: TFRecord framing, CRC-32C, tensorflow.Example parsing and the batch
assembly of task_sparse_gcn.py:93-166 / example_model/sparse.py:47-62, through the C ABI's host entry points.

Pins: CRC-32C against the RFC 3720 (iSCSI) test vectors -- the same vectors TensorFlow's own crc32c_test uses;
the Example wire format in BOTH directions against the google.protobuf runtime on a dynamically built copy of
the public tensorflow.Example schema; the batch against the worked example of the reference's own docstring
(kgcn/data_util.py:703-731)."""
import os
import struct

import numpy as np
import pytest

from kgcn_b200 import _lib, data_util, tfrecords
from kgcn_b200.data_util import DataLoadError


def crc32c_bitwise(data):
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
    return c ^ 0xFFFFFFFF


def test_crc32c_known_answers():
    crc = lambda b: _lib.lib.kgcn_crc32c(b, len(b))
    assert crc(b"123456789") == 0xE3069283
    assert crc(bytes(32)) == 0x8A9136AA                      # RFC 3720 B.4
    assert crc(b"\xff" * 32) == 0x62A8AB43
    assert crc(bytes(range(32))) == 0x46DD794E
    assert crc(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert crc(b"") == 0
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 64, 65, 1000):                  # head / slicing-by-8 body / tail, odd alignments
        buf = rng.integers(0, 256, size=n + 3, dtype=np.uint8)
        for shift in (0, 1, 3):
            view = buf[shift:shift + n]
            assert _lib.lib.kgcn_crc32c(view.ctypes.data, n) == crc32c_bitwise(view.tobytes())
    c = crc(b"foo")
    assert _lib.lib.kgcn_crc32c_masked(b"foo", 3) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def example_class():
    """tensorflow.Example / Features / Feature / {Bytes,Float,Int64}List (tensorflow/core/example/*.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fdp = descriptor_pb2.FileDescriptorProto(name="kgcn_test_example.proto", package="tensorflow", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def field(msg, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None):
        f = msg.field.add(name=name, number=number, type=ftype, label=label)
        if type_name:
            f.type_name = type_name
        if packed is not None:
            f.options.packed = packed
        return f

    field(fdp.message_type.add(name="BytesList"), "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    field(fdp.message_type.add(name="FloatList"), "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    field(fdp.message_type.add(name="Int64List"), "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    feat = fdp.message_type.add(name="Feature")
    feat.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List"))):
        field(feat, n, i + 1, T.TYPE_MESSAGE, type_name=".tensorflow." + t).oneof_index = 0
    feats = fdp.message_type.add(name="Features")
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    field(entry, "key", 1, T.TYPE_STRING)
    field(entry, "value", 2, T.TYPE_MESSAGE, type_name=".tensorflow.Feature")
    field(feats, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.Features.FeatureEntry")
    field(fdp.message_type.add(name="Example"), "features", 1, T.TYPE_MESSAGE, type_name=".tensorflow.Features")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.Example"))


def random_molecule(rng, n=None, input_dim=6):
    n = int(rng.integers(2, 9)) if n is None else n
    adj = np.eye(n, dtype=np.float32)
    for i in range(n - 1):
        adj[i, i + 1] = adj[i + 1, i] = 1.0
    if n > 3 and rng.random() < 0.5:
        adj[0, n - 1] = adj[n - 1, 0] = 1.0
    feat = np.zeros((n, input_dim), np.float32)
    feat[np.arange(n), rng.integers(0, input_dim, size=n)] = rng.integers(1, 5, size=n)
    return adj, feat


def test_parser_reads_protobuf_runtime_output(tmp_path):
    """Records serialised by the google.protobuf runtime (the encoder TensorFlow itself uses) parse to the same values."""
    Example = example_class()
    rng = np.random.default_rng(3)
    records, want = [], []
    for r in range(17):
        ints = rng.integers(-2**62, 2**62, size=int(rng.integers(0, 40))).tolist() + [0, -1, 2**63 - 1, -2**63]
        floats = rng.standard_normal(int(rng.integers(0, 30))).astype(np.float32)
        ex = Example()
        ex.features.feature["ints"].int64_list.value.extend(ints)
        ex.features.feature["floats"].float_list.value.extend(floats.tolist())
        ex.features.feature["name"].bytes_list.value.append(b"mol%d" % r)       # a kind the reader must step over
        if r % 3 == 0:
            ex.features.feature["sometimes"].int64_list.value.extend([r, r + 1])
        if r == 5:
            ex.features.feature["empty"].int64_list.SetInParent()
        records.append(ex.SerializeToString())
        want.append((ints, floats))
    path = str(tmp_path / "pb.tfrecords")
    tfrecords.write_tfrecords(path, records)
    tfr = tfrecords.TFRecordFile(path)
    assert len(tfr) == 17 and tfr.record(4) == records[4]
    ints, counts = tfr.gather("ints", "int64")
    assert ints.dtype == np.int64 and counts.tolist() == [len(w[0]) for w in want]
    assert ints.tolist() == [v for w in want for v in w[0]]
    floats, counts = tfr.gather("floats", "float32")
    assert floats.dtype == np.float32 and counts.tolist() == [len(w[1]) for w in want]
    assert floats.tobytes() == np.concatenate([w[1] for w in want]).tobytes()
    some, counts = tfr.gather("sometimes", "int64")                               # VarLen: absent key -> no values
    assert counts.tolist() == [2 if r % 3 == 0 else 0 for r in range(17)]
    assert tfr.gather("empty", "int64")[1].sum() == 0 and tfr.gather("nope", "float32")[0].size == 0
    sub, counts = tfr.gather("ints", "int64", records=[16, 2])
    assert sub.tolist() == want[16][0] + want[2][0] and counts.tolist() == [len(want[16][0]), len(want[2][0])]
    with pytest.raises(DataLoadError, match="another kind"):
        tfr.gather("ints", "float32")
    with pytest.raises(DataLoadError, match="another kind"):
        tfr.gather("name", "int64")


def test_unpacked_encoding_and_duplicate_keys(tmp_path):
    """proto2-style unpacked repeated scalars (one tag per value) and a repeated map key (last one wins)."""
    v, lf = tfrecords._varint, tfrecords._len_field
    ints = b"".join(v((1 << 3) | 0) + v(x) for x in (5, -7, 1 << 40))
    floats = b"".join(v((1 << 3) | 5) + struct.pack("<f", x) for x in (0.5, -2.0))
    mixed = lf(1, v(1) + v(2)) + v((1 << 3) | 0) + v(3)                         # a packed run then an unpacked value
    entry = lambda key, feat: lf(1, lf(1, key) + lf(2, feat))
    body = (entry(b"i", lf(3, ints)) + entry(b"f", lf(2, floats)) + entry(b"m", lf(3, mixed))
            + entry(b"dup", lf(3, lf(1, v(1)))) + entry(b"dup", lf(3, lf(1, v(2) + v(3)))))
    path = str(tmp_path / "raw.tfrecords")
    tfrecords.write_tfrecords(path, [lf(1, body)])
    tfr = tfrecords.TFRecordFile(path)
    assert tfr.gather("i", "int64")[0].tolist() == [5, -7, 1 << 40]
    assert tfr.gather("f", "float32")[0].tolist() == [0.5, -2.0]
    assert tfr.gather("m", "int64")[0].tolist() == [1, 2, 3]
    assert tfr.gather("dup", "int64")[0].tolist() == [2, 3]
    Example = example_class()                                                    # the runtime agrees on all of it
    ex = Example.FromString(lf(1, body))
    assert list(ex.features.feature["m"].int64_list.value) == [1, 2, 3]
    assert list(ex.features.feature["dup"].int64_list.value) == [2, 3]


def test_writer_matches_protobuf_runtime():
    """convert_to_example's bytes parse in the google.protobuf runtime, and equal its deterministic serialisation."""
    Example = example_class()
    rng = np.random.default_rng(5)
    for _ in range(8):
        adj, feat = random_molecule(rng)
        blob = tfrecords.convert_to_example(adj, feat, label_data=np.array([1.0, np.nan, 0.0]), label_mask=np.array([1, 0, 1]))
        ex = Example.FromString(blob)
        f = ex.features.feature
        r, c = np.nonzero(adj)
        assert list(f["adj_row"].int64_list.value) == r.tolist() and list(f["adj_column"].int64_list.value) == c.tolist()
        assert list(f["adj_elem_len"].int64_list.value) == [len(r)]
        deg = adj.sum(0)
        assert list(f["adj_degrees"].int64_list.value) == [0 if i == j else int(deg[i]) for i, j in zip(r, c)]
        fr, fc = np.nonzero(feat)
        assert list(f["feature_row"].int64_list.value) == fr.tolist()
        assert np.array(f["feature_values"].float_list.value, np.float32).tolist() == feat[fr, fc].tolist()
        assert list(f["size"].int64_list.value) == list(feat.shape)
        assert list(f["label"].int64_list.value) == [1, 0, 0] and list(f["mask_label"].int64_list.value) == [1, 0, 1]
        assert ex.SerializeToString(deterministic=True) == blob


def test_framing_and_corruption(tmp_path):
    path = str(tmp_path / "a.tfrecords")
    recs = [b"", b"x", bytes(range(200)) * 3]
    tfrecords.write_tfrecords(path, recs)
    raw = open(path, "rb").read()
    # framing, field by field, with an independent bitwise CRC
    pos = 0
    mask = lambda c: (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    for r in recs:
        n, = struct.unpack_from("<Q", raw, pos)
        assert n == len(r) and struct.unpack_from("<I", raw, pos + 8)[0] == mask(crc32c_bitwise(raw[pos:pos + 8]))
        assert raw[pos + 12:pos + 12 + n] == r
        assert struct.unpack_from("<I", raw, pos + 12 + n)[0] == mask(crc32c_bitwise(r))
        pos += 16 + n
    assert pos == len(raw)
    tfr = tfrecords.TFRecordFile(path)
    assert [tfr.record(i) for i in range(len(tfr))] == recs
    empty = str(tmp_path / "empty.tfrecords")
    open(empty, "wb").close()
    assert len(tfrecords.TFRecordFile(empty)) == 0
    for name, blob in (("flip_data", raw[:50] + bytes([raw[50] ^ 1]) + raw[51:]),
                       ("flip_len", bytes([raw[0] ^ 2]) + raw[1:]),
                       ("cut", raw[:-3]), ("cut_header", raw + b"\x01\x02\x03")):
        p = str(tmp_path / name)
        open(p, "wb").write(blob)
        with pytest.raises(DataLoadError, match="corrupted|truncated"):
            tfrecords.TFRecordFile(p)
    p = str(tmp_path / "flip_data")
    assert len(tfrecords.TFRecordFile(p, verify_crc=False)) == 3                 # framing is intact, only a payload bit flipped
    bad = str(tmp_path / "notproto.tfrecords")
    tfrecords.write_tfrecords(bad, [b"\x0a\xff\xff"])                            # length prefix runs past the record
    with pytest.raises(DataLoadError, match="not a valid Example"):
        tfrecords.TFRecordFile(bad).gather("x", "int64")


def test_docstring_example_through_files(tmp_path):
    """The worked example of kgcn/data_util.py:703-731 (molecules of 2 and 3 atoms), written with
    convert_to_example, read back and batched as example_model/sparse.py:47-62 does."""
    a0 = np.ones((2, 2), np.float32)
    a1 = np.array([[1, 1, 0], [1, 1, 1], [0, 1, 1]], np.float32)
    f0 = np.zeros((2, 10), np.float32); f0[0, 2], f0[1, 3] = 4, 5
    f1 = np.zeros((3, 10), np.float32); f1[0, 1], f1[1, 2], f1[2, 3] = 1, 2, 3
    path = str(tmp_path / "0_train_.tfrecords")
    tfrecords.write_tfrecords(path, [tfrecords.convert_to_example(a0, f0, np.array([1.0]), np.array([1])),
                                     tfrecords.convert_to_example(a1, f1, np.array([0.0]), np.array([1]))])
    ds = tfrecords.SparseDataset(str(tmp_path / "*_train_.tfrecords"), task_num=1)
    assert ds.info == {"num_elements": 2, "input_dim": 10}
    parsed, = list(ds.batches(2))
    assert parsed["size"].tolist() == [[2, 10], [3, 10]] and parsed["label"].tolist() == [[1], [0]]
    assert parsed["adj_row"].values.tolist() == [0, 0, 1, 1, 0, 0, 1, 1, 1, 2, 2]
    assert parsed["adj_column"].values.tolist() == [0, 1, 0, 1, 0, 1, 0, 1, 2, 1, 2]
    assert parsed["adj_elem_len"][:, 0].tolist() == [4, 7] and parsed["adj_row"].row_splits.tolist() == [0, 4, 11]
    chans, feat = tfrecords.block_diagonal_batch(parsed, normalize=False, split_adj=False)
    idx, val, shape = chans[0]
    dense = np.zeros(shape, np.float32)
    dense[idx[:, 0], idx[:, 1]] = val
    want = np.zeros((5, 5), np.float32); want[:2, :2] = a0; want[2:, 2:] = a1
    assert np.array_equal(dense, want)
    assert np.array_equal(feat, np.concatenate([f0, f1]))
    # the same arrays fed by hand give the same batch (file layer adds nothing)
    chans2, feat2 = data_util.construct_batched_adjacency_and_feature_matrices(
        [2, 3], parsed["adj_row"].values, parsed["adj_column"].values, parsed["adj_values"].values, [4, 7],
        parsed["adj_degrees"].values, parsed["feature_row"].values, parsed["feature_column"].values,
        parsed["feature_values"].values, [2, 3], 10, normalize=False)
    assert np.array_equal(chans2[0][0], idx) and np.array_equal(feat2, feat)
    # degree split (sparse.py with split_adj): 5 degree channels + identity
    chans6, _ = tfrecords.block_diagonal_batch(parsed, max_degree=5, normalize=False, split_adj=True)
    assert len(chans6) == 6 and np.array_equal(chans6[5][0][:, 0], np.arange(5))
    total = sum(len(c[1]) for c in chans6[:5])
    assert total == 11 - 5                                                        # diagonal entries carry degree 0 -> no channel


def test_batches_span_files_and_fixed_len_errors(tmp_path):
    rng = np.random.default_rng(9)
    mols = [random_molecule(rng) for _ in range(11)]
    blobs = [tfrecords.convert_to_example(a, f, np.array([float(i % 2), 1.0]), np.array([1, i % 2])) for i, (a, f) in enumerate(mols)]
    tfrecords.save_tfrecords(str(tmp_path), blobs[:5], blobs[5:7], blobs[7:], 3)
    assert sorted(os.listdir(tmp_path)) == ["3_eval_.tfrecords", "3_test_.tfrecords", "3_train_.tfrecords"]
    ds = tfrecords.SparseDataset([str(tmp_path / "3_train_.tfrecords"), str(tmp_path / "3_eval_.tfrecords"),
                                  str(tmp_path / "3_test_.tfrecords")], task_num=2)
    assert len(ds) == 11
    got = list(ds.batches(4))
    assert [b["size"].shape[0] for b in got] == [4, 4, 3]
    sizes = np.concatenate([b["size"][:, 0] for b in got])
    assert sizes.tolist() == [a.shape[0] for a, _ in mols]
    second = got[1]                                                               # records 4..7: train[4], eval[0..1], test[0]
    assert second["adj_row"].values.tolist() == np.concatenate([np.nonzero(a)[0] for a, _ in mols[4:8]]).tolist()
    assert second["adj_row"].row_splits.tolist() == np.concatenate([[0], np.cumsum([np.count_nonzero(a) for a, _ in mols[4:8]])]).tolist()
    assert second["mask_label"].tolist() == [[1, i % 2] for i in range(4, 8)]
    # every batch assembles into a block-diagonal matrix whose blocks are the molecules
    for b, lo in zip(got, (0, 4, 8)):
        chans, feat = tfrecords.block_diagonal_batch(b, normalize=False)
        idx, val, shape = chans[0]
        dense = np.zeros(shape, np.float32); dense[idx[:, 0], idx[:, 1]] = val
        off = 0
        for a, f in mols[lo:lo + b["size"].shape[0]]:
            n = a.shape[0]
            assert np.array_equal(dense[off:off + n, off:off + n], a) and np.array_equal(feat[off:off + n], f)
            off += n
        assert dense.sum() == sum(a.sum() for a, _ in mols[lo:lo + b["size"].shape[0]])
    with pytest.raises(DataLoadError, match="holds 2 values, the spec needs 3"):      # FixedLenFeature([task_num]) mismatch
        list(tfrecords.SparseDataset(str(tmp_path / "3_train_.tfrecords"), task_num=3).batches(2))
    unlabeled = str(tmp_path / "u.tfrecords")
    tfrecords.write_tfrecords(unlabeled, [tfrecords.convert_to_example(*mols[0])])
    with pytest.raises(DataLoadError, match="'label' of record 0 holds 0 values"):
        list(tfrecords.SparseDataset(unlabeled, task_num=2).batches(1))
    with pytest.raises(DataLoadError, match="no tfrecords file"):
        tfrecords.SparseDataset(str(tmp_path / "*.nothing"), task_num=1)


@pytest.mark.gpu
def test_tfrecords_to_block_diagonal_graphconv(tmp_path):
    """File -> parsed batch -> block-diagonal CSR -> GraphConv(C = max_degree + 1) -> relu -> per-molecule sum on the
    GPU (example_model/sparse.py:47-93 with split_adj), against the oracle fed by the same batch."""
    import torch
    from kgcn_b200 import layers, ops
    from kgcn_b200.csr import BatchedCSR
    from oracle import ref_layers as R
    rng = np.random.default_rng(21)
    mols = [random_molecule(rng, input_dim=8) for _ in range(24)]
    path = str(tmp_path / "0_train_.tfrecords")
    tfrecords.write_tfrecords(path, [tfrecords.convert_to_example(a, f, np.array([1.0]), np.array([1])) for a, f in mols])
    parsed, = list(tfrecords.SparseDataset(path, task_num=1).batches(24))
    for kw, C in ((dict(normalize=True), 1), (dict(normalize=False, split_adj=True, max_degree=3), 4)):
        chans, feat = tfrecords.block_diagonal_batch(parsed, **kw)
        adjs = [[(np.asarray(i), np.asarray(v, np.float32), s) for i, v, s in chans]]
        x = feat.astype(np.float32)[None]
        w = (0.3 * rng.standard_normal((C, 8, 16))).astype(np.float32)
        b = (0.1 * rng.standard_normal((C, 16))).astype(np.float32)
        sizes = parsed["size"][:, 0]
        ref = R.segment_sum(np.maximum(R.graph_conv(x, adjs, w, b), 0)[0], sizes)
        conv = layers.GraphConv(16, C, activation="relu")
        xt = torch.as_tensor(x).cuda()
        csr = BatchedCSR.from_coo_lists(adjs)
        conv(xt, adj=csr)
        with torch.no_grad():
            for c in range(C):
                conv.w[c].copy_(torch.as_tensor(w[c]).cuda()); conv.bias[c].copy_(torch.as_tensor(b[c]).cuda())
        out = ops.segment_sum(conv(xt, adj=csr)[0], sizes)
        np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


def test_parser_survives_mutated_records(tmp_path):
    """Memory safety of the native wire-format reader: truncated, bit-flipped and random records either parse or are
    rejected with DataLoadError -- never a crash or an out-of-bounds read (every length is checked against the record)."""
    rng = np.random.default_rng(123)
    adj, feat = random_molecule(rng, n=6)
    good = tfrecords.convert_to_example(adj, feat, np.array([1.0, 0.0]), np.array([1, 1]))
    records = [good[:k] for k in range(0, len(good), 7)]                               # truncations
    for _ in range(300):                                                               # bit flips / byte splices
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        records.append(bytes(b))
    records += [rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8).tobytes() for _ in range(200)]
    records += [b"\x0a" + b"\xff" * 9 + b"\x01", b"\x0a\x80", b"\x0a\x05\x0a\x03\x0a\x01", b"\x08" + b"\xff" * 12]  # varint / length edge cases
    path = str(tmp_path / "fuzz.tfrecords")
    tfrecords.write_tfrecords(path, records)
    tfr = tfrecords.TFRecordFile(path)
    assert len(tfr) == len(records)
    parsed = rejected = 0
    for i in range(len(tfr)):
        for key, kind in (("adj_row", "int64"), ("adj_values", "float32"), ("label", "int64"), ("nope", "float32")):
            try:
                values, counts = tfr.gather(key, kind, records=[i])
                assert values.shape[0] == counts.sum() and counts.shape == (1,)
                parsed += 1
            except DataLoadError:
                rejected += 1
    assert parsed > 0 and rejected > 0
    # random bytes as a FILE: framing errors are reported, not followed
    for _ in range(50):
        p = str(tmp_path / "junk")
        open(p, "wb").write(rng.integers(0, 256, size=int(rng.integers(1, 400)), dtype=np.uint8).tobytes())
        for verify in (True, False):
            try:
                tfrecords.TFRecordFile(p, verify_crc=verify)
            except DataLoadError:
                pass
