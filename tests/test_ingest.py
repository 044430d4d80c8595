"""CPU: host-side ingest is BIT-EXACT against vectors produced by the reference's own
kgcn/data_util.py + kgcn/feed.py (tests/golden/ingest_*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden, unflatten_adjs
from kgcn_b200 import csr as csr_mod
from kgcn_b200 import data_util, feed

CASES = [
    ("sample", "plain", {}), ("sample", "norm", {"normalize_adj_flag": True}), ("sample", "split", {"split_adj_flag": True}),
    ("sample", "order2", {"order": 2}), ("sample", "split_norm", {"split_adj_flag": True, "normalize_adj_flag": True}),
    ("sample_multiadj", "plain", {}), ("sample_multiadj", "norm", {"normalize_adj_flag": True}),
    ("synthetic", "plain", {}), ("synthetic", "norm", {"normalize_adj_flag": True}), ("synthetic", "split", {"split_adj_flag": True}),
    ("synthetic_sparse", "plain", {}),
]


def raw_data(rec):
    data = {}
    if "in_dense_adj" in rec:
        data["dense_adj"] = rec["in_dense_adj"]
    if "in_multi_dense_adj" in rec:
        data["multi_dense_adj"] = [list(m) for m in rec["in_multi_dense_adj"]]
    if "in_adj_counts" in rec:
        data["adj"] = [row[0] for row in unflatten_adjs(rec, "in_adj_")]
        for a, v0 in zip(data["adj"], np.split(rec["in_adj_values"], np.cumsum(rec["in_adj_counts"].reshape(-1))[:-1])):
            pass
    if "in_max_node_num" in rec:
        data["max_node_num"] = rec["in_max_node_num"]
    return data


@pytest.mark.parametrize("name,var,cfg", CASES)
def test_build_adjs_matches_reference(name, var, cfg):
    rec = load_golden("ingest_%s_%s" % (name, var))
    adjs, enabled, channels = data_util.build_adjs(raw_data(rec), cfg)
    want = unflatten_adjs(rec, "adj_")
    assert channels == int(rec["adj_channel_num"])
    np.testing.assert_array_equal(enabled, rec["enabled_node_nums"])
    assert enabled.dtype == np.int32
    assert len(adjs) == len(want)
    for g in range(len(want)):
        for c in range(channels):
            np.testing.assert_array_equal(np.asarray(adjs[g][c][0]).reshape(-1, 2), want[g][c][0])
            got_v = np.asarray(adjs[g][c][1], np.float32)
            assert got_v.tobytes() == want[g][c][1].tobytes(), (name, var, g, c)     # bit-exact fp32 values
            assert [int(adjs[g][c][2][0]), int(adjs[g][c][2][1])] == want[g][c][2]


@pytest.mark.parametrize("name,var", [("sample", "plain"), ("sample", "norm"), ("sample", "split"), ("sample_multiadj", "plain"),
                                      ("synthetic", "plain"), ("synthetic", "split")])
@pytest.mark.parametrize("tag", ["full", "short"])
def test_construct_feed_matches_reference(name, var, tag):
    rec = load_golden("ingest_%s_%s" % (name, var))
    data = {"adjs": unflatten_adjs(rec, "adj_"), "features": rec["features"], "labels": rec["in_label"],
            "enabled_node_nums": rec["enabled_node_nums"]}
    bi = rec[tag + "_feed_batch_idx"].tolist()
    bs = rec[tag + "_feed_features"].shape[0]
    keys = ["adjs", "features", "labels", "mask", "enabled_node_nums", "dropout_rate", "is_train"]
    fd = feed.construct_feed(bi, keys, data, batch_size=bs, config={"task": "classification"})
    want = unflatten_adjs(rec, tag + "_feed_")
    for b in range(bs):
        for c in range(len(want[0])):
            got = fd["adjs"][b][c]
            np.testing.assert_array_equal(np.asarray(got.indices).reshape(-1, 2), want[b][c][0])
            assert np.asarray(got.values, np.float32).tobytes() == want[b][c][1].tobytes()
            assert [int(got.dense_shape[0]), int(got.dense_shape[1])] == want[b][c][2]
    for k in ("features", "mask", "labels", "enabled_node_nums"):
        assert fd[k].dtype == rec[tag + "_feed_" + k].dtype, k
        np.testing.assert_array_equal(fd[k], rec[tag + "_feed_" + k])
    if tag == "short":  # feed.py:123-129,149-150,216-217
        assert fd["mask"][len(bi):].sum() == 0 and (fd["features"][len(bi):] == 0).all()
        assert fd["adjs"][-1][0].indices.shape == (0, 2)


def csr_reference(counts, indices, values, n_rows, transpose):
    """Independent stable CSR build with numpy (argsort kind='stable')."""
    rowptr, col, val, perm = [0], [], [], []
    pos = 0
    for n in counts.reshape(-1):
        i = indices[pos:pos + n]
        r, c = (i[:, 1], i[:, 0]) if transpose else (i[:, 0], i[:, 1])
        order = np.argsort(r, kind="stable")
        col.append(c[order]); val.append(values[pos:pos + n][order]); perm.append(order + pos)
        cnt = np.bincount(r, minlength=n_rows)
        rowptr.extend((rowptr[-1] + np.cumsum(cnt)).tolist())
        pos += n
    return (np.array(rowptr, np.int32), np.concatenate(col).astype(np.int32), np.concatenate(val).astype(np.float32),
            np.concatenate(perm).astype(np.int32))


@pytest.mark.parametrize("transpose", [False, True])
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_pack_host_is_stable_and_bit_exact(transpose, dtype):
    rng = np.random.default_rng(7)
    counts = rng.integers(0, 40, size=(13, 3))
    counts[4] = 0                                               # a fully empty (padded) graph
    nnz = int(counts.sum())
    idx = rng.integers(0, 11, size=(nnz, 2)).astype(dtype)      # unsorted, with duplicates
    val = rng.standard_normal(nnz).astype(np.float32)
    got = csr_mod.pack_host(counts, idx, val, 11, 11, transpose=transpose, want_perm=True)
    want = csr_reference(counts, idx.astype(np.int64), val, 11, transpose)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)


def test_pack_host_golden_fixture_roundtrip():
    rec = load_golden("ingest_synthetic_plain")
    counts, idx, val = rec["adj_counts"], rec["adj_indices"], rec["adj_values"]
    rowptr, col, v, _ = csr_mod.pack_host(counts, idx, val, 10, 10)
    dense = rec["in_dense_adj"]
    for g in (0, 1, 57, 199):
        a = np.zeros((10, 10), np.float32)
        for i in range(10):
            s, e = rowptr[g * 10 + i], rowptr[g * 10 + i + 1]
            a[i, col[s:e]] = v[s:e]
        np.testing.assert_array_equal(a, dense[g].astype(np.float32))


def test_pack_host_errors():
    from kgcn_b200 import KgcnError, KgcnIndexError
    with pytest.raises(KgcnIndexError):
        csr_mod.pack_host([1], np.array([[0, 5]], np.int32), np.ones(1, np.float32), 5, 5)
    with pytest.raises(KgcnIndexError):
        csr_mod.pack_host([1], np.array([[-1, 0]], np.int64), np.ones(1, np.float32), 5, 5)
    with pytest.raises(ValueError):
        csr_mod.pack_host([2], np.array([[0, 0]], np.int32), np.ones(1, np.float32), 5, 5)
    with pytest.raises(KgcnError):
        csr_mod.pack_host([0], np.zeros((0, 2), np.int32), np.zeros(0, np.float32), 0, 5)
    with pytest.raises(ValueError):  # ragged dense_shapes are rejected, not silently padded
        csr_mod.flatten_coo([[(np.zeros((0, 2), np.int32), np.zeros(0, np.float32), [3, 3])],
                             [(np.zeros((0, 2), np.int32), np.zeros(0, np.float32), [4, 4])]])


def test_flat_dataset_batching_matches_construct_feed():
    rec = load_golden("ingest_synthetic_split")
    adjs = unflatten_adjs(rec, "adj_")
    ds = feed.FlatGraphDataset.from_adjs(adjs, rec["features"], rec["in_label"], 10)
    bi = [5, 3, 199, 0, 42]
    host = ds.host_batch(bi, batch_size=8)
    fd = feed.construct_feed(bi, ["adjs", "features", "mask"], {"adjs": adjs, "features": rec["features"]}, batch_size=8)
    counts, idx, val, shape = csr_mod.flatten_coo(fd["adjs"])
    want = csr_mod.pack_host(counts, idx, val, 10, 10)
    np.testing.assert_array_equal(host["rowptr"], want[0])
    np.testing.assert_array_equal(host["col"], want[1])
    np.testing.assert_array_equal(host["val"], want[2])
    np.testing.assert_array_equal(host["features"], fd["features"])
    np.testing.assert_array_equal(host["mask"], fd["mask"])
