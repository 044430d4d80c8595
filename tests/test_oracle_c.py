"""CPU: the C restatement (oracle/graphconv_ref.c, the timed CPU baseline) agrees with the numpy
oracle (oracle/ref_layers.py) and the KAT vectors."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import ref_layers as R

cref = pytest.importorskip("oracle.cref")


def flat(adjs):
    counts = np.array([[a[0].shape[0] for a in row] for row in adjs], np.int64)
    idx = np.concatenate([np.asarray(a[0], np.int32).reshape(-1, 2) for row in adjs for a in row], 0)
    val = np.concatenate([np.asarray(a[1], np.float32) for row in adjs for a in row], 0)
    return counts, idx, val


def random_batch(rng, B, N, C, F):
    adjs = []
    for _ in range(B):
        row = []
        for _ in range(C):
            nnz = int(rng.integers(0, 3 * N))
            row.append((rng.integers(0, N, size=(nnz, 2)).astype(np.int32), rng.standard_normal(nnz).astype(np.float32), [N, N]))
        adjs.append(row)
    return adjs, rng.standard_normal((B, N, F)).astype(np.float32)


def test_kat1_through_c():
    rec = load_golden("ingest_sample_plain")
    kat = load_golden("kat")
    W = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
    y = cref.graphconv_fwd(rec["adj_counts"], rec["adj_indices"], rec["adj_values"], rec["features"], W[None],
                           np.array([[0.5, -0.5]], np.float32))
    np.testing.assert_array_equal(y, kat["kat1_y"])


@pytest.mark.parametrize("B,N,C,fi,fo,act", [(7, 9, 2, 5, 6, 0), (20, 32, 1, 64, 64, 2), (6, 50, 3, 75, 50, 1)])
def test_graphconv_fwd_matches_numpy(B, N, C, fi, fo, act):
    rng = np.random.default_rng(fi)
    adjs, x = random_batch(rng, B, N, C, fi)
    w = [R.glorot_uniform(rng, fi, fo) for _ in range(C)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32) for _ in range(C)]
    want = R.activation(R.graph_conv(x, adjs, w, b, fast=False), act)
    got = cref.graphconv_fwd(*flat(adjs), x, np.stack(w), np.concatenate(b, 0), act=act, n_threads=3)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("C,dims", [(1, [8, 5]), (2, [6, 6, 4])])
def test_train_step_matches_numpy(C, dims):
    rng = np.random.default_rng(9)
    B, N, F = 9, 7, 6
    adjs, x = random_batch(rng, B, N, C, F)
    labels = np.eye(3, dtype=np.float32)[rng.integers(0, 3, B)]
    mask = np.ones(B, np.float32); mask[-1] = 0
    p = R.init_network(rng, F, dims, C, 3)
    fw, grads = R.network_grad(p, x, adjs, labels, mask, act="sigmoid")
    net = cref.RefNet(F, dims, C, 3, act=2)
    net.load_oracle_params(p)
    stats, logits = net.train_step(*flat(adjs), x, labels, mask, N, apply_update=False, n_threads=2)
    np.testing.assert_allclose(logits, fw["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(stats[0], fw["cost_sum"], rtol=1e-5)
    assert stats[1] == fw["correct_count"]
    for l in range(len(dims)):
        np.testing.assert_allclose(net.view(net.grads, "conv%d/kernel" % l), np.stack(grads["conv_w"][l]), rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(net.view(net.grads, "conv%d/bias" % l), np.concatenate(grads["conv_b"][l], 0), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(net.view(net.grads, "dense/kernel"), grads["out_w"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(net.view(net.grads, "dense/bias"), grads["out_b"], rtol=1e-3, atol=1e-6)
