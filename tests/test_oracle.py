"""CPU: pin the oracle (oracle/ref_layers.py) -- known-answer vectors, tier cross-check, gradients."""
import numpy as np
import pytest
import torch

from conftest import load_golden, unflatten_adjs
from oracle import ref_layers as R

W = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
B0 = np.array([[0.5, -0.5]], np.float32)
# SURVEY.md Appendix B, KAT1 (hand-derived from kgcn/layers.py:105-116 on example_jbl/sample.jbl)
KAT1 = np.array([[[9, 9], [1.5, 1.5], [1.5, 1.5]], [[3.5, 3.5], [1.5, 1.5], [0, 0]], [[5.5, 5.5], [5.5, 5.5], [5, 5]],
                 [[7.5, 7.5], [7.5, 7.5], [9, 9]], [[13, 13], [0, 0], [11, 11]]], np.float32)
KAT1_GATHER = np.array([[12, 12], [5, 5], [16, 16], [24, 24], [24, 24]], np.float32)


def sample():
    rec = load_golden("ingest_sample_plain")
    return unflatten_adjs(rec, "adj_"), rec["features"].astype(np.float32)


def test_kat1_graphconv_gather_adjoint():
    adjs, x = sample()
    y = R.graph_conv(x, adjs, [W], [B0])
    np.testing.assert_array_equal(y, KAT1)
    np.testing.assert_array_equal(R.graph_gather(y), KAT1_GATHER)
    a4 = adjs[4][0]
    np.testing.assert_array_equal(R.sparse_dense_matmul(a4[0], a4[1], a4[2], y[4], adjoint_a=True),
                                  np.array([[13, 13], [24, 24], [11, 11]], np.float32))
    kat = load_golden("kat")
    np.testing.assert_array_equal(kat["kat1_y"], KAT1)


def test_kat2_multichannel_bias_times_degree():
    rec = load_golden("ingest_sample_multiadj_plain")
    adjs, x = unflatten_adjs(rec, "adj_"), rec["features"].astype(np.float32)
    y = R.graph_conv(x, adjs, [W, -W], [B0, np.array([[1, 2]], np.float32)])
    np.testing.assert_array_equal(y[0], np.array([[3, 3], [1.5, 1.5], [1.5, 1.5]], np.float32))
    np.testing.assert_array_equal(y[3], np.array([[1.5, 1.5], [1.5, 1.5], [3, 3]], np.float32))


def test_kat3_normalize_adj_values():
    rec = load_golden("ingest_sample_norm")
    adjs = unflatten_adjs(rec, "adj_")
    np.testing.assert_array_equal(adjs[4][0][0], np.array([[0, 0], [0, 1], [2, 1], [2, 2]]))
    np.testing.assert_allclose(adjs[4][0][1], np.array([1, 0.70710677, 0.70710677, 1], np.float32), rtol=0, atol=1e-7)
    np.testing.assert_allclose(adjs[0][0][1], np.full(4, 0.70710677, np.float32), rtol=0, atol=1e-7)


def random_batch(rng, B, N, C, F, density=0.15, dup=True):
    adjs = []
    for _ in range(B):
        row = []
        for _ in range(C):
            nnz = rng.integers(0, max(2, int(N * N * density)))
            idx = rng.integers(0, N, size=(nnz, 2)).astype(np.int32)   # unsorted, duplicates allowed
            if not dup and nnz:
                idx = np.unique(idx, axis=0)
            row.append((idx, rng.standard_normal(idx.shape[0]).astype(np.float32), [N, N]))
        adjs.append(row)
    return adjs, rng.standard_normal((B, N, F)).astype(np.float32)


def test_tiers_agree_and_duplicates_accumulate():
    rng = np.random.default_rng(0)
    adjs, x = random_batch(rng, 6, 9, 2, 5)
    w = [rng.standard_normal((5, 7)).astype(np.float32) for _ in range(2)]
    b = [rng.standard_normal((1, 7)).astype(np.float32) for _ in range(2)]
    slow, fast = R.graph_conv(x, adjs, w, b), R.graph_conv(x, adjs, w, b, fast=True)
    np.testing.assert_allclose(fast, slow, rtol=1e-5, atol=1e-5)
    out = R.sparse_dense_matmul(np.array([[1, 0], [1, 0]]), np.array([2.0, 3.0], np.float32), [2, 2], np.array([[1.0], [0.0]], np.float32))
    np.testing.assert_array_equal(out, np.array([[0.0], [5.0]], np.float32))


def test_out_of_range_index_raises():
    with pytest.raises(IndexError):
        R.sparse_dense_matmul(np.array([[0, 3]]), np.ones(1, np.float32), [3, 3], np.ones((3, 2), np.float32))


def test_bias_before_aggregation_identity():
    """A.(XW+b) == (A.X)W + rowsum(A) (x) b  -- SURVEY Appendix A.1; isolated rows output 0, not b."""
    rng = np.random.default_rng(1)
    adjs, x = random_batch(rng, 4, 8, 1, 6)
    w, b = rng.standard_normal((6, 3)).astype(np.float32), rng.standard_normal((1, 3)).astype(np.float32)
    y = R.graph_conv(x, adjs, [w], [b])
    for g in range(4):
        a = np.zeros((8, 8), np.float32)
        np.add.at(a, (adjs[g][0][0][:, 0], adjs[g][0][0][:, 1]), adjs[g][0][1])
        np.testing.assert_allclose(y[g], (a @ x[g]) @ w + a.sum(1, keepdims=True) * b, rtol=1e-4, atol=1e-4)
        empty = np.where(np.bincount(adjs[g][0][0][:, 0], minlength=8) == 0)[0]
        assert (y[g][empty] == 0).all()


def _torch_graph_conv(x, adjs, w, b):
    outs = []
    for g in range(x.shape[0]):
        acc = 0
        for c in range(len(w)):
            idx = torch.as_tensor(np.asarray(adjs[g][c][0]).reshape(-1, 2).T.astype(np.int64))
            a = torch.sparse_coo_tensor(idx, torch.as_tensor(adjs[g][c][1]), tuple(adjs[g][c][2])).to_dense()
            acc = acc + a @ (x[g] @ w[c] + b[c])
        outs.append(acc)
    return torch.stack(outs)


def test_network_grad_matches_torch_autograd():
    rng = np.random.default_rng(2)
    B, N, C, F = 5, 7, 2, 6
    adjs, x = random_batch(rng, B, N, C, F)
    p = R.init_network(rng, F, [8, 5], C, 3, dense_dim=4)
    labels = np.eye(3, dtype=np.float32)[rng.integers(0, 3, B)]
    mask = np.array([1, 1, 1, 1, 0], np.float32)
    fw, grads = R.network_grad(p, x, adjs, labels, mask, act="sigmoid")

    tp = {k: ([[torch.tensor(a, requires_grad=True) for a in layer] for layer in v] if k.startswith("conv") else torch.tensor(v, requires_grad=True))
          for k, v in p.items()}
    h = torch.tensor(x)
    for w, b in zip(tp["conv_w"], tp["conv_b"]):
        h = torch.sigmoid(_torch_graph_conv(h, adjs, w, b))
    h = torch.sigmoid(h @ tp["gd_w"] + tp["gd_b"])
    logits = h.sum(1) @ tp["out_w"] + tp["out_b"]
    cost = torch.tensor(mask) * -(torch.tensor(labels) * torch.log_softmax(logits, 1)).sum(1)
    cost.mean().backward()
    np.testing.assert_allclose(fw["logits"], logits.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fw["cost_opt"], cost.mean().item(), rtol=1e-5)
    np.testing.assert_allclose(fw["cost_sum"], cost.sum().item(), rtol=1e-5)
    for k in ("gd_w", "gd_b", "out_w", "out_b"):
        np.testing.assert_allclose(grads[k], tp[k].grad.numpy(), rtol=2e-4, atol=2e-6)
    for layer in range(2):
        for c in range(C):
            np.testing.assert_allclose(grads["conv_w"][layer][c], tp["conv_w"][layer][c].grad.numpy(), rtol=2e-4, atol=2e-6)
            np.testing.assert_allclose(grads["conv_b"][layer][c], tp["conv_b"][layer][c].grad.numpy(), rtol=2e-4, atol=2e-6)


def test_plugin_gradient_restatement():
    """bspmm_grad (bspmm_call.py:44-54) against torch autograd on dense equivalents."""
    rng = np.random.default_rng(3)
    adjs, x = random_batch(rng, 3, 6, 1, 4, dup=False)
    sp = [a[0] for a in adjs]
    dy = [rng.standard_normal((6, 4)).astype(np.float32) for _ in range(3)]
    dvals, db = R.bspmm_grad(sp, list(x), dy)
    for t in range(3):
        idx = torch.as_tensor(sp[t][0].T.astype(np.int64))
        v = torch.tensor(sp[t][1], requires_grad=True)
        b = torch.tensor(x[t], requires_grad=True)
        a = torch.zeros(6, 6).index_put((idx[0], idx[1]), v, accumulate=True)
        (a @ b).backward(torch.tensor(dy[t]))
        np.testing.assert_allclose(db[t], b.grad.numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(dvals[t], v.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.stack(R.bconv([[a[0], a[0]] for a in adjs], [[xx, 2 * xx] for xx in x])),
                               3 * np.stack(R.bspmm(sp, list(x))), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.stack(R.bspmdt(sp, x.reshape(-1, 4))), np.stack(R.bspmm(sp, list(x))))


def test_graph_dense_masking_and_gather_padding():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 5, 4)).astype(np.float32)
    k, b = rng.standard_normal((4, 2)).astype(np.float32), rng.standard_normal(2).astype(np.float32)
    full = R.graph_dense(x, k, b, act="sigmoid")
    masked = R.graph_dense(x, k, b, act="sigmoid", enabled_node_nums=[5, 2, 0])
    np.testing.assert_array_equal(masked[0], full[0])
    assert (masked[1, 2:] == 0).all() and (masked[2] == 0).all()
    np.testing.assert_array_equal(masked[1, :2], full[1, :2])
    # padded (all-zero input) rows still contribute sigmoid(bias) to the readout (layers.py:255-262,164)
    xz = np.zeros((1, 4, 4), np.float32)
    np.testing.assert_allclose(R.graph_gather(R.graph_dense(xz, k, b, act="sigmoid"))[0], 4 / (1 + np.exp(-b)), rtol=1e-6)
