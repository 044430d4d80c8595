"""GraphMaxPooling, GraphBatchNormalization (both variants), the block-diagonal
(tfrecords-style) batch + per-molecule readout, integrated gradients.

CPU half: the oracle restatements against hand-derived known answers and torch-CPU autograd, and the host-side
block-diagonal builder against the worked example in the reference's own docstring (kgcn/data_util.py:703-731).
GPU half (-m gpu): the CUDA kernels through the C ABI against the oracle."""
import numpy as np
import pytest
import torch

from oracle import ref_layers as R


def rand_unique_adjs(rng, B, N, C, density=0.25, full_rows=True):
    """list[B][C] of sorted, duplicate-free COO triples (what tf.sparse_tensor_to_dense accepts)."""
    adjs = []
    for b in range(B):
        row = []
        for c in range(C):
            dense = (rng.random((N, N)) < density) * rng.standard_normal((N, N))
            if full_rows and N > 1:
                dense[0, :] = rng.standard_normal(N) - 2.0        # a completely stored row: no implicit zero
                dense[1 % N, :] = 0.0                             # an empty row
            idx = np.argwhere(dense != 0).astype(np.int32)
            row.append((idx, dense[idx[:, 0], idx[:, 1]].astype(np.float32), [N, N]))
        adjs.append(row)
    return adjs


def dense_of(adjs, B, C, N):
    a = np.zeros((B, C, N, N), np.float32)
    for b in range(B):
        for c in range(C):
            idx, val, _ = adjs[b][c]
            a[b, c, idx[:, 0], idx[:, 1]] = val
    return a


def torch_maxpool(x, a):
    """dense restatement with autograd: d[b,c,i,j,k] = A[b,c,i,j] * x[b,j,k]; amax over j shares ties evenly."""
    d = a[:, :, :, :, None] * x[:, None, None, :, :]
    return d.amax(dim=3).sum(dim=1)


# ------------------------------------------------------------------------------------------- CPU
def test_maxpool_oracle_known_answer():
    # sample.jbl g4: A = [[1,1,0],[0,0,0],[0,1,1]], x = [[1,-2],[3,4],[-5,6]]
    a = np.array([[1, 1, 0], [0, 0, 0], [0, 1, 1]], np.float32)
    idx = np.argwhere(a != 0)
    x = np.array([[[1, -2], [3, 4], [-5, 6]]], np.float32)
    y = R.graph_max_pooling(x, [[(idx, a[idx[:, 0], idx[:, 1]], [3, 3])]])
    # row 0: max(1*1, 1*3, 0) = 3 ; max(-2, 4, 0) = 4.  row 1: empty -> 0.  row 2: max(0, 3, -5) = 3 ; max(0, 4, 6) = 6
    np.testing.assert_array_equal(y[0], np.array([[3, 4], [0, 0], [3, 6]], np.float32))
    # a fully stored row has no implicit zero: the maximum may be negative
    a2 = np.array([[-1.0, -2.0], [0.0, 0.0]], np.float32)
    idx2 = np.array([[0, 0], [0, 1]])
    y2 = R.graph_max_pooling(np.array([[[1.0], [1.0]]], np.float32), [[(idx2, a2[idx2[:, 0], idx2[:, 1]], [2, 2])]])
    np.testing.assert_array_equal(y2[0, :, 0], np.array([-1.0, 0.0], np.float32))


def test_maxpool_oracle_equals_dense_torch():
    rng = np.random.default_rng(3)
    B, N, C, F = 3, 6, 2, 5
    adjs = rand_unique_adjs(rng, B, N, C)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    ref = torch_maxpool(torch.tensor(x), torch.tensor(dense_of(adjs, B, C, N))).numpy()
    np.testing.assert_allclose(R.graph_max_pooling(x, adjs), ref, rtol=0, atol=0)


def test_batchnorm_oracle():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((4, 5, 3)).astype(np.float32)
    n = np.array([5, 2, 0, 3])
    gamma, beta = np.array([1.0, 2.0, 0.5], np.float32), np.array([0.0, -1.0, 1.0], np.float32)
    # moving statistics at their initial values: y = gamma * x / sqrt(1 + 1e-3) + beta on enabled rows, 0 elsewhere
    y, _, _ = R.graph_batch_normalization(x, gamma, beta, np.zeros(3), np.ones(3), n)
    want = gamma * x / np.sqrt(1.001) + beta
    for b in range(4):
        np.testing.assert_allclose(y[b, :n[b]], want[b, :n[b]], rtol=1e-6)
        assert (y[b, n[b]:] == 0).all()
    # batch statistics: enabled rows end up with mean beta / variance ~ gamma^2
    y, mean, var = R.graph_batch_normalization(x, gamma, beta, None, None, n, batch_statistics=True)
    rows = np.concatenate([y[b, :n[b]] for b in range(4)])
    np.testing.assert_allclose(rows.mean(0), beta, atol=1e-5)
    np.testing.assert_allclose(rows.var(0), gamma ** 2 * var / (var + 1e-3), rtol=1e-4)


def test_block_diagonal_docstring_example():
    """The worked example of kgcn/data_util.py:703-731 (two molecules of 2 and 3 atoms)."""
    from kgcn_b200 import data_util
    size = [2, 3]
    adj_row = [0, 0, 1, 1, 0, 0, 1, 1, 1, 2, 2]
    adj_col = [0, 1, 0, 1, 0, 1, 0, 1, 2, 1, 2]
    chans, feat = data_util.construct_batched_adjacency_and_feature_matrices(
        size, adj_row, adj_col, np.ones(11, np.float32), [4, 7], None, [0, 1, 0, 1, 2], [2, 3, 1, 2, 3],
        np.array([4, 5, 1, 2, 3], np.int64), [2, 3], 10, normalize=False, split_adj=False)
    assert len(chans) == 1
    dense = R._dense_adj(chans[0])
    np.testing.assert_array_equal(dense, np.array([[1, 1, 0, 0, 0], [1, 1, 0, 0, 0], [0, 0, 1, 1, 0], [0, 0, 1, 1, 1],
                                                   [0, 0, 0, 1, 1]], np.float32))
    want = np.zeros((5, 10), np.int64)
    want[0, 2], want[1, 3], want[2, 1], want[3, 2], want[4, 3] = 4, 5, 1, 2, 3
    np.testing.assert_array_equal(feat, want)
    # normalize: A[i,j] / sqrt(d_j) / sqrt(d_i) with d = column sums (data_util.py:790-802)
    chans, _ = data_util.construct_batched_adjacency_and_feature_matrices(
        size, adj_row, adj_col, np.ones(11, np.float32), [4, 7], None, [0], [0], [1.0], [1, 0], 10, normalize=True)
    d = dense.sum(0)
    np.testing.assert_allclose(R._dense_adj(chans[0]), dense / np.sqrt(d)[None, :] / np.sqrt(d)[:, None], rtol=1e-6)
    # split_adj: one channel per clipped degree 1..max_degree plus the identity (:803-821)
    deg = [2, 2, 2, 2, 2, 2, 3, 3, 3, 2, 2]
    chans, _ = data_util.construct_batched_adjacency_and_feature_matrices(
        size, adj_row, adj_col, np.ones(11, np.float32), [4, 7], deg, [0], [0], [1.0], [1, 0], 10, max_degree=2,
        normalize=False, split_adj=True)
    assert len(chans) == 3 and chans[0][0].shape[0] == 0 and chans[1][0].shape[0] == 11   # degree 3 clips to 2
    np.testing.assert_array_equal(R._dense_adj(chans[2]), np.eye(5, dtype=np.float32))
    with pytest.raises(data_util.DataLoadError):
        data_util.construct_batched_adjacency_and_feature_matrices(size, [0, 2], [0, 0], [1, 1], [2, 0], None, [0], [0], [1.0],
                                                                    [1, 0], 4, normalize=False)


def test_segment_sum_oracle():
    x = np.arange(12, dtype=np.float32).reshape(6, 2)
    np.testing.assert_array_equal(R.segment_sum(x, [2, 0, 3]), np.array([[2, 4], [0, 0], [18, 21]], np.float32))


# ------------------------------------------------------------------------------------------- GPU
def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,C,F", [(5, 3, 1, 4), (7, 10, 2, 3), (33, 32, 1, 64), (6, 50, 3, 50), (2, 17, 1, 8)])
def test_maxpool_cuda_forward_backward(B, N, C, F):
    from kgcn_b200 import layers
    from kgcn_b200.csr import BatchedCSR
    rng = np.random.default_rng(B * 100 + N)
    adjs = rand_unique_adjs(rng, B, N, C)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    x[0, :, 0] = 0.0     # exact ties with the implicit zeros
    if N > 2:
        x[0, 2, :] = x[0, 1, :]
    csr = BatchedCSR.from_coo_lists(adjs)
    layer = layers.GraphMaxPooling(C)
    xt = dev(x).requires_grad_(True)
    y = layer(xt, adj=csr)
    ref_y = R.graph_max_pooling(x, adjs)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), ref_y)          # products and maxima are exact: bit-equal
    dy = rng.standard_normal((B, N, F)).astype(np.float32)
    y.backward(dev(dy))
    xc = torch.tensor(x, requires_grad=True)
    torch_maxpool(xc, torch.tensor(dense_of(adjs, B, C, N))).backward(torch.tensor(dy))
    np.testing.assert_allclose(xt.grad.cpu().numpy(), xc.grad.numpy(), rtol=1e-5, atol=1e-6)
    # inference call (no workspace) gives the same forward
    with torch.no_grad():
        np.testing.assert_array_equal(layer(dev(x), adj=csr).cpu().numpy(), ref_y)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,F,masked", [(4, 5, 3, True), (64, 50, 64, True), (9, 32, 50, False), (300, 10, 7, True)])
@pytest.mark.parametrize("batch_stats", [False, True])
def test_graph_batch_norm_cuda(B, N, F, masked, batch_stats):
    from kgcn_b200 import layers
    rng = np.random.default_rng(B + N + F)
    x = (rng.standard_normal((B, N, F)) * 2 + 0.5).astype(np.float32)
    n = rng.integers(0, N + 1, size=B).astype(np.int32) if masked else None
    gamma = rng.uniform(0.5, 1.5, F).astype(np.float32)
    beta = rng.standard_normal(F).astype(np.float32)
    layer = layers.GraphBatchNormalization(batch_statistics=batch_stats)
    xt = dev(x).requires_grad_(True)
    layer(xt, enabled_node_nums=None if n is None else dev(n, torch.int32), max_node_num=N)      # builds
    with torch.no_grad():
        layer.gamma.copy_(dev(gamma))
        layer.beta.copy_(dev(beta))
        layer.moving_mean.zero_()
        layer.moving_variance.fill_(1.0)
    y = layer(xt, enabled_node_nums=None if n is None else dev(n, torch.int32), max_node_num=N)
    ref, mean, var = R.graph_batch_normalization(x, gamma, beta, np.zeros(F), np.ones(F), n, batch_statistics=batch_stats)
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
    if n is not None:
        keep = np.arange(N)[None, :] < n[:, None]
        assert (y.detach().cpu().numpy()[~keep] == 0).all()
    if batch_stats:   # moving averages moved towards the batch statistics (momentum 0.99)
        np.testing.assert_allclose(layer.moving_mean.cpu().numpy(), 0.01 * mean, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(layer.moving_variance.cpu().numpy(), 0.99 + 0.01 * var, rtol=1e-4)
    # gradients against torch-CPU autograd of the same formula
    dy = rng.standard_normal((B, N, F)).astype(np.float32)
    y.backward(dev(dy))
    xc = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    gc = torch.tensor(gamma, dtype=torch.float64, requires_grad=True)
    bc = torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    keep_t = torch.ones(B, N, dtype=torch.bool) if n is None else torch.tensor(np.arange(N)[None, :] < n[:, None])
    rows = xc[keep_t]
    if batch_stats:
        m, v = rows.mean(0), rows.var(0, unbiased=False)
    else:
        m, v = torch.zeros(F, dtype=torch.float64), torch.ones(F, dtype=torch.float64)
    yc = ((xc - m) / torch.sqrt(v + 1e-3) * gc + bc) * keep_t[:, :, None]
    yc.backward(torch.tensor(dy, dtype=torch.float64))
    for got, want in ((xt.grad, xc.grad), (layer.gamma.grad, gc.grad), (layer.beta.grad, bc.grad)):
        want = want.numpy()
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-4, atol=2e-5 * max(1.0, float(np.abs(want).max())))


@pytest.mark.gpu
def test_block_diagonal_model_forward_backward():
    """example_model/sparse.py in miniature: block-diagonal batch (B = 1, N = all atoms) -> GraphConv -> relu ->
    per-molecule segment sum, against the oracle on the same inputs."""
    from kgcn_b200 import data_util, layers, ops
    from kgcn_b200.csr import BatchedCSR
    rng = np.random.default_rng(11)
    sizes = rng.integers(3, 12, size=20)
    rows, cols, lens, frow, fcol, flen = [], [], [], [], [], []
    for s in sizes:
        a = np.eye(s, dtype=bool)
        for i in range(s - 1):
            a[i, i + 1] = a[i + 1, i] = True
        idx = np.argwhere(a)
        rows += idx[:, 0].tolist(); cols += idx[:, 1].tolist(); lens.append(len(idx))
        frow += list(range(s)); fcol += rng.integers(0, 8, size=s).tolist(); flen.append(s)
    chans, feat = data_util.construct_batched_adjacency_and_feature_matrices(
        sizes, rows, cols, np.ones(len(rows), np.float32), lens, None, frow, fcol, np.ones(len(frow), np.float32), flen, 8,
        normalize=True)
    n = int(sizes.sum())
    adjs = [[chans[0]]]
    x = feat.astype(np.float32)[None]                                   # [1, n, 8]
    w = rng.standard_normal((1, 8, 16)).astype(np.float32) * 0.3
    b = rng.standard_normal((1, 16)).astype(np.float32) * 0.1
    ref_h = np.maximum(R.graph_conv(x, adjs, w, b), 0)
    ref_out = R.segment_sum(ref_h[0], sizes)

    csr = BatchedCSR.from_coo_lists(adjs)
    conv = layers.GraphConv(16, 1, activation="relu")
    xt = dev(x).requires_grad_(True)
    conv(xt, adj=csr)
    with torch.no_grad():
        conv.w[0].copy_(dev(w[0])); conv.bias[0].copy_(dev(b[0]))
    h = conv(xt, adj=csr)
    out = ops.segment_sum(h[0], sizes)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref_out, rtol=1e-5, atol=1e-5 * np.abs(ref_out).max())
    # backward of the readout broadcasts the molecule's gradient to its atoms
    g = rng.standard_normal(ref_out.shape).astype(np.float32)
    out.backward(dev(g))
    hc = torch.tensor(ref_h[0], requires_grad=True)
    seg = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    torch.zeros(len(sizes), 16).index_add(0, seg, hc).backward(torch.tensor(g))
    assert xt.grad is not None and conv.w[0].grad is not None
    dh_ref = hc.grad.numpy() * (ref_h[0] > 0)
    # d x = A^T (dh . W^T): check through the oracle's graph_conv_grad
    dx_ref = R.graph_conv_grad(x, adjs, w, b, dh_ref[None])[0]
    np.testing.assert_allclose(xt.grad.cpu().numpy(), dx_ref, rtol=1e-4, atol=1e-5 * max(1.0, float(np.abs(dx_ref).max())))


@pytest.mark.gpu
def test_integrated_gradients_completeness():
    """sum of IG ~ score(x) - score(0) (the reference's own check, kgcn/visualization.py:254-257), and the
    adjacency-value attribution through the registered d-values gradient of the batched SpMM."""
    from kgcn_b200 import bconv_call, layers, visualization
    from kgcn_b200.csr import BatchedCSR
    rng = np.random.default_rng(2)
    torch.manual_seed(0)
    B, N, F = 6, 8, 5
    adjs = rand_unique_adjs(rng, B, N, 1, density=0.4, full_rows=False)
    csr = BatchedCSR.from_coo_lists(adjs)
    x = (0.3 * rng.standard_normal((B, N, F))).astype(np.float32)
    conv, gather = layers.GraphConv(4, 1, activation="tanh"), layers.GraphGather()
    xt = dev(x)
    conv(xt, adj=csr)

    def score(features, _values):
        return torch.tanh(gather(conv(features, adj=csr))).sum()

    res = visualization.integrated_gradients(score, xt, None, divide_number=256, method="ig")
    with torch.no_grad():
        full, zero = score(xt, None).item(), score(torch.zeros_like(xt), None).item()
    assert abs(res["sum_of_IG"] - (full - zero)) < 0.05 * max(1.0, abs(full - zero))
    assert res["features"].shape == xt.shape
    grad = visualization.integrated_gradients(score, xt, None, method="grad")["features"]
    xg = xt.clone().requires_grad_(True)
    np.testing.assert_allclose(grad.cpu().numpy(), torch.autograd.grad(score(xg, None), xg)[0].cpu().numpy(), rtol=1e-6)

    # adjacency values: y = sum_b sum(A_b . h_b) is linear in the values -> IG = values * d score / d values, exactly
    counts = [len(a[0][1]) for a in adjs]
    flat = dev(np.concatenate([a[0][1] for a in adjs]))
    h = [dev(rng.standard_normal((N, 3)).astype(np.float32)) for _ in range(B)]

    def score_adj(_features, values):
        parts = torch.split(values, counts)
        sp = [[(adjs[b][0][0], parts[b], [N, N])] for b in range(B)]
        out = bconv_call.BatchedConv().call(sp, [[h[b]] for b in range(B)])
        return torch.stack(out).sum()

    res = visualization.integrated_gradients(score_adj, xt, flat, divide_number=4, method="ig")
    want = np.concatenate([np.asarray(a[0][1]) * h[b].cpu().numpy().sum(1)[a[0][0][:, 1]] for b, a in enumerate(adjs)])
    np.testing.assert_allclose(res["adjs"].cpu().numpy(), want, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bspmm", "bconv", "batched"])
def test_plugin_mode_graphconv_returns_adjacency_value_gradients(mode):
    """A plugin-mode GraphConv (kgcn/layers.py:68-104) fed triples whose ``values`` are tensors on the tape: the registered
    d-values gradient of the plugin op (kgcn/bspmm_call.py:49-54, bconv_call.py:62-67) reaches them -- what the adjacency
    attribution of kgcn/visualization.py needs.  Checked against float64 torch-CPU autograd on dense adjacencies."""
    import types
    from kgcn_b200 import layers as L
    rng = np.random.default_rng(5)
    B, N, C, F, H = 4, 7, 2, 5, 6
    adjs = rand_unique_adjs(rng, B, N, C, density=0.4, full_rows=False)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    L.load_bspmm(types.SimpleNamespace(bspmm=mode == "bspmm", bconv=mode == "bconv", batched=mode == "batched"))
    try:
        conv = L.GraphConv(H, C)
        vals = [[dev(a[1]).requires_grad_(True) for a in row] for row in adjs]
        sp = [[(a[0], v, a[2]) for a, v in zip(row, vrow)] for row, vrow in zip(adjs, vals)]
        xt = dev(x).requires_grad_(True)
        y = conv(xt, adj=sp)
        dy = rng.standard_normal((B, N, H)).astype(np.float32)
        y.backward(dev(dy))
    finally:
        L.load_bspmm(types.SimpleNamespace(bspmm=False, bconv=False, batched=False))
    w64 = [torch.tensor(conv.w[c].detach().cpu().numpy(), dtype=torch.float64) for c in range(C)]
    b64 = [torch.tensor(conv.bias[c].detach().cpu().numpy(), dtype=torch.float64) for c in range(C)]
    x64 = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    v64 = [[torch.tensor(a[1], dtype=torch.float64, requires_grad=True) for a in row] for row in adjs]
    out = []
    for b in range(B):
        acc = 0
        for c in range(C):
            idx = np.asarray(adjs[b][c][0]).reshape(-1, 2)
            A = torch.zeros(N, N, dtype=torch.float64).index_put((torch.as_tensor(idx[:, 0]).long(), torch.as_tensor(idx[:, 1]).long()),
                                                                 v64[b][c], accumulate=True)
            acc = acc + A @ (x64[b] @ w64[c] + b64[c])
        out.append(acc)
    torch.stack(out).backward(torch.tensor(dy, dtype=torch.float64))
    np.testing.assert_allclose(xt.grad.cpu().numpy(), x64.grad.numpy(), rtol=1e-4, atol=1e-5)
    for b in range(B):
        for c in range(C):
            if len(adjs[b][c][1]):
                assert vals[b][c].grad is not None, (mode, b, c)
                np.testing.assert_allclose(vals[b][c].grad.cpu().numpy(), v64[b][c].grad.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,C,F", [(5, 3, 1, 4), (7, 10, 2, 3), (33, 32, 1, 64), (9, 50, 3, 50), (6, 64, 2, 128), (2, 300, 1, 48),
                                     (40, 32, 2, 64)])
def test_gin_aggregate_fused_epsilon_forward_backward(B, N, C, F):
    """GINAggregate (kgcn/layers.py:459-471) with the epsilon term inside the SpMM launch (tile kernel flat / general
    paths and the row kernel, by shape): forward against the oracle, gradients (x and every epsilon) against
    torch-CPU autograd of the dense formula."""
    from kgcn_b200 import layers, ops
    from kgcn_b200.csr import BatchedCSR
    rng = np.random.default_rng(B * 100 + N + C)
    adjs = rand_unique_adjs(rng, B, N, C, density=min(0.3, 6.0 / N), full_rows=False)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    eps = rng.uniform(-0.7, 0.7, C).astype(np.float32)
    csr = BatchedCSR.from_coo_lists(adjs)
    gin = layers.GINAggregate(C)
    xt = dev(x).requires_grad_(True)
    gin(xt, adj=csr)
    with torch.no_grad():
        for c in range(C):
            gin.epsilon[c].fill_(float(eps[c]))
    y = gin(xt, adj=csr)
    ref = R.gin_aggregate(x, adjs, eps)
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
    dy = rng.standard_normal((B, N, F)).astype(np.float32)
    y.backward(dev(dy))
    a = torch.tensor(dense_of(adjs, B, C, N), dtype=torch.float64)
    xc = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ec = torch.tensor(eps, dtype=torch.float64, requires_grad=True)
    yc = sum(ec[c] * xc + torch.einsum("bij,bjf->bif", a[:, c], xc) for c in range(C))
    yc.backward(torch.tensor(dy, dtype=torch.float64))
    want = xc.grad.numpy()
    np.testing.assert_allclose(xt.grad.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    got_eps = np.array([float(gin.epsilon[c].grad) for c in range(C)])
    np.testing.assert_allclose(got_eps, ec.grad.numpy(), rtol=2e-4, atol=2e-4 * max(1.0, float(np.abs(ec.grad.numpy()).max())))
    # the reduction helper on its own: odd sizes exercise the zero-padded folding
    for shape in ((1, 1, 1), (3, 5, 7), (33, 37, 9)):
        p, q = rng.standard_normal(shape).astype(np.float32), rng.standard_normal(shape).astype(np.float32)
        got = float(ops.dot_all(dev(p), dev(q)))
        assert abs(got - float((p.astype(np.float64) * q).sum())) <= 1e-4 * max(1.0, float(np.abs(p * q).sum()))
