"""CPU oracle: numpy restatement of kGCN's GraphConv / GraphDense / GraphGather maths.

TEST INFRASTRUCTURE ONLY -- only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  The product path
(``kgcn_b200/``) never does and fails loudly when its CUDA library is missing.

PARITY PIN (arithmetic half): the reference ships no golden outputs, known-answer vectors or tests for these layers
 and its arithmetic primitives live in un-vendored TensorFlow 1.15.0 (requirements.yaml:84),
which is not installable here.  What pins this file: (i) golden layer outputs produced by the reference's OWN
``kgcn/layers.py`` / ``kgcn/legacy/layers.py`` imported and executed unchanged under ``oracle/tf_numpy.py`` (its
loops, op order, indexing, bias placement, padding rules; only the ~25 TensorFlow primitives are numpy restatements)
on batches ingested by the reference's own ``data_util`` / ``feed`` code -- ``oracle/make_layer_golden.py`` ->
``tests/golden/layers_*.npz``, reproduced bit for bit by ``tests/test_reference_layers.py``; that run also confirms
the hand-derived known-answer vectors KAT1-KAT3 of SURVEY.md Appendix B; (ii) a slow "faithful" tier (O1) that follows
the reference's operation order literally and a fast tier (O2, scipy CSR) cross-checked against it; (iii) the
integer/ingest half pinned by the reference's numpy code run under ``oracle/tf_stub.py`` (``oracle/make_golden.py``
-> ``tests/golden/ingest_*.npz``).  What remains unpinned: TensorFlow's own kernels were never executed (summation
order inside ``tf.matmul`` and Keras' BatchNormalization learning-phase behaviour are taken from documentation).

All citations are relative to /root/reference (clinfo/kGCN @ 32328d5).  float32 throughout.
"""
import numpy as np

try:  # O2 tier only
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None

F32 = np.float32


# --------------------------------------------------------------------------------------------
# activations used after GraphConv / GraphDense in the shipped models
# (example_model/model.py:43,45,50,53 sigmoid; sparse_infer.py:43-58 relu/tanh)
# --------------------------------------------------------------------------------------------
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3
ACT_IDS = {None: 0, "none": 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}


def activation(x, act):
    act = ACT_IDS[act] if not isinstance(act, int) else act
    x = np.asarray(x, F32)
    if act == ACT_NONE:
        return x
    if act == ACT_RELU:
        return np.maximum(x, F32(0))
    if act == ACT_SIGMOID:
        return (F32(1) / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)
    if act == ACT_TANH:
        return np.tanh(x, dtype=F32)
    raise ValueError(act)


def activation_grad_from_output(y, act):
    """d act / d pre-activation expressed through the activation's *output* y."""
    act = ACT_IDS[act] if not isinstance(act, int) else act
    y = np.asarray(y, F32)
    if act == ACT_NONE:
        return np.ones_like(y)
    if act == ACT_RELU:
        return (y > 0).astype(F32)
    if act == ACT_SIGMOID:
        return (y * (F32(1) - y)).astype(F32)
    if act == ACT_TANH:
        return (F32(1) - y * y).astype(F32)
    raise ValueError(act)


# --------------------------------------------------------------------------------------------
# tf.sparse_tensor_dense_matmul  (call sites kgcn/layers.py:113,393,468)
# --------------------------------------------------------------------------------------------
def _check_indices(indices, shape):
    indices = np.asarray(indices).reshape(-1, 2)
    if indices.size and (indices.min() < 0 or (indices[:, 0] >= shape[0]).any() or (indices[:, 1] >= shape[1]).any()):
        # TF raises InvalidArgumentError for an out-of-range index [TF-semantics, SURVEY 8c]
        raise IndexError("sparse index out of range for dense_shape %r" % (tuple(shape),))
    return indices


def sparse_dense_matmul(indices, values, dense_shape, b, adjoint_a=False):
    """O1 (faithful): zero-init output, walk the nnz **in storage order**,
    ``out[i, :] += v * b[j, :]`` in fp32; duplicates accumulate; indices need not be sorted."""
    shape = (int(dense_shape[0]), int(dense_shape[1]))
    indices = _check_indices(indices, shape)
    values = np.asarray(values, F32)
    b = np.asarray(b, F32)
    rows_out, inner = (shape[1], shape[0]) if adjoint_a else shape
    if b.shape[0] != inner:
        raise ValueError("inner dimension mismatch: A is %r (adjoint=%r), b is %r" % (shape, adjoint_a, b.shape))
    out = np.zeros((rows_out, b.shape[1]), F32)
    for e in range(indices.shape[0]):
        i, j = int(indices[e, 0]), int(indices[e, 1])
        if adjoint_a:
            i, j = j, i
        out[i] += values[e] * b[j]
    return out


def sparse_dense_matmul_fast(indices, values, dense_shape, b, adjoint_a=False):
    """O2 (fast): scipy CSR; differs from O1 only by summation order."""
    shape = (int(dense_shape[0]), int(dense_shape[1]))
    indices = _check_indices(indices, shape)
    a = _sp.coo_matrix((np.asarray(values, F32), (indices[:, 0], indices[:, 1])), shape=shape).tocsr()
    if adjoint_a:
        a = a.T.tocsr()
    return np.asarray(a @ np.asarray(b, F32), F32)


# --------------------------------------------------------------------------------------------
# GraphConv  (kgcn/layers.py:32-119; default branch :105-116)
# --------------------------------------------------------------------------------------------
def graph_conv(x, adjs, w, bias, fast=False):
    """``out[b] = add_n_c( spdm(adj[b][c], x[b] @ w[c] + bias[c]) )``; returns [B, N, F_out].

    x: [B, N, F_in]; adjs: list[B][C] of (indices[nnz,2], values[nnz], dense_shape[2]);
    w: list[C] of [F_in, F_out]; bias: list[C] of [1, F_out] (layers.py:53-61).
    The bias is added BEFORE aggregation (layers.py:112-113) so a row with no incident edge
    outputs exactly 0.
    """
    spdm = sparse_dense_matmul_fast if fast else sparse_dense_matmul
    x = np.asarray(x, F32)
    out = []
    for b in range(x.shape[0]):                      # layers.py:108
        acc = None
        for c in range(len(w)):                      # layers.py:109
            idx, val, shp = adjs[b][c][0], adjs[b][c][1], adjs[b][c][2]
            fw = (x[b] @ np.asarray(w[c], F32) + np.asarray(bias[c], F32).reshape(1, -1)).astype(F32)  # :112
            el = spdm(idx, val, shp, fw)             # :113
            acc = el if acc is None else (acc + el).astype(F32)   # tf.add_n, left to right (:115)
        out.append(acc)
    return np.stack(out)                             # :116


def graph_conv_grad(x, adjs, w, bias, dy, want_dvalues=False):
    """Backward of :func:`graph_conv` as TF autodiff produces it (gradient of
    SparseTensorDenseMatMul == bspmm_call.py:44-54): ``dfw = A^T dy``; ``dW_c = sum_b x_b^T dfw``;
    ``dbias_c = sum_b colsum(dfw)``; ``dx_b = sum_c dfw W_c^T``;
    ``dA_val[e] = <dy[row_e], fw[col_e]>``."""
    x = np.asarray(x, F32)
    dy = np.asarray(dy, F32)
    C = len(w)
    dw = [np.zeros_like(np.asarray(w[c], F32)) for c in range(C)]
    db = [np.zeros((1, np.asarray(w[c]).shape[1]), F32) for c in range(C)]
    dx = np.zeros_like(x)
    dvals = []
    for b in range(x.shape[0]):
        row = []
        for c in range(C):
            idx, val, shp = adjs[b][c][0], adjs[b][c][1], adjs[b][c][2]
            dfw = sparse_dense_matmul(idx, val, shp, dy[b], adjoint_a=True)
            dw[c] += x[b].T @ dfw
            db[c] += dfw.sum(axis=0, keepdims=True, dtype=F32)
            dx[b] += dfw @ np.asarray(w[c], F32).T
            if want_dvalues:
                fw = (x[b] @ np.asarray(w[c], F32) + np.asarray(bias[c], F32).reshape(1, -1)).astype(F32)
                ii = np.asarray(idx).reshape(-1, 2)
                row.append((dy[b][ii[:, 0]] * fw[ii[:, 1]]).sum(axis=1, dtype=F32))
        dvals.append(row)
    return (dx, dw, db, dvals) if want_dvalues else (dx, dw, db)


# --------------------------------------------------------------------------------------------
# GINAggregate default branch (kgcn/layers.py:459-471) and BatchGraphConv (layers.py:388-395)
# --------------------------------------------------------------------------------------------
def gin_aggregate(x, adjs, epsilon):
    x = np.asarray(x, F32)
    out = []
    for b in range(x.shape[0]):
        acc = None
        for c in range(len(epsilon)):
            el = sparse_dense_matmul(adjs[b][c][0], adjs[b][c][1], adjs[b][c][2], x[b])
            term = (F32(epsilon[c]) * x[b] + el).astype(F32)
            acc = term if acc is None else (acc + term).astype(F32)
        out.append(acc)
    return np.stack(out)


def batch_graph_conv(x, adj, w, bias):
    """Block-diagonal form: ``relu(A (x W + b))`` with x [sumN, F_in] (layers.py:388-395)."""
    net = (np.asarray(x, F32) @ np.asarray(w, F32) + np.asarray(bias, F32).reshape(1, -1)).astype(F32)
    net = sparse_dense_matmul(adj[0], adj[1], adj[2], net)
    return np.maximum(net, F32(0))


# --------------------------------------------------------------------------------------------
# GraphDense (kgcn/layers.py:223-265) and GraphGather (layers.py:156-167)
# --------------------------------------------------------------------------------------------
def graph_dense(x, kernel, bias=None, act=None, enabled_node_nums=None):
    """Keras Dense = ``act(x K + bias)``.  Without ``enabled_node_nums`` it is applied to ALL
    B*N rows, padding included (layers.py:255-262).  With it, only the first n_b rows of each
    graph go through Dense and the rest are exactly zero (layers.py:243-254)."""
    x = np.asarray(x, F32)
    B, N, _ = x.shape
    kernel = np.asarray(kernel, F32)
    flat = x.reshape(-1, x.shape[2]) @ kernel
    if bias is not None:
        flat = flat + np.asarray(bias, F32).reshape(1, -1)
    out = activation(flat.astype(F32), act).reshape(B, N, kernel.shape[1])
    if enabled_node_nums is not None:
        n = np.asarray(enabled_node_nums).reshape(-1)
        keep = (np.arange(N)[None, :] < n[:, None])
        out = np.where(keep[:, :, None], out, F32(0)).astype(F32)
    return out


def graph_gather(x):
    """``tf.reduce_sum(inputs, axis=1)`` over all N rows, padding included (layers.py:164)."""
    x = np.asarray(x, F32)
    out = np.zeros((x.shape[0], x.shape[2]), F32)
    for i in range(x.shape[1]):
        out += x[:, i, :]
    return out


# --------------------------------------------------------------------------------------------
# Plugin-op contracts (absent .so's): Bspmm / Bconv / Bspmdt forward + registered gradients
# --------------------------------------------------------------------------------------------
def bspmm(sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False):
    """bspmm_call.py:10-15: N independent ``A_t . B_t``."""
    out = []
    for a, b in zip(sp_matrices, dense_matrices):
        b = np.asarray(b, F32)
        out.append(sparse_dense_matmul(a[0], a[1], a[2], b.T if adjoint_b else b, adjoint_a=adjoint_a))
    return out


def bspmm_grad(sp_matrices, dense_matrices, grads, adjoint_a=False, adjoint_b=False):
    """bspmm_call.py:21-57.  Returns (a_values_grads list, b_grads list)."""
    b_grads = bspmm(sp_matrices, grads, adjoint_a=True, adjoint_b=False)       # :44 (as written)
    if adjoint_b:
        b_grads = [g.T for g in b_grads]                                       # :46-47
    a_values_grads = []
    for t, (a, b) in enumerate(zip(sp_matrices, dense_matrices)):
        idx = np.asarray(a[0]).reshape(-1, 2)
        rows, cols = idx[:, 0], idx[:, 1]
        g = np.asarray(grads[t], F32)
        bb = np.asarray(b, F32)
        bb = bb.T if adjoint_b else bb
        pa = g[rows if not adjoint_a else cols]                                # :52
        pb = bb[cols if not adjoint_a else rows]                               # :53
        a_values_grads.append((pa * pb).sum(axis=1, dtype=F32))                # :54
    return a_values_grads, b_grads


def bconv(sp_matrices, dense_matrices, adjoint_a=False, adjoint_b=False):
    """bconv_call.py:10-21: ``out_b = sum_c A[b][c] . B[b][c]`` (channel sum fused)."""
    out = []
    for a_row, d_row in zip(sp_matrices, dense_matrices):
        parts = bspmm(a_row, d_row, adjoint_a=adjoint_a, adjoint_b=adjoint_b)
        acc = parts[0]
        for p in parts[1:]:
            acc = (acc + p).astype(F32)
        out.append(acc)
    return out


def bconv_grad(sp_matrices, dense_matrices, grads, adjoint_a=False, adjoint_b=False):
    """bconv_call.py:28-70 (_bconv_grad).  ``sp_matrices`` / ``dense_matrices`` are list[B][C], ``grads`` list[B] (one
    per output).  The incoming gradient of output b is replicated for its C channels (:46, batch-major / channel-minor
    like the inputs), the dense gradients are ONE Bspmm over all B*C matrices with ``adjoint_a = not adj_a`` (:57), and
    the value gradients gather rows of dY and of the dense operand per stored entry (:62-67).
    Returns (a_values_grads list[B*C], b_grads list[B*C]) in the flat order of the op's inputs."""
    C = len(sp_matrices[0])
    flat_sp = [a for row in sp_matrices for a in row]
    flat_d = [d for row in dense_matrices for d in row]
    addn_grad = [g for g in grads for _ in range(C)]                           # :46
    b_grads = bspmm(flat_sp, addn_grad, adjoint_a=not adjoint_a, adjoint_b=adjoint_b)   # :57
    if adjoint_b:
        b_grads = [g.T for g in b_grads]                                       # :59-60
    a_values_grads = []
    for t, (a, b) in enumerate(zip(flat_sp, flat_d)):
        idx = np.asarray(a[0]).reshape(-1, 2)
        rows, cols = idx[:, 0], idx[:, 1]
        g = np.asarray(addn_grad[t], F32)
        bb = np.asarray(b, F32)
        bb = bb.T if adjoint_b else bb
        pa = g[rows if not adjoint_a else cols]                                # :65
        pb = bb[cols if not adjoint_a else rows]                               # :66
        a_values_grads.append((pa * pb).sum(axis=1, dtype=F32))                # :67
    return a_values_grads, b_grads


def bspmdt(sp_matrices, dense, adjoint_a=False, adjoint_b=False):
    """batched_call.py:21-26: N COO matrices times one stacked dense [N*rows, cols]."""
    n = len(sp_matrices)
    dense = np.asarray(dense, F32)
    rows = dense.shape[0] // n
    return bspmm(sp_matrices, [dense[t * rows:(t + 1) * rows] for t in range(n)],
                 adjoint_a=adjoint_a, adjoint_b=adjoint_b)


# --------------------------------------------------------------------------------------------
# The measured network: L x [GraphConv -> act] (-> GraphDense -> act) -> GraphGather -> Dense
# -> softmax cross-entropy with batch mask.  Follows example_model/model.py:41-69 minus the
# BatchNormalization/Dropout rows (both are inference-mode identities up to
# a constant scale under the reference trainer, Appendix A.10).
# --------------------------------------------------------------------------------------------
def _dense_adj(triple):
    idx, val, shape = triple
    a = np.zeros(tuple(int(v) for v in shape), F32)
    idx = np.asarray(idx).reshape(-1, 2)
    a[idx[:, 0], idx[:, 1]] = np.asarray(val, F32)      # sparse_tensor_to_dense: unique, sorted indices only
    return a


def graph_max_pooling(x, adjs):
    """GraphMaxPooling (kgcn/layers.py:122-153), literally: for every molecule b, channel c and feature k
    ``d = to_dense(A[b][c] * x[b, :, k])`` (``A * v`` scales column j by v[j]; absent entries stay 0, :143-144),
    ``el = reduce_max(d, axis=1)`` (:145), features stacked (:148), channels added (:149)."""
    x = np.asarray(x, F32)
    B, N, F = x.shape
    out = np.zeros((B, N, F), F32)
    for b in range(B):
        for c in range(len(adjs[b])):
            a = _dense_adj(adjs[b][c])
            for k in range(F):
                d = (a * x[b, :, k][None, :]).astype(F32)
                out[b, :, k] += d.max(axis=1)
    return out


def graph_batch_normalization(x, gamma, beta, mean, var, enabled_node_nums=None, eps=1e-3, batch_statistics=False):
    """GraphBatchNormalization (kgcn/layers.py:186-219): the first enabled_node_nums[b] rows of every molecule are
    stacked (:202-204), normalised per feature (:205; moving statistics as the Keras layer does under the
    reference trainer, or the batch's own statistics for the legacy tf.layers variant, legacy/layers.py:202),
    split and zero-padded back (:206-211).  Returns (y, mean used, biased variance used)."""
    x = np.asarray(x, F32)
    B, N, F = x.shape
    n = np.full(B, N, np.int64) if enabled_node_nums is None else np.asarray(enabled_node_nums, np.int64)
    rows = np.concatenate([x[b, :n[b]] for b in range(B)], 0).astype(np.float64)
    if batch_statistics:
        mean = rows.mean(0) if rows.shape[0] else np.zeros(F)
        var = rows.var(0) if rows.shape[0] else np.zeros(F)
    mean, var = np.asarray(mean, np.float64), np.asarray(var, np.float64)
    y = np.zeros((B, N, F), F32)
    for b in range(B):
        y[b, :n[b]] = ((x[b, :n[b]].astype(np.float64) - mean) / np.sqrt(var + eps) * np.asarray(gamma, np.float64)
                       + np.asarray(beta, np.float64)).astype(F32)
    return y, mean.astype(F32), var.astype(F32)


def segment_sum(x, sizes):
    """The per-molecule readout of the block-diagonal model (example_model/sparse.py:79-90)."""
    x = np.asarray(x, F32)
    out, pos = [], 0
    for s in sizes:
        acc = np.zeros(x.shape[1], F32)
        for i in range(int(s)):
            acc = acc + x[pos + i]
        out.append(acc)
        pos += int(s)
    return np.stack(out) if out else np.zeros((0, x.shape[1]), F32)


def glorot_uniform(rng, fan_in, fan_out):
    lim = np.sqrt(6.0 / (fan_in + fan_out))       # layers.py:54-57 'glorot_uniform'
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(F32)


def init_network(rng, feature_dim, conv_dims, channels, label_dim, dense_dim=None):
    p = {"conv_w": [], "conv_b": []}
    f = feature_dim
    for d in conv_dims:
        p["conv_w"].append([glorot_uniform(rng, f, d) for _ in range(channels)])
        # the reference initialises biases to zero (layers.py:58-61); small random values here
        # make every parity test sensitive to the bias-times-degree rule.
        p["conv_b"].append([rng.uniform(-0.1, 0.1, size=(1, d)).astype(F32) for _ in range(channels)])
        f = d
    if dense_dim:
        p["gd_w"] = glorot_uniform(rng, f, dense_dim)
        p["gd_b"] = rng.uniform(-0.1, 0.1, size=(dense_dim,)).astype(F32)
        f = dense_dim
    p["out_w"] = glorot_uniform(rng, f, label_dim)
    p["out_b"] = rng.uniform(-0.1, 0.1, size=(label_dim,)).astype(F32)
    return p


def network_forward(p, x, adjs, labels, mask, act="sigmoid", fast=True, keep=False):
    """Returns dict(logits, prediction, cost_opt, cost_sum, correct_count[, acts])."""
    h = np.asarray(x, F32)
    acts = [h]
    for w, b in zip(p["conv_w"], p["conv_b"]):
        h = activation(graph_conv(h, adjs, w, b, fast=fast), act)
        acts.append(h)
    if "gd_w" in p:
        h = graph_dense(h, p["gd_w"], p["gd_b"], act=act)
        acts.append(h)
    g = graph_gather(h)
    logits = (g @ p["out_w"] + p["out_b"][None, :]).astype(F32)
    z = logits - logits.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z, dtype=F32).sum(axis=1, keepdims=True, dtype=F32))
    logp = (z - lse).astype(F32)
    labels = np.asarray(labels, F32)
    mask = np.asarray(mask, F32)
    cost = (mask * -(labels * logp).sum(axis=1, dtype=F32)).astype(F32)     # model.py:58-60
    pred = np.exp(logp, dtype=F32)
    out = {
        "logits": logits, "prediction": pred,
        "cost_opt": F32(cost.mean(dtype=F32)),                              # model.py:61
        "cost_sum": F32(cost.sum(dtype=F32)),                               # model.py:64
        "correct_count": F32((mask * (pred.argmax(1) == labels.argmax(1))).sum(dtype=F32)),  # :66-69
    }
    if keep:
        out["acts"], out["gathered"] = acts, g
    return out


def network_grad(p, x, adjs, labels, mask, act="sigmoid"):
    """Gradients of ``cost_opt`` w.r.t. every parameter (what AdamOptimizer.minimize sees,
    core.py:121-127).  Returns (forward dict, grads dict shaped like ``p``)."""
    fw = network_forward(p, x, adjs, labels, mask, act=act, fast=True, keep=True)
    acts, g = fw["acts"], fw["gathered"]
    labels = np.asarray(labels, F32)
    mask = np.asarray(mask, F32)
    B = labels.shape[0]
    # d mean(mask * xent) / d logits = mask/B * (softmax * sum(labels) - labels)
    dlogits = ((mask / F32(B))[:, None] * (fw["prediction"] * labels.sum(axis=1, keepdims=True) - labels)).astype(F32)
    grads = {"out_w": (g.T @ dlogits).astype(F32), "out_b": dlogits.sum(axis=0, dtype=F32)}
    dg = (dlogits @ p["out_w"].T).astype(F32)
    h = acts[-1]
    dh = np.broadcast_to(dg[:, None, :], h.shape).astype(F32)               # grad of reduce_sum(axis=1)
    k = len(acts) - 1
    if "gd_w" in p:
        du = (dh * activation_grad_from_output(h, act)).astype(F32)
        hin = acts[k - 1]
        grads["gd_w"] = (hin.reshape(-1, hin.shape[2]).T @ du.reshape(-1, du.shape[2])).astype(F32)
        grads["gd_b"] = du.reshape(-1, du.shape[2]).sum(axis=0, dtype=F32)
        dh = (du @ p["gd_w"].T).astype(F32)
        k -= 1
        h = acts[k]
    grads["conv_w"], grads["conv_b"] = [None] * len(p["conv_w"]), [None] * len(p["conv_w"])
    for layer in range(len(p["conv_w"]) - 1, -1, -1):
        du = (dh * activation_grad_from_output(acts[layer + 1], act)).astype(F32)
        dx, dw, db = graph_conv_grad(acts[layer], adjs, p["conv_w"][layer], p["conv_b"][layer], du)
        grads["conv_w"][layer], grads["conv_b"][layer] = dw, db
        dh = dx
    grads["x"] = dh
    return fw, grads
