"""Host-side ingest: ``.jbl``-style dictionaries -> per-graph / per-channel COO adjacency lists.

Mirror of the adjacency half of ``kgcn/data_util.py`` (clinfo/kGCN @ 32328d5): same function names,
same outputs bit for bit (checked against vectors produced by the reference's own code,
``tests/golden/ingest_*.npz``), rewritten on vectorised numpy.  One-time preprocessing, so it
stays on the host; the per-step work (batching + CSR packing + upload) is in ``feed.py``/``csr.py``.
"""
import numpy as np


class DataLoadError(Exception):
    """Raised for inputs the loader cannot use (kgcn/data_util.py:143-152)."""


def dense_to_sparse(dense):
    """dense [R, C] -> (indices [nnz,2] int32 row-major sorted, values float32, shape) -- data_util.py:40-45."""
    dense = np.asarray(dense)
    r, c = np.nonzero(dense)
    idx = np.stack([r, c], axis=1).astype(np.int32).reshape(-1, 2)
    return idx, dense[r, c].astype(np.float32), np.array(dense.shape)


def align_size(adjs, max_n):
    """Relabel every matrix as [max_n, max_n]; indices are untouched (data_util.py:30-37)."""
    for i in range(len(adjs)):
        for ch in range(len(adjs[i])):
            a = list(adjs[i][ch])
            a[2] = [max_n, max_n]
            adjs[i][ch] = a


def check_adj(adj):
    """True if ``adj`` is a single (indices, values, shape) triple rather than a list of them (data_util.py:48-55)."""
    try:
        return len(adj) == 3 and len(adj[2]) == 2 and not isinstance(adj[2][0], (np.ndarray, list))
    except TypeError:
        return False


def _to_dense(adj, dtype):
    idx = np.asarray(adj[0]).reshape(-1, 2)
    out = np.zeros((int(adj[2][0]), int(adj[2][1])), dtype)
    np.add.at(out, (idx[:, 0], idx[:, 1]), np.asarray(adj[1]).astype(dtype))
    return out


def high_order_adj(adj, order):
    """Pattern of A^order with all values 1, (row, col)-sorted (data_util.py:58-73)."""
    if order <= 1:
        return adj
    a = _to_dense(adj, np.float64)
    b = a
    for _ in range(order - 1):
        b = b @ a
    r, c = np.nonzero(b)
    return (np.stack([r, c], 1).astype(np.int32).reshape(-1, 2), np.ones(r.shape[0], np.float32),
            np.array(b.shape, np.int64))


def split_adj(adjs, min_deg=1, max_deg=5):
    """Degree-split channels: per input channel emit (max_deg-min_deg+1) degree channels plus one
    self-loop channel; every channel starts with a dummy ([0,0], 0.0) entry unless its first real
    entry is [0,0] (data_util.py:76-122)."""
    n_ch = (max_deg - min_deg + 1) + 1
    self_ch = n_ch - 1
    for gid, adj_set in enumerate(adjs):
        out = []
        for adj in adj_set:
            idx = np.asarray(adj[0]).reshape(-1, 2)
            val = np.asarray(adj[1]).astype(np.float32)
            shape = adj[2]
            deg = np.bincount(idx[:, 0], minlength=int(shape[0]))
            ch_of_row = np.minimum(deg, max_deg) - min_deg
            entry_ch = np.where(idx[:, 0] == idx[:, 1], self_ch, ch_of_row[idx[:, 0]])
            for k in range(n_ch):
                sel = entry_ch == k
                e, v = idx[sel], val[sel]
                if e.shape[0] == 0 or not (e[0, 0] == 0 and e[0, 1] == 0):
                    e = np.concatenate([np.zeros((1, 2), idx.dtype), e], 0)
                    v = np.concatenate([np.zeros((1,), np.float32), v], 0)
                out.append([e.astype(np.int32), v.astype(np.float32), shape])
        adjs[gid] = out
    return adjs


def normalize_adj(adjs):
    """Kipf-style D^-1/2 A D^-1/2 with column-sum degrees of the binarised pattern
    (data_util.py:125-140).  Bit-exact with the reference as it runs on scipy >= 1.8: there
    ``A_tilde / sqrt(d)[:,None] / sqrt(d)`` is evaluated as two sparse multiplies by fp32
    reciprocals, which keeps the COO entry order, duplicates and explicit zeros (e.g. the
    ``[0,0] -> 0.0`` dummies of ``split_adj``)."""
    out = []
    for adj_set in adjs:
        row = []
        for adj in adj_set:
            idx = np.asarray(adj[0]).reshape(-1, 2)
            val = np.array(adj[1])
            val[val > 0] = 1
            n_r, n_c = int(adj[2][0]), int(adj[2][1])
            deg = np.zeros(n_c, val.dtype)
            np.add.at(deg, idx[:, 1], val)
            deg[deg == 0] = 1
            recip = np.true_divide(1.0, np.sqrt(deg))
            v = (val * recip[idx[:, 0]]) * recip[idx[:, 1]] if n_r == n_c else None
            if v is None:
                raise DataLoadError("normalize_adj needs square adjacency matrices")
            row.append((idx.astype(np.int32), np.asarray(v, np.float32), np.array([n_r, n_c])))
        out.append(row)
    return out


def build_adjs(data, config):
    """The adjacency block of ``build_data`` (data_util.py:396-424).

    Returns (adjs list[G][C], enabled_node_nums int32[G], adj_channel_num)."""
    order = config.get("order", 1)
    if "multi_dense_adj" not in data:
        if "adj" in data:
            adjs = list(data["adj"])
        elif "dense_adj" in data:
            adjs = [dense_to_sparse(m) for m in data["dense_adj"]]
        else:
            raise DataLoadError("adj or dense_adj are required for GCN")
        max_n = int(data["max_node_num"])
        if check_adj(adjs[0]):
            adjs = [[high_order_adj(a, o) for o in range(1, order + 1)] for a in adjs]
        enabled = [a[0][2][0] for a in adjs]
        align_size(adjs, max_n)
    else:
        enabled = [max(len(m) for m in mats) for mats in data["multi_dense_adj"]]
        adjs = [[dense_to_sparse(m) for m in mats] for mats in data["multi_dense_adj"]]
    if config.get("split_adj_flag", False):
        adjs = split_adj(adjs)
    if config.get("normalize_adj_flag", False):
        adjs = normalize_adj(adjs)
    return adjs, np.array(enabled, dtype=np.int32), len(adjs[0])


def construct_batched_adjacency_and_feature_matrices(size, adj_row, adj_column, adj_values, adj_elem_len, adj_degrees,
                                                     feature_row, feature_column, feature_values, feature_elem_len,
                                                     input_dim, max_degree=5, normalize=True, split_adj=False):
    """Block-diagonal batch of the reference's tfrecords path (kgcn/data_util.py:698-845), as host numpy.

    The reference builds this with a tf.scan per batch; here it is two prefix sums.  ``size[m]`` nodes per
    molecule, ``adj_elem_len[m]`` stored entries per molecule, rows / columns local to their molecule.  Returns
    ``(channels, features)``: ``channels`` is a list of ``(indices int64 [nnz, 2], values float32 [nnz],
    [n, n])`` triples over the ``n = sum(size)`` rows of the batch (one triple; ``max_degree + 1`` with
    ``split_adj``: entries whose ``adj_degrees`` clipped to ``[0, max_degree]`` equals 1..max_degree, then the
    identity, :803-821), ``features`` the dense ``[n, input_dim]`` matrix of the stacked sparse feature entries
    (:832-845).  ``normalize`` (:790-802): ``A[i, j] / sqrt(d[j]) / sqrt(d[i])`` with ``d`` = column sums, in
    float32 like the TensorFlow ops."""
    size = np.asarray(size, np.int64).reshape(-1)
    adj_elem_len = np.asarray(adj_elem_len, np.int64).reshape(-1)
    adj_row = np.asarray(adj_row, np.int64).reshape(-1)
    adj_column = np.asarray(adj_column, np.int64).reshape(-1)
    adj_values = np.asarray(adj_values, np.float32).reshape(-1)
    if size.shape != adj_elem_len.shape or adj_elem_len.sum() != adj_row.shape[0] or adj_row.shape != adj_column.shape:
        raise DataLoadError("block-diagonal batch: size / adj_elem_len / adj_row / adj_column are inconsistent")
    n = int(size.sum())
    offset = np.cumsum(size) - size                       # tf.cumsum(size, exclusive=True), :761
    entry_off = np.repeat(offset, adj_elem_len)            # every entry of molecule m is shifted by offset[m], :764-789
    diagonal_row, diagonal_col = adj_row + entry_off, adj_column + entry_off
    if (adj_row < 0).any() or (adj_column < 0).any() or (adj_row >= np.repeat(size, adj_elem_len)).any() or \
            (adj_column >= np.repeat(size, adj_elem_len)).any():
        raise DataLoadError("block-diagonal batch: adjacency index outside its molecule")
    shape = [n, n]
    if normalize:
        degree_hat = np.zeros(n, np.float32)
        np.add.at(degree_hat, diagonal_col, adj_values)    # tf.sparse.reduce_sum(axis=0): column sums
        root = np.sqrt(degree_hat, dtype=np.float32)
        values = ((adj_values / root[diagonal_col]).astype(np.float32) / root[diagonal_row]).astype(np.float32)
        channels = [(np.stack([diagonal_row, diagonal_col], 1), values, shape)]
    elif split_adj:
        deg = np.clip(np.asarray(adj_degrees, np.int64).reshape(-1), 0, max_degree)
        channels = []
        for degree in range(1, max_degree + 1):
            keep = deg == degree
            channels.append((np.stack([diagonal_row[keep], diagonal_col[keep]], 1), adj_values[keep], shape))
        eye = np.arange(n, dtype=np.int64)
        channels.append((np.stack([eye, eye], 1), np.ones(n, np.float32), shape))   # connection to self, :821
    else:
        channels = [(np.stack([diagonal_row, diagonal_col], 1), adj_values, shape)]

    feature_elem_len = np.asarray(feature_elem_len, np.int64).reshape(-1)
    feature_row = np.asarray(feature_row, np.int64).reshape(-1)
    feature_column = np.asarray(feature_column, np.int64).reshape(-1)
    feature_values = np.asarray(feature_values).reshape(-1)
    stacked_row = feature_row + np.repeat(offset, feature_elem_len)
    if (feature_column < 0).any() or (feature_column >= input_dim).any() or (stacked_row >= n).any():
        raise DataLoadError("block-diagonal batch: feature index out of range")
    flat = stacked_row * int(input_dim) + feature_column
    if np.unique(flat).shape[0] != flat.shape[0]:      # tf.sparse_tensor_to_dense rejects repeated indices
        raise DataLoadError("block-diagonal batch: repeated feature index")
    features = np.zeros((n, int(input_dim)), feature_values.dtype)
    features[stacked_row, feature_column] = feature_values
    return channels, features
