"""Host-side ingest: ``.jbl``-style dictionaries -> per-graph / per-channel COO adjacency lists.

Mirror of ``kgcn/data_util.py`` (clinfo/kGCN @ 32328d5) for the graph path -- the adjacency builders,
``build_data`` / ``load_data`` (``all_data`` members and ``info`` fields), shuffling and the train / validation
split: same function names, same outputs bit for bit (checked against vectors produced by the reference's own code,
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


class dotdict(dict):
    """``d.key`` access with ``None`` for a missing key -- the container ``build_data`` returns (data_util.py:14-18)."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__


# per-sample members of ``all_data`` that shuffling / splitting permutes (data_util.py:155-179, 622-643)
_PER_SAMPLE = ("features", "nodes", "adjs", "labels", "mask_label", "node_label", "mask_node_label", "label_list", "sequences",
               "sequences_vec", "sequences_vec_range", "sequences_len", "enabled_node_nums")
_MODAL_NAMES = ("vector_modal", "profeat", "dragon", "chemical_fp")
_POS_WEIGHT_EPS = 0.01


def _say(verbose, *args):
    if verbose:
        print(*args)


def build_data(config, data, prohibit_shuffle=False, verbose=True, test_mode=False):
    """``.jbl`` dictionary -> ``(all_data, info)`` with the members and ``info`` fields of ``kgcn/data_util.py:374-561``
    (what ``kgcn/core.py``, ``kgcn/feed.py`` and the model files read).  The adjacency block is :func:`build_adjs`."""
    features = data["feature"] if ("feature" in data and config["with_feature"]) else None
    if features is not None and len(features) == 0:
        features = None
    nodes = np.array(data["node"], np.int32) if ("node" in data and config["with_node_embedding"]) else None
    if nodes is not None and len(nodes) == 0:
        nodes = None
    try:
        adjs, enabled_node_nums, adj_channel_num = build_adjs(data, config)
    except DataLoadError:
        _say(verbose, "[INFO] no graph")
        adjs, enabled_node_nums, adj_channel_num = None, None, 1

    labels = data.get("label")
    mask_label = data.get("mask_label")
    if "label_sparse" in data:
        labels = np.array(data["label_sparse"].todense())
    if "mask_label_sparse" in data:
        mask_label = np.array(data["mask_label_sparse"].todense())
    label_list = None
    if "label_list" in data:
        label_list = data["test_label_list"] if test_mode else data["label_list"]
    vector_modal, vector_modal_name = [], {}
    for name in _MODAL_NAMES:
        if name in data:
            vector_modal_name[name] = len(vector_modal)
            vector_modal.append(data[name])

    all_data = dotdict(
        features=features, nodes=nodes, adjs=adjs, labels=np.array(labels) if labels is not None else None,
        mask_label=mask_label, node_label=data.get("node_label"), mask_node_label=data.get("mask_node_label"),
        label_list=label_list, num=len(adjs) if adjs is not None else max(len(v) for v in vector_modal),
        sequences=data.get("sequence"), sequences_vec=data.get("sequence_vec"), sequences_vec_range=data.get("sequence_vec_range"),
        sequences_len=np.array(data["sequence_length"], np.int32) if "sequence" in data else None,
        sequence_symbol=np.array(data["sequence_symbol"]) if "sequence_symbol" in data else None,
        vector_modal=vector_modal, enabled_node_nums=enabled_node_nums)
    if config["shuffle_data"] and not prohibit_shuffle:
        _say(verbose, "[INFO] data_shuffle is done")
        all_data = shuffle_data(all_data)

    info = dotdict(all_node_num=None)
    if features is not None:      # features: #graphs x #nodes(graph) x #features
        info.feature_dim, info.graph_node_num, info.feature_enabled = features.shape[2], features.shape[1], True
    elif nodes is not None:       # nodes: #graphs x #nodes(graph)
        info.feature_dim, info.graph_node_num, info.feature_enabled = 0, nodes.shape[1], False
        info.all_node_num = data["node_num"]
    elif adjs is not None:
        raise DataLoadError("feature or node are required: please confirm input data and configuration")
    sequences, sequences_vec = all_data.sequences, all_data.sequences_vec
    info.sequence_max_length = sequences.shape[1] if sequences is not None else 0
    info.sequence_symbol_num = data["sequence_symbol_num"] if sequences is not None else 0
    info.sequences_vec_dim = 0
    if sequences_vec is not None:
        info.sequence_max_length, info.sequences_vec_dim = sequences_vec.shape[1], sequences_vec.shape[2]
    if all_data.sequences_vec_range is not None:
        info.sequences_vec_dim = len(data["sequence_vec_name"])
    info.graph_num = len(adjs) if adjs is not None else 0
    info.adj_channel_num = adj_channel_num
    if labels is not None:
        shape = np.shape(labels)
        info.label_dim = data["label_dim"] if "label_dim" in data else (shape[1] if len(shape) >= 2 else 1)
        expected = info.graph_num if adjs is not None else all_data.num
        if shape[0] != expected:
            _say(True, "[ERROR] %d labels for %d samples" % (shape[0], expected))
    elif all_data.node_label is not None:      # node_label: graph_num x node_num x label_dim
        info.label_dim = np.shape(all_data.node_label)[2]
        _say(verbose, "[INFO] node centric mode")
    else:
        info.label_dim = data.get("label_dim")
    if (features is not None and features.shape[0] != info.graph_num) or (nodes is not None and nodes.shape[0] != info.graph_num):
        raise DataLoadError("the numbers of feature matrices, node lists and adjacency matrices differ: please confirm input data")
    _say(verbose, "[OK] checking #graphs")
    info.vector_modal_dim = [modal.shape[1] for modal in vector_modal]
    info.vector_modal_name = vector_modal_name
    info.graph_index_list = data.get("graph_index_list")
    if all_data.mask_label is not None and all_data.labels is not None:      # class balance for weighted losses
        positive = np.nansum(all_data.labels, axis=0)
        negative = np.nansum(all_data.mask_label, axis=0) - positive
        info.pos_weight = (negative + _POS_WEIGHT_EPS) / (positive + _POS_WEIGHT_EPS)
    if "class_weight" in data:
        info.class_weight = data["class_weight"]
    elif all_data.labels is not None:      # labels: #data x #class
        info.class_weight = (np.nansum(all_data.labels) + _POS_WEIGHT_EPS) / (np.nansum(all_data.labels, axis=0) + _POS_WEIGHT_EPS)
    if "mol_info" in data:
        info.mol_info = data["mol_info"]
    _say(verbose, "graphs=%s feature_dim=%s graph_node_num=%s all_node_num=%s label_dim=%s adj_channel_num=%s"
         % (info.graph_num, info.feature_dim, info.graph_node_num, info.all_node_num, info.label_dim, info.adj_channel_num))
    return all_data, info


def load_data(config, filename="data.jbl", prohibit_shuffle=False, test_mode=False, verbose=True):
    """``joblib.load`` + :func:`build_data` (data_util.py:368-371)."""
    import joblib
    _say(verbose, "[LOAD]", filename)
    return build_data(config, joblib.load(filename), prohibit_shuffle=prohibit_shuffle, test_mode=test_mode, verbose=verbose)


def shuffle_data(data):
    """One permutation from numpy's GLOBAL generator applied to every per-sample member (data_util.py:155-179): the
    same ``np.random.seed`` gives the same order as the reference."""
    idx = list(range(data.num))
    np.random.shuffle(idx)
    if data.adjs is not None:
        data.adjs = _object_array(data.adjs)
    for key in _PER_SAMPLE:
        if data[key] is not None:
            data[key] = data[key][idx]
    if data.vector_modal is not None:
        data.vector_modal = [m[idx] for m in data.vector_modal]
    return data


def _object_array(items):
    """``np.array(list of per-graph channel lists)`` without numpy trying to broadcast ragged triples."""
    out = np.empty(len(items), dtype=object)
    for i, item in enumerate(items):
        out[i] = item
    return out


def _take(value, indices):
    if isinstance(value, np.ndarray):
        return value[indices]
    picked = [value[i] for i in indices]
    try:
        return np.array(picked)
    except ValueError:          # ragged per-graph adjacency lists: keep them as an object array
        return _object_array(picked)


def split_data(all_data, valid_data_rate=0.2, indices_for_train_data=None, indices_for_valid_data=None):
    """Train / validation split of ``all_data`` (data_util.py:597-644): ``int(num * rate)`` validation samples taken
    from the tail of one global-generator shuffle of ``arange(num)``, unless both index lists are given."""
    if all_data.get("label_list") is not None:
        return split_label_list(all_data, valid_data_rate, indices_for_train_data, indices_for_valid_data)
    if indices_for_train_data is None or indices_for_valid_data is None:
        valid_num = int(all_data.num * valid_data_rate)
        indices = np.arange(all_data.num)
        np.random.shuffle(indices)
        indices_for_train_data, indices_for_valid_data = indices[:all_data.num - valid_num], indices[all_data.num - valid_num:]
    train_data, valid_data = dotdict(), dotdict()
    for key in all_data.keys() - {"sequence_symbol", "num"}:
        value = all_data[key]
        for part, indices in ((train_data, indices_for_train_data), (valid_data, indices_for_valid_data)):
            if value is None:
                part[key] = None
            elif key == "vector_modal":
                part[key] = [np.array([modal[i] for i in indices]) for modal in value]
            else:
                part[key] = _take(value, indices)
    train_data.num, valid_data.num = len(indices_for_train_data), len(indices_for_valid_data)
    return train_data, valid_data


def split_label_list(all_data, valid_data_rate=0.2, indices_for_train_data=None, indices_for_valid_data=None):
    """Node / edge prediction data: only ``label_list [tasks, items, fields]`` is split, along its second axis
    (data_util.py:662-695)."""
    if indices_for_train_data is None or indices_for_valid_data is None:
        n = len(all_data.label_list[0])
        valid_num = int(n * valid_data_rate)
        nid = np.array(list(range(n)))
        np.random.shuffle(nid)
        indices_for_train_data, indices_for_valid_data = nid[:n - valid_num], nid[n - valid_num:]
    train_data, valid_data = dotdict(all_data), dotdict(all_data)
    train_data["label_list"] = all_data["label_list"][:, indices_for_train_data, :]
    valid_data["label_list"] = all_data["label_list"][:, indices_for_valid_data, :]
    return train_data, valid_data


def load_and_split_data(config, filename="data.jbl", valid_data_rate=0.2):
    all_data, info = load_data(config, filename)
    return (all_data,) + split_data(all_data, valid_data_rate) + (info,)


def build_and_split_data(config, data, valid_data_rate=0.2):
    all_data, info = build_data(config, data)
    return (all_data,) + split_data(all_data, valid_data_rate) + (info,)


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
