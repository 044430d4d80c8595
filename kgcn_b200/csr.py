"""BatchedCSR: the device-resident form of a batch of kGCN adjacency matrices.

The reference feeds ``batch_size * adj_channel_num`` separate ``SparseTensorValue`` triples per step
(kgcn/default_model.py:10, kgcn/feed.py:112-126).  Here the same list-of-lists becomes three flat
device arrays (graph-major / channel-minor, the flatten order of kgcn/bconv_call.py:11-15) plus
the transposed copy that ``adjoint_a=True`` (kgcn/bspmm_call.py:44) needs.  Layout: DESIGN.md
section 3 and include/kgcn_b200.h.
"""
import numpy as np
import torch

from . import _lib


def _triple(sp):
    """Accept SparseTensorValue-like objects, (indices, values, dense_shape) tuples/lists."""
    if hasattr(sp, "indices") and hasattr(sp, "dense_shape"):
        return sp.indices, sp.values, sp.dense_shape
    return sp[0], sp[1], sp[2]


def flatten_coo(adjs):
    """list[B][C] (or flat list[N]) of COO triples -> (counts[B,C], indices[nnz,2], values[nnz], shape).

    All matrices must share one dense_shape (kgcn/data_util.py:30-37 ``align_size`` guarantees it
    for GraphConv inputs).  Empty ``[0,2]`` index arrays -- the short-last-batch dummies of
    kgcn/feed.py:123-126 -- are fine.
    """
    if len(adjs) == 0:
        raise ValueError("empty adjacency list")
    nested = isinstance(adjs[0], (list, tuple)) and len(adjs[0]) > 0 and not _looks_like_triple(adjs[0])
    rows = adjs if nested else [[a] for a in adjs]
    B, C = len(rows), len(rows[0])
    counts = np.zeros((B, C), np.int64)
    idx_parts, val_parts = [], []
    shape = None
    for b, row in enumerate(rows):
        if len(row) != C:
            raise ValueError("graph %d has %d channels, expected %d" % (b, len(row), C))
        for c, sp in enumerate(row):
            i, v, s = _triple(sp)
            if torch.is_tensor(i):
                i = i.detach().cpu().numpy()
            if torch.is_tensor(v):
                v = v.detach().cpu().numpy()
            i = np.asarray(i).reshape(-1, 2)
            v = np.asarray(v, np.float32).reshape(-1)
            if i.shape[0] != v.shape[0]:
                raise ValueError("graph %d channel %d: %d indices but %d values" % (b, c, i.shape[0], v.shape[0]))
            s = (int(s[0]), int(s[1]))
            if shape is None:
                shape = s
            elif s != shape:
                raise ValueError("all adjacency matrices of a batch must share one dense_shape; got %r and %r "
                                 "(the reference aligns them with data_util.align_size)" % (shape, s))
            counts[b, c] = i.shape[0]
            idx_parts.append(i)
            val_parts.append(v)
    indices = np.concatenate(idx_parts, 0) if idx_parts else np.zeros((0, 2), np.int32)
    if indices.dtype not in (np.int32, np.int64):
        indices = indices.astype(np.int64)
    values = np.concatenate(val_parts, 0) if val_parts else np.zeros((0,), np.float32)
    return counts, np.ascontiguousarray(indices), np.ascontiguousarray(values), shape


def _looks_like_triple(obj):
    if hasattr(obj, "indices") and hasattr(obj, "dense_shape"):
        return True
    if isinstance(obj, (list, tuple)) and len(obj) == 3:
        try:
            return len(obj[2]) == 2 and np.ndim(obj[2][0]) == 0 and np.ndim(obj[1]) <= 1 and np.ndim(obj[0]) == 2
        except TypeError:
            return False
    return False


def pack_host(counts, indices, values, n_rows, n_cols, transpose=False, want_perm=False):
    """Run the C-ABI host packer; returns numpy (rowptr, col, val, perm|None)."""
    counts = np.asarray(counts, np.int64)
    n_mat = int(counts.size)
    nnz_off = np.zeros(n_mat + 1, np.int64)
    np.cumsum(counts.reshape(-1), out=nnz_off[1:])
    nnz = int(nnz_off[-1])
    indices = np.ascontiguousarray(indices)
    if indices.dtype not in (np.int32, np.int64):
        indices = indices.astype(np.int64)
    values = np.ascontiguousarray(values, np.float32)
    if indices.shape[0] != nnz or values.shape[0] != nnz:
        raise ValueError("counts sum to %d but %d indices / %d values given" % (nnz, indices.shape[0], values.shape[0]))
    out_rows = n_cols if transpose else n_rows
    rowptr = np.empty(n_mat * out_rows + 1, np.int32)
    col = np.empty(max(nnz, 1), np.int32)
    val = np.empty(max(nnz, 1), np.float32)
    perm = np.empty(max(nnz, 1), np.int32) if want_perm else None
    _lib.check(_lib.lib.kgcn_pack_coo_host(
        n_mat, int(n_rows), int(n_cols), nnz_off.ctypes.data, indices.ctypes.data,
        1 if indices.dtype == np.int64 else 0, values.ctypes.data, 1 if transpose else 0,
        rowptr.ctypes.data, col.ctypes.data, val.ctypes.data, perm.ctypes.data if want_perm else None))
    return rowptr, col[:nnz], val[:nnz], (perm[:nnz] if want_perm else None)


class BatchedCSR:
    """Device CSR of ``n_graphs * channels`` matrices ``[n_rows, n_cols]`` (+ its transpose)."""

    def __init__(self, n_graphs, channels, n_rows, n_cols, rowptr, col, val, rowptr_t, col_t, val_t,
                 perm=None, perm_t=None):
        self.n_graphs, self.channels, self.n_rows, self.n_cols = int(n_graphs), int(channels), int(n_rows), int(n_cols)
        self.rowptr, self.col, self.val = rowptr, col, val
        self.rowptr_t, self.col_t, self.val_t = rowptr_t, col_t, val_t
        self.perm, self.perm_t = perm, perm_t
        self.nnz = int(val.numel())

    @property
    def device(self):
        return self.rowptr.device

    @classmethod
    def from_flat(cls, counts, indices, values, n_rows, n_cols, device="cuda", want_perm=False, pin=False):
        counts = np.asarray(counts, np.int64)
        if counts.ndim == 1:
            counts = counts[:, None]
        B, C = counts.shape
        fwd = pack_host(counts, indices, values, n_rows, n_cols, False, want_perm)
        bwd = pack_host(counts, indices, values, n_rows, n_cols, True, want_perm)

        def up(a):
            if a is None:
                return None
            t = torch.from_numpy(a)
            if pin:
                t = t.pin_memory()
            return t.to(device, non_blocking=pin)

        return cls(B, C, n_rows, n_cols, up(fwd[0]), up(fwd[1]), up(fwd[2]), up(bwd[0]), up(bwd[1]), up(bwd[2]),
                   up(fwd[3]), up(bwd[3]))

    @classmethod
    def from_coo_lists(cls, adjs, device="cuda", want_perm=False):
        counts, indices, values, shape = flatten_coo(adjs)
        return cls.from_flat(counts, indices, values, shape[0], shape[1], device=device, want_perm=want_perm)

    def transposed(self):
        """The adjoint batch (kgcn/bspmm_call.py:44 ``adjoint_a=True``) -- O(1), buffers are shared."""
        return BatchedCSR(self.n_graphs, self.channels, self.n_cols, self.n_rows, self.rowptr_t, self.col_t,
                          self.val_t, self.rowptr, self.col, self.val, self.perm_t, self.perm)

    def with_values(self, val, val_t):
        return BatchedCSR(self.n_graphs, self.channels, self.n_rows, self.n_cols, self.rowptr, self.col, val,
                          self.rowptr_t, self.col_t, val_t, self.perm, self.perm_t)


_PACK_CACHE = {}


def as_batched_csr(adj, device):
    """``adj`` argument of a layer call -> BatchedCSR.  list-of-lists inputs are packed once and
    cached on the list object's identity, so the 3+ GraphConv layers of a model that all receive
    the same ``placeholders['adjs']`` (example_model/model.py:42-46) pack it a single time."""
    if isinstance(adj, BatchedCSR):
        return adj
    key = id(adj)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0] is adj:
        return hit[1]
    csr = BatchedCSR.from_coo_lists(adj, device=device)
    _PACK_CACHE.clear()  # keep exactly one batch alive
    _PACK_CACHE[key] = (adj, csr)
    return csr
