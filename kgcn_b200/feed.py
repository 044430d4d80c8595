"""Per-step feed construction -- mirror of ``kgcn/feed.py`` (``construct_feed``, lines 91-234) for the
keys on the graph-convolution path, plus the packed form the CUDA library consumes.

The reference builds ``batch_size * channels`` SparseTensorValue triples and a dense feature copy
per step in Python (feed.py:112-133).  ``construct_feed`` keeps that contract (same keys, shapes,
dtypes, short-last-batch padding: empty sparse dummies with the last real shape, zero features,
``mask = 0``, ``enabled_node_nums = 0``); :class:`FlatGraphDataset` does the same batching on
pre-flattened arrays without a Python loop and hands back pinned host buffers ready for upload.
"""
import collections

import numpy as np
import torch

from .csr import BatchedCSR, pack_host

SparseTensorValue = collections.namedtuple("SparseTensorValue", ["indices", "values", "dense_shape"])

GRAPH_KEYS = ("adjs", "features", "nodes", "labels", "mask", "mask_label", "mask_node", "node_label", "mask_node_label",
              "dropout_rate", "is_train", "enabled_node_nums")


def construct_feed(batch_idx, placeholders, data, batch_size=None, dropout_rate=0.0, is_train=False, info=None,
                   config=None, **kwargs):
    """Returns ``{key: value}`` for every key of ``placeholders`` this path knows (feed.py:110-218).
    ``data`` needs ``adjs`` (list[G][C] of triples), ``features`` [G,N,F] and optionally ``labels``,
    ``mask_label``, ``nodes``, ``enabled_node_nums``."""
    if batch_size is None:
        batch_size = len(batch_idx)
    n_real = len(batch_idx)
    get = (lambda k: data.get(k)) if isinstance(data, dict) else (lambda k: getattr(data, k, None))
    feed = {}
    for key in placeholders:
        if key == "adjs":
            adjs, rows, b_shape = get("adjs"), [], None
            for b in range(batch_size):
                row = []
                for ch in range(len(adjs[0])):
                    if b < n_real:
                        a = adjs[batch_idx[b]][ch]
                        b_shape = a[2]
                        row.append(SparseTensorValue(a[0], a[1], a[2]))
                    else:  # feed.py:123-126
                        row.append(SparseTensorValue(np.zeros((0, 2), np.int32), np.zeros((0,), np.float32), b_shape))
                rows.append(row)
            feed[key] = rows
        elif key == "features" and get("features") is not None:
            f = get("features")
            tmp = np.zeros((batch_size, f.shape[1], f.shape[2]), np.float32)
            tmp[:n_real] = f[batch_idx]
            feed[key] = tmp
        elif key == "nodes" and get("features") is None and get("nodes") is not None:
            n = get("nodes")
            tmp = np.zeros((batch_size, n.shape[1]), np.int32)
            tmp[:n_real] = n[batch_idx]
            feed[key] = tmp
        elif key == "labels" and get("labels") is not None:
            lab = get("labels")
            lab = lab[:, None] if lab.ndim == 1 else lab
            regression = config is not None and config.get("task") == "regression"
            tmp = np.zeros((batch_size, lab.shape[1]), np.float32 if regression else np.int32)
            tmp[:n_real] = lab[batch_idx]
            feed[key] = tmp
        elif key == "mask":
            m = np.zeros((batch_size,), np.float32)
            m[:n_real] = 1
            feed[key] = m
        elif key == "mask_label" and get("mask_label") is not None:
            ml = get("mask_label")
            ml = ml[:, None] if ml.ndim == 1 else ml
            tmp = np.zeros((batch_size, ml.shape[1]), np.float32)
            tmp[:n_real] = ml[batch_idx]
            feed[key] = tmp
        elif key == "node_label" and get("node_label") is not None:          # feed.py:160-163 (node-centric models)
            nl = np.asarray(get("node_label"))
            tmp = np.zeros((batch_size, nl.shape[1], nl.shape[2]), np.float32)
            tmp[:n_real] = nl[batch_idx]
            feed[key] = tmp
        elif key == "mask_node_label" and get("mask_node_label") is not None:  # feed.py:164-170
            # the reference allocates [batch, nodes] and assigns [n, nodes, labels] rows into it, which only works for a
            # mask without a label axis; a mask that carries one (the shipped sample_node_label.jbl) keeps it here
            ml = np.asarray(get("mask_node_label"))
            tmp = np.zeros((batch_size,) + ml.shape[1:], np.float32)
            tmp[:n_real] = ml[batch_idx]
            feed[key] = tmp
        elif key == "mask_node" and get("enabled_node_nums") is not None:      # feed.py:209-214: 1 for the real atoms
            n_nodes = int(info.graph_node_num) if info is not None else int(np.asarray(get("features")).shape[1])
            tmp = np.zeros((batch_size, n_nodes), np.float32)
            lengths = np.asarray(get("enabled_node_nums"))[batch_idx].reshape(-1)
            tmp[:n_real] = np.arange(n_nodes)[None, :] < lengths[:, None]
            feed[key] = tmp
        elif key == "dropout_rate":
            feed[key] = dropout_rate
        elif key == "is_train":
            feed[key] = is_train
        elif key == "enabled_node_nums" and get("enabled_node_nums") is not None:
            tmp = np.zeros((batch_size,), np.int32)
            tmp[:n_real] = np.squeeze(np.asarray(get("enabled_node_nums"))[batch_idx])
            feed[key] = tmp
    return feed


class PackedBatch:
    """A feed on the device: BatchedCSR + feature / label / mask tensors."""

    def __init__(self, csr, features, labels=None, mask=None, enabled_node_nums=None):
        self.csr, self.features, self.labels, self.mask = csr, features, labels, mask
        self.enabled_node_nums = enabled_node_nums


def pack_feed(feed, device="cuda"):
    """``construct_feed`` output -> :class:`PackedBatch` (one CSR pack + one upload per array)."""
    def up(key, dtype):
        v = feed.get(key)
        return None if v is None else torch.as_tensor(np.asarray(v), dtype=dtype).to(device)

    csr = BatchedCSR.from_coo_lists(feed["adjs"], device=device)
    return PackedBatch(csr, up("features", torch.float32), up("labels", torch.float32), up("mask", torch.float32),
                       up("enabled_node_nums", torch.int32))


class FlatGraphDataset:
    """Whole dataset flattened once: ``counts[G,C]``, ``offsets[G*C+1]``, ``indices[nnz,2] int32``,
    ``values[nnz] f32``, ``features[G,N,F] f32``, ``labels[G,L] f32``.  ``host_batch`` slices a batch
    with vectorised numpy (no per-graph Python), packs CSR + transposed CSR through the C-ABI host
    packer into pinned buffers; ``upload`` issues the H2D copies on the current stream."""

    def __init__(self, counts, indices, values, features, labels, n_nodes, enabled_node_nums=None):
        self.counts = np.ascontiguousarray(counts, np.int64)
        self.G, self.C = self.counts.shape
        self.offsets = np.zeros(self.G * self.C + 1, np.int64)
        np.cumsum(self.counts.reshape(-1), out=self.offsets[1:])
        self.indices = np.ascontiguousarray(indices, np.int32).reshape(-1, 2)
        self.values = np.ascontiguousarray(values, np.float32)
        self.features = np.ascontiguousarray(features, np.float32)
        self.labels = None if labels is None else np.ascontiguousarray(labels, np.float32)
        self.n_nodes = int(n_nodes)
        self.enabled_node_nums = enabled_node_nums

    @classmethod
    def from_adjs(cls, adjs, features, labels, n_nodes, enabled_node_nums=None):
        from .csr import flatten_coo
        counts, indices, values, _ = flatten_coo(adjs)
        return cls(counts, indices, values, features, labels, n_nodes, enabled_node_nums)

    def host_batch(self, batch_idx, batch_size=None):
        batch_idx = np.asarray(batch_idx, np.int64)
        n_real = batch_idx.shape[0]
        batch_size = n_real if batch_size is None else batch_size
        C = self.C
        counts = np.zeros((batch_size, C), np.int64)
        counts[:n_real] = self.counts[batch_idx]
        # gather the COO ranges of the selected graphs (all channels of a graph are contiguous)
        starts = self.offsets[batch_idx * C]
        lens = self.counts[batch_idx].sum(axis=1)
        total = int(lens.sum())
        if total:
            seg_off = np.zeros(n_real + 1, np.int64)
            np.cumsum(lens, out=seg_off[1:])
            pos = np.arange(total, dtype=np.int64) - np.repeat(seg_off[:-1], lens) + np.repeat(starts, lens)
            indices, values = self.indices[pos], self.values[pos]
        else:
            indices, values = np.zeros((0, 2), np.int32), np.zeros((0,), np.float32)
        N = self.n_nodes
        fwd = pack_host(counts, indices, values, N, N, transpose=False)
        bwd = pack_host(counts, indices, values, N, N, transpose=True)
        feats = np.zeros((batch_size,) + self.features.shape[1:], np.float32)
        feats[:n_real] = self.features[batch_idx]
        out = {"rowptr": fwd[0], "col": fwd[1], "val": fwd[2], "rowptr_t": bwd[0], "col_t": bwd[1], "val_t": bwd[2],
               "features": feats}
        mask = np.zeros((batch_size,), np.float32)
        mask[:n_real] = 1
        out["mask"] = mask
        if self.labels is not None:
            lab = np.zeros((batch_size, self.labels.shape[1]), np.float32)
            lab[:n_real] = self.labels[batch_idx]
            out["labels"] = lab
        return out

    def upload(self, host, device="cuda", pinned=None):
        """H2D of one ``host_batch`` result.  ``pinned``: optional dict of preallocated pinned staging
        tensors (reused across steps); returns (PackedBatch, bytes copied)."""
        dev, nbytes = {}, 0
        for k, a in host.items():
            t = torch.from_numpy(np.ascontiguousarray(a))
            if pinned is not None:
                buf = pinned.get(k)
                if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                    buf = torch.empty(max(t.numel(), 1), dtype=t.dtype).pin_memory()
                    pinned[k] = buf
                stage = buf[:t.numel()].view(t.shape) if t.numel() else t
                if t.numel():
                    stage.copy_(t)
                t = stage
            dev[k] = t.to(device, non_blocking=True)
            nbytes += t.numel() * t.element_size()
        B = host["features"].shape[0]
        csr = BatchedCSR(B, self.C, self.n_nodes, self.n_nodes, dev["rowptr"], dev["col"], dev["val"], dev["rowptr_t"],
                         dev["col_t"], dev["val_t"])
        return PackedBatch(csr, dev["features"], dev.get("labels"), dev["mask"]), nbytes
