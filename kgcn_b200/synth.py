"""Synthetic workloads of BASELINE.json's configs (shapes per SURVEY.md section 8d).

* :func:`ring_graphs` -- config 2: the generator of ``data_generator/synth_generator_ring.py:11-91``
  scaled from D=10 to D=N nodes: class 0 is a ring of ``p = floor(3N/5)`` nodes, class 1 a ring of
  ``q = N/2`` (6 / 5 at N=10), self-loops on ring nodes, every (noise node, ring node) pair is an
  edge with probability 0.1, symmetric; a FRESH adjacency per sample (the reference script aliases
  one array per class).
* :func:`random_molecule_coo` -- configs 3-5: molecules with ``n ~ U{8..N}`` atoms padded to N, a
  random spanning tree plus ``floor(0.1 n)`` ring-closure bonds, symmetric, diagonal 1
  (``kgcn/preprocessing/utils.py:147-154``), bonds split over C edge types.

All generators are vectorised numpy and emit the flattened COO form (``counts[B,C]``,
``indices[nnz,2] int32`` row-major sorted per matrix like ``data_util.dense_to_sparse``,
``values[nnz] f32``).
"""
import numpy as np


def _dense_to_flat(dense):
    """dense [B, C, N, N] -> counts[B,C], indices[nnz,2], values[nnz] (graph-major, channel-minor)."""
    B, C, N, _ = dense.shape
    b, c, i, j = np.nonzero(dense)
    counts = np.bincount(b * C + c, minlength=B * C).reshape(B, C).astype(np.int64)
    indices = np.stack([i, j], 1).astype(np.int32)
    values = dense[b, c, i, j].astype(np.float32)
    return counts, indices, values


def ring_graphs(rng, B, N, feature_dim, one_hot_features=False, chunk=8192):
    """Returns dict(counts, indices, values, features [B,N,F] f32, labels [B,2] f32)."""
    p, q = (3 * N) // 5, N // 2
    labels_int = (np.arange(B) % 2).astype(np.int64)
    rng.shuffle(labels_int)
    parts = []
    for s in range(0, B, chunk):
        lab = labels_int[s:s + chunk]
        nb = lab.shape[0]
        ring = np.where(lab == 0, p, q)                                        # [nb]
        node = np.arange(N)
        on_ring = node[None, :] < ring[:, None]                                # [nb, N]
        dense = np.zeros((nb, N, N), np.uint8)
        bi = np.arange(nb)[:, None]
        nxt = np.where(node[None, :] + 1 < ring[:, None], node[None, :] + 1, 0)
        dense[bi, node[None, :], node[None, :]] = on_ring                      # self loops on ring nodes
        bsel, isel = np.nonzero(on_ring)
        dense[bsel, isel, nxt[bsel, isel]] = 1
        dense[bsel, nxt[bsel, isel], isel] = 1
        noise = (rng.random((nb, N, N)) < 0.1) & (~on_ring)[:, :, None] & on_ring[:, None, :]   # (noise i, ring j)
        dense |= noise.astype(np.uint8)
        dense |= noise.transpose(0, 2, 1).astype(np.uint8)
        parts.append(_dense_to_flat(dense[:, None]))
    counts = np.concatenate([p_[0] for p_ in parts], 0)
    indices = np.concatenate([p_[1] for p_ in parts], 0)
    values = np.concatenate([p_[2] for p_ in parts], 0)
    if one_hot_features:  # generator Level 1: one-hot of (node index mod F)  (synth_generator_ring.py:86-91)
        features = np.zeros((B, N, feature_dim), np.float32)
        features[:, np.arange(N), np.arange(N) % feature_dim] = 1.0
    else:
        features = rng.standard_normal((B, N, feature_dim), dtype=np.float32)
    labels = np.eye(2, dtype=np.float32)[labels_int]
    return {"counts": counts, "indices": indices, "values": values, "features": features, "labels": labels}


def random_molecule_coo(rng, B, N, C=1, min_atoms=8, type_probs=(0.7, 0.2, 0.1), return_sizes=False):
    """Flattened COO of B random molecules with C bond types.  Self-loops live in channel 0."""
    n_atoms = rng.integers(min(min_atoms, N), N + 1, size=B)
    node = np.arange(N)
    real = node[None, :] < n_atoms[:, None]                                    # [B, N]
    dense = np.zeros((B, C, N, N), np.float32)
    probs = np.asarray(type_probs[:C], np.float64)
    probs = probs / probs.sum()

    def add_bonds(b, i, j):
        keep = i != j
        b, i, j = b[keep], i[keep], j[keep]
        t = rng.choice(C, size=b.shape[0], p=probs) if C > 1 else np.zeros(b.shape[0], np.int64)
        dense[b, t, i, j] = 1.0
        dense[b, t, j, i] = 1.0

    # spanning tree: node i >= 1 bonds to a uniformly random earlier node
    parent = np.floor(rng.random((B, N)) * np.maximum(node, 1)[None, :]).astype(np.int64)
    bsel, isel = np.nonzero(real & (node[None, :] >= 1))
    add_bonds(bsel, isel, parent[bsel, isel])
    # ring closures: floor(0.1 n) random extra bonds
    n_extra = (n_atoms // 10).astype(np.int64)
    max_extra = int(n_extra.max()) if B else 0
    if max_extra:
        u = np.floor(rng.random((B, max_extra)) * n_atoms[:, None]).astype(np.int64)
        v = np.floor(rng.random((B, max_extra)) * n_atoms[:, None]).astype(np.int64)
        bsel, ksel = np.nonzero(np.arange(max_extra)[None, :] < n_extra[:, None])
        add_bonds(bsel, u[bsel, ksel], v[bsel, ksel])
    # a bond keeps a single type: if two draws hit the same pair with different types keep the lowest
    if C > 1:
        seen = np.zeros((B, N, N), bool)
        for c in range(C):
            dense[:, c][seen] = 0.0
            seen |= dense[:, c] > 0
    bsel, isel = np.nonzero(real)
    dense[bsel, 0, isel, isel] = 1.0                                           # diag = 1 on real atoms
    out = _dense_to_flat(dense)
    return out + (n_atoms.astype(np.int32),) if return_sizes else out


def atom_like_features(rng, B, N, n_atoms=None, blocks=(44, 11, 7, 2, 5, 1, 5)):
    """75-dim rows made of one-hot blocks mimicking ``atom_features``
    (kgcn/preprocessing/utils.py:20-54, 75-dim with --use_deepchem_feature); padded atoms are 0."""
    F = int(sum(blocks))
    feats = np.zeros((B, N, F), np.float32)
    off = 0
    bi, ni = np.meshgrid(np.arange(B), np.arange(N), indexing="ij")
    for w in blocks:
        k = rng.integers(0, w, size=(B, N))
        feats[bi, ni, off + k] = 1.0
        off += w
    if n_atoms is not None:
        feats *= (np.arange(N)[None, :] < np.asarray(n_atoms)[:, None])[:, :, None]
    return feats
