#!/usr/bin/env python
"""Golden LAYER outputs produced by the reference's own ``kgcn/layers.py`` (and ``kgcn/legacy/layers.py``), executed
unchanged under the numpy-eager TensorFlow stand-in ``oracle/tf_numpy.py``, on batches ingested by the reference's own
``kgcn/data_util.py`` + ``kgcn/feed.py``.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python oracle/make_layer_golden.py        ->  tests/golden/layers_*.npz

Every file stores the inputs (features, flattened adjacency lists, enabled_node_nums), the weights the reference layer
objects created or were given, and the outputs of their ``call``.  The tests require the CPU oracle
(``oracle/ref_layers.py``) to reproduce the outputs bit for bit (same primitives, same order) and the CUDA path to
reproduce them within the stated fp32 tolerance.  What "reference" means here is spelled out in ``oracle/tf_numpy.py``:
the control flow is the reference's, the TensorFlow primitives are numpy restatements.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("KGCN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import tf_numpy  # noqa: E402

tf = tf_numpy.install()
sys.path.insert(0, REF)
import kgcn.data_util as du  # noqa: E402  (the reference's modules)
import kgcn.feed as rfeed  # noqa: E402
import kgcn.layers as rl  # noqa: E402
import kgcn.legacy.layers as rl_legacy  # noqa: E402

CFG = {"with_feature": True, "with_node_embedding": False, "normalize_adj_flag": False, "split_adj_flag": False,
       "shuffle_data": False}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def flatten_adjs(adjs, prefix="adj_"):
    G, C = len(adjs), len(adjs[0])
    counts = np.zeros((G, C), np.int64)
    shapes = np.zeros((G, C, 2), np.int64)
    idx, val = [], []
    for g in range(G):
        for c in range(C):
            i = np.asarray(adjs[g][c].indices).reshape(-1, 2)
            counts[g, c] = i.shape[0]
            shapes[g, c] = np.asarray(adjs[g][c].dense_shape).reshape(2)
            idx.append(i.astype(np.int64))
            val.append(np.asarray(adjs[g][c].values, np.float32).reshape(-1))
    return {prefix + "counts": counts, prefix + "indices": np.concatenate(idx, 0), prefix + "values": np.concatenate(val, 0),
            prefix + "shapes": shapes}


def fed_batch(fname, cfg, batch_idx, batch_size):
    """features, adjs (as fed to the placeholders), enabled_node_nums of one batch -- the reference's own ingest."""
    all_data, info = quiet(du.load_data, dict(CFG, **cfg), os.path.join(REF, "example_jbl", fname), prohibit_shuffle=True)
    C = info.adj_channel_num
    ph = {"adjs": [[("adj", b, c) for c in range(C)] for b in range(batch_size)], "features": "features", "labels": "labels",
          "mask": "mask", "enabled_node_nums": "enabled_node_nums", "dropout_rate": "dropout_rate", "is_train": "is_train"}
    fd = rfeed.construct_feed(batch_idx, ph, all_data, batch_size=batch_size, dropout_rate=0.0, is_train=False, info=info,
                              config={"task": "classification"})
    adjs = [[tf.SparseTensorValue(*fd[("adj", b, c)]) for c in range(C)] for b in range(batch_size)]
    return tf_numpy.T(fd["features"]), adjs, np.asarray(fd["enabled_node_nums"]), info


def set_weights(layer, ws, bs):
    for c in range(len(ws)):
        layer.w[c][...] = ws[c]
        layer.bias[c][...] = bs[c]


def random_adjs(rng, B, N, C, max_nnz, unique):
    adjs = []
    for _ in range(B):
        row = []
        for _ in range(C):
            if unique:    # sorted, duplicate-free (what tf.sparse_tensor_to_dense accepts): GraphMaxPooling
                dense = (rng.random((N, N)) < 0.3) * rng.standard_normal((N, N))
                dense[0, :] = rng.standard_normal(N) - 1.5      # a fully stored row (no implicit zero)
                idx = np.argwhere(dense != 0)
                row.append(tf.SparseTensorValue(idx.astype(np.int64), dense[idx[:, 0], idx[:, 1]].astype(np.float32), [N, N]))
            else:         # storage order kept, duplicates allowed: the sparse matmul accumulates them
                nnz = int(rng.integers(0, max_nnz + 1))
                row.append(tf.SparseTensorValue(rng.integers(0, N, size=(nnz, 2)).astype(np.int64),
                                                rng.standard_normal(nnz).astype(np.float32), [N, N]))
        adjs.append(row)
    return adjs


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    W = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
    b = np.array([[0.5, -0.5]], np.float32)

    # ---- 1. sample.jbl, KAT weights: the reference code itself produces the hand-derived KAT1 ----
    x, adjs, enabled, info = fed_batch("sample.jbl", {}, [0, 1, 2, 3, 4], 5)
    conv = rl.GraphConv(2, 1)
    conv(x, adj=adjs)
    set_weights(conv, [W], [b])
    y = conv(x, adj=adjs)
    rec = dict(flatten_adjs(adjs), features=np.asarray(x), w=np.stack([W]), bias=np.stack([b]), y=np.asarray(y),
               gather=np.asarray(rl.GraphGather()(y)))
    np.savez_compressed(os.path.join(out_dir, "layers_sample_kat.npz"), **rec)
    print("layers_sample_kat: y[0] =", np.asarray(y)[0].tolist(), " y[4] =", np.asarray(y)[4].tolist())

    # ---- 2. sample_multiadj.jbl, two channels, KAT2 weights ----
    x, adjs, enabled, info = fed_batch("sample_multiadj.jbl", {}, [0, 1, 2, 3], 4)
    conv = rl.GraphConv(2, 2)
    conv(x, adj=adjs)
    set_weights(conv, [W, -W], [b, np.array([[1, 2]], np.float32)])
    y = conv(x, adj=adjs)
    rec = dict(flatten_adjs(adjs), features=np.asarray(x), w=np.stack([W, -W]), bias=np.stack([b, np.array([[1, 2]], np.float32)]),
               y=np.asarray(y))
    np.savez_compressed(os.path.join(out_dir, "layers_multiadj_kat.npz"), **rec)
    print("layers_multiadj_kat: y[0] =", np.asarray(y)[0].tolist())

    # ---- 3. synthetic.jbl: the layer stack of example_model/model.py:41-56 (short batch: 7 graphs padded to 10) ----
    tf_numpy.seed(1234)
    x, adjs, enabled, info = fed_batch("synthetic.jbl", {}, list(range(190, 197)), 10)
    N = info.graph_node_num
    c1, c2 = rl.GraphConv(50, 1), rl.GraphConv(50, 1)
    bn, gd, gg = rl.GraphBatchNormalization(), rl.GraphDense(50), rl.GraphGather()
    sig = lambda t: tf_numpy.T(1.0 / (1.0 + np.exp(-np.asarray(t, np.float32), dtype=np.float32)))
    h1 = sig(c1(x, adj=adjs))
    h2 = c2(h1, adj=adjs)
    h3 = sig(bn(h2, max_node_num=N, enabled_node_nums=enabled))
    h4 = sig(gd(h3))
    out = gg(h4)
    rec = dict(flatten_adjs(adjs), features=np.asarray(x), enabled_node_nums=enabled,
               w1=np.stack(c1.w), b1=np.stack(c1.bias), w2=np.stack(c2.w), b2=np.stack(c2.bias), gd_kernel=np.asarray(gd.kernel),
               gd_bias=np.asarray(gd.bias), h1=np.asarray(h1), h2=np.asarray(h2), h3=np.asarray(h3), h4=np.asarray(h4),
               gathered=np.asarray(out))
    # GraphDense with enabled_node_nums (layers.py:241-253): padded rows become exact zeros
    gd2 = rl.GraphDense(6)
    h5 = gd2(h3, max_node_num=N, enabled_node_nums=enabled)
    rec.update(gd2_kernel=np.asarray(gd2.kernel), gd2_bias=np.asarray(gd2.bias), h5=np.asarray(h5))
    # legacy batch-statistics normalisation (kgcn/legacy/layers.py:186-216)
    rec["h2_bn_legacy"] = np.asarray(rl_legacy.GraphBatchNormalization()(h2, max_node_num=N, enabled_node_nums=enabled))
    np.savez_compressed(os.path.join(out_dir, "layers_synthetic_stack.npz"), **rec)
    print("layers_synthetic_stack: gathered[0,:3] =", np.asarray(out)[0, :3].tolist(), " enabled =", enabled.tolist())

    # ---- 4. random COO batches: unsorted / duplicated entries, three channels; GIN; max pooling; block-diagonal conv ----
    rng = np.random.default_rng(7)
    tf_numpy.seed(7)
    B, N, C, F, H = 6, 7, 3, 5, 9
    x = tf_numpy.T(rng.standard_normal((B, N, F)))
    adjs = random_adjs(rng, B, N, C, 20, unique=False)
    conv = rl.GraphConv(H, C)
    conv(x, adj=adjs)
    for c in range(C):
        conv.bias[c][...] = rng.uniform(-0.5, 0.5, (1, H)).astype(np.float32)
    y = conv(x, adj=adjs)
    gin = rl.GINAggregate(C)
    gin(x, adj=adjs)
    eps = rng.uniform(-0.5, 0.5, C).astype(np.float32)
    gin.epsilon = [tf_numpy.T(e) for e in eps]
    rec = dict(flatten_adjs(adjs), features=np.asarray(x), w=np.stack(conv.w), bias=np.stack(conv.bias), y=np.asarray(y),
               gin_eps=eps, gin_y=np.asarray(gin(x, adj=adjs)))
    uadjs = random_adjs(rng, B, N, 2, 0, unique=True)
    rec.update(flatten_adjs(uadjs, "uadj_"))
    rec["maxpool_y"] = np.asarray(rl.GraphMaxPooling(2)(x, adj=uadjs))
    # BatchGraphConv (layers.py:363-398): one sparse matrix over all rows of the batch, relu inside the layer
    big = tf.SparseTensorValue(np.concatenate([np.asarray(a[0].indices) + g * N for g, a in enumerate(adjs)]),
                               np.concatenate([a[0].values for a in adjs]), [B * N, B * N])
    bgc = rl.BatchGraphConv(H)
    flat = tf_numpy.T(np.asarray(x).reshape(B * N, F))
    bgc([flat, big])
    rec.update(bgc_w=np.asarray(bgc.w), bgc_bias=np.asarray(bgc.bias), bgc_y=np.asarray(bgc([flat, big])))
    np.savez_compressed(os.path.join(out_dir, "layers_random.npz"), **rec)
    print("layers_random: y[0,0,:3] =", np.asarray(y)[0, 0, :3].tolist(), " maxpool[0,0,:3] =", rec["maxpool_y"][0, 0, :3].tolist())


if __name__ == "__main__":
    main()
