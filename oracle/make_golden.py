#!/usr/bin/env python
"""Generate the golden ingest vectors under ``tests/golden/`` by running the REFERENCE's own
numpy code (``/root/reference/kgcn/data_util.py`` + ``kgcn/feed.py``) under ``oracle/tf_stub``.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python oracle/make_golden.py

The GPU box has no ``/root/reference``; tests there read only the committed ``.npz`` files.
Each ``.npz`` stores the raw fixture content (the inputs) and what the reference produced for
it (the expected outputs), flattened as

    counts[G, C]  (nnz per graph/channel), indices[sum, 2], values[sum], shapes[G, C, 2]

Variants: plain / normalize_adj_flag / split_adj_flag / order=2 of ``build_data``
(data_util.py:374-424) and ``construct_feed`` (feed.py:91-234) for a full and a short batch.
Also stores the arithmetic known-answer vectors KAT1-KAT3 evaluated by
``oracle/ref_layers.py`` on the reference-ingested fixtures.
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

from oracle import tf_stub  # noqa: E402

tf_stub.install()
sys.path.insert(0, REF)
import joblib  # noqa: E402
import kgcn.data_util as du  # noqa: E402  (the reference's module)
import kgcn.feed as rfeed  # noqa: E402

from oracle import ref_layers as R  # noqa: E402


def flatten_adjs(adjs):
    G, C = len(adjs), len(adjs[0])
    counts = np.zeros((G, C), np.int64)
    shapes = np.zeros((G, C, 2), np.int64)
    idx, val = [], []
    for g in range(G):
        for c in range(C):
            i = np.asarray(adjs[g][c][0]).reshape(-1, 2)
            counts[g, c] = i.shape[0]
            shapes[g, c] = np.asarray(adjs[g][c][2]).reshape(2)
            idx.append(i.astype(np.int64))
            val.append(np.asarray(adjs[g][c][1], np.float32).reshape(-1))
    return {"counts": counts, "indices": np.concatenate(idx, 0), "values": np.concatenate(val, 0), "shapes": shapes}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


BASE_CFG = {"with_feature": True, "with_node_embedding": False, "normalize_adj_flag": False,
            "split_adj_flag": False, "shuffle_data": False}
VARIANTS = {
    "plain": {},
    "norm": {"normalize_adj_flag": True},
    "split": {"split_adj_flag": True},
    "order2": {"order": 2},
    "split_norm": {"split_adj_flag": True, "normalize_adj_flag": True},
}


class FakePlaceholders(dict):
    """feed.py only iterates ``placeholders.items()`` and uses the values as dict keys."""


def run_feed(all_data, info, batch_idx, batch_size):
    C = info.adj_channel_num
    ph = {
        "adjs": [[("adj", b, c) for c in range(C)] for b in range(batch_size)],
        "features": "features", "labels": "labels", "mask": "mask",
        "enabled_node_nums": "enabled_node_nums", "dropout_rate": "dropout_rate", "is_train": "is_train",
    }
    fd = rfeed.construct_feed(batch_idx, ph, all_data, batch_size=batch_size, dropout_rate=0.0,
                              is_train=False, info=info, config={"task": "classification"})
    adjs = [[fd[("adj", b, c)] for c in range(C)] for b in range(batch_size)]
    out = {"feed_" + k: v for k, v in flatten_adjs([[(a.indices, a.values, a.dense_shape) for a in row] for row in adjs]).items()}
    out["feed_features"] = fd["features"]
    out["feed_mask"] = fd["mask"]
    out["feed_labels"] = fd["labels"]
    out["feed_enabled_node_nums"] = fd["enabled_node_nums"]
    out["feed_batch_idx"] = np.asarray(batch_idx, np.int64)
    return out


def object_array(items):
    out = np.empty(len(items), dtype=object)
    for i, item in enumerate(items):
        out[i] = item
    return out


def build_data_goldens(out_dir):
    """``build_data`` beyond the adjacency block -- the ``info`` fields, labels / masks / class weights -- and the
    reference's ``split_data`` / ``shuffle_data`` under a fixed ``np.random.seed``  ->  tests/golden/build_data_*.npz.

    With the installed numpy (2.x) the reference's ``np.array(list of ragged adjacency triples)`` raises, so its own
    split / shuffle cannot run on the list ``build_data`` returns; they are run unchanged on the same data with
    ``adjs`` pre-wrapped in a 1-D object array (the ``isinstance(..., np.ndarray)`` branch of data_util.py:636-638)."""
    import json
    fixtures = [("sample", "sample.jbl", {}), ("sample_multiadj", "sample_multiadj.jbl", {}), ("sample_multitask", "sample_multitask.jbl", {}),
                ("sample_node_label", "sample_node_label.jbl", {}), ("synthetic", "synthetic.jbl", {}),
                ("synthetic_norm", "synthetic.jbl", {"normalize_adj_flag": True})]
    for name, fname, extra in fixtures:
        raw = joblib.load(os.path.join(REF, "example_jbl", fname))
        cfg = dict(BASE_CFG, **extra)
        all_data, info = quiet(du.build_data, cfg, raw, prohibit_shuffle=True)
        rec = {"raw_keys": np.array(sorted(raw.keys()))}
        for k, v in raw.items():
            rec["in_" + k] = np.asarray(v)
        scalars = {k: (None if info[k] is None else (bool(info[k]) if isinstance(info[k], (bool, np.bool_)) else int(info[k])))
                   for k in ("all_node_num", "feature_dim", "graph_node_num", "feature_enabled", "sequence_max_length",
                             "sequence_symbol_num", "sequences_vec_dim", "graph_num", "adj_channel_num", "label_dim")}
        scalars["vector_modal_dim"], scalars["vector_modal_name"] = list(info.vector_modal_dim), dict(info.vector_modal_name)
        scalars["info_keys"] = sorted(info.keys())
        scalars["all_data_keys"] = sorted(all_data.keys())
        scalars["none_members"] = sorted(k for k in all_data if all_data[k] is None)
        rec["info_json"] = np.array(json.dumps(scalars, sort_keys=True))
        for k in ("pos_weight", "class_weight"):
            if info.get(k) is not None:
                rec["info_" + k] = np.asarray(info[k])
        for k in ("labels", "mask_label", "node_label", "mask_node_label", "sequences", "sequences_len", "enabled_node_nums", "features"):
            if all_data[k] is not None:
                rec["all_" + k] = np.asarray(all_data[k])
        rec["all_num"] = np.int64(all_data.num)
        # node-level feed keys (feed.py:160-163, 209-214) and the label mask, short batch (padded by one)
        bi = [2, 0, 1]
        keys = ["mask_node", "enabled_node_nums"] + [k for k in ("node_label", "mask_label") if all_data[k] is not None]
        fd = rfeed.construct_feed(bi, {k: k for k in keys}, all_data, batch_size=len(bi) + 1, info=info, config={"task": "classification"})
        rec["feed_keys"], rec["feed_batch_idx"] = np.array(keys), np.asarray(bi, np.int64)
        for k in keys:
            rec["feed_" + k] = np.asarray(fd[k])
        # split_data, seed 7, 40 % validation
        all_data.adjs = object_array(all_data.adjs)
        np.random.seed(7)
        train, valid = quiet(du.split_data, all_data, 0.4)
        for tag, part in (("train", train), ("valid", valid)):
            rec["split_%s_num" % tag] = np.int64(part.num)
            for k in ("features", "labels", "mask_label", "node_label", "enabled_node_nums", "sequences_len"):
                if part[k] is not None:
                    rec["split_%s_%s" % (tag, k)] = np.asarray(part[k])
            for k, v in flatten_adjs(list(part.adjs)).items():
                rec["split_%s_adj_%s" % (tag, k)] = v
        # explicit index lists
        tr2, va2 = quiet(du.split_data, all_data, 0.4, indices_for_train_data=[2, 0], indices_for_valid_data=[1, 3])
        rec["split_explicit_train_enabled"], rec["split_explicit_valid_enabled"] = np.asarray(tr2.enabled_node_nums), np.asarray(va2.enabled_node_nums)
        # shuffle_data, seed 3
        np.random.seed(3)
        shuffled = quiet(du.shuffle_data, all_data)
        for k in ("features", "labels", "mask_label", "node_label", "enabled_node_nums", "sequences_len"):
            if shuffled[k] is not None:
                rec["shuffle_" + k] = np.asarray(shuffled[k])
        for k, v in flatten_adjs(list(shuffled.adjs)).items():
            rec["shuffle_adj_" + k] = v
        path = os.path.join(out_dir, "build_data_%s.npz" % name)
        np.savez_compressed(path, **rec)
        print("wrote", os.path.relpath(path, ROOT), scalars["graph_num"], "graphs, label_dim", scalars["label_dim"])


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    fixtures = {
        "sample": ("sample.jbl", ["plain", "norm", "split", "order2", "split_norm"], ([0, 1, 2, 3, 4], 5), ([3, 4], 4)),
        "sample_multiadj": ("sample_multiadj.jbl", ["plain", "norm"], ([0, 1, 2, 3], 4), ([2], 3)),
        "synthetic": ("synthetic.jbl", ["plain", "norm", "split"], (list(range(10)), 10), (list(range(190, 197)), 10)),
        "synthetic_sparse": ("synthetic_sparse.jbl", ["plain"], None, None),
    }
    for name, (fname, variants, full, short) in fixtures.items():
        raw = joblib.load(os.path.join(REF, "example_jbl", fname))
        for var in variants:
            cfg = dict(BASE_CFG, **VARIANTS[var])
            if name == "synthetic_sparse":
                # COO `adj` form with `node` ids; build_data wants a `node_num` key (data_util.py:496)
                # that this shipped fixture lacks, so it is supplied here (max id + 1).
                cfg["with_feature"] = False
                cfg["with_node_embedding"] = True
                data = dict(raw, node_num=int(np.max(raw["node"])) + 1)
                all_data, info = quiet(du.build_data, cfg, data, prohibit_shuffle=True)
            else:
                all_data, info = quiet(du.load_data, cfg, os.path.join(REF, "example_jbl", fname), prohibit_shuffle=True)
            rec = {}
            # inputs
            for k in ("feature", "dense_adj", "label", "max_node_num", "node"):
                if k in raw:
                    rec["in_" + k] = np.asarray(raw[k])
            if "multi_dense_adj" in raw:
                rec["in_multi_dense_adj"] = np.asarray(raw["multi_dense_adj"])
            if "adj" in raw:
                for k, v in flatten_adjs([[a] for a in raw["adj"]]).items():
                    rec["in_adj_" + k] = v
            # expected outputs of build_data
            for k, v in flatten_adjs(all_data.adjs).items():
                rec["adj_" + k] = v
            rec["enabled_node_nums"] = np.asarray(all_data.enabled_node_nums)
            rec["adj_channel_num"] = np.int64(info.adj_channel_num)
            rec["graph_node_num"] = np.int64(info.graph_node_num)
            if all_data.features is not None:
                rec["features"] = np.asarray(all_data.features)
            if full is not None and var in ("plain", "norm", "split"):
                for tag, (bi, bs) in (("full", full), ("short", short)):
                    for k, v in run_feed(all_data, info, bi, bs).items():
                        rec[tag + "_" + k] = v
            path = os.path.join(out_dir, "ingest_%s_%s.npz" % (name, var))
            np.savez_compressed(path, **rec)
            print("wrote", os.path.relpath(path, ROOT), "C=%d" % info.adj_channel_num, "nnz=%d" % rec["adj_values"].shape[0])

    # ---- arithmetic known-answer vectors, via reference ingest + O1 ----
    kat = {}
    all_data, info = quiet(du.load_data, dict(BASE_CFG), os.path.join(REF, "example_jbl", "sample.jbl"), prohibit_shuffle=True)
    W = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
    b = np.array([[0.5, -0.5]], np.float32)
    y = R.graph_conv(all_data.features, all_data.adjs, [W], [b])
    kat["kat1_y"] = y
    kat["kat1_gather"] = R.graph_gather(y)
    a4 = all_data.adjs[4][0]
    kat["kat1_adjoint_g4"] = R.sparse_dense_matmul(a4[0], a4[1], a4[2], y[4], adjoint_a=True)
    md, minfo = quiet(du.load_data, dict(BASE_CFG), os.path.join(REF, "example_jbl", "sample_multiadj.jbl"), prohibit_shuffle=True)
    kat["kat2_y"] = R.graph_conv(md.features, md.adjs, [W, -W], [b, np.array([[1, 2]], np.float32)])
    np.savez_compressed(os.path.join(out_dir, "kat.npz"), **kat)
    print("wrote tests/golden/kat.npz")
    for k, v in kat.items():
        print(k, v.tolist())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build_data":      # only the build_data_*.npz files
        build_data_goldens(os.path.join(ROOT, "tests", "golden"))
    else:
        main()
        build_data_goldens(os.path.join(ROOT, "tests", "golden"))
