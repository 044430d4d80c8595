#!/usr/bin/env python
"""Real-weights golden: the network of the reference's ``example_model/model_rxn_3layer.py`` (the model its shipped
checkpoint ``model/reaction/model.best.ckpt.*`` was trained with), executed with THOSE weights.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python oracle/make_ckpt_golden.py        ->  tests/golden/ckpt_reaction.npz, tests/golden/ckpt_reaction_manifest.json

What runs: the reference's own ``kgcn/legacy/layers.py`` classes (GraphConv, GraphBatchNormalization, GraphDense,
GraphGather), called in the order and with the arguments of ``model_rxn_3layer.py:50-88``, unchanged, under the numpy
stand-in for the TensorFlow primitives (``oracle/tf_numpy.py``).  The weights come out of the TensorFlow-written
checkpoint through the oracle's OWN minimal bundle reader (``oracle/tf_bundle_min.py``, which shares no code with the
product's ``kgcn_b200/tf_checkpoint.py``); every tensor's bytes are checked there against the CRC-32C that TensorFlow
stored next to them, and the manifest records name / dtype / shape / masked CRC of all 56 entries.  The product's reader
is then checked AGAINST this one (same names, dtypes, shapes, offsets, checksums, values) -- the golden does not depend on it.
The molecules are synthetic (75 atom features, the ``--use_deepchem_feature`` width the first kernel ``[75, 128]``
implies); the file also stores the re-serialised bytes' digests: writing the parsed checkpoint back with
``save_checkpoint`` reproduces TensorFlow's ``.index`` and ``.data`` files byte for byte.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("KGCN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import tf_numpy  # noqa: E402

tf = tf_numpy.install()
sys.path.insert(0, REF)
import kgcn.legacy.layers as rl_legacy  # noqa: E402  (the reference's module)

from oracle import tf_bundle_min  # noqa: E402
from kgcn_b200 import synth, tf_checkpoint  # noqa: E402  (synth: numpy generators; tf_checkpoint: only as the thing CHECKED below)

PREFIX = os.path.join(REF, "model", "reaction", "model.best.ckpt")


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    header, entries = tf_bundle_min.read_index(PREFIX)
    every = tf_bundle_min.read_tensors(PREFIX)                 # verifies every tensor's CRC-32C
    manifest = {"prefix": "model/reaction/model.best.ckpt", "num_shards": header["num_shards"],
                "entries": [{"name": n, "dtype": np.dtype(e["dtype"]).name, "shape": list(e["shape"]), "offset": e["offset"],
                             "size": e["size"], "crc32c_masked": e["crc32c"]} for n, e in entries.items()]}
    # the product's reader against the oracle's (not the other way round)
    reader = tf_checkpoint.load_checkpoint(PREFIX)
    assert list(reader.entries) == list(entries) and reader.num_shards == header["num_shards"]
    for n, e in entries.items():
        p = reader.entries[n]
        assert (np.dtype(p.dtype), list(p.shape), p.offset, p.size, p.crc32c) == (np.dtype(e["dtype"]), list(e["shape"]), e["offset"], e["size"], e["crc32c"]), n
        assert np.array_equal(reader.get_tensor(n), every[n]) and reader.get_tensor(n).dtype == every[n].dtype, n
    with tempfile.TemporaryDirectory() as tmp:
        tf_checkpoint.save_checkpoint(os.path.join(tmp, "again"), every)
        for ext in (".index", ".data-00000-of-00001"):
            ours, theirs = open(os.path.join(tmp, "again" + ext), "rb").read(), open(PREFIX + ext, "rb").read()
            assert ours == theirs, "re-serialised %s differs from TensorFlow's file" % ext
            manifest["sha256" + ext] = hashlib.sha256(theirs).hexdigest()
    json.dump(manifest, open(os.path.join(out_dir, "ckpt_reaction_manifest.json"), "w"), indent=1)

    W = {n: v for n, v in every.items()
         if n.rsplit("/", 1)[-1] not in ("Adam", "Adam_1") and n not in ("beta1_power", "beta2_power")}   # trained variables only
    rng = np.random.default_rng(2024)
    B, N, C = 12, 50, 1
    counts, idx, val, n_atoms = synth.random_molecule_coo(rng, B, N, C, return_sizes=True)
    feats = synth.atom_like_features(rng, B, N, n_atoms)
    adjs, pos = [], 0
    for b in range(B):
        k = int(counts[b, 0])
        adjs.append([tf.SparseTensorValue(idx[pos:pos + k].astype(np.int64), val[pos:pos + k], [N, N])])
        pos += k
    enabled = np.asarray(n_atoms, np.int32)

    def conv(name, x):
        layer = rl_legacy.GraphConv(128, C)
        layer(x, adj=adjs)                                     # builds kernel0 / bias0
        layer.w[0][...] = W["rollout/%s/kernel0" % name]
        layer.bias[0][...] = W["rollout/%s/bias0" % name]
        return layer(x, adj=adjs)

    def bn(name, x):
        tf_numpy.BN_VARIABLES.append({k: W["rollout/%s/%s" % (name, k)] for k in ("gamma", "beta", "moving_mean", "moving_variance")})
        return rl_legacy.GraphBatchNormalization()(x, max_node_num=N, enabled_node_nums=enabled)

    relu = tf.nn.relu
    x = tf_numpy.T(feats)
    h1 = relu(bn("batch_normalization", conv("graph_conv", x)))             # model_rxn_3layer.py:55-63
    h2 = relu(bn("batch_normalization_1", conv("graph_conv_1", h1)))        # :65-73
    h3 = relu(bn("batch_normalization_2", conv("graph_conv_2", h2)))        # :75-83
    gd = rl_legacy.GraphDense(128)
    gd(h3)
    gd.kernel[...] = W["rollout/graph_dense/kernel"]
    gd.bias[...] = W["rollout/graph_dense/bias"]
    h4 = relu(gd(h3))                                                       # :85-86
    g = rl_legacy.GraphGather()(h4)                                         # :88
    logits = (np.asarray(g, np.float32) @ W["rollout/dense/kernel"] + W["rollout/dense/bias"]).astype(np.float32)   # :89 K.layers.Dense
    assert not tf_numpy.BN_VARIABLES
    rec = {"adj_counts": counts, "adj_indices": idx.astype(np.int64), "adj_values": val,
           "adj_shapes": np.tile(np.array([N, N], np.int64), (B, C, 1)), "features": feats, "enabled_node_nums": enabled,
           "h1": np.asarray(h1), "h3": np.asarray(h3), "h4": np.asarray(h4), "gathered": np.asarray(g), "logits": logits,
           "top1": logits.argmax(1)}
    rec.update({"var:" + k: v for k, v in W.items()})
    np.savez_compressed(os.path.join(out_dir, "ckpt_reaction.npz"), **rec)
    print("ckpt_reaction: %d variables, logits %s, top-1 classes %s, atoms %s" % (len(W), logits.shape, rec["top1"].tolist(), enabled.tolist()))
    print("h3 range %.3f..%.3f  gathered range %.3f..%.3f" % (rec["h3"].min(), rec["h3"].max(), rec["gathered"].min(), rec["gathered"].max()))


if __name__ == "__main__":
    main()
