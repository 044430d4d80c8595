#!/usr/bin/env python
"""Golden outputs of BASELINE config 1 -- ``example_config/sample.json`` -> ``example_model.model:GCN`` on
``example_jbl/synthetic.jbl`` -- produced by the reference's OWN files end to end: ``kgcn/data_util.load_data`` ->
``GCN.build_placeholders`` (``kgcn/default_model.py``) -> ``kgcn/feed.construct_feed`` -> ``GCN.build_model``
(``example_model/model.py:30-73``, which instantiates ``kgcn/layers.py``'s GraphConv x3, GraphBatchNormalization,
GraphDense, GraphGather and a Keras Dense) -- the way ``CoreModel.build`` / ``fit`` drive them (``kgcn/core.py:138-166,
247-281``), executed unchanged under the numpy-eager TensorFlow stand-in (``oracle/tf_numpy.py``: control flow the
reference's, TF primitives numpy restatements).

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs ``/root/reference``):

    python oracle/make_c1_golden.py     ->  tests/golden/c1_sample_json.npz  +  tests/fixtures/c1/ (three INPUT files)

The three files the config names are USER INPUTS of the drop-in (a model definition, a config, a dataset); they are
copied byte for byte to ``tests/fixtures/c1/`` so the GPU box, which has no ``/root/reference``, can load the same model
file through ``kgcn_b200.compat.ModelRunner``, the same dataset through ``kgcn_b200.data_util.load_data`` and feed it
through ``kgcn_b200.feed.construct_feed`` (``tests/test_c1_reference_model.py``).

Stored: the variables the reference layers created (TensorFlow names), and for three batches of ten molecules (two full,
one short: 7 of 10, the tail of the dataset) the fed features / labels / mask / enabled_node_nums and the model's
logits, prediction, cost_opt, cost_sum and correct_count.
"""
import contextlib
import io
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("KGCN_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import tf_numpy  # noqa: E402

tf = tf_numpy.install_model_level()
sys.path.insert(0, REF)
import kgcn.data_util as du  # noqa: E402  (the reference's modules)
import kgcn.feed as rfeed  # noqa: E402
import example_model.model as ref_model  # noqa: E402

FILES = ["example_model/model.py", "example_config/sample.json", "example_jbl/synthetic.jbl"]
# gcn.py:84-129 get_default_config, the keys this path reads
DEFAULTS = {"with_feature": True, "with_node_embedding": False, "embedding_dim": 10, "normalize_adj_flag": False,
            "split_adj_flag": False, "order": 1, "shuffle_data": False, "task": "classification"}
BATCHES = [list(range(0, 10)), list(range(10, 20)), list(range(193, 200))]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def bind(placeholders, fd):
    """The values sess.run would substitute for each placeholder (unfed ones -> None)."""
    out = {}
    for key, ph in placeholders.items():
        if key == "adjs":
            out[key] = [[tf.SparseTensorValue(*fd[p]) for p in row] for row in ph]
        else:
            v = fd.get(ph)
            out[key] = tf_numpy.T(v) if isinstance(v, np.ndarray) and v.dtype.kind == "f" else v
    return out


def main(out_root=None):
    """``out_root``: write under this directory instead of the repository (tests regenerate into a scratch directory)."""
    root = out_root or ROOT
    fix = os.path.join(root, "tests", "fixtures", "c1")
    for rel in FILES:
        dst = os.path.join(fix, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        os.chmod(dst, 0o644)
    with open(os.path.join(REF, "example_config", "sample.json")) as f:
        config = dict(DEFAULTS, **json.load(f))
    batch_size = int(config["batch_size"])
    all_data, info = quiet(du.load_data, config, os.path.join(REF, config["dataset"]), prohibit_shuffle=True)

    tf_numpy.seed(20261017)
    model = ref_model.GCN()
    placeholders = model.build_placeholders(info, config, batch_size)
    rec = {"batch_size": np.int64(batch_size), "n_batches": np.int64(len(BATCHES)), "num": np.int64(all_data.num)}
    # one throw-away execution creates the variables (glorot kernels, zero biases, gamma 1 / beta 0); the zero / one
    # initial values are then perturbed so the goldens also pin where every bias, gamma and beta enters
    fd = rfeed.construct_feed(BATCHES[0], placeholders, all_data, batch_size=batch_size, info=info, config=config)
    model.build_model(bind(placeholders, fd), info, config, batch_size)
    rng = np.random.default_rng(7)
    for name, value in _VARIABLES.items():
        if name.endswith("gamma"):
            value[...] = 1.0 + 0.2 * rng.standard_normal(value.shape)
        elif "bias" in name or name.endswith("beta"):
            value[...] = 0.3 * rng.standard_normal(value.shape)
    for k, idx in enumerate(BATCHES):
        fd = rfeed.construct_feed(idx, placeholders, all_data, batch_size=batch_size, dropout_rate=0.0, is_train=False,
                                  info=info, config=config)
        bound = bind(placeholders, fd)
        _, prediction, cost_opt, cost_sum, metrics = model.build_model(bound, info, config, batch_size)
        rec.update({"b%d_idx" % k: np.asarray(idx, np.int64), "b%d_features" % k: np.asarray(bound["features"], np.float32),
                    "b%d_labels" % k: np.asarray(bound["labels"]), "b%d_mask" % k: np.asarray(bound["mask"], np.float32),
                    "b%d_enabled_node_nums" % k: np.asarray(bound["enabled_node_nums"]),
                    "b%d_logits" % k: np.asarray(model.out, np.float32), "b%d_prediction" % k: np.asarray(prediction, np.float32),
                    "b%d_cost_opt" % k: np.float32(cost_opt), "b%d_cost_sum" % k: np.float32(cost_sum),
                    "b%d_correct_count" % k: np.float32(metrics["correct_count"])})
    for name, value in _VARIABLES.items():
        rec["var/" + name] = np.asarray(value, np.float32)
    out = os.path.join(root, "tests", "golden", "c1_sample_json.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, **rec)
    print("wrote", out, "variables:", sorted(_VARIABLES))
    for k in range(len(BATCHES)):
        print("batch", k, "cost_sum", rec["b%d_cost_sum" % k], "correct", rec["b%d_correct_count" % k])


# ---- variable reuse across eager executions, under TensorFlow's names -------------------------------------------
# Graph-mode TF builds the model once; the eager stand-in re-executes build_model per batch, so every execution must
# see the SAME variables.  The stand-in's Layer.add_weight is wrapped: the k-th layer of a class gets TF's scope name
# (graph_conv, graph_conv_1, ...; Keras to_snake_case) and its variables are created once and then re-served.
_LAYERS, _VARIABLES, _COUNTS = [], {}, {}
_orig_call, _orig_add = tf_numpy.Layer.__call__, tf_numpy.Layer.add_weight


def _snake(name):
    import re
    return re.sub(r"([a-z])([A-Z])", r"\1_\2", re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)).lower()


def _call(self, inputs, *a, **k):
    if not getattr(self, "_scope", None):
        base = _snake(type(self).__name__)
        n = _COUNTS.get(base, 0)
        _COUNTS[base] = n + 1
        self._scope = base if n == 0 else "%s_%d" % (base, n)
        _LAYERS.append(self)
    return _orig_call(self, inputs, *a, **k)


def _add_weight(self, name=None, shape=(), initializer="zeros", trainable=True, **kw):
    key = "%s/%s" % (self._scope, name)
    if key not in _VARIABLES:
        _VARIABLES[key] = _orig_add(self, name=name, shape=shape, initializer=initializer, trainable=trainable, **kw)
    return _VARIABLES[key]


tf_numpy.Layer.__call__ = _call
tf_numpy.Layer.add_weight = _add_weight
_build_model = ref_model.GCN.build_model


def _build_model_fresh_scope(self, *a, **k):
    _COUNTS.clear()                  # each execution names its layers from graph_conv again -> same variables
    return _build_model(self, *a, **k)


ref_model.GCN.build_model = _build_model_fresh_scope

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
