"""BASELINE config 1 end to end on the reference's own files: ``example_config/sample.json`` ->
``example_model.model:GCN`` on ``example_jbl/synthetic.jbl`` (tests/fixtures/c1/, byte copies made by
oracle/make_c1_golden.py, which also ran them through the reference's data_util / feed / layers under the numpy TF
stand-in and stored the outputs in tests/golden/c1_sample_json.npz).

The GPU test drives them the way kgcn/core.py does: load_data -> build_placeholders -> construct_feed -> build_model
(``ModelRunner.run`` = ``sess.run([prediction, cost, metrics])``, core.py:276-281) and three optimizer steps
(``ModelRunner.train_step`` = ``sess.run([train_step, cost_sum, metrics])``, core.py:267-269, Adam of core.py:121-127).
"""
import filecmp
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import ref_layers as R

FIX = os.path.join(ROOT, "tests", "fixtures", "c1")
REFERENCE = "/root/reference"
FILES = ["example_model/model.py", "example_config/sample.json", "example_jbl/synthetic.jbl"]
# gcn.py:84-129 get_default_config, the keys this path reads
DEFAULTS = {"with_feature": True, "with_node_embedding": False, "embedding_dim": 10, "normalize_adj_flag": False,
            "split_adj_flag": False, "order": 1, "shuffle_data": False, "task": "classification"}


def load_config():
    with open(os.path.join(FIX, "example_config", "sample.json")) as f:
        return dict(DEFAULTS, **json.load(f))


def ingest():
    from kgcn_b200 import data_util
    config = load_config()
    all_data, info = data_util.load_data(config, os.path.join(FIX, config["dataset"]), prohibit_shuffle=True, verbose=False)
    return config, all_data, info


def oracle_logits(P, fd):
    """example_model/model.py:40-56 on the oracle's layer functions (inference-mode BN, Dropout = identity)."""
    adjs, x = fd["adjs"], fd["features"]
    h = R.activation(R.graph_conv(x, adjs, [P["graph_conv/kernel0"]], [P["graph_conv/bias0"]]), "sigmoid")
    h = R.activation(R.graph_conv(h, adjs, [P["graph_conv_1/kernel0"]], [P["graph_conv_1/bias0"]]), "sigmoid")
    h = R.graph_conv(h, adjs, [P["graph_conv_2/kernel0"]], [P["graph_conv_2/bias0"]])
    f = h.shape[-1]
    h = R.graph_batch_normalization(h, P["batch_normalization/gamma"], P["batch_normalization/beta"], np.zeros(f, np.float32),
                                    np.ones(f, np.float32), enabled_node_nums=fd["enabled_node_nums"])[0]
    h = R.activation(h, "sigmoid")
    h = R.graph_dense(h, P["graph_dense/kernel"], P["graph_dense/bias"], act="sigmoid")
    return R.graph_gather(h) @ P["dense/kernel"] + P["dense/bias"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "example_model")), reason="reference tree not present")
def test_fixture_files_are_the_reference_files():
    for rel in FILES:
        assert filecmp.cmp(os.path.join(FIX, rel), os.path.join(REFERENCE, rel), shallow=False), rel


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "example_model")), reason="reference tree not present")
def test_golden_regenerates_identically_from_the_reference(tmp_path):
    """oracle/make_c1_golden.py run again on /root/reference (in a fresh interpreter: it installs the numpy TensorFlow stand-in on
    sys.modules) reproduces every array of the committed golden bit for bit."""
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "make_c1_golden.py"), str(tmp_path)], check=True, stdout=subprocess.DEVNULL)
    new = np.load(os.path.join(str(tmp_path), "tests", "golden", "c1_sample_json.npz"))
    old = load_golden("c1_sample_json")
    assert sorted(new.files) == sorted(old)
    for k in old:
        assert new[k].dtype == old[k].dtype and np.array_equal(new[k], old[k]), k


def test_own_ingest_feeds_what_the_reference_fed_and_oracle_reproduces_the_model():
    """kgcn_b200's load_data + DefaultModel placeholders + construct_feed give the arrays the reference fed (bit for
    bit); the oracle's layer functions on that feed reproduce the reference model's logits."""
    from kgcn_b200 import feed
    from kgcn_b200.default_model import DefaultModel
    g = load_golden("c1_sample_json")
    config, all_data, info = ingest()
    B = int(g["batch_size"])
    assert all_data.num == int(g["num"]) and B == config["batch_size"]
    keys = ["adjs", "nodes", "labels", "mask", "dropout_rate", "enabled_node_nums", "is_train", "features"]
    ph = DefaultModel().get_placeholders(info, config, B, keys)
    P = {k[4:]: g[k] for k in g if k.startswith("var/")}
    for k in range(int(g["n_batches"])):
        fd = feed.construct_feed(list(g["b%d_idx" % k]), ph, all_data, batch_size=B, info=info, config=config)
        for name in ("features", "labels", "mask", "enabled_node_nums"):
            got, want = np.asarray(fd[name]), g["b%d_%s" % (k, name)]
            assert got.dtype == want.dtype and np.array_equal(got, want), (k, name)
        logits = oracle_logits(P, fd)
        np.testing.assert_allclose(logits, g["b%d_logits" % k], rtol=2e-6, atol=2e-6)


def torch_reference_trajectory(P0, feeds, lr, steps):
    """float64 torch-CPU autograd of the same network + TensorFlow's Adam: cost_sum before each update, final variables."""
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P0.items()}
    m = {k: torch.zeros_like(v) for k, v in P.items()}
    v2 = {k: torch.zeros_like(v) for k, v in P.items()}
    costs = []
    for t in range(1, steps + 1):
        fd = feeds[(t - 1) % len(feeds)]
        B, N = fd["features"].shape[:2]
        A = torch.zeros(B, N, N, dtype=torch.float64)
        for b in range(B):
            idx, val, _ = fd["adjs"][b][0]
            idx = np.asarray(idx).reshape(-1, 2)
            A[b].index_put_((torch.as_tensor(idx[:, 0]).long(), torch.as_tensor(idx[:, 1]).long()), torch.as_tensor(val, dtype=torch.float64),
                            accumulate=True)
        h = torch.tensor(fd["features"], dtype=torch.float64)
        for i, name in enumerate(["graph_conv", "graph_conv_1", "graph_conv_2"]):
            h = A @ (h @ P[name + "/kernel0"] + P[name + "/bias0"])
            if i < 2:
                h = torch.sigmoid(h)
        keep = (torch.arange(N)[None, :] < torch.as_tensor(fd["enabled_node_nums"])[:, None]).double()[:, :, None]
        h = (h / np.sqrt(1.0 + 1e-3) * P["batch_normalization/gamma"] + P["batch_normalization/beta"]) * keep
        h = torch.sigmoid(h)
        h = torch.sigmoid(h @ P["graph_dense/kernel"] + P["graph_dense/bias"])
        logits = h.sum(1) @ P["dense/kernel"] + P["dense/bias"]
        cost = torch.as_tensor(fd["mask"], dtype=torch.float64) * \
            -(torch.as_tensor(fd["labels"], dtype=torch.float64) * torch.log_softmax(logits, 1)).sum(1)
        costs.append(float(cost.sum()))
        grads = torch.autograd.grad(cost.mean(), list(P.values()))
        with torch.no_grad():
            for (k, p), gk in zip(P.items(), grads):
                m[k] = 0.9 * m[k] + 0.1 * gk
                v2[k] = 0.999 * v2[k] + 0.001 * gk * gk
                p -= lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t) * m[k] / (v2[k].sqrt() + 1e-8)
    return costs, {k: p.detach().numpy() for k, p in P.items()}


@pytest.mark.gpu
def test_reference_model_file_runs_and_trains_on_gpu():
    from kgcn_b200 import compat, feed, ops
    g = load_golden("c1_sample_json")
    config, all_data, info = ingest()
    B = int(g["batch_size"])
    P0 = {k[4:]: g[k] for k in g if k.startswith("var/")}
    launched = []
    real_apply = ops.GraphConvFunction.apply

    def spy(x, w, bias, csr, act, flags):
        launched.append(int(act))
        return real_apply(x, w, bias, csr, act, flags)

    runner = compat.ModelRunner(config["model.py"], info, config, B, search_path=FIX)
    try:
        ops.GraphConvFunction.apply = spy
        assert sorted(runner.placeholders) == sorted(["adjs", "nodes", "labels", "mask", "dropout_rate", "enabled_node_nums", "is_train",
                                                      "features"])
        runner.store.initial_values = {k: v.copy() for k, v in P0.items()}
        feeds = []
        for k in range(int(g["n_batches"])):
            fd = feed.construct_feed(list(g["b%d_idx" % k]), runner.placeholders, all_data, batch_size=B, info=info, config=config)
            feeds.append(fd)
            del launched[:]
            out = runner.run(fd)
            # the tf.sigmoid after GraphConv 1 and 2 (model.py:42-45) ran inside the layer launches, the third has none
            assert launched == [ops.act_id("sigmoid"), ops.act_id("sigmoid"), ops.act_id(None)]
            scale = np.abs(g["b%d_logits" % k]).max()
            np.testing.assert_allclose(out["model"].out.detach().cpu().numpy(), g["b%d_logits" % k], rtol=1e-5, atol=1e-5 * scale)
            np.testing.assert_allclose(out["prediction"].detach().cpu().numpy(), g["b%d_prediction" % k], rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(float(out["cost_sum"].detach()), float(g["b%d_cost_sum" % k]), rtol=1e-5)
            np.testing.assert_allclose(float(out["cost_opt"].detach()), float(g["b%d_cost_opt" % k]), rtol=1e-5)
            assert float(out["metrics"]["correct_count"]) == float(g["b%d_correct_count" % k])
        assert sorted(runner.named_parameters()) == sorted(P0)
        # three optimizer steps, the third on the short batch (7 of 10 molecules, mask 0 / no enabled rows on the padding)
        steps = 3
        order = [feeds[0], feeds[1], feeds[2]]
        want_costs, want_P = torch_reference_trajectory(P0, order, float(config["learning_rate"]), steps)
        got_costs = [float(runner.train_step(order[t])["cost_sum"]) for t in range(steps)]
        np.testing.assert_allclose(got_costs, want_costs, rtol=2e-4)
        for k, p in runner.named_parameters().items():
            want = want_P[k]
            np.testing.assert_allclose(p.detach().cpu().numpy(), want, rtol=0, atol=2e-4 * max(1.0, np.abs(want).max()), err_msg=k)
            assert np.abs(p.detach().cpu().numpy() - P0[k]).max() > 1e-3, k       # every variable was updated
    finally:
        ops.GraphConvFunction.apply = real_apply
        for name in [m for m in sys.modules if m.startswith("example_model")]:
            sys.modules.pop(name, None)
        compat.uninstall()
