"""TensorFlow V2 checkpoint reader / writer and real-weights parity.

Pins.  (1) The reference's shipped checkpoint (model/reaction/model.best.ckpt.*, written by TensorFlow) parses with
every table-block and per-tensor CRC-32C that TensorFlow stored verifying, and writing the parsed content back
reproduces TensorFlow's ``.index`` and ``.data`` files BYTE FOR BYTE (authoring container; sha256 in the manifest).
(2) Without the reference tree: the trained variables committed in tests/golden/ckpt_reaction.npz re-serialise to
tensors whose masked CRC-32C equal the values TensorFlow stored (manifest).  (3) The network those weights belong to,
run by the reference's own legacy layer classes (oracle/make_ckpt_golden.py), is reproduced by the oracle bit for bit
and by the CUDA path -- through a TF-style model file restored by variable NAME -- within the stated tolerance."""
import hashlib
import json
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_golden, unflatten_adjs
from kgcn_b200 import tf_checkpoint as ckpt
from oracle import ref_layers as R

REF_PREFIX = "/root/reference/model/reaction/model.best.ckpt"
needs_reference = pytest.mark.skipif(not os.path.exists(REF_PREFIX + ".index"), reason="reference tree not present")


def manifest():
    return json.load(open(os.path.join(GOLDEN, "ckpt_reaction_manifest.json")))


def golden_variables():
    g = load_golden("ckpt_reaction")
    return g, {k[4:]: v for k, v in g.items() if k.startswith("var:")}


@needs_reference
def test_reference_checkpoint_parses_and_rewrites_byte_for_byte(tmp_path):
    reader = ckpt.load_checkpoint(REF_PREFIX)                       # verify_crc=True: block + tensor checksums
    m = manifest()
    assert reader.num_shards == 1 and len(reader.entries) == 56 == len(m["entries"])
    for e in m["entries"]:
        got = reader.entries[e["name"]]
        assert (np.dtype(got.dtype).name, got.shape, got.offset, got.size, got.crc32c) == \
            (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c_masked"])
    shapes = reader.get_variable_to_shape_map()
    assert shapes["rollout/graph_conv/kernel0"] == [75, 128] and shapes["rollout/graph_conv_2/bias0"] == [1, 128]
    assert shapes["beta1_power"] == [] and reader.get_tensor("beta1_power").shape == ()
    assert reader.has_tensor("rollout/dense/kernel/Adam_1") and not reader.has_tensor("rollout/dense/kernel/adam")
    trained = reader.tensors()
    assert len(trained) == 22 and all("Adam" not in k for k in trained)
    g, want = golden_variables()
    assert sorted(want) == sorted(trained) and all(np.array_equal(trained[k], want[k]) for k in want)
    ckpt.save_checkpoint(str(tmp_path / "again"), {n: reader.get_tensor(n) for n in reader.entries})
    for ext in (".index", ".data-00000-of-00001"):
        ours = open(str(tmp_path / "again") + ext, "rb").read()
        assert ours == open(REF_PREFIX + ext, "rb").read()
        assert hashlib.sha256(ours).hexdigest() == m["sha256" + ext]


def test_product_writer_and_reader_against_the_oracles_independent_bundle_reader(tmp_path):
    """oracle/tf_bundle_min.py shares no code with kgcn_b200/tf_checkpoint.py: what the product writes must parse there
    (table blocks, footer, protos, both CRC layers), and on the reference's own file both readers must agree."""
    from oracle import tf_bundle_min
    g, variables = golden_variables()
    prefix = str(tmp_path / "model.ckpt")
    ckpt.save_checkpoint(prefix, variables)
    header, entries = tf_bundle_min.read_index(prefix)
    tensors = tf_bundle_min.read_tensors(prefix)                     # verifies block and tensor CRC-32C with its own table
    stored = {e["name"]: e for e in manifest()["entries"]}
    assert header["num_shards"] == 1 and sorted(tensors) == sorted(variables)
    for name, value in variables.items():
        assert tensors[name].dtype == value.dtype and np.array_equal(tensors[name], value)
        assert entries[name]["crc32c"] == stored[name]["crc32c_masked"] and list(entries[name]["shape"]) == list(value.shape)
    assert tf_bundle_min.crc32c(b"123456789") == 0xE3069283          # RFC 3720 check value
    if os.path.exists(REF_PREFIX + ".index"):
        theirs, reader = tf_bundle_min.read_tensors(REF_PREFIX), ckpt.load_checkpoint(REF_PREFIX)
        assert list(theirs) == list(reader.entries)
        assert all(np.array_equal(theirs[n], reader.get_tensor(n)) for n in theirs)


def test_trained_variables_roundtrip_and_match_tensorflows_checksums(tmp_path):
    g, variables = golden_variables()
    prefix = str(tmp_path / "model.ckpt")
    ckpt.save_checkpoint(prefix, variables)
    reader = ckpt.load_checkpoint(prefix)
    stored = {e["name"]: e for e in manifest()["entries"]}
    assert sorted(reader.entries) == sorted(variables)
    for name, value in variables.items():
        got = reader.get_tensor(name)
        assert got.dtype == value.dtype and got.shape == value.shape and np.array_equal(got, value)
        assert reader.entries[name].crc32c == stored[name]["crc32c_masked"]          # the CRC TensorFlow computed
        assert reader.entries[name].size == stored[name]["size"] and reader.entries[name].shape == stored[name]["shape"]
    with pytest.raises(KeyError):
        reader.get_tensor("rollout/nope")


def test_other_dtypes_scalars_and_multi_block_tables(tmp_path, monkeypatch):
    monkeypatch.setattr(ckpt, "BLOCK_SIZE", 256)                     # many data blocks -> separators in the index block
    rng = np.random.default_rng(0)
    tensors = {"layer_%03d/%s" % (i, leaf): rng.standard_normal((3, i % 5 + 1)).astype(np.float32)
               for i in range(60) for leaf in ("kernel", "bias")}
    tensors.update({"global_step": np.array(7, np.int64), "flags": np.array([True, False]), "half": np.arange(4, dtype=np.float16),
                    "empty": np.zeros((0, 3), np.float32), "d": rng.standard_normal(5), "\xff\xfe": np.arange(3, dtype=np.int32),
                    "layer_003": np.arange(2, dtype=np.uint8)})
    prefix = str(tmp_path / "many")
    ckpt.save_checkpoint(prefix, tensors)
    pairs = ckpt.read_table(prefix + ".index")
    assert [k for k, _ in pairs] == [b""] + sorted(n.encode("utf-8") for n in tensors)
    reader = ckpt.load_checkpoint(prefix)
    for name, value in tensors.items():
        got = reader.get_tensor(name)
        assert got.dtype == np.asarray(value).dtype and got.shape == np.asarray(value).shape and np.array_equal(got, value)
    assert ckpt._shortest_separator(b"abcdef", b"abzz") == b"abd" and ckpt._shortest_separator(b"ab", b"abc") == b"ab"
    assert ckpt._shortest_separator(b"a\xffb", b"b") == b"a\xffb" and ckpt._short_successor(b"\xff\xffq") == b"\xff\xffr"


def test_corruption_is_detected(tmp_path):
    _, variables = golden_variables()
    small = {k: variables[k] for k in ("rollout/graph_conv/bias0", "rollout/graph_dense/bias")}
    prefix = str(tmp_path / "c")
    ckpt.save_checkpoint(prefix, small)
    index, data = open(prefix + ".index", "rb").read(), open(prefix + ".data-00000-of-00001", "rb").read()

    def variant(name, new_index=None, new_data=None):
        p = str(tmp_path / name)
        open(p + ".index", "wb").write(index if new_index is None else new_index)
        open(p + ".data-00000-of-00001", "wb").write(data if new_data is None else new_data)
        return p

    with pytest.raises(ckpt.CheckpointError, match="tensor bytes fail their CRC-32C"):
        ckpt.load_checkpoint(variant("flip_data", new_data=data[:100] + bytes([data[100] ^ 4]) + data[101:])).get_tensor("rollout/graph_conv/bias0")
    flipped = variant("flip_data2", new_data=data[:100] + bytes([data[100] ^ 4]) + data[101:])
    assert ckpt.load_checkpoint(flipped, verify_crc=False).get_tensor("rollout/graph_conv/bias0").shape == (1, 128)
    with pytest.raises(ckpt.CheckpointError, match="block at byte 0 fails its CRC-32C"):
        ckpt.load_checkpoint(variant("flip_index", new_index=index[:20] + bytes([index[20] ^ 1]) + index[21:]))
    with pytest.raises(ckpt.CheckpointError, match="bad table magic"):
        ckpt.load_checkpoint(variant("magic", new_index=index[:-1] + b"\x00"))
    with pytest.raises(ckpt.CheckpointError, match="truncated"):
        ckpt.load_checkpoint(variant("short", new_data=data[:600])).get_tensor("rollout/graph_dense/bias")
    with pytest.raises(FileNotFoundError):
        ckpt.load_checkpoint(str(tmp_path / "absent"))


def oracle_forward(g, W):
    adjs = unflatten_adjs(g, "adj_")
    n = g["enabled_node_nums"]
    h, hidden = g["features"], []
    for conv, bn in (("graph_conv", "batch_normalization"), ("graph_conv_1", "batch_normalization_1"), ("graph_conv_2", "batch_normalization_2")):
        h = R.graph_conv(h, adjs, [W["rollout/%s/kernel0" % conv]], [W["rollout/%s/bias0" % conv]])
        h, _, _ = R.graph_batch_normalization(h, W["rollout/%s/gamma" % bn], W["rollout/%s/beta" % bn], None, None, n, batch_statistics=True)
        h = np.maximum(h, 0)
        hidden.append(h)
    h4 = np.maximum(R.graph_dense(h, W["rollout/graph_dense/kernel"], W["rollout/graph_dense/bias"]), 0)
    gathered = R.graph_gather(h4)
    logits = (gathered @ W["rollout/dense/kernel"] + W["rollout/dense/bias"]).astype(np.float32)
    return hidden, h4, gathered, logits


def test_oracle_reproduces_reference_layers_with_trained_weights():
    g, W = golden_variables()
    hidden, h4, gathered, logits = oracle_forward(g, W)
    np.testing.assert_array_equal(hidden[0], g["h1"])
    np.testing.assert_array_equal(hidden[2], g["h3"])
    np.testing.assert_array_equal(h4, g["h4"])
    np.testing.assert_array_equal(gathered, g["gathered"])
    np.testing.assert_array_equal(logits, g["logits"])
    n = g["enabled_node_nums"]
    assert all((g["h3"][b, n[b]:] == 0).all() for b in range(len(n)))        # padded atoms stay exact zeros through BN


def test_variable_store_restores_by_name_before_and_after_creation():
    import torch
    from kgcn_b200.compat import facade
    store = facade.VariableStore()
    make = lambda: torch.nn.Parameter(torch.zeros(2, 3))
    store._scope.append("rollout")
    assert store.layer_name("GraphConv") == "rollout/graph_conv" and store.layer_name("GraphConv") == "rollout/graph_conv_1"
    assert store.layer_name("BatchNormalization") == "rollout/batch_normalization" and store.scoped("my_bn") == "rollout/my_bn"
    # Keras' to_snake_case keeps acronyms together: TF names example_model/model_gin.py's layer scope "gin_aggregate"
    assert store.layer_name("GINAggregate") == "rollout/gin_aggregate" and store.layer_name("GINAggregate") == "rollout/gin_aggregate_1"
    assert store.layer_name("GraphBatchNormalization") == "rollout/graph_batch_normalization"
    store.initial_values = {"rollout/graph_conv/kernel0": np.arange(6, dtype=np.float32).reshape(2, 3)}
    p = store.get("rollout/graph_conv/kernel0", make)                            # created -> pending value applied
    assert p.detach().numpy().tolist() == [[0, 1, 2], [3, 4, 5]] and not store.initial_values
    store.assign("rollout/graph_conv/kernel0", np.ones((2, 3), np.float32))
    assert float(store.get("rollout/graph_conv/kernel0", make).sum()) == 6.0
    with pytest.raises(ValueError, match="has shape"):
        store.assign("rollout/graph_conv/kernel0", np.ones((3, 2), np.float32))


@pytest.mark.gpu
def test_tf_style_model_restored_from_checkpoint_matches_reference_layers(tmp_path):
    """A TF-style model file + saver.restore by variable name + the CUDA path == the reference's layer classes run with
    the reference's trained weights.  Tolerance: 5e-4 of max|ref| on hidden activations and logits (three
    3xTF32 GraphConv + batch-statistics normalisation blocks; the measured error is printed)."""
    import torch
    from kgcn_b200 import compat, feed
    g, W = golden_variables()
    prefix = str(tmp_path / "model.best.ckpt")
    ckpt.save_checkpoint(prefix, W)                                              # == TensorFlow's bytes for these tensors (CPU tests)
    B, N, F = g["features"].shape
    L = g["logits"].shape[1]
    labels = np.eye(L, dtype=np.float32)[g["top1"]]
    data = {"adjs": unflatten_adjs(g, "adj_"), "features": g["features"], "labels": labels, "enabled_node_nums": g["enabled_node_nums"]}
    info = types.SimpleNamespace(adj_channel_num=1, graph_node_num=N, feature_dim=F, label_dim=L, feature_enabled=True)
    compat.install()
    try:
        runner = compat.ModelRunner("models.tf_style_rxn:Rollout", info, {}, B, search_path=os.path.join(ROOT, "tests"))
        found = runner.restore(prefix)                                           # before the first run: applied on creation
        assert len(found) == 22
        fd = feed.construct_feed(list(range(B)), runner.placeholders, data, batch_size=B, config={"task": "classification"})
        out = runner.run(fd)
        assert sorted(runner.named_variables()) == sorted(W)                     # TensorFlow's variable names, all 22
        assert sorted(set(runner.named_variables()) - set(runner.named_parameters())) == sorted(k for k in W if "moving_" in k)
        for k, v in runner.named_variables().items():
            if "moving_" not in k:                                               # (batch statistics update the moving averages)
                np.testing.assert_array_equal(v.detach().cpu().numpy(), W[k])
        logits = out["model"].out.detach().cpu().numpy()
        gathered = out["model"].gathered.detach().cpu().numpy()
        err_g = np.abs(gathered - g["gathered"]).max() / np.abs(g["gathered"]).max()
        err_l = np.abs(logits - g["logits"]).max() / np.abs(g["logits"]).max()
        print("real-weights parity: gathered %.2e, logits %.2e (relative to max)" % (err_g, err_l))
        assert err_g <= 5e-4 and err_l <= 5e-4
        margin = np.sort(g["logits"], 1)
        clear = (margin[:, -1] - margin[:, -2]) > 2e-3 * np.abs(g["logits"]).max()
        assert (logits.argmax(1)[clear] == g["top1"][clear]).all()
        assert float(out["metrics"]["correct_count"]) >= float(clear.sum())     # labels were set to the reference's top-1
        # saver.save -> a TensorFlow-format checkpoint again, same trained values
        runner.save(str(tmp_path / "resaved"))
        again = ckpt.load_checkpoint(str(tmp_path / "resaved")).tensors()
        assert sorted(again) == sorted(W) and all(np.array_equal(again[k], W[k]) for k in W if "moving_" not in k)
        # strict restore: a model variable the file lacks is an error (TF: NotFoundError)
        ckpt.save_checkpoint(str(tmp_path / "partial"), {k: v for k, v in W.items() if "graph_dense" not in k})
        with pytest.raises(KeyError, match="graph_dense"):
            runner.restore(str(tmp_path / "partial"))
    finally:
        for name in [m for m in sys.modules if m.startswith("models")]:
            sys.modules.pop(name, None)
        compat.uninstall()


def test_trainer_checkpoint_names_and_roundtrip(tmp_path):
    """The step loop's flat buffers <-> a TensorFlow-format checkpoint with TensorFlow's names (host logic only)."""
    import torch
    from kgcn_b200.trainer import NetSpec, Trainer
    spec = NetSpec(5, [6, 4], 7, channels=2, label_dim=3, dense_dim=8)
    a = Trainer(spec, 4, device="cpu", seed=1)
    g = torch.Generator().manual_seed(0)
    a.adam_m.copy_(torch.randn(a.n_params, generator=g))
    a.adam_v.copy_(torch.rand(a.n_params, generator=g))
    a.step_state[0] = 37
    prefix = str(tmp_path / "run" / "model.ckpt")
    os.makedirs(os.path.dirname(prefix))
    a.save_checkpoint(prefix, scope="net")
    reader = ckpt.load_checkpoint(prefix)
    shapes = reader.get_variable_to_shape_map()
    assert shapes["net/graph_conv/kernel0"] == [5, 6] and shapes["net/graph_conv/bias1"] == [1, 6]      # layers.py:54-61
    assert shapes["net/graph_conv_1/kernel1"] == [6, 4] and shapes["net/graph_dense/kernel"] == [4, 8]
    assert shapes["net/dense/kernel"] == [8, 3] and shapes["net/dense/bias/Adam_1"] == [3] and shapes["global_step"] == []
    assert len(shapes) == 3 * (2 * 2 * 2 + 4) + 3
    np.testing.assert_allclose(reader.get_tensor("beta1_power"), 0.9 ** 38, rtol=1e-5)
    np.testing.assert_array_equal(reader.get_tensor("net/graph_conv_1/kernel1"), a.views["conv1/kernel"][1].numpy())
    b = Trainer(spec, 4, device="cpu", seed=2)
    assert b.load_checkpoint(prefix, scope="net") == 37 and b.steps_done() == 37
    for name in a.views:          # (the flat buffers also hold alignment padding, which is not a variable)
        for x, y in ((a.views, b.views), (a.mviews, b.mviews), (a.vviews, b.vviews)):
            assert torch.equal(x[name], y[name]), name
    # weights only (e.g. a file written by tf.train.Saver(var_list=trainable)): moments and step start afresh
    a.save_checkpoint(prefix + ".weights", scope="net", with_slots=False)
    assert len(ckpt.load_checkpoint(prefix + ".weights").entries) == 2 * 2 * 2 + 4
    c = Trainer(spec, 4, device="cpu", seed=3)
    c.adam_m.fill_(1.0)
    assert c.load_checkpoint(prefix + ".weights", scope="net") == 0
    assert all(torch.equal(c.views[n], a.views[n]) for n in a.views) and float(c.adam_m.abs().sum()) == 0.0 and c.steps_done() == 0
    # without global_step the count is recovered from beta2_power (what TensorFlow's own Adam slots provide)
    tensors = {n: reader.get_tensor(n) for n in reader.entries if n != "global_step"}
    ckpt.save_checkpoint(prefix + ".tf", tensors)
    assert Trainer(spec, 4, device="cpu").load_checkpoint(prefix + ".tf", scope="net") == 37
    with pytest.raises(KeyError, match="other/graph_conv/kernel0"):
        b.load_checkpoint(prefix, scope="other")


@pytest.mark.gpu
def test_trainer_resumes_bit_exactly_from_checkpoint(tmp_path):
    """3 steps, save, 2 more steps == 3 steps, save | new process state: load, 2 steps -- bit for bit (the kernels are
    deterministic and the Adam step counter, moments and weights all travel through the file)."""
    import torch
    from kgcn_b200 import synth
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    rng = np.random.default_rng(4)
    B, N, F = 64, 32, 64
    spec = NetSpec(F, [64, 64], N)
    batches = []
    for _ in range(5):
        d = synth.ring_graphs(rng, B, N, F)
        batches.append(DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N))
    a = Trainer(spec, B, seed=7)
    for b in batches[:3]:
        a.step_eager(b)
    prefix = str(tmp_path / "model.ckpt")
    a.save_checkpoint(prefix)
    for b in batches[3:]:
        a.step_eager(b)
    r = Trainer(spec, B, seed=99)
    assert r.load_checkpoint(prefix) == 3
    for b in batches[3:]:
        r.step_eager(b)
    torch.cuda.synchronize()
    assert r.steps_done() == a.steps_done() == 5
    assert torch.equal(r.params, a.params) and torch.equal(r.adam_m, a.adam_m) and torch.equal(r.adam_v, a.adam_v)
    assert not torch.equal(r.params, Trainer(spec, B, seed=7).params)
