"""CPU: ``load_data`` / ``build_data`` / ``split_data`` / ``shuffle_data`` (kgcn/data_util.py:155-179, 368-644) are BIT-EXACT
against the reference's own functions run on its shipped fixtures (tests/golden/build_data_*.npz, written by
``python oracle/make_golden.py build_data``): every ``info`` field, every ``all_data`` member, and -- with the same
``np.random.seed`` -- the same train / validation split and the same shuffle."""
import json

import joblib
import numpy as np
import pytest

from conftest import load_golden, unflatten_adjs
from kgcn_b200 import data_util

BASE_CFG = {"with_feature": True, "with_node_embedding": False, "normalize_adj_flag": False, "split_adj_flag": False,
            "shuffle_data": False}
FIXTURES = [("sample", {}), ("sample_multiadj", {}), ("sample_multitask", {}), ("sample_node_label", {}), ("synthetic", {}),
            ("synthetic_norm", {"normalize_adj_flag": True})]


def raw_dict(rec):
    raw = {}
    for key in rec["raw_keys"].tolist():
        v = rec["in_" + key]
        raw[key] = v.item() if v.ndim == 0 else v
    return raw


def same_adjs(got, rec, prefix):
    want = unflatten_adjs(rec, prefix)
    assert len(got) == len(want)
    for g in range(len(want)):
        assert len(got[g]) == len(want[g])
        for c in range(len(want[g])):
            np.testing.assert_array_equal(np.asarray(got[g][c][0]).reshape(-1, 2), want[g][c][0])
            assert np.asarray(got[g][c][1], np.float32).tobytes() == want[g][c][1].tobytes()
            assert [int(got[g][c][2][0]), int(got[g][c][2][1])] == want[g][c][2]


def same_members(part, rec, prefix, keys):
    for k in keys:
        name = prefix + k
        if name in rec:
            got = np.asarray(part[k])
            assert got.dtype == rec[name].dtype and got.shape == rec[name].shape, k
            assert got.tobytes() == rec[name].tobytes(), k
        else:
            assert part[k] is None, k


@pytest.mark.parametrize("name,extra", FIXTURES)
def test_load_data_matches_reference(name, extra, tmp_path):
    rec = load_golden("build_data_" + name)
    path = str(tmp_path / (name + ".jbl"))
    joblib.dump(raw_dict(rec), path)
    all_data, info = data_util.load_data(dict(BASE_CFG, **extra), path, prohibit_shuffle=True, verbose=False)
    want = json.loads(str(rec["info_json"]))
    assert sorted(info.keys()) == want["info_keys"]
    assert sorted(all_data.keys()) == want["all_data_keys"]
    assert sorted(k for k in all_data if all_data[k] is None) == want["none_members"]
    for k in ("all_node_num", "feature_dim", "graph_node_num", "feature_enabled", "sequence_max_length", "sequence_symbol_num",
              "sequences_vec_dim", "graph_num", "adj_channel_num", "label_dim"):
        assert info[k] == want[k], k
    assert list(info.vector_modal_dim) == want["vector_modal_dim"] and dict(info.vector_modal_name) == want["vector_modal_name"]
    for k in ("pos_weight", "class_weight"):
        if "info_" + k in rec:
            assert np.asarray(info[k]).tobytes() == rec["info_" + k].tobytes(), k       # float64, bit for bit
        else:
            assert info.get(k) is None
    assert all_data.num == int(rec["all_num"]) and info.missing_attribute is None       # dotdict: None for unknown names
    same_members(all_data, rec, "all_", ("labels", "mask_label", "node_label", "mask_node_label", "sequences", "sequences_len",
                                         "enabled_node_nums", "features"))


@pytest.mark.parametrize("name,extra", FIXTURES)
def test_split_and_shuffle_follow_the_reference_generator(name, extra):
    rec = load_golden("build_data_" + name)
    cfg = dict(BASE_CFG, **extra)
    members = ("features", "labels", "mask_label", "node_label", "enabled_node_nums", "sequences_len")
    all_data, _ = data_util.build_data(cfg, raw_dict(rec), prohibit_shuffle=True, verbose=False)
    np.random.seed(7)
    train, valid = data_util.split_data(all_data, 0.4)
    for tag, part in (("train", train), ("valid", valid)):
        assert part.num == int(rec["split_%s_num" % tag])
        same_members(part, rec, "split_%s_" % tag, members)
        same_adjs(list(part.adjs), rec, "split_%s_adj_" % tag)
    assert train.num + valid.num == all_data.num and valid.num == int(all_data.num * 0.4)
    tr2, va2 = data_util.split_data(all_data, 0.4, indices_for_train_data=[2, 0], indices_for_valid_data=[1, 3])
    np.testing.assert_array_equal(tr2.enabled_node_nums, rec["split_explicit_train_enabled"])
    np.testing.assert_array_equal(va2.enabled_node_nums, rec["split_explicit_valid_enabled"])
    assert tr2.num == 2 and len(tr2.adjs) == 2
    # shuffle: through build_data's own switch (config["shuffle_data"]) with the reference's seed
    np.random.seed(3)
    shuffled, _ = data_util.build_data(dict(cfg, shuffle_data=True), raw_dict(rec), verbose=False)
    same_members(shuffled, rec, "shuffle_", members)
    same_adjs(list(shuffled.adjs), rec, "shuffle_adj_")
    np.random.seed(3)
    kept, _ = data_util.build_data(dict(cfg, shuffle_data=True), raw_dict(rec), prohibit_shuffle=True, verbose=False)
    same_members(kept, rec, "all_", ("labels", "enabled_node_nums"))                     # prohibit_shuffle wins


@pytest.mark.parametrize("name,extra", FIXTURES)
def test_node_level_feed_keys_match_reference(name, extra):
    """``mask_node`` (1 for a graph's real atoms), ``node_label``, ``mask_label`` of a short batch, as the reference's
    construct_feed builds them (feed.py:148-163, 209-218) from what build_data returned."""
    from kgcn_b200 import feed
    rec = load_golden("build_data_" + name)
    all_data, info = data_util.build_data(dict(BASE_CFG, **extra), raw_dict(rec), prohibit_shuffle=True, verbose=False)
    keys, bi = rec["feed_keys"].tolist(), rec["feed_batch_idx"].tolist()
    fd = feed.construct_feed(bi, keys, all_data, batch_size=len(bi) + 1, info=info, config={"task": "classification"})
    for k in keys:
        assert fd[k].dtype == rec["feed_" + k].dtype and fd[k].shape == rec["feed_" + k].shape, k
        assert fd[k].tobytes() == rec["feed_" + k].tobytes(), k
    assert (fd["mask_node"].sum(1) == fd["enabled_node_nums"]).all() and fd["mask_node"][-1].sum() == 0
    if all_data.mask_node_label is not None:      # the reference itself raises on its 3-D fixture mask (see feed.py here)
        got = feed.construct_feed(bi, ["mask_node_label"], all_data, batch_size=len(bi) + 1, info=info)["mask_node_label"]
        assert got.shape == (len(bi) + 1,) + np.shape(all_data.mask_node_label)[1:] and (got[-1] == 0).all()
        np.testing.assert_array_equal(got[:len(bi)], np.asarray(all_data.mask_node_label)[bi])


def test_build_data_without_graph_and_error_paths():
    rec = load_golden("build_data_sample")
    raw = raw_dict(rec)
    no_graph = {k: v for k, v in raw.items() if k not in ("dense_adj", "max_node_num")}
    # (graph-free data is the multimodal "vector only" case: with features present the reference's #graphs check fails too)
    all_data, info = data_util.build_data(dict(BASE_CFG, with_feature=False), dict(no_graph, vector_modal=np.ones((5, 7))), verbose=False)
    assert all_data.adjs is None and all_data.enabled_node_nums is None and all_data.num == 5      # data_util.py:424-428
    assert info.graph_num == 0 and info.adj_channel_num == 1 and info.vector_modal_dim == [7] and info.vector_modal_name == {"vector_modal": 0}
    with pytest.raises(data_util.DataLoadError, match="feature or node"):                            # :501-503
        data_util.build_data(dict(BASE_CFG, with_feature=False), raw, verbose=False)
    with pytest.raises(data_util.DataLoadError, match="differ"):                                     # :549-557
        data_util.build_data(dict(BASE_CFG), dict(raw, feature=raw["feature"][:3]), verbose=False)
    all_data, train, valid, info = data_util.build_and_split_data(dict(BASE_CFG), raw, valid_data_rate=0.2)
    assert (all_data.num, train.num, valid.num) == (5, 4, 1)
