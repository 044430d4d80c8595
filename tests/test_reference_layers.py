"""Golden LAYER outputs produced by the reference's own kgcn/layers.py (run unchanged under oracle/tf_numpy.py by
oracle/make_layer_golden.py, on batches ingested by the reference's own data_util/feed code): the CPU oracle must
reproduce them (it is what every other parity test compares against), and so must the CUDA path (-m gpu)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, unflatten_adjs
from oracle import ref_layers as R

FWD_TOL = 1e-5   # |got - ref| <= FWD_TOL * |ref| + FWD_TOL * max|ref|   ("Precision")


def close(got, ref, tol=FWD_TOL):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * max(float(np.abs(ref).max()), 1e-30))


def sig(x):
    return R.activation(x, "sigmoid")


# ------------------------------------------------------------------------------------------- CPU: oracle == reference
def test_reference_code_confirms_the_known_answer_vectors():
    """kgcn/layers.py itself yields KAT1 / KAT2 of SURVEY.md Appendix B (they were derived by hand before)."""
    g, kat = load_golden("layers_sample_kat"), load_golden("kat")
    np.testing.assert_array_equal(g["y"], kat["kat1_y"])
    np.testing.assert_array_equal(g["gather"], kat["kat1_gather"])
    assert (g["y"][1, 2] == 0).all() and (g["y"][4, 1] == 0).all()          # isolated nodes never see the bias
    np.testing.assert_array_equal(load_golden("layers_multiadj_kat")["y"], kat["kat2_y"])


@pytest.mark.parametrize("name", ["layers_sample_kat", "layers_multiadj_kat", "layers_random"])
def test_oracle_graphconv_equals_reference_layer(name):
    g = load_golden(name)
    adjs = unflatten_adjs(g, "adj_")
    C = g["w"].shape[0]
    y = R.graph_conv(g["features"], adjs, [g["w"][c] for c in range(C)], [g["bias"][c] for c in range(C)])
    np.testing.assert_array_equal(y, g["y"])                                 # same primitives, same order: bit-equal
    close(R.graph_conv(g["features"], adjs, [g["w"][c] for c in range(C)], [g["bias"][c] for c in range(C)], fast=True), g["y"], 2e-6)
    if "gather" in g:
        np.testing.assert_array_equal(R.graph_gather(y), g["gather"])


def test_oracle_stack_equals_reference_layers():
    """example_model/model.py:41-56 in miniature: conv, conv, masked BN, GraphDense, GraphGather."""
    g = load_golden("layers_synthetic_stack")
    adjs, n = unflatten_adjs(g, "adj_"), g["enabled_node_nums"]
    h1 = sig(R.graph_conv(g["features"], adjs, [g["w1"][0]], [g["b1"][0]]))
    np.testing.assert_array_equal(h1, g["h1"])
    h2 = R.graph_conv(h1, adjs, [g["w2"][0]], [g["b2"][0]])
    np.testing.assert_array_equal(h2, g["h2"])
    F = h2.shape[2]
    bn, _, _ = R.graph_batch_normalization(h2, np.ones(F), np.zeros(F), np.zeros(F), np.ones(F), n)
    close(sig(bn), g["h3"], 1e-6)
    assert (g["h3"][7:] == 0.5).all()                                        # padded graphs: sigmoid(0)
    close(R.graph_dense(g["h3"], g["gd_kernel"], g["gd_bias"], act="sigmoid"), g["h4"], 1e-6)
    close(R.graph_gather(g["h4"]), g["gathered"], 1e-6)
    h5 = R.graph_dense(g["h3"], g["gd2_kernel"], g["gd2_bias"], enabled_node_nums=n)
    close(h5, g["h5"], 1e-6)
    assert (g["h5"][7:] == 0).all()                                          # masked GraphDense: exact zeros (layers.py:249-253)
    legacy, _, _ = R.graph_batch_normalization(h2, np.ones(F), np.zeros(F), None, None, n, batch_statistics=True)
    close(legacy, g["h2_bn_legacy"], 1e-5)


def test_oracle_gin_maxpool_blockdiag_equal_reference_layers():
    g = load_golden("layers_random")
    adjs = unflatten_adjs(g, "adj_")
    np.testing.assert_array_equal(R.gin_aggregate(g["features"], adjs, g["gin_eps"]), g["gin_y"])
    np.testing.assert_array_equal(R.graph_max_pooling(g["features"], unflatten_adjs(g, "uadj_")), g["maxpool_y"])
    B, N, F = g["features"].shape
    idx = np.concatenate([adjs[b][0][0] + b * N for b in range(B)])
    val = np.concatenate([adjs[b][0][1] for b in range(B)])
    y = R.batch_graph_conv(g["features"].reshape(B * N, F), (idx, val, [B * N, B * N]), g["bgc_w"], g["bgc_bias"])
    np.testing.assert_array_equal(y, g["bgc_y"])


# ------------------------------------------------------------------------------------------- GPU: CUDA == reference
def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def load_conv(layer, w, b):
    with torch.no_grad():
        for c in range(w.shape[0]):
            layer.w[c].copy_(dev(w[c]))
            layer.bias[c].copy_(dev(b[c]).reshape(layer.bias[c].shape))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["layers_sample_kat", "layers_multiadj_kat", "layers_random"])
def test_cuda_graphconv_equals_reference_layer(name):
    from kgcn_b200 import layers
    from kgcn_b200.csr import BatchedCSR
    g = load_golden(name)
    adjs = unflatten_adjs(g, "adj_")
    C, _, H = g["w"].shape
    csr = BatchedCSR.from_coo_lists(adjs)
    for flags in (0, 1):      # fused and reference-order paths of the library
        conv = layers.GraphConv(H, C, reference_order=bool(flags))
        x = dev(g["features"])
        conv(x, adj=csr)
        load_conv(conv, g["w"], g["bias"])
        close(conv(x, adj=csr), g["y"])
    if "gather" in g:
        close(layers.GraphGather()(dev(g["y"])), g["gather"])


@pytest.mark.gpu
def test_cuda_stack_equals_reference_layers():
    from kgcn_b200 import layers
    from kgcn_b200.csr import BatchedCSR
    g = load_golden("layers_synthetic_stack")
    adjs, n = unflatten_adjs(g, "adj_"), g["enabled_node_nums"]
    csr = BatchedCSR.from_coo_lists(adjs)
    N = g["features"].shape[1]
    c1, c2 = layers.GraphConv(50, 1, activation="sigmoid"), layers.GraphConv(50, 1)
    bn, gd, gg = layers.GraphBatchNormalization(), layers.GraphDense(50, activation="sigmoid"), layers.GraphGather()
    gd2, bn_legacy = layers.GraphDense(6), layers.GraphBatchNormalization(batch_statistics=True)
    x = dev(g["features"])
    nd = dev(n, torch.int32)
    h1 = c1(x, adj=csr)
    load_conv(c1, g["w1"], g["b1"])
    h1 = c1(x, adj=csr)
    close(h1, g["h1"])
    c2(h1, adj=csr)
    load_conv(c2, g["w2"], g["b2"])
    h2 = c2(h1, adj=csr)
    close(h2, g["h2"])
    h3 = torch.sigmoid(bn(h2, max_node_num=N, enabled_node_nums=nd))
    close(h3, g["h3"])
    gd(h3)
    with torch.no_grad():
        gd.kernel.copy_(dev(g["gd_kernel"])); gd.bias.copy_(dev(g["gd_bias"]))
    h4 = gd(h3)
    close(h4, g["h4"])
    close(gg(h4), g["gathered"])
    gd2(h3, max_node_num=N, enabled_node_nums=nd)
    with torch.no_grad():
        gd2.kernel.copy_(dev(g["gd2_kernel"])); gd2.bias.copy_(dev(g["gd2_bias"]))
    h5 = gd2(h3, max_node_num=N, enabled_node_nums=nd)
    close(h5, g["h5"])
    assert (h5[7:] == 0).all()
    close(bn_legacy(h2, max_node_num=N, enabled_node_nums=nd), g["h2_bn_legacy"], 5e-5)


@pytest.mark.gpu
def test_cuda_gin_maxpool_blockdiag_equal_reference_layers():
    from kgcn_b200 import layers
    from kgcn_b200.csr import BatchedCSR
    g = load_golden("layers_random")
    adjs = unflatten_adjs(g, "adj_")
    x = dev(g["features"])
    csr = BatchedCSR.from_coo_lists(adjs)
    gin = layers.GINAggregate(len(g["gin_eps"]))
    gin(x, adj=csr)
    with torch.no_grad():
        for c, e in enumerate(g["gin_eps"]):
            gin.epsilon[c].fill_(float(e))
    close(gin(x, adj=csr), g["gin_y"])
    ucsr = BatchedCSR.from_coo_lists(unflatten_adjs(g, "uadj_"))
    np.testing.assert_array_equal(layers.GraphMaxPooling(2)(x, adj=ucsr).cpu().numpy(), g["maxpool_y"])
    B, N, F = g["features"].shape
    idx = np.concatenate([adjs[b][0][0] + b * N for b in range(B)])
    val = np.concatenate([adjs[b][0][1] for b in range(B)])
    bgc = layers.BatchGraphConv(g["bgc_w"].shape[1])
    flat = dev(g["features"].reshape(B * N, F))
    big = (idx, val, [B * N, B * N])
    bgc([flat, big])
    with torch.no_grad():
        bgc.w.copy_(dev(g["bgc_w"])); bgc.bias.copy_(dev(g["bgc_bias"]).reshape(bgc.bias.shape))
    close(bgc([flat, big]), g["bgc_y"])
