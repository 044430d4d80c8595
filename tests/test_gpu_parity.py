"""GPU parity: every CUDA entry point, called through the C-ABI wrappers, against the CPU oracle
(oracle/ref_layers.py) on the same seeded inputs, plus the committed golden fixtures.

Stated tolerances (fp32, BASELINE.json north_star "within a stated fp32 tolerance"):
  forward values     |got - ref| <= 1e-5 * |ref| + 1e-5 * max|ref|
  parameter grads    |got - ref| <= 1e-4 * |ref| + 1e-4 * max|ref|   (sums over B*N rows)
Index / gather / packing work is bit-exact (assert_array_equal).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, unflatten_adjs
from oracle import ref_layers as R

pytestmark = pytest.mark.gpu


def close(got, ref, tol=1e-5):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * max(scale, 1e-30))


def random_batch(rng, B, N, C, F, max_nnz=None, empty_graphs=()):
    max_nnz = max_nnz if max_nnz is not None else max(2, int(N * 3.5))
    adjs = []
    for b in range(B):
        row = []
        for _ in range(C):
            nnz = 0 if b in empty_graphs else int(rng.integers(0, max_nnz + 1))
            idx = rng.integers(0, N, size=(nnz, 2)).astype(np.int32)         # unsorted, duplicates allowed
            row.append((idx, rng.standard_normal(nnz).astype(np.float32), [N, N]))
        adjs.append(row)
    return adjs, rng.standard_normal((B, N, F)).astype(np.float32)


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


@pytest.fixture(scope="module")
def K():
    import kgcn_b200
    from kgcn_b200 import batched_call, bconv_call, bspmm_call, csr, feed, layers, ops
    return dict(pkg=kgcn_b200, csr=csr, ops=ops, layers=layers, feed=feed, bspmm=bspmm_call, bconv=bconv_call,
                batched=batched_call)


# ---------------------------------------------------------------------------------------------
# batched SpMM (Bspmm / Bconv / Bspmdt)
# ---------------------------------------------------------------------------------------------
SPMM_SHAPES = [  # B, N, C, F          what it exercises
    (5, 3, 1, 4),    # sample.jbl shape, 48-byte tile (TMA-staged, VEC=4, LPR=1)
    (7, 10, 1, 3),   # synthetic.jbl shape: F=3 -> VEC=1, tile 120 B not 16B-multiple -> gather path
    (33, 32, 1, 64),  # C2
    (9, 50, 3, 50),  # C4-like: VEC=2, LPR=32
    (9, 50, 1, 75),  # C3 layer 1: VEC=1, 3 column chunks
    (6, 64, 2, 128),  # C5
    (3, 40, 1, 256),  # sparse_infer width: two 128-column chunks
    (2, 700, 1, 48),  # tile 134 KB > staging budget -> gather path (large-N graphs)
]


@pytest.mark.parametrize("B,N,C,F", SPMM_SHAPES)
def test_bspmm_layouts_match_oracle(K, B, N, C, F):
    rng = np.random.default_rng(B * 1000 + N)
    adjs, x = random_batch(rng, B, N, C, F, empty_graphs=(1,))
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    xc = rng.standard_normal((B, C, N, F)).astype(np.float32)
    # shared right-hand side, channels summed (aggregate-first GraphConv / GIN)
    want = np.stack([sum(R.sparse_dense_matmul(*adjs[b][c], x[b]) for c in range(C)) for b in range(B)])
    close(K["ops"].bspmm(csr, dev(x), "shared_sum"), want)
    # per-channel right-hand sides, summed (Bconv)
    want = np.stack(R.bconv([[a for a in row] for row in adjs], [[xc[b, c] for c in range(C)] for b in range(B)]))
    close(K["ops"].bspmm(csr, dev(xc), "sum"), want)
    # fully independent products (Bspmm), and the adjoint through the transposed CSR
    flat = [a for row in adjs for a in row]
    want = np.stack(R.bspmm(flat, list(xc.reshape(B * C, N, F)))).reshape(B, C, N, F)
    close(K["ops"].bspmm(csr, dev(xc), "per_matrix"), want)
    want_t = np.stack(R.bspmm(flat, list(xc.reshape(B * C, N, F)), adjoint_a=True)).reshape(B, C, N, F)
    close(K["ops"].bspmm(csr.transposed(), dev(xc), "per_matrix"), want_t)


def test_bspmm_long_rows_and_duplicates(K):
    """Rows with more entries than a lane group (>32) and heavy duplication."""
    rng = np.random.default_rng(5)
    N, F = 16, 64
    idx = np.stack([np.zeros(200, np.int32), rng.integers(0, N, 200).astype(np.int32)], 1)   # 200 entries in row 0
    idx = np.concatenate([idx, np.array([[3, 3]] * 70, np.int32)])                            # 70 duplicates of (3,3)
    val = rng.standard_normal(idx.shape[0]).astype(np.float32)
    x = rng.standard_normal((1, N, F)).astype(np.float32)
    csr = K["csr"].BatchedCSR.from_coo_lists([[(idx, val, [N, N])]])
    close(K["ops"].bspmm(csr, dev(x), "shared_sum"), R.sparse_dense_matmul(idx, val, [N, N], x[0])[None])


def test_bspmm_rectangular(K):
    rng = np.random.default_rng(6)
    Rr, Kc, F = 12, 20, 32
    mats = []
    for _ in range(4):
        nnz = 30
        idx = np.stack([rng.integers(0, Rr, nnz), rng.integers(0, Kc, nnz)], 1).astype(np.int64)
        mats.append((idx, rng.standard_normal(nnz).astype(np.float32), [Rr, Kc]))
    dense = [rng.standard_normal((Kc, F)).astype(np.float32) for _ in range(4)]
    out = K["bspmm"].BatchedSpMM().call(mats, [dev(d) for d in dense])
    for o, w in zip(out, R.bspmm(mats, dense)):
        close(o, w)
    dy = [rng.standard_normal((Rr, F)).astype(np.float32) for _ in range(4)]
    out_t = K["bspmm"].BatchedSpMM().call(mats, [dev(d) for d in dy], adjoint_a=True)
    for o, w in zip(out_t, R.bspmm(mats, dy, adjoint_a=True)):
        close(o, w)


def test_golden_kat_through_cuda(K):
    """KAT1 / KAT2 on the reference-ingested fixtures: exact small integers."""
    W = np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32)
    b0 = np.array([[0.5, -0.5]], np.float32)
    rec = load_golden("ingest_sample_plain")
    adjs, x = unflatten_adjs(rec, "adj_"), rec["features"].astype(np.float32)
    kat = load_golden("kat")
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    for flags in (0, 1):
        y = K["ops"].graphconv_fwd(csr, dev(x), dev(W[None]), dev(b0), 0, flags)
        np.testing.assert_array_equal(y.cpu().numpy(), kat["kat1_y"])
    g = K["layers"].GraphGather()(y)
    np.testing.assert_array_equal(g.cpu().numpy(), kat["kat1_gather"])
    # adjoint check on the asymmetric graph 4
    one = K["csr"].BatchedCSR.from_coo_lists([adjs[4]])
    yt = K["ops"].bspmm(one.transposed(), y[4:5].contiguous(), "shared_sum")
    np.testing.assert_array_equal(yt[0].cpu().numpy(), kat["kat1_adjoint_g4"])
    rec = load_golden("ingest_sample_multiadj_plain")
    adjs, x = unflatten_adjs(rec, "adj_"), rec["features"].astype(np.float32)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    w2, b2 = np.stack([W, -W]), np.array([[0.5, -0.5], [1, 2]], np.float32)
    for flags in (0, 1):
        y = K["ops"].graphconv_fwd(csr, dev(x), dev(w2), dev(b2), 0, flags)
        np.testing.assert_array_equal(y.cpu().numpy(), kat["kat2_y"])


# ---------------------------------------------------------------------------------------------
# GraphConv forward / backward
# ---------------------------------------------------------------------------------------------
CONV_SHAPES = [  # B, N, C, F_in, F_out
    (10, 10, 1, 3, 50),    # C1 (sample.json on synthetic.jbl)
    (40, 32, 1, 64, 64),   # C2
    (12, 50, 1, 75, 50),   # C3 layer 1
    (12, 50, 1, 50, 50),   # C3 layers 2-3
    (12, 50, 3, 75, 50),   # C4
    (8, 64, 1, 128, 128),  # C5
    (3, 7, 6, 5, 9),       # degree-split channels, odd sizes
    (1, 300, 1, 16, 24),   # B = 1, large N (block-diagonal use)
]


@pytest.mark.parametrize("B,N,C,fi,fo", CONV_SHAPES)
@pytest.mark.parametrize("act", ["none", "sigmoid", "relu", "tanh"])
@pytest.mark.parametrize("flags", [0, 1])
def test_graphconv_forward(K, B, N, C, fi, fo, act, flags):
    rng = np.random.default_rng(N * 31 + fo)
    adjs, x = random_batch(rng, B, N, C, fi, empty_graphs=(B - 1,) if B > 1 else ())
    w = [R.glorot_uniform(rng, fi, fo) for _ in range(C)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32) for _ in range(C)]
    want = R.activation(R.graph_conv(x, adjs, w, b, fast=True), act)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    y = K["ops"].graphconv_fwd(csr, dev(x), dev(np.stack(w)), dev(np.concatenate(b)), R.ACT_IDS[act], flags)
    close(y, want)
    if B > 1:  # an empty (padded) graph outputs act(0) everywhere: the bias never reaches it
        assert torch.all(y[B - 1] == float(R.activation(np.zeros(1, np.float32), act)[0]))


# shapes the warp-specialised kernel (graphconv_fused_v4.cu: F_in % 32 == 0, F_out % 4 == 0, N <= 128) takes, with the
# cases its pipeline has to get right: short last tiles, graphs that do not fill the 128 lanes, several slabs per
# aggregation warp (K = C * F_in > 64), a single Z buffer (K large), dense tiles whose CSR slice does not fit the
# stage (entries are then read from global memory), more CTAs than graphs per CTA, one graph only.
V4_SHAPES = [  # B, N, C, F_in, F_out, max_nnz per matrix
    (1, 32, 1, 64, 64, None),
    (7, 32, 1, 64, 64, None),
    (1024, 32, 1, 64, 64, None),     # C2 at full size
    (149, 32, 1, 64, 64, None),      # 148 CTAs + 1 graph
    (333, 50, 1, 32, 20, None),      # 2 graphs per tile, 100 of 128 lanes
    (40, 128, 1, 64, 36, None),      # one graph per tile
    (61, 17, 2, 32, 48, None),       # 7 graphs per tile, two channels
    (90, 32, 3, 64, 64, None),       # K = 192: three slabs per aggregation warp, single Z buffer
    (33, 16, 1, 32, 64, 16 * 16),    # dense graphs (~10 entries per row > stage capacity): un-staged CSR path
    (20, 64, 1, 96, 128, None),      # three slabs, 4 column slabs in the epilogue
    (12, 24, 1, 160, 8, None),       # five slabs, F_out < 32 (second epilogue column phase idle)
    (9, 64, 1, 128, 128, None),      # C5 width: [W ; bias] hi / lo is 160 KB -> two output-column slices (grid.y = 2), 1 graph per tile
    (300, 64, 1, 128, 128, None),    # the same with several tiles per CTA (74 CTAs per slice)
    (11, 50, 3, 64, 64, None),       # C4 hidden layer: K = 192 -> one accumulator + one Z buffer in tensor memory
    (11, 50, 1, 96, 64, None),       # C3 first layer after padding 75 -> 96
    (5, 32, 1, 256, 128, None),      # K = 256: column slices with a single Z buffer
    (11, 50, 3, 96, 64, None),       # C4 first layer after padding: K = 288 exceeds tensor memory -> channel groups {0,1} + {2}
    (300, 50, 3, 96, 64, None),      # the same with several tiles per CTA
]


@pytest.mark.parametrize("B,N,C,fi,fo,max_nnz", V4_SHAPES)
@pytest.mark.parametrize("act", ["none", "sigmoid"])
def test_graphconv_forward_v4_pipeline(K, B, N, C, fi, fo, max_nnz, act):
    rng = np.random.default_rng(B * 7 + N)
    adjs, x = random_batch(rng, B, N, C, fi, max_nnz=max_nnz, empty_graphs=(B // 2,) if B > 2 else ())
    w = [R.glorot_uniform(rng, fi, fo) for _ in range(C)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32) for _ in range(C)]
    want = R.activation(R.graph_conv(x, adjs, w, b, fast=True), act)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    xd, wd, bd = dev(x), dev(np.stack(w)), dev(np.concatenate(b))
    y = K["ops"].graphconv_fwd(csr, xd, wd, bd, R.ACT_IDS[act], 0)
    close(y, want)
    # fused == decomposed reference-order path (exact-fp32 GEMM + SpMM) of the library itself
    close(y, K["ops"].graphconv_fwd(csr, xd, wd, bd, R.ACT_IDS[act], 1).cpu().numpy())
    # deterministic: bit-identical on a second launch
    assert torch.equal(y, K["ops"].graphconv_fwd(csr, xd, wd, bd, R.ACT_IDS[act], 0))
    # no bias
    y0 = K["ops"].graphconv_fwd(csr, xd, wd, None, R.ACT_IDS[act], 0)
    close(y0, R.activation(R.graph_conv(x, adjs, w, [np.zeros((1, fo), np.float32)] * C, fast=True), act))


def test_graphconv_forward_faithful_order_small(K):
    """Against the slow O1 tier (per-nnz storage order) on a small case."""
    rng = np.random.default_rng(11)
    adjs, x = random_batch(rng, 6, 9, 2, 5)
    w = [R.glorot_uniform(rng, 5, 7) for _ in range(2)]
    b = [rng.uniform(-0.5, 0.5, (1, 7)).astype(np.float32) for _ in range(2)]
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    for flags in (0, 1):
        close(K["ops"].graphconv_fwd(csr, dev(x), dev(np.stack(w)), dev(np.concatenate(b)), 0, flags),
              R.graph_conv(x, adjs, w, b, fast=False))


@pytest.mark.parametrize("B,N,C,fi,fo", [(10, 10, 1, 3, 50), (24, 32, 1, 64, 64), (9, 50, 3, 75, 50), (5, 64, 2, 128, 128)])
@pytest.mark.parametrize("act", ["none", "sigmoid", "relu"])
def test_graphconv_backward(K, B, N, C, fi, fo, act):
    rng = np.random.default_rng(fi + fo)
    adjs, x = random_batch(rng, B, N, C, fi)
    w = [R.glorot_uniform(rng, fi, fo) for _ in range(C)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32) for _ in range(C)]
    y = R.activation(R.graph_conv(x, adjs, w, b, fast=True), act)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    du = dy * R.activation_grad_from_output(y, act)
    dx, dw, db = R.graph_conv_grad(x, adjs, w, b, du)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    gdx, gdw, gdb = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy))
    close(gdx, dx, 1e-4)
    close(gdw, np.stack(dw), 1e-4)
    close(gdb, np.concatenate(db), 1e-4)
    _, gdw2, _ = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), need_dx=False)
    assert torch.equal(gdw, gdw2)   # deterministic reduction order


FUSED_BWD_SHAPES = [  # B, N, fi, fo, max_nnz: single-channel shapes the one-kernel tensor-core backward takes
    (24, 32, 64, 64, None),    # C2
    (700, 32, 64, 64, None),   # more tiles than CTAs: the dW accumulator lives across tiles
    (7, 50, 36, 48, None),     # one graph per tile (50 of 64 rows), lane groups wider than the row, odd graph count
    (5, 64, 64, 32, None),     # full 64-row graphs
    (33, 16, 8, 128, None),    # 4 graphs per tile, last tile short, widest output
    (6, 32, 64, 64, 900),      # more entries per tile than the stage holds (entries read from global memory)
    (3, 1, 4, 4, 1),           # degenerate
]


@pytest.mark.parametrize("B,N,fi,fo,max_nnz", FUSED_BWD_SHAPES)
@pytest.mark.parametrize("act", ["none", "sigmoid", "relu", "tanh"])
@pytest.mark.parametrize("bcast", [False, True])
def test_graphconv_backward_fused(K, B, N, fi, fo, max_nnz, act, bcast):
    """Fused backward (default) vs the oracle and vs the decomposed exact-fp32 path (REFERENCE_ORDER flag)."""
    rng = np.random.default_rng(B * 7 + fo)
    adjs, x = random_batch(rng, B, N, 1, fi, max_nnz=max_nnz, empty_graphs=(0,))
    w = [R.glorot_uniform(rng, fi, fo)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32)]
    y = R.activation(R.graph_conv(x, adjs, w, b, fast=True), act)
    dy = rng.standard_normal((B, fo) if bcast else y.shape).astype(np.float32)
    dy_full = np.broadcast_to(dy[:, None, :], y.shape) if bcast else dy
    du = dy_full * R.activation_grad_from_output(y, act)
    dx, dw, db = R.graph_conv_grad(x, adjs, w, b, du)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    base = 2 if bcast else 0   # KGCN_FLAG_DY_BROADCAST
    gdx, gdw, gdb = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), flags=base)
    rdx, rdw, rdb = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), flags=base | 1)
    scale = max(1.0, float(np.abs(dw[0]).max()))
    close(gdx, dx, 1e-4)
    close(gdx, rdx.cpu().numpy(), 1e-4)
    np.testing.assert_allclose(gdw.cpu().numpy(), np.stack(dw), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(gdw.cpu().numpy(), rdw.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(gdb.cpu().numpy(), np.concatenate(db), rtol=1e-4, atol=2e-5 * scale)
    _, gdw2, gdb2 = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), need_dx=False,
                                           flags=base)
    assert torch.equal(gdw, gdw2) and torch.equal(gdb, gdb2)   # same tiles, same accumulation order


SPLIT_BWD_SHAPES = [  # B, N, C, fi, fo, max_nnz: widths that are multiples of 32 -> dx by the fused layer kernel on
    # (A^T, dU, W^T) + dW / dbias by graphconv_fused_dw_kernel (concurrent roles, all channels in one launch)
    (24, 32, 1, 64, 64, None),      # C2
    (700, 32, 1, 64, 64, None),     # more tiles than CTAs: the dW accumulator lives across tiles
    (9, 50, 3, 64, 64, None),       # C4 hidden layer: three channels, 50-row graphs (56-row operand chunks)
    (9, 50, 3, 96, 64, None),       # C4 first layer after padding (X as separate hi / lo operands, N = 192)
    (9, 50, 1, 96, 64, None),       # C3 first layer after padding 75 -> 96: Xhi / Xlo as separate operands
    (7, 64, 1, 128, 128, None),     # C5 width: 32-row operand chunks, dx in two column slices
    (300, 64, 1, 128, 128, None),   # several tiles per CTA at C5 width
    (33, 16, 1, 32, 128, None),     # 4 graphs per tile, last tile short
    (6, 32, 1, 64, 64, 900),        # more entries per tile than the stage holds (entries read from global memory)
    (5, 128, 1, 64, 32, None),      # 128-node graphs: two operand chunks per tile
    (3, 1, 2, 32, 32, 1),           # degenerate
]


@pytest.mark.parametrize("B,N,C,fi,fo,max_nnz", SPLIT_BWD_SHAPES)
@pytest.mark.parametrize("act", ["none", "sigmoid", "relu"])
@pytest.mark.parametrize("bcast", [False, True])
def test_graphconv_backward_split(K, B, N, C, fi, fo, max_nnz, act, bcast):
    """Backward as two streaming launches vs the oracle and vs the decomposed exact-fp32 path (REFERENCE_ORDER flag)."""
    rng = np.random.default_rng(B * 7 + fo + C)
    adjs, x = random_batch(rng, B, N, C, fi, max_nnz=max_nnz, empty_graphs=(0,))
    w = [R.glorot_uniform(rng, fi, fo) for _ in range(C)]
    b = [rng.uniform(-0.5, 0.5, (1, fo)).astype(np.float32) for _ in range(C)]
    y = R.activation(R.graph_conv(x, adjs, w, b, fast=True), act)
    dy = rng.standard_normal((B, fo) if bcast else y.shape).astype(np.float32)
    dy_full = np.broadcast_to(dy[:, None, :], y.shape) if bcast else dy
    du = dy_full * R.activation_grad_from_output(y, act)
    dx, dw, db = R.graph_conv_grad(x, adjs, w, b, du)
    csr = K["csr"].BatchedCSR.from_coo_lists(adjs)
    base = 2 if bcast else 0   # KGCN_FLAG_DY_BROADCAST
    gdx, gdw, gdb = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), flags=base)
    rdx, rdw, rdb = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), flags=base | 1)
    scale = max(1.0, float(np.abs(np.stack(dw)).max()))
    close(gdx, dx, 1e-4)
    close(gdx, rdx.cpu().numpy(), 1e-4)
    np.testing.assert_allclose(gdw.cpu().numpy(), np.stack(dw), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(gdw.cpu().numpy(), rdw.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(gdb.cpu().numpy(), np.concatenate(db), rtol=1e-4, atol=2e-5 * scale)
    _, gdw2, gdb2 = K["ops"].graphconv_bwd(csr, dev(x), dev(np.stack(w)), R.ACT_IDS[act], dev(y), dev(dy), need_dx=False,
                                           flags=base)
    assert torch.equal(gdw, gdw2) and torch.equal(gdb, gdb2)   # same tiles, same accumulation order


# ---------------------------------------------------------------------------------------------
# GraphDense / GraphGather
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,fi,fo", [(5, 3, 4, 2), (16, 50, 50, 50), (8, 32, 64, 256), (3, 10, 50, 2)])
@pytest.mark.parametrize("masked", [False, True])
def test_graphdense_fwd_bwd(K, B, N, fi, fo, masked):
    rng = np.random.default_rng(fi * fo)
    x = rng.standard_normal((B, N, fi)).astype(np.float32)
    k, b = R.glorot_uniform(rng, fi, fo), rng.uniform(-0.5, 0.5, fo).astype(np.float32)
    en = rng.integers(0, N + 1, B).astype(np.int32) if masked else None
    want = R.graph_dense(x, k, b, act="sigmoid", enabled_node_nums=en)
    layer = K["layers"].GraphDense(fo, activation="sigmoid")
    xt = dev(x).requires_grad_(True)
    y = layer(xt, enabled_node_nums=None if en is None else dev(en, torch.int32))
    with torch.no_grad():
        layer.kernel.copy_(dev(k)); layer.bias.copy_(dev(b))
    y = layer(xt, enabled_node_nums=None if en is None else dev(en, torch.int32))
    close(y, want)
    if masked:
        keep = np.arange(N)[None, :] < en[:, None]
        assert torch.all(y[torch.as_tensor(~keep).cuda()] == 0)          # exact zeros (layers.py:249-253)
    dy = rng.standard_normal(want.shape).astype(np.float32)
    y.backward(dev(dy))
    du = dy * R.activation_grad_from_output(want, "sigmoid")
    if masked:
        du = du * keep[:, :, None]
    close(xt.grad, du @ k.T, 1e-4)
    close(layer.kernel.grad, x.reshape(-1, fi).T @ du.reshape(-1, fo), 1e-4)
    close(layer.bias.grad, du.reshape(-1, fo).sum(0), 1e-4)


@pytest.mark.parametrize("B,N,F", [(5, 3, 2), (64, 50, 50), (7, 32, 64), (3, 64, 130)])
def test_graphgather(K, B, N, F):
    rng = np.random.default_rng(F)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    xt = dev(x).requires_grad_(True)
    out = K["layers"].GraphGather()(xt)
    close(out, R.graph_gather(x))
    dg = rng.standard_normal((B, F)).astype(np.float32)
    out.backward(dev(dg))
    np.testing.assert_array_equal(xt.grad.cpu().numpy(), np.broadcast_to(dg[:, None, :], x.shape))


# ---------------------------------------------------------------------------------------------
# plugin-op wrappers with the reference call signatures + registered gradients
# ---------------------------------------------------------------------------------------------
def test_plugin_ops_forward_and_gradients(K):
    rng = np.random.default_rng(21)
    B, C, N, F = 4, 2, 8, 16
    adjs, _ = random_batch(rng, B, N, C, F)
    dense = [[rng.standard_normal((N, F)).astype(np.float32) for _ in range(C)] for _ in range(B)]
    STV = K["feed"].SparseTensorValue
    # Bconv: list[B][C] in, list[B] out, channel sum fused
    sp = [[STV(*a) for a in row] for row in adjs]
    out = K["bconv"].BatchedConv().call(sp, [[dev(d) for d in row] for row in dense])
    for o, w in zip(out, R.bconv(adjs, dense)):
        close(o, w)
    # Bspmm with gradients w.r.t. dense AND sparse values (bspmm_call.py:44-54)
    flat_sp = [a for row in adjs for a in row]
    flat_d = [d for row in dense for d in row]
    vals = [dev(a[1]).requires_grad_(True) for a in flat_sp]
    dts = [dev(d).requires_grad_(True) for d in flat_d]
    outs = K["bspmm"].BatchedSpMM().call([STV(a[0], v, a[2]) for a, v in zip(flat_sp, vals)], dts)
    dy = [rng.standard_normal((N, F)).astype(np.float32) for _ in flat_sp]
    torch.autograd.backward(outs, [dev(g) for g in dy])
    want_dv, want_db = R.bspmm_grad(flat_sp, flat_d, dy)
    for t in range(len(flat_sp)):
        close(outs[t], R.bspmm(flat_sp, flat_d)[t])
        close(dts[t].grad, want_db[t], 1e-4)
        if flat_sp[t][1].shape[0]:
            close(vals[t].grad, want_dv[t], 1e-4)
    # Bspmdt: one stacked dense [N_mats*rows, cols]
    first = [row[0] for row in adjs]
    stacked = np.concatenate([row[0] for row in dense], 0)
    st = dev(stacked).requires_grad_(True)
    outs = K["batched"].BatchedSpMDT().call([STV(*a) for a in first], st)
    for o, w in zip(outs, R.bspmdt(first, stacked)):
        close(o, w)
    torch.autograd.backward(outs, [dev(g) for g in dy[:B]])
    close(st.grad, np.concatenate(R.bspmm(first, dy[:B], adjoint_a=True), 0), 1e-4)


@pytest.mark.parametrize("adjoint_a", [False, True])
@pytest.mark.parametrize("B,C,N,F", [(4, 2, 8, 16), (3, 3, 50, 50), (5, 1, 32, 64)])
def test_bconv_registered_gradient(K, B, C, N, F, adjoint_a):
    """``_bconv_grad`` (kgcn/bconv_call.py:28-70): dY replicated over the channels, dense gradients through
    ``adjoint_a = not adj_a``, value gradients per stored entry -- through BatchedConv().call and autograd, i.e. the
    "sum"-layout backward of ops.BspmmFunction."""
    rng = np.random.default_rng(100 * B + C)
    adjs, _ = random_batch(rng, B, N, C, F)
    dense = [[rng.standard_normal((N, F)).astype(np.float32) for _ in range(C)] for _ in range(B)]
    STV = K["feed"].SparseTensorValue
    vals = [[dev(a[1]).requires_grad_(True) for a in row] for row in adjs]
    dts = [[dev(d).requires_grad_(True) for d in row] for row in dense]
    sp = [[STV(a[0], v, a[2]) for a, v in zip(row, vrow)] for row, vrow in zip(adjs, vals)]
    outs = K["bconv"].BatchedConv().call(sp, dts, adjoint_a=adjoint_a)
    for o, w in zip(outs, R.bconv(adjs, dense, adjoint_a=adjoint_a)):
        close(o, w)
    dy = [rng.standard_normal((N, F)).astype(np.float32) for _ in range(B)]
    torch.autograd.backward(outs, [dev(g) for g in dy])
    want_dv, want_db = R.bconv_grad(adjs, dense, dy, adjoint_a=adjoint_a)
    t = 0
    for b in range(B):
        for c in range(C):
            close(dts[b][c].grad, want_db[t], 1e-4)
            if adjs[b][c][1].shape[0]:
                close(vals[b][c].grad, want_dv[t], 1e-4)
            t += 1


@pytest.mark.parametrize("B,N,fi,fo", [(6, 10, 3, 16), (4, 50, 75, 64), (1, 200, 32, 32)])
def test_batch_graph_conv_backward(K, B, N, fi, fo):
    """BatchGraphConv (kgcn/layers.py:363-398), ``relu(A (x W + b))`` on one block-diagonal matrix: gradients of x, W and
    bias against float64 torch-CPU autograd of the oracle's formula (oracle/ref_layers.batch_graph_conv)."""
    rng = np.random.default_rng(fi + fo)
    adjs, x = random_batch(rng, B, N, 1, fi)
    idx = np.concatenate([adjs[b][0][0] + b * N for b in range(B)])
    val = np.concatenate([adjs[b][0][1] for b in range(B)])
    big = (idx, val, [B * N, B * N])
    w, bias = R.glorot_uniform(rng, fi, fo), rng.uniform(-0.5, 0.5, fo).astype(np.float32)
    layer = K["layers"].BatchGraphConv(fo)
    xt = dev(x.reshape(B * N, fi)).requires_grad_(True)
    layer([xt, big])
    with torch.no_grad():
        layer.w.copy_(dev(w)); layer.bias.copy_(dev(bias).reshape(layer.bias.shape))
    y = layer([xt, big])
    close(y, R.batch_graph_conv(x.reshape(B * N, fi), big, w, bias))
    dy = rng.standard_normal((B * N, fo)).astype(np.float32)
    y.backward(dev(dy))
    A = torch.zeros(B * N, B * N, dtype=torch.float64)
    A.index_put_((torch.as_tensor(idx[:, 0]).long(), torch.as_tensor(idx[:, 1]).long()), torch.as_tensor(val, dtype=torch.float64), accumulate=True)
    x64 = torch.tensor(x.reshape(B * N, fi), dtype=torch.float64, requires_grad=True)
    w64, b64 = torch.tensor(w, dtype=torch.float64, requires_grad=True), torch.tensor(bias, dtype=torch.float64, requires_grad=True)
    torch.relu(A @ (x64 @ w64 + b64)).backward(torch.tensor(dy, dtype=torch.float64))
    close(xt.grad, x64.grad.numpy(), 1e-4)
    close(layer.w.grad, w64.grad.numpy(), 1e-4)
    close(layer.bias.grad.reshape(-1), b64.grad.numpy(), 1e-4)


@pytest.mark.parametrize("mode", ["bspmm", "bconv", "batched"])
def test_layer_plugin_branches_equal_default(K, mode):
    """--bspmm / --bconv / --batched (layers.py:68-104) compute the same function as the default branch."""
    import types
    L = K["layers"]
    rng = np.random.default_rng(33)
    adjs, x = random_batch(rng, 6, 12, 2, 10)
    L.load_bspmm(types.SimpleNamespace(bspmm=False, bconv=False, batched=False))
    torch.manual_seed(0)
    ref_layer = L.GraphConv(14, 2)
    want = ref_layer(dev(x), adj=adjs)
    L.load_bspmm(types.SimpleNamespace(bspmm=mode == "bspmm", bconv=mode == "bconv", batched=mode == "batched"))
    try:
        layer = L.GraphConv(14, 2)
        layer(dev(x), adj=adjs)
        with torch.no_grad():
            for c in range(2):
                layer.w[c].copy_(ref_layer.w[c]); layer.bias[c].copy_(ref_layer.bias[c])
        close(layer(dev(x), adj=adjs), want.detach().cpu().numpy())
    finally:
        L.load_bspmm(types.SimpleNamespace(bspmm=False, bconv=False, batched=False))


def test_device_pack_equals_host_pack(K):
    from kgcn_b200 import _lib
    rng = np.random.default_rng(8)
    counts = rng.integers(0, 60, size=(50, 2))
    counts[7] = 0
    nnz = int(counts.sum())
    idx = rng.integers(0, 17, size=(nnz, 2)).astype(np.int32)
    val = rng.standard_normal(nnz).astype(np.float32)
    off = np.zeros(counts.size + 1, np.int64)
    np.cumsum(counts.reshape(-1), out=off[1:])
    d_off, d_idx, d_val = dev(off, torch.int64), dev(idx, torch.int32), dev(val)     # keep alive across the launches
    for transpose in (0, 1):
        want = K["csr"].pack_host(counts, idx, val, 17, 17, transpose=bool(transpose), want_perm=True)
        rowptr = torch.empty(counts.size * 17 + 1, dtype=torch.int32, device="cuda")
        col = torch.empty(nnz, dtype=torch.int32, device="cuda")
        v = torch.empty(nnz, dtype=torch.float32, device="cuda")
        perm = torch.empty(nnz, dtype=torch.int32, device="cuda")
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib.kgcn_pack_coo_device(counts.size, 17, 17, d_off.data_ptr(), d_idx.data_ptr(),
                                                 d_val.data_ptr(), transpose, rowptr.data_ptr(), col.data_ptr(), v.data_ptr(),
                                                 perm.data_ptr(), flag.data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert int(flag.item()) == 0
        for g, w in zip((rowptr, col, v, perm), want):
            np.testing.assert_array_equal(g.cpu().numpy(), w)
    bad = idx.copy(); bad[3, 1] = 17
    d_bad = dev(bad, torch.int32)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.kgcn_pack_coo_device(counts.size, 17, 17, d_off.data_ptr(), d_bad.data_ptr(),
                                             d_val.data_ptr(), 0, rowptr.data_ptr(), col.data_ptr(), v.data_ptr(), perm.data_ptr(),
                                             flag.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert int(flag.item()) == 3   # KGCN_ERR_INDEX_RANGE


# ---------------------------------------------------------------------------------------------
# whole network through the layer API + autograd vs the oracle network
# ---------------------------------------------------------------------------------------------
def test_network_forward_backward_through_layers(K):
    L = K["layers"]
    rng = np.random.default_rng(77)
    B, N, C, F = 16, 10, 1, 3
    rec = load_golden("ingest_synthetic_plain")
    adjs = unflatten_adjs(rec, "adj_")[:B]
    x = rec["features"][:B].astype(np.float32)
    labels = rec["in_label"][:B].astype(np.float32)
    mask = np.ones(B, np.float32); mask[-3:] = 0
    p = R.init_network(rng, F, [50, 50, 50], C, 2, dense_dim=50)
    fw, grads = R.network_grad(p, x, adjs, labels, mask, act="sigmoid")

    convs = [L.GraphConv(50, C, activation="sigmoid") for _ in range(3)]
    gd, gather = L.GraphDense(50, activation="sigmoid"), L.GraphGather()
    out_w, out_b = dev(p["out_w"]).requires_grad_(True), dev(p["out_b"]).requires_grad_(True)

    def forward():
        h = dev(x)
        for conv in convs:
            h = conv(h, adj=adjs)
        return gather(gd(h)) @ out_w + out_b

    forward()
    with torch.no_grad():
        for i, conv in enumerate(convs):
            conv.w[0].copy_(dev(p["conv_w"][i][0])); conv.bias[0].copy_(dev(p["conv_b"][i][0]))
        gd.kernel.copy_(dev(p["gd_w"])); gd.bias.copy_(dev(p["gd_b"]))
    logits = forward()
    close(logits, fw["logits"], 1e-4)
    cost = dev(mask) * -(dev(labels) * torch.log_softmax(logits, 1)).sum(1)
    cost.mean().backward()
    close(cost.sum(), fw["cost_sum"], 1e-4)
    for i, conv in enumerate(convs):
        close(conv.w[0].grad, grads["conv_w"][i][0], 1e-3)
        close(conv.bias[0].grad, grads["conv_b"][i][0], 1e-3)
    close(gd.kernel.grad, grads["gd_w"], 1e-3)
    close(out_w.grad, grads["out_w"], 1e-3)


# ---------------------------------------------------------------------------------------------
# full BASELINE sizes: size-independent properties instead of the (slow) oracle
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N,F,C", [(1024, 32, 64, 1), (512, 50, 75, 3), (512, 64, 128, 1)])
def test_full_size_properties(K, B, N, F, C):
    from kgcn_b200 import synth
    rng = np.random.default_rng(1234)
    counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
    csr = K["csr"].BatchedCSR.from_flat(counts, idx, val, N, N)
    g = torch.Generator(device="cuda").manual_seed(1)
    x1 = torch.randn(B, N, F, device="cuda", generator=g)
    x2 = torch.randn(B, N, F, device="cuda", generator=g)
    spmm = lambda c, x: K["ops"].bspmm(c, x, "shared_sum")
    y1, y2 = spmm(csr, x1), spmm(csr, x2)
    # linearity
    close(spmm(csr, 2.0 * x1 - 3.0 * x2), (2.0 * y1 - 3.0 * y2).cpu().numpy(), 1e-4)
    # adjoint identity  <A x, y> = <x, A^T y>
    lhs = (y1.double() * x2.double()).sum().item()
    rhs = (x1.double() * spmm(csr.transposed(), x2).double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), abs(rhs), 1.0) + 1e-2
    # A . ones = row sums of the values (checks every CSR entry is visited exactly once)
    ones = torch.ones(B, N, 4, device="cuda")
    deg = np.zeros((B, N), np.float64)
    mat = np.repeat(np.arange(B * C), counts.reshape(-1))
    np.add.at(deg, (mat // C, idx[:, 0]), val.astype(np.float64))
    close(spmm(csr, ones)[:, :, 0], deg.astype(np.float32), 1e-5)
    # fused layer == decomposed reference-order layer at full size
    w = torch.randn(C, F, 64, device="cuda", generator=g) * 0.1
    b = torch.randn(C, 64, device="cuda", generator=g) * 0.1
    ya = K["ops"].graphconv_fwd(csr, x1, w, b, 2, 0)
    yb = K["ops"].graphconv_fwd(csr, x1, w, b, 2, 1)
    close(ya, yb.cpu().numpy(), 1e-5)
