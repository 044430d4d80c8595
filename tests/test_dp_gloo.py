"""CPU, world_size 2 over gloo: the data-parallel plumbing of the step loop -- contiguous sharding
of the global batch, loss scale 1/B_global folded into local gradients, ONE sum all-reduce of the
flat gradient buffer -- reproduces the single-process gradient.  Local
gradients come from the CPU oracle here (no GPU in this tier); the GPU tier checks the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kgcn_b200.trainer import NetSpec, shard_range
from oracle import ref_layers as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make(B=12, N=6, F=5, seed=3):
    rng = np.random.default_rng(seed)
    adjs = []
    for _ in range(B):
        nnz = int(rng.integers(1, 14))
        adjs.append([(rng.integers(0, N, size=(nnz, 2)).astype(np.int32), rng.standard_normal(nnz).astype(np.float32), [N, N])])
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    labels = np.eye(2, dtype=np.float32)[rng.integers(0, 2, B)]
    p = R.init_network(rng, F, [7, 4], 1, 2)
    return adjs, x, labels, p


def _flat(grads):
    parts = []
    for w, b in zip(grads["conv_w"], grads["conv_b"]):
        parts += [np.stack(w).ravel(), np.concatenate(b, 0).ravel()]
    parts += [grads["out_w"].ravel(), grads["out_b"].ravel()]
    return np.concatenate(parts).astype(np.float32)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    adjs, x, labels, p = _make()
    B = x.shape[0]
    lo, hi = shard_range(B, rank, world)
    mask = np.ones(hi - lo, np.float32)
    _, g = R.network_grad(p, x[lo:hi], adjs[lo:hi], labels[lo:hi], mask)
    # network_grad differentiates the LOCAL mean; rescale to the global mean (1/B_global) as the
    # trainer does through inv_batch, then sum-all-reduce the flat buffer once.
    flat = torch.from_numpy(_flat(g) * ((hi - lo) / B))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put(flat.numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process_gradient():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    adjs, x, labels, p = _make()
    _, g = R.network_grad(p, x, adjs, labels, np.ones(x.shape[0], np.float32))
    np.testing.assert_allclose(got, _flat(g), rtol=1e-4, atol=1e-6)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 512, 4096, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_parameter_layout_is_16_byte_aligned():
    spec = NetSpec(75, [50, 50, 50], 50, channels=3, dense_dim=50)
    names = [n for n, _ in spec.param_shapes()]
    assert names[0] == "conv0/kernel" and names[-1] == "dense/bias" and "graph_dense/kernel" in names


def _tfrecord_worker(rank, world, port, path, out):
    from kgcn_b200 import tfrecords
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = tfrecords.SparseDataset(path, task_num=1)
    rng = np.random.default_rng(5)
    w = rng.standard_normal((1, 6, 4)).astype(np.float32)
    b = rng.standard_normal((1, 4)).astype(np.float32)
    total = torch.zeros(4, dtype=torch.float64)
    n_local = 0
    for parsed in ds.batches(5, rank=rank, world_size=world):
        sizes = parsed["size"][:, 0]
        n_local += len(sizes)
        if len(sizes) == 0:
            continue
        chans, feat = tfrecords.block_diagonal_batch(parsed, normalize=True)
        h = R.graph_conv(feat.astype(np.float32)[None], [[chans[0]]], w, b)
        total += torch.from_numpy(R.segment_sum(np.maximum(h, 0)[0], sizes).astype(np.float64).sum(0))
    count = torch.tensor([n_local])
    dist.all_reduce(total)
    dist.all_reduce(count)
    if rank == 0:
        out.put((total.numpy(), int(count)))
    dist.destroy_process_group()


def test_two_ranks_shard_tfrecord_batches_without_overlap(tmp_path):
    """File-fed block-diagonal path at world_size 2: the ranks' shards of every global batch partition it, and the
    per-molecule readouts summed over ranks equal the single-process ones (degree normalisation is per molecule
    block, so sharding a block-diagonal batch changes nothing)."""
    from kgcn_b200 import tfrecords
    rng = np.random.default_rng(1)
    blobs = []
    for _ in range(13):                                   # 13 records, batch 5 -> global batches of 5, 5, 3
        n = int(rng.integers(2, 7))
        adj = np.eye(n, dtype=np.float32)
        for i in range(n - 1):
            adj[i, i + 1] = adj[i + 1, i] = 1.0
        feat = np.zeros((n, 6), np.float32)
        feat[np.arange(n), rng.integers(0, 6, n)] = 1.0
        blobs.append(tfrecords.convert_to_example(adj, feat, np.array([1.0]), np.array([1])))
    path = str(tmp_path / "0_train_.tfrecords")
    tfrecords.write_tfrecords(path, blobs)
    ctx = mp.get_context("spawn")
    results = []
    for world in (1, 2):
        out, port = ctx.Queue(), _free_port()
        procs = [ctx.Process(target=_tfrecord_worker, args=(r, world, port, path, out)) for r in range(world)]
        for pr in procs:
            pr.start()
        results.append(out.get(timeout=120))
        for pr in procs:
            pr.join(timeout=60)
            assert pr.exitcode == 0
    (single, n1), (double, n2) = results
    assert n1 == n2 == 13
    np.testing.assert_allclose(double, single, rtol=1e-6)
