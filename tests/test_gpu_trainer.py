"""GPU: the step loop (forward + head + backward + Adam through the C ABI, eager and CUDA-graph
replay) against the oracle network (oracle/ref_layers.py network_grad) on the same inputs."""
import numpy as np
import pytest
import torch

from oracle import ref_layers as R

pytestmark = pytest.mark.gpu


def close(got, ref, tol):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = np.asarray(ref)
    scale = float(np.abs(ref).max()) if ref.size else 0.0
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * max(scale, 1e-30))


def make_case(B, N, C, F, conv_dims, dense_dim, seed):
    from kgcn_b200 import synth
    rng = np.random.default_rng(seed)
    counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
    x = rng.standard_normal((B, N, F)).astype(np.float32)
    labels = np.eye(2, dtype=np.float32)[rng.integers(0, 2, B)]
    mask = np.ones(B, np.float32)
    mask[-2:] = 0
    adjs, pos = [], 0
    for b in range(B):
        row = []
        for c in range(C):
            n = int(counts[b, c])
            row.append((idx[pos:pos + n], val[pos:pos + n], [N, N]))
            pos += n
        adjs.append(row)
    p = R.init_network(rng, F, conv_dims, C, 2, dense_dim=dense_dim)
    return counts, idx, val, x, labels, mask, adjs, p


def numpy_adam(p, g, lr=0.01, b1=0.9, b2=0.999, eps=1e-8, t=1):
    m = (1 - b1) * g
    v = (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps)


@pytest.mark.parametrize("B,N,C,F,conv_dims,dense_dim", [
    (16, 10, 1, 3, [50, 50, 50], 50),     # C1 network (example_model/model.py without BN/Dropout)
    (24, 32, 1, 64, [64, 64], None),      # C2
    (10, 50, 3, 75, [50, 50, 50], None),  # C4
    (12, 50, 1, 75, [50, 50, 50], None),  # C3
    (6, 64, 1, 128, [128, 128], None),    # C5 (per-GPU network)
])
@pytest.mark.parametrize("flags", [0, 1])
def test_step_matches_oracle(B, N, C, F, conv_dims, dense_dim, flags):
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    counts, idx, val, x, labels, mask, adjs, p = make_case(B, N, C, F, conv_dims, dense_dim, seed=B + N)
    fw, grads = R.network_grad(p, x, adjs, labels, mask, act="sigmoid")
    spec = NetSpec(F, conv_dims, N, channels=C, label_dim=2, dense_dim=dense_dim, act="sigmoid")
    tr = Trainer(spec, B, lr=0.01, flags=flags)
    tr.load_oracle_params(p)
    batch = DeviceBatch.from_host(counts, idx, val, x, labels, N, mask=mask, pad_to=tr.dims[0])
    assert tr.padded == (flags == 0 and any(d % 32 for d in [F] + conv_dims))   # Tox21-style widths run zero-padded to 32s
    assert tr.fused_step == (flags == 0 and dense_dim is None)
    tr.forward_eager(batch)
    close(tr.logits, fw["logits"], 1e-4)
    close(tr.prediction, fw["prediction"], 1e-4)
    cost_sum, correct = tr.read_stats()
    assert abs(cost_sum - float(fw["cost_sum"])) <= 1e-4 * abs(float(fw["cost_sum"]))
    assert correct == float(fw["correct_count"])
    before = {k: v.clone() for k, v in tr.views.items()}
    tr.step_eager(batch)
    for i in range(len(conv_dims)):
        close(tr.gviews["conv%d/kernel" % i], np.stack(grads["conv_w"][i]), 1e-3)
        close(tr.gviews["conv%d/bias" % i], np.concatenate(grads["conv_b"][i], 0), 1e-3)
    if dense_dim:
        close(tr.gviews["graph_dense/kernel"], grads["gd_w"], 1e-3)
        close(tr.gviews["graph_dense/bias"], grads["gd_b"], 1e-3)
    close(tr.gviews["dense/kernel"], grads["out_w"], 1e-3)
    close(tr.gviews["dense/bias"], grads["out_b"], 1e-3)
    # Adam, TF formulation, first step
    for k, v in tr.views.items():
        want = numpy_adam(before[k].cpu().numpy(), tr.gviews[k].cpu().numpy())
        close(v, want, 1e-5)
    assert int(tr.step_state[0].item()) == 1 and int(tr.step_state[1].item()) == 0


def test_graph_replay_equals_eager():
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    counts, idx, val, x, labels, mask, adjs, p = make_case(32, 32, 1, 64, [64, 64], None, seed=5)
    spec = NetSpec(64, [64, 64], 32)
    batch = DeviceBatch.from_host(counts, idx, val, x, labels, 32, mask=mask)
    a, b = Trainer(spec, 32), Trainer(spec, 32)
    a.load_oracle_params(p); b.load_oracle_params(p)
    b.capture("k", batch)                      # warm-up inside capture() runs forward + backward only: no update
    assert int(b.step_state[0].item()) == 0
    for _ in range(3):
        a.step_eager(batch)
        b.replay("k")
    torch.cuda.synchronize()
    assert torch.equal(a.params, b.params)     # same kernels, same order, deterministic reductions
    assert int(b.step_state[0].item()) == 3
    sa, sb = a.read_stats(), b.read_stats()      # cost_sum is accumulated with float atomics: order-dependent last bits
    assert abs(sa[0] - sb[0]) <= 1e-5 * abs(sa[0]) and sa[1] == sb[1]


@pytest.mark.parametrize("name,B,N,C,F,conv_dims", [
    ("c2", 1024, 32, 1, 64, [64, 64]),          # BASELINE configs at their FULL per-GPU batch
    ("c3", 512, 50, 1, 75, [50, 50, 50]),
    ("c4", 512, 50, 3, 75, [50, 50, 50]),
    ("c5", 512, 64, 1, 128, [128, 128]),
])
def test_full_batch_training_steps_match_the_oracle(name, B, N, C, F, conv_dims):
    """Two training steps at the BASELINE batch sizes -- the chained launches, channel groups, padded widths, the v5 kernel,
    the fused head, the weight-gradient launch and the reduce + Adam tail exactly as bench.py runs them -- against the C
    restatement of the reference's per-molecule path (oracle/graphconv_ref.c, pinned to the numpy oracle and through it to
    the reference's own layers in tests/test_oracle_c.py): logits, cost, every gradient and the parameters after Adam."""
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    from oracle import cref
    spec = NetSpec(F, conv_dims, N, channels=C)
    tr = Trainer(spec, B, seed=7)
    assert tr.fused_step and tr.chain and (tr.step_chain or name == "c5")
    ref = cref.RefNet(F, conv_dims, C, 2, act=2)
    for k in ref.offsets:
        ref.view(ref.params, k)[...] = tr.views[k].detach().cpu().numpy().reshape(ref.offsets[k][1])
    for step in range(2):
        counts, idx, val, x, labels, mask, _, _ = make_case(B, N, C, F, conv_dims, None, seed=100 + step)
        batch = DeviceBatch.from_host(counts, idx, val, x, labels, N, mask=mask, pad_to=tr.dims[0])
        stats, logits = ref.train_step(counts, idx, val, x, labels, mask, N, lr=tr.lr)
        tr.step_eager(batch)
        torch.cuda.synchronize()
        close(tr.logits, logits, 1e-4)
        cost_sum, correct = tr.read_stats()
        assert abs(cost_sum - float(stats[0])) <= 1e-4 * abs(float(stats[0]))
        assert abs(correct - float(stats[1])) <= 1.0          # an argmax tie may flip one molecule
        for k in ref.offsets:
            want = ref.view(ref.grads, k)
            close(tr.gviews[k].reshape(want.shape), want, 2e-3)       # sums over B * N rows in fp32 (tf32 x3 products)
        for k in ref.offsets:
            want = ref.view(ref.params, k)
            got = tr.views[k].detach().cpu().numpy().reshape(want.shape)
            # Adam's first steps move every parameter by ~lr whatever the gradient's size: compare on that scale
            np.testing.assert_allclose(got, want, rtol=0, atol=0.05 * tr.lr + 1e-6, err_msg="%s after step %d" % (k, step + 1))


@pytest.mark.parametrize("B,N,F,conv_dims", [(700, 32, 64, [64, 64]), (300, 50, 75, [50, 50, 50])])
def test_multi_step_graph_equals_eager_steps(B, N, F, conv_dims):
    """Trainer.capture_many: several steps over different resident batches in ONE graph -- every step after the first is
    launched with KGCN_FLAG_INPUTS_STABLE, i.e. loads and aggregates its first tiles while the previous step's reduce + Adam
    launch still runs -- leaves exactly the parameters of the same steps run one by one, replay after replay."""
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    spec = NetSpec(F, conv_dims, N)
    a, b = Trainer(spec, B, seed=11), Trainer(spec, B, seed=11)
    assert a.step_chain and torch.equal(a.params, b.params)
    batches = []
    for k in range(4):
        counts, idx, val, x, labels, mask, _, _ = make_case(B, N, 1, F, conv_dims, None, seed=20 + k)
        batches.append(DeviceBatch.from_host(counts, idx, val, x, labels, N, mask=mask, pad_to=a.dims[0]))
    b.capture_many("epoch", batches)
    assert int(b.step_state[0].item()) == 0
    for rep in range(3):
        for batch in batches:
            a.step_eager(batch)
        b.replay("epoch")
        torch.cuda.synchronize()
        assert torch.equal(a.params, b.params), rep
    assert int(b.step_state[0].item()) == 12
    sa, sb = a.read_stats(), b.read_stats()
    assert abs(sa[0] - sb[0]) <= 1e-5 * abs(sa[0]) and sa[1] == sb[1]


@pytest.mark.parametrize("B,N,C,F,conv_dims", [
    (700, 32, 1, 64, [64, 64]),        # C2 widths, 5 graphs per CTA: two tiles per job
    (40, 50, 1, 75, [50, 50, 50]),     # C3: three chained layers on padded widths
    (300, 50, 3, 75, [50, 50, 50]),    # C4: the first layer runs as two chained jobs over channel groups {0,1} + {2}
    (9, 32, 1, 32, [64]),              # a single layer
])
def test_chained_launches_equal_per_layer_launches(B, N, C, F, conv_dims, monkeypatch):
    """All forward layers in ONE launch and all dx layers in ONE launch (CTA-local hand-off between layers) give bit-identical
    parameters to one launch per layer: same kernels, same tiles, same accumulation order."""
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    counts, idx, val, x, labels, mask, adjs, p = make_case(B, N, C, F, conv_dims, None, seed=B + N)
    spec = NetSpec(F, conv_dims, N, channels=C, label_dim=2, act="sigmoid")
    monkeypatch.setenv("KGCN_CHAIN", "1")
    monkeypatch.setenv("KGCN_STEP_CHAIN", "0")
    a = Trainer(spec, B, seed=3)
    monkeypatch.setenv("KGCN_CHAIN", "0")
    b = Trainer(spec, B, seed=3)
    monkeypatch.setenv("KGCN_CHAIN", "1")
    monkeypatch.setenv("KGCN_STEP_CHAIN", "1")
    c = Trainer(spec, B, seed=3)          # + the readout head fused into the last forward layer's epilogue
    assert not b.chain and a.fused_step and b.fused_step and not a.step_chain
    assert a.chain and c.step_chain
    batch = DeviceBatch.from_host(counts, idx, val, x, labels, N, mask=mask, pad_to=a.dims[0])
    for _ in range(3):
        a.step_eager(batch)
        b.step_eager(batch)
        c.step_eager(batch)
    torch.cuda.synchronize()
    assert torch.equal(a.logits, b.logits)
    assert torch.equal(a.grads, b.grads)
    assert torch.equal(a.params, b.params)
    # the fused head adds the node rows per warp quarter instead of in index order: equal up to fp32 summation order
    close(c.logits, a.logits.cpu().numpy(), 1e-5)
    close(c.grads, a.grads.cpu().numpy(), 1e-4)
    close(c.params, a.params.cpu().numpy(), 1e-4)
    sa, sc = a.read_stats(), c.read_stats()
    assert abs(sa[0] - sc[0]) <= 1e-4 * abs(sa[0]) and sa[1] == sc[1]
    assert int(c.step_state[0].item()) == 3


@pytest.mark.parametrize("B,N,C,F,conv_dims", [
    (700, 32, 1, 64, [64, 64]),          # C2 widths: ragged last tile (5 graphs per CTA, 4 + 1), one stored G
    (1024, 32, 1, 64, [64, 64]),         # C2 at the full batch
    (300, 50, 1, 75, [50, 50, 50]),      # C3 on padded widths: two stored G (layers 2 and 1), tiles of 2 graphs = 100 of 128 rows
    (37, 20, 1, 32, [32, 64, 32]),       # unequal widths: hard job boundaries, a different plan per weight-gradient job
    (512, 64, 1, 128, [128, 128]),       # C5's shard: the wide-layer (v5) step launch stores G, 32-row operand chunks copy it
    (21, 64, 1, 128, [128, 128]),        # the same kernels on a batch smaller than the grid
    (512, 50, 3, 75, [50, 50, 50]),      # C4: three bond types, G = [G_0 | G_1 | G_2] is 192 wide (its own stage layout in the copying jobs)
    (45, 32, 2, 64, [64, 64]),           # two channels, small batch
])
def test_stored_aggregate_equals_second_gather(B, N, C, F, conv_dims, monkeypatch):
    """The dx jobs of the step launch store G_l = A^T . dU_l and the weight-gradient launch copies it (kgcn_gcn_step_chain_g_f32 /
    kgcn_graphconv_chain_dw_g_f32) -- against KGCN_GSAVE=0, where the weight-gradient launch gathers G_l again: the same per-row
    accumulation order and the same tf32 split, so gradients and parameters must be BIT-identical, and the stored G itself equals
    the batched SpMM of the library on (A^T, dU_l)."""
    from kgcn_b200 import ops
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    counts, idx, val, x, labels, mask, adjs, p = make_case(B, N, C, F, conv_dims, None, seed=B + N)
    spec = NetSpec(F, conv_dims, N, channels=C, label_dim=2, act="sigmoid")
    monkeypatch.setenv("KGCN_GSAVE", "2")          # 2: also with several channels (off by default there: measured slower)
    a = Trainer(spec, B, seed=3)
    monkeypatch.setenv("KGCN_GSAVE", "0")
    b = Trainer(spec, B, seed=3)
    assert a.step_chain and b.step_chain and a.g_save is not None and b.g_save is None
    batch = DeviceBatch.from_host(counts, idx, val, x, labels, N, mask=mask, pad_to=a.dims[0])
    for _ in range(4):
        a.step_eager(batch)
        b.step_eager(batch)
    torch.cuda.synchronize()
    assert torch.equal(a.grads, b.grads)
    assert torch.equal(a.params, b.params)
    for part_a, part_b in zip(a.partials, b.partials):
        assert torch.equal(part_a, part_b)
    csr_t = batch.csr.transposed()
    for l in range(1, len(conv_dims)):
        f = a.dims[l + 1]
        want = torch.empty(B, C, N, f, device="cuda")          # per channel: A_c^T . dU_l (dU shared by the channels)
        ops.bspmm_raw(csr_t, a.du[l], N * f, 0, want, C * N * f, N * f, f)
        torch.cuda.synchronize()
        close(a.g_save[l].view(B, N, C, f), want.permute(0, 2, 1, 3).cpu().numpy(), 1e-6)


def test_training_reduces_loss_on_ring_task():
    """C2 generator (ring-size classification) is learnable: cost_sum falls over 60 steps."""
    from kgcn_b200 import synth
    from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
    rng = np.random.default_rng(0)
    d = synth.ring_graphs(rng, 256, 32, 64, one_hot_features=True)
    batch = DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], 32)
    tr = Trainer(NetSpec(64, [64, 64], 32), 256, lr=0.01)
    tr.step_eager(batch)
    first = tr.read_stats()[0]
    for _ in range(60):
        tr.step_eager(batch)
    last = tr.read_stats()[0]
    assert np.isfinite(last) and last < 0.7 * first, (first, last)


def test_host_fed_pipeline_matches_resident_training():
    """H2D + device-side CSR pack + step (2-slot pipelined) == the same steps on resident batches."""
    from kgcn_b200 import synth
    from kgcn_b200.trainer import DeviceBatch, HostFedPipeline, NetSpec, Trainer
    rng = np.random.default_rng(3)
    spec = NetSpec(64, [64, 64], 32)
    data = [synth.ring_graphs(rng, 64, 32, 64) for _ in range(5)]
    a, b = Trainer(spec, 64, seed=7), Trainer(spec, 64, seed=7)
    want = []
    for d in data:
        a.step_eager(DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], 32))
        want.append(a.read_stats())
    pipe = HostFedPipeline(b, max_nnz=max(d["values"].shape[0] for d in data) + 64, depth=2)
    hosts = [pipe.pin_host_batch(d["counts"], d["indices"], d["values"], d["features"], d["labels"]) for d in data]
    got = list(pipe.run_many(hosts))            # eager device part (no graph captured)
    for (c0, k0), (c1, k1) in zip(want, got):
        assert abs(c0 - c1) <= 1e-4 * abs(c0) and k0 == k1
    assert torch.allclose(a.params, b.params, rtol=1e-5, atol=1e-7)
    assert all(int(s.d_flag.item()) == 0 for s in pipe.slots)
