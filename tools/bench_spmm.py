#!/usr/bin/env python
"""CUDA-graph timing of the standalone batched SpMM kernel: usage bench_spmm.py B N C F"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import ops, synth
from kgcn_b200.csr import BatchedCSR
B, N, C, F = (int(v) for v in sys.argv[1:5])
rng = np.random.default_rng(1234)
rot = 24 if B <= 2048 else 4
csrs, xs, ys = [], [], []
for _ in range(rot):
    if C == 1 and N == 32:
        d = synth.ring_graphs(rng, B, N, F); counts, idx, val = d["counts"], d["indices"], d["values"]
    else:
        counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
    csrs.append(BatchedCSR.from_flat(counts, idx, val, N, N))
    xs.append(torch.randn(B, N, F, device="cuda")); ys.append(torch.empty(B, N, F, device="cuda"))
run = lambda: [ops.bspmm_raw(csrs[i], xs[i], N * F, 0, ys[i], N * F, 0, F) for i in range(rot)]
run(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
best = 1e9
for trial in range(5):
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(40): g.replay()
    e.record(); torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) * 1e3 / (40 * rot))
nnz = np.mean([c.nnz for c in csrs])
nbytes = 8 * B * N * F + 8 * nnz + 4 * C * B * (N + 1)
print("spmm B=%d N=%d C=%d F=%d G=%s  %.2f us  %.0f GB/s  %.1f%% of 6556" % (B, N, C, F, os.environ.get("KGCN_SPMM_G", "auto"), best, nbytes / best / 1e3, nbytes / best / 1e3 / 65.562))
