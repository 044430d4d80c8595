#!/usr/bin/env python
"""CUDA-graph timing of one fused GraphConv layer (no host launch overhead): usage bench_layer.py B N C F_in F_out"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import ops, synth
from kgcn_b200.csr import BatchedCSR
B, N, C, fi, fo = (int(v) for v in sys.argv[1:6])
rng = np.random.default_rng(1234)
rot = 24 if B <= 2048 else 4
if C == 1 and N == 32:
    d = synth.ring_graphs(rng, B, N, fi); counts, idx, val = d["counts"], d["indices"], d["values"]
else:
    counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
csr = BatchedCSR.from_flat(counts, idx, val, N, N)
xs = [torch.randn(B, N, fi, device="cuda") for _ in range(rot)]
ys = [torch.empty(B, N, fo, device="cuda") for _ in range(rot)]
w = torch.randn(C, fi, fo, device="cuda") * 0.1
b = torch.randn(C, fo, device="cuda") * 0.1
for flags, tag in ((0, "fused"), (1, "decomposed")):
    run = lambda: [ops.graphconv_fwd(csr, xs[i], w, b, 2, flags, out=ys[i]) for i in range(rot)]
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    s.record()
    for _ in range(reps): g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / (reps * rot)
    nbytes = 4 * B * N * (fi + fo) + 8 * csr.nnz + 4 * C * B * (N + 1) + 4 * C * fi * fo + 4 * C * fo
    print("%-10s B=%d N=%d C=%d %d->%d  BM=%s  %8.2f us/layer  %7.1f GB/s (%.1f%% of 6556)  %.3g mol/s" %
          (tag, B, N, C, fi, fo, os.environ.get("KGCN_FUSED_BM", "auto"), us, nbytes / us / 1e3, nbytes / us / 1e3 / 65.562, B / us * 1e6))
