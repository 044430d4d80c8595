#!/usr/bin/env python
"""A few backward calls of one layer shape, for `ncu --metrics gpu__time_duration.sum` (per-kernel durations).
usage: bwd_kernels.py [B N C F_in F_out]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import _lib, ops, synth
from kgcn_b200.csr import BatchedCSR

a = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else [1024, 32, 1, 64, 64]
B, N, C, fi, fo = a
rng = np.random.default_rng(0)
bs = []
for _ in range(3):
    if C > 1 or N != 32:
        counts, indices, values = synth.random_molecule_coo(rng, B, N, C)
        feats = rng.standard_normal((B, N, fi)).astype(np.float32)
    else:
        d = synth.ring_graphs(rng, B, N, fi)
        counts, indices, values, feats = d["counts"], d["indices"], d["values"], d["features"]
    bs.append((BatchedCSR.from_flat(counts, indices, values, N, N), torch.as_tensor(feats).cuda()))
w = torch.randn(C, fi, fo, device="cuda") * 0.1
bias = torch.zeros(C, fo, device="cuda")
y = torch.rand(B, N, fo, device="cuda")
dy = torch.randn(B, N, fo, device="cuda")
dg = torch.randn(B, fo, device="cuda")
for it in range(4):
    csr, x = bs[it % 3]
    ops.graphconv_fwd(csr, x, w, bias, 2, 0)
    ops.graphconv_bwd(csr, x, w, 2, y, dy)
    ops.graphconv_bwd(csr, x, w, 2, y, dg, flags=_lib.FLAG_DY_BROADCAST)
    ops.graphconv_bwd(csr, x, w, 2, y, dy, need_dx=False)
torch.cuda.synchronize()
