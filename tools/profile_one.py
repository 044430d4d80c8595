#!/usr/bin/env python
"""Tiny driver for ncu: launches one kernel family a few times on BASELINE shapes.
usage: profile_one.py {spmm|layer|layer_ref|bwd} [B N C F_in F_out] [iters]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import ops, synth  # noqa: E402
from kgcn_b200.csr import BatchedCSR  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "spmm"
B, N, C, fi, fo = (int(v) for v in sys.argv[2:7]) if len(sys.argv) >= 7 else (1024, 32, 1, 64, 64)
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 4
rng = np.random.default_rng(1234)
if C == 1 and N == 32:
    d = synth.ring_graphs(rng, B, N, fi)
    counts, idx, val = d["counts"], d["indices"], d["values"]
else:
    counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
csr = BatchedCSR.from_flat(counts, idx, val, N, N)
rot = 4
xs = [torch.randn(B, N, fi, device="cuda") for _ in range(rot)]
hs = [torch.randn(B, N, fo, device="cuda") for _ in range(rot)]
ys = [torch.empty(B, N, fo, device="cuda") for _ in range(rot)]
w = torch.randn(C, fi, fo, device="cuda") * 0.1
b = torch.randn(C, fo, device="cuda") * 0.1
for i in range(iters):
    k = i % rot
    if what == "spmm":
        ops.bspmm_raw(csr, hs[k], N * fo, 0, ys[k], N * fo, 0, fo)
    elif what == "layer":
        ops.graphconv_fwd(csr, xs[k], w, b, 2, 0, out=ys[k])
    elif what == "layer_ref":
        ops.graphconv_fwd(csr, xs[k], w, b, 2, 1, out=ys[k])
    elif what == "bwd":
        ops.graphconv_bwd(csr, xs[k], w, 2, ys[k], hs[k])
torch.cuda.synchronize()
print("done", what, B, N, C, fi, fo, "nnz", csr.nnz)
