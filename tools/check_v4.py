#!/usr/bin/env python
"""Fused v4 kernel vs the decomposed reference-order path (exact-fp32 FFMA GEMM + SpMM) on a few shapes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import ops, synth
from kgcn_b200.csr import BatchedCSR

shapes = [(5, 32, 1, 64, 64, 2), (1024, 32, 1, 64, 64, 2), (333, 32, 1, 64, 64, 1), (64, 50, 1, 64, 64, 3), (100, 17, 2, 32, 48, 0),
          (257, 32, 3, 64, 64, 2), (40, 64, 1, 128, 128, 1), (1500, 10, 1, 32, 16, 2), (3, 128, 1, 64, 8, 0), (4096, 32, 1, 64, 64, 2)]
rng = np.random.default_rng(7)
bad = 0
for (B, N, C, fi, fo, act) in shapes:
    if C == 1 and N == 32:
        d = synth.ring_graphs(rng, B, N, fi); counts, idx, val = d["counts"], d["indices"], d["values"]
        val = (val * rng.uniform(0.5, 1.5, size=val.shape)).astype(np.float32)
    else:
        counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
    csr = BatchedCSR.from_flat(counts, idx, val, N, N)
    x = torch.randn(B, N, fi, device="cuda")
    w = torch.randn(C, fi, fo, device="cuda") * 0.2
    b = torch.randn(C, fo, device="cuda") * 0.3
    y = ops.graphconv_fwd(csr, x, w, b, act, 0)
    ref = ops.graphconv_fwd(csr, x, w, b, act, 1)
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item()
    tol = 1e-5 * ref.abs().max().item() + 1e-6
    ok = err <= 2 * tol and bool(torch.isfinite(y).all())
    bad += not ok
    print("B=%d N=%d C=%d %d->%d act=%d  max|err|=%.3g (max|ref|=%.3g) %s" % (B, N, C, fi, fo, act, err, ref.abs().max().item(), "ok" if ok else "FAIL"), flush=True)
print("FAILED" if bad else "ALL OK")
sys.exit(1 if bad else 0)
