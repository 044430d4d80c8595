#!/usr/bin/env python
"""Per-phase cycle breakdown of the fused GraphConv kernel (tuning aid; uses the kgcn_debug_fused_times hook)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import _lib, ops, synth
from kgcn_b200.csr import BatchedCSR
B, N, C, fi, fo = (int(v) for v in sys.argv[1:6])
rng = np.random.default_rng(1234)
d = synth.ring_graphs(rng, B, N, fi)
csr = BatchedCSR.from_flat(d["counts"], d["indices"], d["values"], N, N)
x = torch.randn(B, N, fi, device="cuda"); y = torch.empty(B, N, fo, device="cuda")
w = torch.randn(C, fi, fo, device="cuda") * 0.1; b = torch.randn(C, fo, device="cuda") * 0.1
dbg = torch.zeros(148 * 2 * 8, dtype=torch.int64, device="cuda")
hook = _lib.lib.kgcn_debug_fused_times; hook.argtypes = [ctypes.c_void_p]; hook.restype = None
for _ in range(3): ops.graphconv_fwd(csr, x, w, b, 2, 0, out=y)
hook(dbg.data_ptr())
ops.graphconv_fwd(csr, x, w, b, 2, 0, out=y)
torch.cuda.synchronize(); hook(None)
t = dbg.cpu().numpy().reshape(-1, 8); t = t[t[:, 7] > 0]
names = ["wait stage", "convert", "aggregate", "mma issue", "mma wait", "epilogue+sync"]
tiles = t[:, 7].mean(); tot = t[:, :6].sum(1).mean()
print("BM=%s  CTAs %d  tiles/CTA %.1f  cycles/CTA %.0f  cycles/tile %.0f" % (os.environ.get("KGCN_FUSED_BM", "auto"), len(t), tiles, tot, tot / tiles))
for i, n in enumerate(names):
    print("  %-14s %8.0f cycles/tile  %5.1f%%" % (n, t[:, i].mean() / tiles, 100 * t[:, i].mean() / tot))
