#!/usr/bin/env python
"""Per-role phase cycle breakdown of the v4 fused GraphConv kernel (tuning aid; kgcn_debug_v4_times hook)."""
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
dbg = torch.zeros(148 * 18, dtype=torch.int64, device="cuda")
hook = _lib.lib.kgcn_debug_v4_times; hook.argtypes = [ctypes.c_void_p]; hook.restype = None
for _ in range(3): ops.graphconv_fwd(csr, x, w, b, 2, 0, out=y)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
hook(dbg.data_ptr())
ops.graphconv_fwd(csr, x, w, b, 2, 0, out=y)
torch.cuda.synchronize(); hook(None)
raw = dbg.cpu().numpy(); extra = raw[148 * 16:].reshape(-1, 2); t = raw[:148 * 16].reshape(-1, 16); keep = t[:, 13] > 0; t = t[keep]; extra = extra[keep]
names = ["A wait stage", "A wait Z free", "A gather", "A unperm+split+st", "A publish", "MMA wait Z", "MMA wait acc", "MMA issue",
         "E wait acc", "E wait store read", "P wait stage free", "total", "setup", "tiles"]
tiles = t[:, 13].mean()
print("groups=%s CTAs %d  tiles/CTA %.1f  total %.0f cycles  setup %.0f  per tile %.0f" % (os.environ.get("KGCN_V4_GROUPS", "auto"), len(t), tiles, t[:, 11].mean(), t[:, 12].mean(), (t[:, 11].mean() - t[:, 12].mean()) / tiles))
for i, n in enumerate(names[:11]):
    print("  %-20s %8.0f cycles/CTA  %8.0f /tile" % (n, t[:, i].mean(), t[:, i].mean() / tiles))
for n, v in (("E tmem ld", t[:, 14]), ("E act", t[:, 15]), ("E sts", extra[:, 0]), ("E fence+store", extra[:, 1])):
    print("  %-20s %8.0f cycles/CTA  %8.0f /tile" % (n, v.mean(), v.mean() / tiles))
