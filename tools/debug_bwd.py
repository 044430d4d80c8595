"""Bring-up aid: fused GraphConv backward vs the decomposed exact-fp32 path on one shape (prints max errors)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kgcn_b200 import csr as csr_mod, ops  # noqa: E402
from kgcn_b200.synth import random_molecule_coo  # noqa: E402


def main():
    B, N, fi, fo = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (24, 32, 64, 64))]
    act = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    bcast = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    rng = np.random.default_rng(0)
    adjs = []
    for b in range(B):
        nnz = int(rng.integers(0, int(N * 3.5)))
        idx = rng.integers(0, N, size=(nnz, 2)).astype(np.int32)
        adjs.append([(idx, rng.standard_normal(nnz).astype(np.float32), [N, N])])
    c = csr_mod.BatchedCSR.from_coo_lists(adjs)
    x = torch.randn(B, N, fi, device="cuda")
    w = torch.randn(1, fi, fo, device="cuda") * 0.2
    y = torch.rand(B, N, fo, device="cuda")
    dy = torch.randn((B, fo) if bcast else (B, N, fo), device="cuda")
    base = 2 if bcast else 0
    for need_dx in (True, False):
        got = ops.graphconv_bwd(c, x, w, act, y, dy, need_dx=need_dx, flags=base)
        ref = ops.graphconv_bwd(c, x, w, act, y, dy, need_dx=need_dx, flags=base | 1)
        torch.cuda.synchronize()
        for name, g, r in zip(("dx", "dw", "db"), got, ref):
            if g is None:
                continue
            err = (g - r).abs().max().item()
            print(f"need_dx={need_dx} {name}: max|ref|={r.abs().max().item():.4g} max err={err:.3g} "
                  f"nan={bool(torch.isnan(g).any())}", flush=True)
            if name == "dw" and err > 1e-3 * r.abs().max().item():
                d = (g - r).abs()[0]
                print("   dw err by row block (8x8 of 64x64 means):")
                hh, ww = d.shape
                blk = d[: hh // 8 * 8, : ww // 8 * 8].reshape(8, hh // 8, 8, ww // 8).mean(dim=(1, 3))
                print(np.array2string(blk.cpu().numpy(), precision=3, suppress_small=True))
                print("   ratio g/r sample:", (g[0, :4, :4] / r[0, :4, :4]).cpu().numpy())


if __name__ == "__main__":
    main()
