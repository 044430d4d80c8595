#!/usr/bin/env python
"""Kernel-level timings on one B200 (CUDA events, rotating buffers so operands come from HBM)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import ops, synth  # noqa: E402
from kgcn_b200.csr import BatchedCSR  # noqa: E402


def timeit(fn, n_rot, iters=200, warm=20):
    for i in range(warm):
        fn(i % n_rot)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i % n_rot)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rot", type=int, default=16)
    args = ap.parse_args()
    peak = 6556.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rng = np.random.default_rng(1234)
    cfgs = [("C2", 1024, 32, 1, 64, 64), ("C3L1", 512, 50, 1, 75, 50), ("C3L2", 512, 50, 1, 50, 50),
            ("C4", 512, 50, 3, 75, 50), ("C5", 512, 64, 1, 128, 128), ("C5x8", 4096, 64, 1, 128, 128),
            ("C2x16", 16384, 32, 1, 64, 64)]
    for name, B, N, C, fi, fo in cfgs:
        if name.startswith("C2"):
            d = synth.ring_graphs(rng, B, N, fi)
            counts, idx, val = d["counts"], d["indices"], d["values"]
        else:
            counts, idx, val = synth.random_molecule_coo(rng, B, N, C)
        csr = BatchedCSR.from_flat(counts, idx, val, N, N)
        nnz = csr.nnz
        rot = args.rot if B <= 4096 else 4
        xs = [torch.randn(B, N, fi, device="cuda") for _ in range(rot)]
        hs = [torch.randn(B, N, fo, device="cuda") for _ in range(rot)]
        ys = [torch.empty(B, N, fo, device="cuda") for _ in range(rot)]
        zs = [torch.empty(B, N, fi, device="cuda") for _ in range(rot)]
        w = torch.randn(C, fi, fo, device="cuda") * 0.1
        b = torch.randn(C, fo, device="cuda") * 0.1
        # standalone SpMM  Y = sum_c A_c X   on F = f_out features (SURVEY Appendix D column bytes_spmm)
        bytes_spmm = 4 * B * N * fo * 2 + 8 * nnz + 4 * C * B * (N + 1)
        t = timeit(lambda i: ops.bspmm_raw(csr, hs[i], N * fo, 0, ys[i], N * fo, 0, fo), rot)
        print("%-6s spmm(F=%d)      %8.2f us  %7.1f GB/s  %5.1f%% of measured %.0f  (%.2f MB, nnz/graph %.1f)" %
              (name, fo, t, bytes_spmm / t / 1e3, 100 * bytes_spmm / t / 1e3 / peak, peak, bytes_spmm / 1e6, nnz / B))
        bytes_layer = 4 * B * N * (fi + fo) + 8 * nnz + 4 * C * B * (N + 1) + 4 * C * fi * fo + 4 * C * fo
        for flags, tag in ((0, "auto"), (1, "ref-order")):
            t = timeit(lambda i: ops.graphconv_fwd(csr, xs[i], w, b, 2, flags, out=ys[i]), rot)
            print("%-6s layer fwd %-9s %8.2f us  %7.1f GB/s  %5.1f%%  -> %.3g molecules/s/layer" %
                  (name, tag, t, bytes_layer / t / 1e3, 100 * bytes_layer / t / 1e3 / peak, B / t * 1e6))
        t = timeit(lambda i: ops.graphconv_bwd(csr, xs[i], w, 2, ys[i], hs[i]), rot, iters=50, warm=5)
        print("%-6s layer bwd           %8.2f us" % (name, t))
        # plain device copy of the same footprint for reference
        t = timeit(lambda i: ys[i].copy_(hs[i]), rot)
        print("%-6s torch copy (same Y) %8.2f us  %7.1f GB/s" % (name, t, 8 * B * N * fo / t / 1e3))


if __name__ == "__main__":
    main()
