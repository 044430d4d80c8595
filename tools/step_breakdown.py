#!/usr/bin/env python
"""Warm (CUDA-graph replayed, rotating batches) time of every C-ABI call of the C2 training step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import _lib
from kgcn_b200._lib import check, lib, ptr
from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
from kgcn_b200 import synth

B, N, F = 1024, 32, 64
ROT = 12
rng = np.random.default_rng(0)
spec = NetSpec(F, [64, 64], N)
tr = Trainer(spec, B)
batches = []
for _ in range(ROT):
    d = synth.ring_graphs(rng, B, N, F)
    batches.append(DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N))
for b in batches[:2]:
    tr.step_eager(b)
torch.cuda.synchronize()
st = lambda: torch.cuda.current_stream().cuda_stream
ws = tr.ws
acts1 = [torch.rand(B, N, 64, device="cuda") for _ in range(ROT)]
acts2 = [torch.rand(B, N, 64, device="cuda") for _ in range(ROT)]
dys = [torch.randn(B, N, 64, device="cuda") for _ in range(ROT)]
dxs = [torch.empty(B, N, 64, device="cuda") for _ in range(ROT)]

def fwd(i):
    b = batches[i]
    check(lib.kgcn_graphconv_fwd_f32(ptr(b.csr.rowptr), ptr(b.csr.col), ptr(b.csr.val), B, 1, N, ptr(b.features), F, ptr(tr.views["conv0/kernel"]),
                                     ptr(tr.views["conv0/bias"]), 64, 2, ptr(acts1[i]), 0, ptr(ws), ws.numel(), st()))
def gather(i):
    check(lib.kgcn_gather_fwd_f32(ptr(acts2[i]), B, N, 64, ptr(tr.gathered), st()))
def head(i):
    b = batches[i]
    check(lib.kgcn_readout_xent_f32(ptr(tr.gathered), B, 64, ptr(tr.views["dense/kernel"]), ptr(tr.views["dense/bias"]), 2, ptr(b.labels), ptr(b.mask),
                                    1.0 / B, ptr(tr.logits), ptr(tr.prediction), ptr(tr.stats), ptr(tr.dlogits), ptr(tr.dgathered),
                                    ptr(tr.gviews["dense/kernel"]), ptr(tr.gviews["dense/bias"]), ptr(ws), ws.numel(), st()))
def gather_head(i):   # what the trainer launches: GraphGather fused into the readout head
    b = batches[i]
    check(lib.kgcn_gather_readout_xent_f32(ptr(acts2[i]), B, N, 64, ptr(tr.gathered), ptr(tr.views["dense/kernel"]), ptr(tr.views["dense/bias"]), 2,
                                           ptr(b.labels), ptr(b.mask), 1.0 / B, ptr(tr.logits), ptr(tr.prediction), ptr(tr.stats), ptr(tr.dlogits),
                                           ptr(tr.dgathered), ptr(tr.gviews["dense/kernel"]), ptr(tr.gviews["dense/bias"]), ptr(ws), ws.numel(), st()))
def bwd2(i):   # last conv layer: dy broadcast from the gather gradient, dx needed
    b = batches[i]
    check(lib.kgcn_graphconv_bwd_f32(ptr(b.csr.rowptr_t), ptr(b.csr.col_t), ptr(b.csr.val_t), B, 1, N, ptr(acts1[i]), 64, ptr(tr.views["conv1/kernel"]), 64, 2,
                                     ptr(acts2[i]), ptr(tr.dgathered), ptr(dxs[i]), ptr(tr.gviews["conv1/kernel"]), ptr(tr.gviews["conv1/bias"]),
                                     _lib.FLAG_DY_BROADCAST, ptr(ws), ws.numel(), st()))
def bwd1(i):   # first conv layer: no dx
    b = batches[i]
    check(lib.kgcn_graphconv_bwd_f32(ptr(b.csr.rowptr_t), ptr(b.csr.col_t), ptr(b.csr.val_t), B, 1, N, ptr(b.features), F, ptr(tr.views["conv0/kernel"]), 64, 2,
                                     ptr(acts1[i]), ptr(dys[i]), None, ptr(tr.gviews["conv0/kernel"]), ptr(tr.gviews["conv0/bias"]), 0, ptr(ws), ws.numel(), st()))
def adam(i):
    tr._optimizer(st())

total = 0.0
for name, fn, mult in (("graphconv_fwd (fused)", fwd, 2), ("gather_fwd (layer API only)", gather, 0), ("readout_xent (layer API only)", head, 0),
                       ("gather + readout_xent (+dW)", gather_head, 1), ("graphconv_bwd L2 (dx, bcast)", bwd2, 1),
                       ("graphconv_bwd L1 (no dx)", bwd1, 1), ("adam", adam, 1)):
    for i in range(ROT): fn(i)
    torch.cuda.synchronize()
    c0 = lib.kgcn_launch_count(); fn(0); n_launch = lib.kgcn_launch_count() - c0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ROT): fn(i)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / (20 * ROT)
    total += us * mult
    print("%-30s %7.2f us  x%d  (%d launches)" % (name, us, mult, n_launch))
print("sum over the step: %.1f us" % total)
