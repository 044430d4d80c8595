#!/usr/bin/env python
"""Warm (graph-replayed, rotating batches) time of the chained launches of the C2 / C3 training step, alone and in pieces.
usage: chain_times.py [c2|c3]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from kgcn_b200._lib import check, lib, ptr
from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[key]
B, N, F, C, ROT = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"], w["channels"], 12
spec = NetSpec(F, w["conv_dims"], N, channels=C)
tr = Trainer(spec, B)
host = bench.make_host_batches(w, ROT, seed=1)
batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, pad_to=tr.dims[0]) for d in host]
L = len(w["conv_dims"])
st = lambda: torch.cuda.current_stream().cuda_stream
print("padded", tr.padded, "fused_step", tr.fused_step, "chain", tr.chain, "step_chain", tr.step_chain, "g_save", tr.g_save is not None)
for b in batches[:2]:
    tr.step_eager(b)
torch.cuda.synchronize()
xp = lambda b: (ctypes.c_void_p * L)(*([b.features.data_ptr()] + [a.data_ptr() for a in tr.acts[1:L]]))

def step_chain(i): tr._launch_step_chain(batches[i], st())   # incl. the stored G when the trainer uses it (KGCN_GSAVE)
def fwd_chain(i): check(lib.kgcn_graphconv_chain_fwd_f32(ptr(batches[i].csr.rowptr), ptr(batches[i].csr.col), ptr(batches[i].csr.val), B, C, N, L,
    tr._dims_c, tr._ldims_c, ptr(batches[i].features), tr._w_ptrs, tr._b_ptrs, tr._y_ptrs, tr.act, st()))
def fwd_one(i): check(lib.kgcn_graphconv_chain_fwd_f32(ptr(batches[i].csr.rowptr), ptr(batches[i].csr.col), ptr(batches[i].csr.val), B, C, N, 1,
    tr._dims_c, tr._ldims_c, ptr(batches[i].features), tr._w_ptrs, tr._b_ptrs, tr._y_ptrs, tr.act, st()))
def dx_chain(i): check(lib.kgcn_graphconv_chain_dx_f32(ptr(batches[i].csr.rowptr_t), ptr(batches[i].csr.col_t), ptr(batches[i].csr.val_t), B, C, N, L,
    tr._dims_c, xp(batches[i]), tr._w_ptrs, tr._du_ptrs, tr.act, st()))
def dw_chain(i): tr._launch_dw_chain(batches[i], st())
def dw_one(i): check(lib.kgcn_graphconv_chain_dw_f32(ptr(batches[i].csr.rowptr_t), ptr(batches[i].csr.col_t), ptr(batches[i].csr.val_t), B, C, N, 1,
    tr._dims_c, xp(batches[i]), tr._du_ptrs, tr._part_ptrs, tr._part_bytes, st()))
def head(i):
    tr._last_nodes = tr.acts[L]; tr._head(batches[i], tr.f_head, st(), train=False)
def head_du(i):
    tr._last_nodes = tr.acts[L]; tr._head_in_chain = False; tr._head(batches[i], tr.f_head, st(), train=True)
def tail(i):
    tr._head_in_chain = tr.step_chain; tr._optimizer(st())
def whole(i): tr.step_eager(batches[i])
def whole_stable(i): tr.step_eager(batches[i], stable=i > 0)

for name, fn in (("step chain (fwd x%d + head + dx x%d)" % (L, L - 1), step_chain), ("fwd chain x%d" % L, fwd_chain), ("fwd single job", fwd_one),
                 ("dx chain x%d" % (L - 1), dx_chain), ("dW chain x%d" % L, dw_chain), ("dW single job", dw_one), ("head kernel (infer)", head), ("head kernel (train: + dU)", head_du),
                 ("tail", tail), ("whole step", whole), ("whole step, inputs stable", whole_stable)):
    if L == 1 and "dx" in name: continue
    if "step chain" in name and not tr.step_chain: continue
    if ("chain" in name or "single job" in name) and not tr.chain: continue
    for i in range(ROT): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ROT): fn(i)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): g.replay()
    e.record(); torch.cuda.synchronize()
    print("%-42s %7.2f us" % (name, s.elapsed_time(e) * 1e3 / (20 * ROT)))
