#!/usr/bin/env python
"""Timeline of the chained step kernel (kgcn_debug_v4_chain_times): per job, when each role reached its milestones, in
microseconds since kernel entry, median over CTAs.  usage: chain_timeline.py [c2|c3]"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kgcn_b200 import _lib
from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[key]
B, N, F, C = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"], w["channels"]
tr = Trainer(NetSpec(F, w["conv_dims"], N, channels=C), B)
host = bench.make_host_batches(w, 16, seed=1)   # 16 batches > L2: the stamped launch finds its inputs in HBM, the weights in L2
batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, pad_to=tr.dims[0]) for d in host]
for _ in range(2):
    for b in batches: tr.step_eager(b)
torch.cuda.synchronize()
dbg = torch.zeros(148 * 128, dtype=torch.int64, device="cuda")
hook = _lib.lib.kgcn_debug_v4_chain_times; hook.argtypes = [ctypes.c_void_p]; hook.restype = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
hook(dbg.data_ptr())
tr.step_chain and tr._step_chain(batches[0], torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); hook(None)
t = dbg.cpu().numpy().reshape(148, 128)
t = t[t[:, 126] > 0]
clk = 1.92e3   # cycles per us at ~1.92 GHz
names = ["agg: job start", "agg: first stage landed", "agg: tile 0 in Z", "agg: last tile in Z", "mma: tile 0 issued", "mma: last tile issued",
         "epi: first accumulator", "epi: tile 0 stored", "epi: last tile stored", "tma: first issue", "epi: B operand staged", "epi: job end", "head: column sums done", "head: barrier 1 passed",
         "head: graph heads done", "head: barrier 2 passed"]
n_jobs = 2 * len(w["conv_dims"]) - 1
print("CTAs %d, kernel %.2f us (median), max %.2f us" % (len(t), np.median(t[:, 127] - t[:, 126]) / clk, (t[:, 127] - t[:, 126]).max() / clk))
order = [0, 9, 10, 1, 2, 4, 6, 12, 13, 14, 15, 7, 3, 5, 8, 11]
for j in range(n_jobs):
    print("job %d" % j)
    for ev in order:
        v = t[:, j * 16 + ev]
        ok = v > 0
        if ok.any(): print("   %-26s %7.2f us" % (names[ev], np.median(v[ok] - t[ok, 126]) / clk))
