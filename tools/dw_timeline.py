#!/usr/bin/env python
"""Timeline of the weight-gradient launch (kgcn_debug_dw_times): per tile of a CTA, when the producer issued its copies, when the
workers saw the stage full / finished the tile's operand chunks, when the MMA warp saw the first chunk / issued the last one;
microseconds since kernel entry, median over CTAs.  The stamps are compiled in only with
`touch kgcn_b200/csrc/graphconv_fused_dw.cu && make -C kgcn_b200/csrc TIMELINE=1` (rebuild without it afterwards).
usage: dw_timeline.py [c2|c3|c5]"""
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
host = bench.make_host_batches(w, 16, seed=1)   # 16 batches > L2: the stamped launch finds its inputs in HBM
batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, pad_to=tr.dims[0]) for d in host]
for _ in range(2):
    for b in batches: tr.step_eager(b)
torch.cuda.synchronize()
dbg = torch.zeros(148 * 256, dtype=torch.int64, device="cuda")
hook = _lib.lib.kgcn_debug_dw_times; hook.argtypes = [ctypes.c_void_p]; hook.restype = None
st = torch.cuda.current_stream().cuda_stream
tr._launch_step_chain(batches[0], st) if tr.step_chain else tr._fwd_bwd(batches[0])
torch.cuda.synchronize()
hook(dbg.data_ptr())
tr._launch_dw_chain(batches[0], st)
torch.cuda.synchronize(); hook(None)
t = dbg.cpu().numpy().reshape(148, 256)
t = t[t[:, 0] > 0]
clk = 1.92e3
rel = lambda col: np.median((t[:, col] - t[:, 0])[t[:, col] > 0]) / clk if (t[:, col] > 0).any() else float("nan")
print("CTAs %d; setup done %.2f us, workers left the job loop %.2f us, partials written %.2f us (median; max %.2f us)"
      % (len(t), rel(3), rel(1), rel(2), ((t[:, 2] - t[:, 0]).max()) / clk))
print("tile   tma issued   stage full   workers done   mma first chunk   mma last issued")
for T in range(30):
    if not (t[:, 8 + 8 * T] > 0).any(): break
    print("%4d   %10.2f   %10.2f   %12.2f   %15.2f   %15.2f" % ((T,) + tuple(rel(8 + 8 * T + e) for e in (0, 1, 2, 3, 4))))
