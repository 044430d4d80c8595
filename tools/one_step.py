#!/usr/bin/env python
"""A few eager training steps of one workload (for ncu captures).  usage: one_step.py [c2|c3|c4|c5] [n_steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer
key = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = bench.WORKLOADS[key]
B, N, F, C = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"], w["channels"]
tr = Trainer(NetSpec(F, w["conv_dims"], N, channels=C), B)
host = bench.make_host_batches(w, 3, seed=1)
batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, pad_to=tr.dims[0]) for d in host]
for i in range(n):
    tr.step_eager(batches[i % 3])
torch.cuda.synchronize()
print("steps", tr.steps_done(), tr.read_stats())
