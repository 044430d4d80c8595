#!/usr/bin/env python
"""Stress check of the chained launches' hand-over protocols (soft job boundaries, on-chip hand-over, stager warps, early start):
N training steps over rotating full-size batches, once with the chained launches and once with one launch per layer (KGCN_CHAIN=0) --
same kernels, same tiles, same accumulation order, so the parameters must stay BIT-IDENTICAL; a stale or torn tile anywhere in the
chain breaks that.  usage: stress_chain.py [c2|c3|c4|c5] [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kgcn_b200.trainer import DeviceBatch, NetSpec, Trainer

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
w = bench.WORKLOADS[key]
B, N, F, C = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"], w["channels"]
spec = NetSpec(F, w["conv_dims"], N, channels=C)
os.environ["KGCN_CHAIN"], os.environ["KGCN_STEP_CHAIN"] = "1", "0"
a = Trainer(spec, B, seed=5)          # chained forward / dx launches (soft boundaries, hand-over), separate head kernel
os.environ["KGCN_CHAIN"] = "0"
b = Trainer(spec, B, seed=5)          # one launch per layer
os.environ["KGCN_CHAIN"], os.environ["KGCN_STEP_CHAIN"] = "1", "1"
c = Trainer(spec, B, seed=5)          # whole graph-local step in one launch (fused head)
d = Trainer(spec, B, seed=5)          # the same as multi-step CUDA graphs (early start under the previous tail)
os.environ["KGCN_GSAVE"] = "0"
e = Trainer(spec, B, seed=5)          # step chain whose weight-gradient launch gathers A^T.dU again instead of reading the dx jobs' copy
os.environ["KGCN_GSAVE"] = "1"
assert a.chain and not b.chain and c.step_chain
host = bench.make_host_batches(w, 8, seed=77)
batches = [DeviceBatch.from_host(h["counts"], h["indices"], h["values"], h["features"], h["labels"], N, pad_to=a.dims[0]) for h in host]
d.capture_many("epoch", batches)
bad, rel8 = 0, None
for s in range(steps):
    bt = batches[s % len(batches)]
    a.step_eager(bt); b.step_eager(bt); c.step_eager(bt); e.step_eager(bt)
    if s % len(batches) == len(batches) - 1:
        d.replay("epoch")
        torch.cuda.synchronize()
        if rel8 is None:   # fused head vs separate head kernel differ in summation order only: compared before Adam's dynamics amplify it
            rel8 = float((c.params - a.params).abs().max() / a.params.abs().max())
        if not torch.equal(a.params, b.params):
            bad += 1
            print("step %d: chained != per-layer, max |diff| %.3e" % (s, float((a.params - b.params).abs().max())))
        if not torch.equal(c.params, e.params):
            bad += 1
            print("step %d: stored G != second gather, max |diff| %.3e" % (s, float((c.params - e.params).abs().max())))
        if not torch.equal(c.params, d.params):
            bad += 1
            print("step %d: step chain eager != multi-step graph, max |diff| %.3e" % (s, float((c.params - d.params).abs().max())))
print("%s: %d steps, %d mismatches; fused-head chain vs separate head kernel after %d steps: max rel. parameter difference %.2e (summation order)"
      % (key, steps, bad, len(batches), rel8))
sys.exit(1 if bad else 0)
