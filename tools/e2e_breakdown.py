#!/usr/bin/env python
"""Where the end-to-end (host-fed) step of bench.py spends its time: usage e2e_breakdown.py [c2 c5 ...]

For each workload, on one GPU: (a) the step's H2D copy alone, back to back on the copy stream; (b) the device part
alone (the slot's CUDA graph: 2 pack launches + the training step), back to back on the compute stream; (c) the pack
launches alone; (d) the pipelined loop bench.py times (HostFedPipeline.run_many).  One JSON line per workload."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    keys = sys.argv[1:] or ["c2", "c5"]
    args = argparse.Namespace(verbose=False, steps=200, warmup=5)
    b = bench.Bench(args)
    torch = b.torch
    from kgcn_b200.trainer import DeviceBatch, HostFedPipeline, NetSpec, Trainer
    for key in keys:
        w = bench.WORKLOADS[key]
        B, N, F, C = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"], w["channels"]
        spec = NetSpec(F, w["conv_dims"], N, channels=C, label_dim=w["label_dim"], act=w["act"])
        tr = Trainer(spec, B, device=b.dev, lr=0.01, world_size=1, seed=1234, rank=0)
        w2 = dict(w, n_rot=int(os.environ.get("E2E_HOST_BATCHES", "4")))   # bench.py rotates w["n_rot"] pinned batches (24 for c2: never cache-resident on the host)
        batches, host = b.make_batches(w2)
        if batches is None:
            batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, device=b.dev,
                                             pad_to=tr.dims[0]) for d in host]
        tr._fwd_bwd(batches[0])
        torch.cuda.synchronize()
        max_nnz = int(max(d["values"].shape[0] for d in host) * 1.1) + 64
        pipe = HostFedPipeline(tr, max_nnz, train=True, depth=int(os.environ.get("KGCN_E2E_DEPTH", "2")))
        pinned = [pipe.pin_host_batch(d["counts"], d["indices"], d["values"], d["features"], d["labels"]) for d in host]
        pipe.capture()
        for _ in pipe.run_many(pinned[i % len(pinned)] for i in range(6)):
            pass
        torch.cuda.synchronize()
        n = 100
        out = {"workload": key, "h2d_bytes": pipe.h2d_bytes(pinned[0])}

        # (a) copies alone
        t0 = time.perf_counter()
        with torch.cuda.stream(pipe.copy_stream):
            for i in range(n):
                slot, h = pipe.slots[i % len(pipe.slots)], pinned[i % len(pinned)]
                slot.d_all.copy_(h["blob"], non_blocking=True)
        pipe.copy_stream.synchronize()
        out["copies_alone_ms"] = (time.perf_counter() - t0) / n * 1e3
        out["copies_alone_gbs"] = out["h2d_bytes"] / out["copies_alone_ms"] / 1e6

        # (b) device part alone (graph replays)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(pipe.compute_stream):
            s.record()
            for i in range(n):
                for g in pipe.slots[i % len(pipe.slots)].graph:
                    g.replay()
            e.record()
        torch.cuda.synchronize()
        out["device_part_ms"] = s.elapsed_time(e) / n

        # (c) the pack launches alone (eager)
        with torch.cuda.stream(pipe.compute_stream):
            pipe._pack(pipe.slots[0])
            s.record()
            for i in range(n):
                pipe._pack(pipe.slots[i % len(pipe.slots)])
            e.record()
        torch.cuda.synchronize()
        out["pack_2_launches_ms"] = s.elapsed_time(e) / n

        # (d) the pipelined loop
        t0 = time.perf_counter()
        for _ in pipe.run_many(pinned[i % len(pinned)] for i in range(n)):
            pass
        torch.cuda.synchronize()
        out["pipelined_ms"] = (time.perf_counter() - t0) / n * 1e3
        print(json.dumps(out), flush=True)
        del pipe, tr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
