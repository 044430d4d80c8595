#!/usr/bin/env python
"""Concurrent pinned host -> device copy bandwidth per rank (the e2e path's limiter), with the default pinned allocation and
with the pinned buffer bound to the GPU's own NUMA node (kgcn_b200.hostmem).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/h2d_concurrent.py

Every rank copies CHUNK bytes (the e2e step's 9.8 MB by default) REPS times, first rank by rank (alone), then all ranks at
once; rank 0 prints one JSON line with the per-rank GB/s, the GPUs' NUMA nodes, this process' CPU / memory affinity and
`nvidia-smi topo -m`."""
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kgcn_b200 import hostmem  # noqa: E402

CHUNK = int(os.environ.get("H2D_CHUNK", 9807568))
REPS = int(os.environ.get("H2D_REPS", 200))


def bandwidth(host, dev, stream):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for _ in range(REPS):
            dev.copy_(host, non_blocking=True)
    stream.synchronize()
    return CHUNK * REPS / (time.perf_counter() - t0) / 1e9


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.empty(CHUNK, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    node = hostmem.gpu_numa_node(local)
    out = {"rank": rank, "gpu_numa_node": node, "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] + sorted(os.sched_getaffinity(0))[-1:],
           "mems_allowed": hostmem.mems_allowed()}
    for kind in ("default", "numa_local"):
        if kind == "numa_local":
            with hostmem.numa_preferred(node):
                host = torch.empty(CHUNK, dtype=torch.uint8).pin_memory()
                host.fill_(1)
        else:
            host = torch.empty(CHUNK, dtype=torch.uint8).pin_memory()
            host.fill_(1)
        out[kind + "_pages_on_node"] = hostmem.node_of_buffer(host)
        bandwidth(host, dev, stream)
        alone = None
        for r in range(world):          # rank by rank
            if world > 1:
                dist.barrier()
            if r == rank:
                alone = bandwidth(host, dev, stream)
        if world > 1:
            dist.barrier()
        out[kind + "_alone_gbs"] = alone
        out[kind + "_concurrent_gbs"] = bandwidth(host, dev, stream)
        del host
    gathered = [None] * world
    if world > 1:
        dist.all_gather_object(gathered, out)
    else:
        gathered = [out]
    if rank == 0:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True).stdout
        print(json.dumps({"chunk_bytes": CHUNK, "reps": REPS, "world": world, "ranks": gathered,
                          "sum_concurrent_default_gbs": sum(g["default_concurrent_gbs"] for g in gathered),
                          "sum_concurrent_numa_local_gbs": sum(g["numa_local_concurrent_gbs"] for g in gathered)}))
        print(topo, file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
