#!/usr/bin/env python
"""bench.py -- molecules/sec of the batched graph-convolution hot path on N B200s (one node).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic ring graphs (data_generator/synth_generator_ring.py
scaled to 32 nodes), batch 1024 molecules per GPU, 64-dim features, 2 x GraphConv(64) + sigmoid ->
GraphGather -> Dense(2) -> softmax cross-entropy.  One "step" = one training pass over one batch:
forward + backward + (N>1: one NCCL all-reduce of the flat gradient buffer) + Adam.  Weak scaling:
the per-GPU batch is fixed, the global batch is 1024 * N.

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each field is measured.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"name": "ring_graphs_b1024_n32_f64_2xGraphConv64", "batch_per_gpu": 1024, "n_nodes": 32, "feature_dim": 64,
            "conv_dims": [64, 64], "channels": 1, "label_dim": 2, "act": "sigmoid"}
N_ROT = 24          # distinct resident batches rotated through the timed loop (24 x 8.9 MB > 126 MB L2)
METRIC, UNIT = "molecules/sec", "molecules/s"
V4_TRAFFIC_BYTES = 9458176.0   # profiles/r01b_v4_B1024.txt: 9.46 MB read, writes still in the 126 MB L2 at kernel end


def make_host_batches(n, seed, B=None):
    from kgcn_b200 import synth
    w = WORKLOAD
    B = B or w["batch_per_gpu"]
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        out.append(synth.ring_graphs(rng, B, w["n_nodes"], w["feature_dim"]))
    return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag = index, period, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
            for k, bit in names.items():
                if mask & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        self.sample()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's per-molecule path (oracle/graphconv_ref.c)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, max_seconds=None):
    from oracle import cref
    w = WORKLOAD
    B = w["batch_per_gpu"]
    host = make_host_batches(min(4, max(1, steps)), seed=1234)
    net = cref.RefNet(w["feature_dim"], w["conv_dims"], w["channels"], w["label_dim"], act=2)
    rng = np.random.default_rng(1234)
    for name, (off, shape) in net.offsets.items():
        if name.endswith("kernel"):
            lim = np.sqrt(6.0 / (shape[-2] + shape[-1]))
            net.view(net.params, name)[...] = rng.uniform(-lim, lim, size=shape)
    mask = np.ones(B, np.float32)
    # all the host cores this process may use -- stated explicitly, because torchrun exports OMP_NUM_THREADS=1 to its
    # workers and the OpenMP default would then time the CPU arm on a single thread
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def one(i):
        d = host[i % len(host)]
        return net.train_step(d["counts"], d["indices"], d["values"], d["features"], d["labels"], mask, w["n_nodes"],
                              n_threads=threads)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        one(i)
        done += 1
        if max_seconds is not None and time.perf_counter() - t0 > max_seconds and done >= 3:
            break
    dt = time.perf_counter() - t0
    return {"value": B * done / dt, "steps": done, "seconds": dt, "cores": threads, "ms_per_step": dt / done * 1e3}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    w = WORKLOAD
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "step": "train: fwd+bwd+Adam", "batch_per_step": w["batch_per_gpu"],
                   "note": "CPU port of the reference's per-molecule GraphConv path (oracle/graphconv_ref.c, OpenMP over "
                           "molecules); TensorFlow itself is not installable here (BASELINE.md section 2)"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": "%d full steps of %d molecules" % (r["steps"], w["batch_per_gpu"])},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    from kgcn_b200 import _lib, ops
    from kgcn_b200.trainer import DeviceBatch, HostFedPipeline, NetSpec, Trainer

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kgcn_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    B, N, F = w["batch_per_gpu"], w["n_nodes"], w["feature_dim"]
    spec = NetSpec(F, w["conv_dims"], N, channels=w["channels"], label_dim=w["label_dim"], act=w["act"])
    tr = Trainer(spec, B, device=dev, lr=0.01, world_size=world, seed=1234)
    host = make_host_batches(N_ROT, seed=1234 + rank)
    batches = [DeviceBatch.from_host(d["counts"], d["indices"], d["values"], d["features"], d["labels"], N, device=dev)
               for d in host]
    nnz_mean = float(np.mean([b.csr.nnz for b in batches]))

    def note(msg):
        if args.verbose:
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    # ---- launches of this library's kernels per step (counted on one eager step) ----
    c0 = _lib.lib.kgcn_launch_count()
    tr.step_eager(batches[0])
    launches_per_step = _lib.lib.kgcn_launch_count() - c0
    c0 = _lib.lib.kgcn_launch_count()
    tr.forward_eager(batches[0])
    launches_per_infer = _lib.lib.kgcn_launch_count() - c0
    torch.cuda.synchronize()
    note("eager step ok, %d launches" % launches_per_step)
    # ---- capture one CUDA graph per resident batch (train) + one inference graph per batch ----
    for i in range(N_ROT):
        tr.capture(("train", i), batches[i])
    for i in range(N_ROT):
        tr.capture(("infer", i), batches[i], train=False)
    note("graphs captured")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(kind, steps, warmup):
        for i in range(warmup):
            tr.replay((kind, i % N_ROT))
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            tr.replay((kind, i % N_ROT))
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_train = timed("train", args.steps, args.warmup)
    clocks = sampler.result()
    note("train timed %.3f ms" % ms_train)
    ms_infer = timed("infer", args.steps, args.warmup)
    note("infer timed")
    cost_sum, correct = tr.read_stats()

    # ---- dominant-kernel roofline: the batched SpMM  Y = A.X  on the step's own shape, timed alone ----
    peaks, peak_kind = measured_peaks()
    ys = [torch.empty(B, N, F, device=dev) for _ in range(N_ROT)]
    st = torch.cuda.current_stream()

    def spmm(i):
        b = batches[i % N_ROT]
        ops.bspmm_raw(b.csr, b.features, N * F, 0, ys[i % N_ROT], N * F, 0, F)

    reps = max(1, min(200, args.steps // 4))

    def time_alone(fn):
        """us per launch of fn(i), i over the rotating batches: graph-replayed back-to-back launches, CUDA events."""
        g = torch.cuda.CUDAGraph()
        for i in range(N_ROT):
            fn(i)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for i in range(N_ROT):
                fn(i)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) * 1e3 / (reps * N_ROT)

    spmm_us = time_alone(spmm)
    bytes_spmm = 4 * B * N * F * 2 + 8 * nnz_mean + 4 * w["channels"] * B * (N + 1)
    achieved = bytes_spmm / spmm_us / 1e3
    roofline = {"kernel": "bspmm_tile_kernel (kgcn_bspmm_f32, Y[b]=A[b].X[b], B=%d N=%d F=%d)" % (B, N, F), "bound": "hbm",
                "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "peak_source": peak_kind + " copy bandwidth (burst)", "frac_of_8TBs_spec": achieved / 8000.0,
                "algorithmic_bytes_per_launch": bytes_spmm, "us_per_launch": spmm_us,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE ncu --set full launch of this kernel on this
                # shape (profiles/r01_spmm_tile_c2.txt): the 8.4 MB output was still in the 126 MB L2 at kernel end
                "traffic": 9386000.0,
                "timing": "CUDA events around %d back-to-back launches (graph replay) over %d rotating batches" % (reps * N_ROT, N_ROT)}

    # context for `frac` at this size: a plain device-to-device copy moving about the same bytes (8.4 MB read + 8.4 MB
    # written), timed the same way.  MEASURED_PEAKS' copy figure is for 2 GiB transfers; at 17 MB a launch is a few
    # microseconds long and its fill / drain is a large part of it.
    def copy_same(i):
        ys[i % N_ROT].copy_(batches[i % N_ROT].features)

    copy_us = time_alone(copy_same)
    copy_bytes = 2 * 4 * B * N * F
    roofline["same_size_copy"] = {"bytes": copy_bytes, "us_per_launch": copy_us, "gbs": copy_bytes / copy_us / 1e3,
                                  "spmm_vs_copy": (bytes_spmm / spmm_us) / (copy_bytes / copy_us)}

    # ---- the fused GraphConv layer kernel (x -> act(A.x.W + deg*b)) timed the same way ----
    w0, b0 = tr.views["conv0/kernel"], tr.views["conv0/bias"]

    def layer(i):
        b = batches[i % N_ROT]
        ops.graphconv_fwd(b.csr, b.features, w0, b0, 2, 0, out=ys[i % N_ROT])

    layer_us = time_alone(layer)
    bytes_layer = 4 * B * N * (F + F) + 8 * nnz_mean + 4 * w["channels"] * B * (N + 1) + 4 * F * F + 4 * F
    fused_roofline = {"kernel": "graphconv_fused_v4_kernel (kgcn_graphconv_fwd_f32: TMA ring -> thread-per-row aggregation into TMEM -> "
                                "tcgen05 TS-mode 3xTF32 with the bias in the GEMM -> sigmoid epilogue)",
                      "bound": "hbm", "achieved": bytes_layer / layer_us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                      "frac": bytes_layer / layer_us / 1e3 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": bytes_layer,
                      "us_per_launch": layer_us, "molecules_per_s_per_layer": B / layer_us * 1e6,
                      # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full launch (profiles/r01b_v4_B1024.txt)
                      "traffic": V4_TRAFFIC_BYTES}

    # ---- the fused GraphConv backward (dU, A^T.dU, dx, dW, dbias + the fixed-order partial reduce), the largest
    # share of the step (profiles/r01b_bench_launches.csv), timed the same way on layer-2's shape with a dense dy ----
    dys = [torch.randn(B, N, F, device=dev) for _ in range(4)]
    acts = [torch.rand(B, N, F, device=dev) for _ in range(4)]

    def layer_bwd(i):
        b = batches[i % N_ROT]
        ops.graphconv_bwd(b.csr, b.features, w0, 2, acts[i % 4], dys[i % 4])

    bwd_us = time_alone(layer_bwd)
    bytes_bwd = 4 * B * N * F * 4 + 8 * nnz_mean + 4 * w["channels"] * B * (N + 1) + 2 * (4 * F * F + 4 * F)
    bwd_roofline = {"kernel": "graphconv_fused_bwd_kernel + splitk_reduce_kernel (kgcn_graphconv_bwd_f32: x, y, dy read once, dx "
                              "written once, dW/dbias per-CTA partials reduced in fixed order)",
                    "bound": "hbm", "achieved": bytes_bwd / bwd_us / 1e3, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": bytes_bwd / bwd_us / 1e3 / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": bytes_bwd,
                    "us_per_launch": bwd_us}

    # ---- end to end from pinned host buffers through the public step call ----
    max_nnz = int(max(d["values"].shape[0] for d in host) * 1.1) + 64
    pipe = HostFedPipeline(tr, max_nnz, train=True, depth=2)
    pinned = [pipe.pin_host_batch(d["counts"], d["indices"], d["values"], d["features"], d["labels"]) for d in host]
    pipe.capture()
    note("e2e pipeline captured")
    e2e_steps = max(10, min(args.steps, 300))
    for _ in pipe.run_many(pinned[i % N_ROT] for i in range(6)):
        pass
    barrier()
    t0 = time.perf_counter()
    e2e_last = None
    for e2e_last in pipe.run_many(pinned[i % N_ROT] for i in range(e2e_steps)):   # every step's stats are read on the host
        pass
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h2d = int(np.mean([pipe.h2d_bytes(p) for p in pinned]))
    e2e = {"value": B * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
           "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
           "path": "pinned host COO+labels (1 packed copy) + features -> H2D (copy stream, 2 slots) -> device CSR pack -> "
                   "train step (CUDA graph) -> D2H cost_sum/correct_count, every step"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(10 ** 6, 2, max_seconds=12.0)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                        "sample": "%d training steps of %d molecules (%.1f s) of the same workload, oracle/graphconv_ref.c"
                                  % (r["steps"], B, r["seconds"])}

    if rank == 0:
        mols = B * world
        line = {
            "metric": METRIC, "value": mols * args.steps / (ms_train * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_train / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "step": "train: fwd+bwd+%sAdam" % ("NCCL grad all-reduce+" if world > 1 else ""),
                       "batch_per_gpu": B, "global_batch": mols, "n_nodes": N, "feature_dim": F, "conv_dims": w["conv_dims"],
                       "nnz_per_graph": nnz_mean / B, "parallelism": "dp%d" % world,
                       "l2": "rotating %d resident batches (%.0f MB of inputs > 126 MB L2)" % (N_ROT, N_ROT * (B * N * F * 4 + 12 * nnz_mean) / 1e6),
                       "launch": "one CUDA graph replay per step"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step), "roofline": roofline, "roofline_fused_layer": fused_roofline, "roofline_fused_bwd": bwd_roofline,
            "cpu_baseline": cpu_baseline,
            "infer": {"value": mols * args.steps / (ms_infer * 1e-3), "unit": UNIT, "ms_per_step": ms_infer / args.steps,
                      "launches_per_step": int(launches_per_infer), "step": "forward only (layers + readout)"},
            "last_step": {"cost_sum": cost_sum, "correct_count": correct},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
