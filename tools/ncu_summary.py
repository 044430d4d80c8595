#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / bench.py quote.
usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]   (runs `ncu -i` locally, no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("== %s" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[hdr.index("Kernel Name")][:110])
        for k in KEYS:
            if k in hdr:
                print("  %-72s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        t, rd, wr = (float(r[hdr.index(k)].replace(",", "")) if k in hdr else 0.0 for k in KEYS[:3])
        ut, ur, uw = (units[hdr.index(k)] for k in KEYS[:3])
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
        traffic = rd * scale.get(ur, 1) + wr * scale.get(uw, 1)
        print("  %-72s %.3f MB  ->  %.1f GB/s DRAM" % ("dram traffic (read+write)", traffic / 1e6, traffic / (t * tsc.get(ut, 1e-6)) / 1e9))
