#!/usr/bin/env python
"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, mean / min us, share of the total.
usage: launch_summary.py launches.csv"""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("unnamed>::", "")
    key = "%s grid=%s block=%s" % (name, r[8], r[7])
    agg.setdefault(key, []).append(float(r[14]) / 1e3)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-78s n=%3d mean %8.2f us  min %8.2f us  share %5.1f%%" % (k[:78], len(v), sum(v) / len(v), min(v), 100 * sum(v) / tot))
print("total %.1f us over %d launches" % (tot, len(rows)))
