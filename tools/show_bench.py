#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (last line starting with '{' in the file)."""
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
d = json.loads(lines[-1])
def show(name, r):
    e2e = (r.get("e2e") or {}).get("value")
    cpu = (r.get("cpu_baseline") or {}).get("value")
    print("%s value %.4g mol/s  %.2f us/step  launches %s  e2e %s  cpu %s  step_frac %.3f  infer %.2f us" % (
        name, r["value"], r["ms_per_step"] * 1e3, r["launches_per_step"], "%.4g" % e2e if e2e else None, "%.4g" % cpu if cpu else None,
        r["roofline_step"]["frac"], r["infer"]["ms_per_step"] * 1e3))
    for k in r["kernels"]:
        print("   %-85s x%d %7.2f us frac %.3f share %.2f" % (k["kernel"][:85], k["launches_per_step"], k["us_per_launch"], k["frac"], k["share_of_step"]))
show("primary n_gpus=%d" % d["n_gpus"], d)
for k, v in d.get("workloads", {}).items():
    show(k, v)
print("last_step", d["last_step"], "dp_check", d.get("dp_check"), "p2p_err", d.get("p2p_error_flag"))
if d.get("roofline_spmm"): print("spmm frac %.3f" % d["roofline_spmm"]["frac"])
