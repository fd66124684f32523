#!/usr/bin/env python
"""Round-2 experiment 1 (round-1 kernels): the C3 headline at SURVEY 8(d)'s run lengths, table-size sensitivity of the
thread-per-trajectory kernel and its state-evaluation rate.  One JSON line per measurement."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from profiles.run_lanes import measure
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout

w = workloads.c3_voltage_search(n_controls=16384, seeds=16)
lt = w["tables"]
lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
for tl in (None, 13, 12, 11, 10, 9):
    if tl is not None:
        os.environ["KMCB200_LTAB_LOG"] = str(tl)
    r = measure(lay, lt, w, 100000, 0, "lanes", steps=1)
    r["ltab_log"] = tl; r["what"] = "c3 1M x 1e5"
    print(json.dumps(r), flush=True)
# evaluation (miss) rate on a contiguous block of members
n = 8192
for tl in (14, 12, 11, 10, 9, 8):
    os.environ["KMCB200_LTAB_LOG"] = str(tl)
    r = lay.run(100000, w["kT"][:n], w["V"][:n], basis=lt.basis, occupation0=w["occupation0"], seed=7, want_misses=True, kernel="lanes")
    print(json.dumps({"what": "miss rate, 8192 contiguous members x 1e5", "ltab_log": tl, "miss_rate": float(r["misses"].mean() / 1e5),
                      "p99": float(np.percentile(r["misses"], 99) / 1e5)}), flush=True)
os.environ.pop("KMCB200_LTAB_LOG", None)
w2 = workloads.c3_voltage_search(n_controls=1024, seeds=16)
for k in ("warp", "lanes"):
    r = measure(lay, lt, w2, 1000000, 0, k, steps=1)
    r["what"] = "c3 65536 x 1e6"
    print(json.dumps(r), flush=True)
lay.close()
