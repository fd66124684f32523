#!/usr/bin/env python
"""Round-2 experiment 5: ensembles of DISTINCT members (one seed per voltage vector: no table sharing) on both kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from profiles.run_lanes import measure
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout

for nc in (1024, 3072, 8192, 65536):
    w = workloads.c3_voltage_search(n_controls=nc, seeds=1)
    lt = w["tables"]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    for k in ("warp", "lanes"):
        r = measure(lay, lt, w, 100000, 0, k, steps=1)
        print(json.dumps({"members": r["members"], "distinct": True, "kernel": k, "hops_per_s": r["hops_per_s"], "ms": r["ms_per_step"]}), flush=True)
    lay.close()
