#!/usr/bin/env python
"""Round-2 experiment 2: generation-9 thread-per-trajectory kernel on C3 -- throughput vs table size / slicing, evaluation rate."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from profiles.run_lanes import measure
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout

hops = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = workloads.c3_voltage_search(n_controls=16384, seeds=16)
lt = w["tables"]
lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
for tl, sl in ((None, None), (None, "1"), (11, None), (13, None), (10, None)):
    for k, v in (("KMCB200_LTAB_LOG", tl), ("KMCB200_LANES_SLICES", sl)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)
    r = measure(lay, lt, w, hops, 0, "lanes", steps=1)
    r["ltab_log"] = tl; r["slices"] = sl; r["what"] = f"c3 1M x {hops}"
    print(json.dumps(r), flush=True)
os.environ.pop("KMCB200_LANES_SLICES", None)
n = 8192
for tl in (12, 11, 10, 9, 8):
    os.environ["KMCB200_LTAB_LOG"] = str(tl)
    r = lay.run(hops, w["kT"][:n], w["V"][:n], basis=lt.basis, occupation0=w["occupation0"], seed=7, want_misses=True, kernel="lanes")
    print(json.dumps({"what": f"evaluation rate, 8192 contiguous members x {hops}", "ltab_log": tl, "miss_rate": float(r["misses"].mean() / hops),
                      "p99": float(np.percentile(r["misses"], 99) / hops)}), flush=True)
os.environ.pop("KMCB200_LTAB_LOG", None)
for nc in (1024, 256):
    w2 = workloads.c3_voltage_search(n_controls=nc, seeds=16)
    for k in ("warp", "lanes"):
        r = measure(lay, lt, w2, hops, 0, k, steps=1)
        r["what"] = f"c3 {nc * 64} x {hops}"
        print(json.dumps(r), flush=True)
lay.close()
