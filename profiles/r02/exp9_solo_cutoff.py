#!/usr/bin/env python
"""Round-2 experiment 9: the latency kernel against the warp-per-trajectory kernel around its cut-off (2 members per SM): C3 with
DISTINCT members, 1e5 hops, host arrays through Layout.run."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout, last_kernel

w = workloads.c3_voltage_search(n_controls=1024, seeds=1)
lt = w["tables"]
lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
for B in [int(a) for a in sys.argv[1:]] or (1, 4, 64, 148, 296, 297, 512, 1024, 1184, 2048, 4096):
    for kernel in ("solo", "warp"):
        kw = dict(basis=lt.basis, occupation0=w["occupation0"], seed=3, kernel=kernel)
        lay.run(100000, w["kT"][:B], w["V"][:B], **kw)  # (same run length: the tables of that size class are allocated)
        t0 = time.perf_counter()
        r = lay.run(100000, w["kT"][:B], w["V"][:B], **kw)
        dt = time.perf_counter() - t0
        print(json.dumps({"members": B, "kernel": last_kernel(), "hops_per_s": B * 1e5 / dt, "ms": dt * 1e3,
                          "finite": bool(np.isfinite(r["time"]).all())}), flush=True)
lay.close()
