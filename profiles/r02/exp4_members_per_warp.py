#!/usr/bin/env python
"""Round-2 experiment 4: members per warp (32 / 16 / 8) of the thread-per-trajectory kernel for ensembles that do not fill the
device's warp slots with full warps, against the warp-per-trajectory kernel.  C3, device-resident inputs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from profiles.run_lanes import measure
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout

for nc, hops in [(int(a), 100000) for a in sys.argv[1:]] or ((1024, 1000000), (1024, 100000), (512, 100000), (384, 100000), (256, 100000)):
    w = workloads.c3_voltage_search(n_controls=nc, seeds=16)
    lt = w["tables"]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    for k, ml in (("warp", None), ("lanes", "5"), ("lanes", "4"), ("lanes", "3"), ("lanes", None)):
        if ml is None:
            os.environ.pop("KMCB200_LANES_MPB_LOG", None)
        else:
            os.environ["KMCB200_LANES_MPB_LOG"] = ml
        r = measure(lay, lt, w, hops, 0, k, steps=1)
        print(json.dumps({"members": r["members"], "hops": hops, "kernel": k, "mpb_log": ml or "auto", "hops_per_s": r["hops_per_s"],
                          "ms": r["ms_per_step"]}), flush=True)
    lay.close()
