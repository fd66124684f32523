#!/usr/bin/env python
"""Round-2 experiment 4: the thread-per-trajectory kernel on ensembles that do not fill the device's warp slots -- one warp per
block of 32 members (halves=0), blocks split where a run starts at member 16 (halves=1), the library's choice (auto) -- against
the warp-per-trajectory kernel.  C3 (runs of 16 seeds) and C4 (runs of 1024 seeds), device-resident inputs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from profiles.run_lanes import measure
from kmc_dn_b200 import workloads
from kmc_dn_b200.ensemble import Layout

def sweep(w, name, hops):
    lt = w["tables"]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    for k, ml in (("warp", None), ("lanes", "0"), ("lanes", "1"), ("lanes", None)):
        if ml is None:
            os.environ.pop("KMCB200_LANES_HALVES", None)
        else:
            os.environ["KMCB200_LANES_HALVES"] = ml
        r = measure(lay, lt, w, hops, 0, k, steps=1)
        print(json.dumps({"workload": name, "members": r["members"], "hops": hops, "kernel": k, "halves": ml or "auto",
                          "hops_per_s": r["hops_per_s"], "ms": r["ms_per_step"]}), flush=True)
    lay.close()


for nc, hops in [(int(a), 100000) for a in sys.argv[1:]] or ((1024, 1000000), (1024, 100000), (512, 100000), (384, 100000), (256, 100000), (192, 100000)):
    w = workloads.c3_voltage_search(n_controls=nc, seeds=16)
    sweep(w, "C3", hops)

sweep(workloads.c4_temperature(), "C4", 100000)
