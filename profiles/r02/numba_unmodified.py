#!/usr/bin/env python
"""SURVEY 8(d) CPU baseline (1), measured where the reference tree exists (the build container): the UNMODIFIED numba
loop `_simulate_discrete_record` (kmc_dopant_networks.py:33-135), one process per core, JIT excluded, on the C3 ensemble.
    python profiles/r02/numba_unmodified.py > profiles/r02/numba_unmodified_container.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench

if __name__ == "__main__":
    sys.argv = ["bench.py"]
    args = bench.parse()
    w = bench.workload(args)
    r = bench.cpu_numba_unmodified(w, float(os.environ.get("SECONDS_PER_CORE", "10")))
    r["host"] = "build container (no GPU)"
    print(json.dumps(r))
