#!/usr/bin/env python
"""The reference's acceptance run over ALL 400 of its fixtures at their own run lengths, on the GPU:
    python profiles/run_acceptance.py [lanes | warp] [stride_1M stride_5M] > gpurun_out/acceptance_r02.json
("lanes" / "warp": the 5 runs of a fixture on the thread-per-trajectory / warp-per-trajectory kernel instead of the library's
choice for 5 members, which is the latency kernel)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmc_dn_b200.validate import acceptance_over_sets  # noqa: E402

t0 = time.time()
kernel = "lanes" if "lanes" in sys.argv[1:] else ("warp" if "warp" in sys.argv[1:] else None)
strides = [int(a) for a in sys.argv[1:] if a.isdigit()]  # optional: stride over the 1e6-hop sets, stride over the 5e6-hop sets
s1, s5 = (strides + [1, 1])[:2]
res = acceptance_over_sets(os.path.join(ROOT, "tests", "golden", "fixtures_all.npz"), stride_5m=s5, stride_1m=s1, kernel=kernel)
res["_strides"] = {"1e6-hop sets": s1, "5e6-hop sets": s5}
from kmc_dn_b200.ensemble import last_kernel  # noqa: E402
res["_kernel"] = (kernel or "library's choice") + " (" + last_kernel() + ")"
res["_seconds"] = time.time() - t0
res["_note"] = ("per-electrode Bhattacharyya distance of 5 GPU runs against the fixtures' stored 5-run mean/stddev "
                "(thesis_indrek/validate_tests.py:80-135); D > 0.9 = extreme.  CPU oracle on the same fixtures: "
                "rnd_min_max D_mean 0.204 extreme 0.018; XOR_wide D_mean 0.384 extreme 0.064")
print(json.dumps(res, indent=1))
