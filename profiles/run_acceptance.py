#!/usr/bin/env python
"""The reference's acceptance run over ALL 400 of its fixtures at their own run lengths, on the GPU:
    python profiles/run_acceptance.py > gpurun_out/acceptance_r01.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kmc_dn_b200.validate import acceptance_over_sets  # noqa: E402

t0 = time.time()
res = acceptance_over_sets(os.path.join(ROOT, "tests", "golden", "fixtures_all.npz"), stride_5m=1)
res["_seconds"] = time.time() - t0
res["_note"] = ("per-electrode Bhattacharyya distance of 5 GPU runs against the fixtures' stored 5-run mean/stddev "
                "(thesis_indrek/validate_tests.py:80-135); D > 0.9 = extreme.  CPU oracle on the same fixtures: "
                "rnd_min_max D_mean 0.204 extreme 0.018; XOR_wide D_mean 0.384 extreme 0.064")
print(json.dumps(res, indent=1))
