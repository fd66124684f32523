#!/usr/bin/env python
"""C3 through the one-process multi-GPU entry (MultiLayout.run: pageable host buffers in, host results out) on all
visible GPUs; prints one JSON line.   python profiles/run_multi.py [--controls 16384] [--hops 10000]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--controls", type=int, default=16384)
    ap.add_argument("--hops", type=int, default=10000)
    ap.add_argument("--per-device", action="store_true", help="weak scaling: --controls voltage vectors PER device")
    args = ap.parse_args()
    from kmc_dn_b200 import _lib, workloads
    from kmc_dn_b200.ensemble import MultiLayout
    ndev = _lib.load().kmcb200_device_count()
    w = workloads.c3_voltage_search(n_controls=args.controls * (ndev if args.per_device else 1), seeds=16, hops=args.hops)
    lt = w["tables"]
    lay = MultiLayout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    kw = dict(basis=lt.basis, occupation0=w["occupation0"], seed=1)
    lay.run(1000, w["kT"], w["V"], **kw)  # warm-up (allocations on every device)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        r = lay.run(args.hops, w["kT"], w["V"], **kw)
        dt = time.perf_counter() - t0
        best = max(best, len(w["V"]) * args.hops / dt)
    print(json.dumps({"what": "C3 via MultiLayout.run, one process", "devices": ndev, "members": int(len(w["V"])),
                      "hops": args.hops, "hops_per_s_e2e": best, "finite": bool((r["time"] > 0).all())}), flush=True)
    lay.close()


if __name__ == "__main__":
    main()
