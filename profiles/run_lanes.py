#!/usr/bin/env python
"""Thread-per-trajectory kernel (csrc/hop_lanes.cu) against the warp-per-trajectory kernel on the bench workload (C3)
and on C4: device-resident inputs, CUDA events, one line of JSON per measurement.

    python profiles/run_lanes.py [--controls 16384] [--hops 10000] [--tlogs 12,13] [--c4] > gpurun_out/lanes.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def measure(lay, lt, w, hops, prehops, kernel, steps=2, seed0=100):
    import torch
    dev = torch.device("cuda", 0)
    B = len(w["V"])
    V = torch.from_numpy(np.ascontiguousarray(w["V"])).to(dev)
    kT = torch.from_numpy(np.ascontiguousarray(w["kT"])).to(dev)
    occ0 = w["occupation0"] if w["occupation0"] is not None else np.zeros(lt.N, dtype=bool)
    occ = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(occ0, (B, lt.N)).astype(np.uint8))).to(dev)
    basis = torch.from_numpy(np.ascontiguousarray(lt.basis)).to(dev)
    t = torch.zeros(B, dtype=torch.float64, device=dev)
    eo = torch.zeros((B, lt.P), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()

    def step(i):
        lay.run_device(B, hops, kT, V, t, eo, basis=basis, occupation0=occ, prehops=prehops, seed=seed0 + i,
                       cuda_stream=st.cuda_stream, kernel=kernel)
    step(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(steps):
        step(1 + i)
    b.record(st)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    cur = (eo.double() / t[:, None]).mean(0).cpu().numpy()
    return {"kernel": kernel, "members": B, "hops": hops, "prehops": prehops, "ms_per_step": ms,
            "hops_per_s": B * (hops + prehops) / (ms * 1e-3), "finite": float(torch.isfinite(t).double().mean()),
            "mean_time": float(t[torch.isfinite(t)].mean()), "mean_current": [float(x) for x in cur]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--controls", type=int, default=16384)
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--hops", type=int, default=10000)
    ap.add_argument("--tlogs", default="")
    ap.add_argument("--c4", action="store_true")
    ap.add_argument("--crossover", default="", help="member counts at which both kernels are timed, e.g. 8192,16384,32768")
    ap.add_argument("--kernels", default="warp,lanes")
    args = ap.parse_args()
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout
    w = workloads.c3_voltage_search(n_controls=args.controls, seeds=args.seeds)
    lt = w["tables"]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    for k in args.kernels.split(","):
        tl = [None] + ([int(x) for x in args.tlogs.split(",")] if (k == "lanes" and args.tlogs) else [])
        for t in tl:
            if t is not None:
                os.environ["KMCB200_LTAB_LOG"] = str(t)
            r = measure(lay, lt, w, args.hops, 0, k)
            r["workload"] = "c3"; r["ltab_log"] = t
            print(json.dumps(r), flush=True)
        os.environ.pop("KMCB200_LTAB_LOG", None)
    for b in [int(x) for x in args.crossover.split(",") if x]:
        wb = workloads.c3_voltage_search(n_controls=max(1, b // (4 * args.seeds)), seeds=args.seeds)
        for k in ("warp", "lanes"):
            r = measure(lay, lt, wb, args.hops, 0, k)
            r["workload"] = "c3-crossover"
            print(json.dumps(r), flush=True)
    lay.close()
    if args.c4:
        w = workloads.c4_temperature()
        lt = w["tables"]
        lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
        for k in args.kernels.split(","):
            r = measure(lay, lt, w, int(w["hops"]), int(w["prehops"]), k, steps=1)
            r["workload"] = "c4"
            print(json.dumps(r), flush=True)
        lay.close()


if __name__ == "__main__":
    main()
