#!/usr/bin/env python
"""Measures every BASELINE.json configuration (C1..C5) on one GPU next to the CPU ports on the
box's host cores, and writes profiles/configs_<tag>.json + .md (the table BASELINE.md section 3 asks for).

    python profiles/run_configs.py --tag r01 [--quick]

GPU numbers are end-to-end through the public host API (Layout.run: H2D + kernel + D2H); the CPU numbers are a
bounded strided sample of the same ensemble on all host cores (C restatements of the reference's loops).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def gpu_rate(w, scale=1.0, reps=2):
    from kmc_dn_b200.ensemble import Layout
    lt = w["tables"]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R,
                 prune_threshold=w.get("prune", 0.0))
    hops = max(1, int(w["hops"] * scale)); pre = int(w["prehops"] * scale)
    kw = dict(basis=lt.basis, prehops=pre, occupation0=w["occupation0"], seed=1)
    lay.run(min(hops, 1000), w["kT"], w["V"], **kw)  # warm-up
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        r = lay.run(hops, w["kT"], w["V"], **kw)
        dt = time.perf_counter() - t0
        best = max(best, len(w["V"]) * (hops + pre) / dt)
    idx = np.linspace(0, len(w["V"]) - 1, min(2048, len(w["V"]))).astype(np.int64)
    s = lay.run(hops, w["kT"][idx], w["V"][idx], want_misses=True, **kw)
    miss = float(s["misses"].mean() / (hops + pre))
    lay.close()
    finite = float(np.isfinite(r["time"]).mean())
    return best, miss, finite, hops, pre


def cpu_rate(w, seconds, semantics="go", use_cache=True, scale=1.0):
    from oracle import oracle
    lt = w["tables"]
    n = os.cpu_count() or 1
    hops = max(1, int((w["hops"] + w["prehops"]) * scale))

    def run(B, h):
        idx = np.linspace(0, len(w["V"]) - 1, B).astype(np.int64)
        V = w["V"][idx]; E = lt.E_constant(V); kT = w["kT"][idx]
        t0 = time.perf_counter()
        if semantics == "go":
            oracle.go_ensemble(lt.N, lt.P, lt.nu, kT, lt.I_0, lt.R, lt.distances, E, lt.transitions_constant, V, h,
                               variant=1, use_cache=use_cache, occupation0=w["occupation0"], seed0=1)
        else:
            oracle.py_ensemble(lt.N, lt.P, lt.nu, kT, lt.I_0, lt.R, lt.distances, E, lt.transitions_constant, V, h,
                               occupation0=w["occupation0"], seed0=1)
        return B * h / (time.perf_counter() - t0)

    if len(w["V"]) == 1:  # single-trajectory latency: one member on one core
        return run(1, hops), 1, hops
    r0 = run(2 * n, max(1, min(hops, 500)))
    B = int(max(n, min(len(w["V"]), r0 * seconds // hops // n * n)))
    if B * hops > r0 * seconds * 4:  # a single member is already too long: shorten the trajectories instead
        hops = max(1, int(r0 * seconds / B))
    return run(B, hops), B, hops


def cpu_rate_pruned(w, seconds, use_cache):
    """The pruned transition list (simulation.go:200-215, wrapperSimulatePruned) exists only in the single-run port:
    one trajectory per host thread (ctypes releases the GIL), a strided handful of members."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    lt = w["tables"]
    n = os.cpu_count() or 1
    idx = np.linspace(0, len(w["V"]) - 1, n).astype(np.int64)
    E = lt.E_constant(w["V"][idx])

    def one(k, h):
        se = np.zeros(lt.N + lt.P); se[lt.N:] = w["V"][idx[k]]
        oracle.go_simulate(lt.N, lt.P, lt.nu, float(w["kT"][idx[k]]), lt.I_0, lt.R, lt.distances, E[k], lt.transitions_constant,
                           se, h, variant=1, occupation=w["occupation0"], use_cache=use_cache, cut=w["prune"], seed=k + 1)

    def run(h):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=n) as ex:
            list(ex.map(lambda k: one(k, h), range(n)))
        return n * h / (time.perf_counter() - t0)

    r0 = run(200)
    hops = int(max(200, min(w["hops"], r0 * seconds / n)))
    return run(hops), n, hops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--only", default="", help="comma-separated config names (default: all)")
    ap.add_argument("--no-cpu", action="store_true", help="GPU numbers only")
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "profiles"))
    args = ap.parse_args()
    from kmc_dn_b200 import workloads
    q = args.quick
    configs = [
        ("C1-single", workloads.c1_basic(B=1), 1.0),   # single-trajectory latency
        ("C1", workloads.c1_basic(B=4096), 1.0),
        ("C2", workloads.c2_grid4x4(seeds=64), 1.0),
        ("C3", workloads.c3_voltage_search(n_controls=1024 if q else 16384, seeds=16, hops=100000), 1.0),  # 1e5 hops
        ("C3-4runs", workloads.c3_voltage_search(n_controls=1, seeds=1, hops=1000000), 1.0),  # latency-bound: 4 trajectories
        ("C3-1e6", workloads.c3_voltage_search(n_controls=64 if q else 1024, seeds=16, hops=1000000), 0.01 if q else 1.0),
        ("C4", workloads.c4_temperature(n_T=64, seeds=64 if q else 1024), 0.1 if q else 1.0),
        ("C5", workloads.c5_scaling(N=256, M=25, B=1024 if q else 8192), 1.0),
        # C5 with the pruned transition list of validate_tests.py:323 (pairs with tc <= 1e-7 max tc dropped)
        ("C5-pruned", dict(workloads.c5_scaling(N=256, M=25, B=1024 if q else 8192), prune=1e-7,
                           name="C5 scaling N=256 P=8, prune_threshold 1e-7"), 1.0),
    ]
    rows = []
    only = [x for x in args.only.split(",") if x]
    for name, w, scale in configs:
        if only and name not in only:
            continue
        g, miss, finite, hops, pre = gpu_rate(w, scale)
        if args.no_cpu:
            print(json.dumps(dict(config=name, workload=w["name"], members=int(len(w["V"])), hops=hops, prehops=pre,
                                  gpu_hops_per_s=g, state_cache_miss_rate=miss, finite_fraction=finite)), flush=True)
            continue
        if w.get("prune"):
            c_cache, Bc, hc = cpu_rate_pruned(w, args.cpu_seconds, True)
            c_nocache, _, _ = cpu_rate_pruned(w, args.cpu_seconds / 2, False)
            c_py = float("nan")  # (the numba loop has no pruned list)
        else:
            c_cache, Bc, hc = cpu_rate(w, args.cpu_seconds, "go", True, scale)
            c_nocache, _, _ = cpu_rate(w, args.cpu_seconds / 2, "go", False, scale)
            c_py, _, _ = cpu_rate(w, args.cpu_seconds / 2, "py", True, scale)
        row = dict(config=name, workload=w["name"], members=int(len(w["V"])), hops=hops, prehops=pre,
                   gpu_hops_per_s=g, state_cache_miss_rate=miss, finite_fraction=finite,
                   cpu_cores=os.cpu_count(), cpu_go_port_cached=c_cache, cpu_go_port_uncached=c_nocache,
                   cpu_numba_port=c_py, cpu_sample=f"{Bc} members x {hc} hops",
                   speedup_vs_cached=g / c_cache, speedup_vs_uncached=g / c_nocache, speedup_vs_numba_port=g / c_py)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.no_cpu:
        return
    out = os.path.join(args.out_dir, f"configs_{args.tag}")
    json.dump(rows, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("| config | members x hops | B200 x1 hops/s (e2e) | miss rate | CPU Go-port +cache | Go-port no cache | numba-port | "
                "x cached | x uncached | x numba |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            mr = "n/a" if r["state_cache_miss_rate"] is None else f"{r['state_cache_miss_rate']:.3f}"
            f.write(f"| {r['workload']} | {r['members']} x {r['prehops']}+{r['hops']} | {r['gpu_hops_per_s']:.3g} | {mr} | "
                    f"{r['cpu_go_port_cached']:.3g} | {r['cpu_go_port_uncached']:.3g} | {r['cpu_numba_port']:.3g} | "
                    f"{r['speedup_vs_cached']:.0f} | {r['speedup_vs_uncached']:.0f} | {r['speedup_vs_numba_port']:.0f} |\n")
        f.write(f"\nCPU: {os.cpu_count()} host cores of the GPU box, one trajectory per thread, bounded strided sample.\n")


if __name__ == "__main__":
    main()
