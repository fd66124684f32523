#!/usr/bin/env python
"""bench.py -- aggregate KMC hops/s of the B200 hop loop on BASELINE.json's voltage-search ensemble (C3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hops H]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hop loop over the whole ensemble (n_controls x 4 inputs x seeds members, each `hops`
hops; SURVEY.md 8(d): 1 048 576 members x 1e5 hops).  Ensemble members are sharded over ranks in contiguous blocks
with global Philox stream numbering (weak scaling: every GPU gets a full C3 ensemble); the only collective is the
final NCCL all_gather of the time / electrode tallies (int32 on the wire), identical in the `value` and `e2e` legs.
A second timed leg (`long_run` in the JSON line) runs the reference's real run length -- 1e6 hops
(voltage_search.py:49,53) -- on a 65 536-member subset, and `e2e_parallelSimulations` times the reference's own
batched plugin call (GoSlice export, simulationWrapper.go:274-316; N = 1 only).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "aggregate KMC hops/s over ensemble"
UNIT = "hops/s"
DTYPE = ("f32 rates (correctly rounded f32 quotient dE/kT, two-float exponent product, ex2.approx) and per-acceptor sums; f64 site energies and total rate; event thresholds 2^-20 fixed point + "
         "exact fp64 tail; dwell = lg2.approx(f32 uniform) x f32(-ln2/total), f32 partial sums per 64 hops, f64 elapsed time")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hops", type=int, default=None, help="hops per member per step (default: SURVEY 8d -- 1e5 for c3, 1e4 for c5)")
    ap.add_argument("--controls", type=int, default=16384, help="control-voltage vectors (x4 inputs x seeds members)")
    ap.add_argument("--seeds", type=int, default=16)
    ap.add_argument("--long-hops", type=int, default=1000000, help="hops per member of the long-run leg (voltage_search.py:49)")
    ap.add_argument("--long-controls", type=int, default=1024, help="control vectors of the long-run leg (x4 inputs x seeds members)")
    ap.add_argument("--no-long", action="store_true", help="skip the long-run leg")
    ap.add_argument("--ps-sims", type=int, default=4096, help="simulations of the parallelSimulations e2e leg (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 = BASELINE.json's metric config (default); c5 = examples/scaling.py, 8192 members of a 256-acceptor "
                         "layout sharded over the GPUs (strong scaling), 1e4 hops")
    a = ap.parse_args()
    if a.hops is None:
        a.hops = 10000 if a.workload == "c5" else 100000  # examples/scaling.py:16 / SURVEY 8(d) C3
    return a


def workload(args, long_run=False):
    from kmc_dn_b200 import workloads
    if args.workload == "c5":
        w = workloads.c5_scaling()
        w["hops"] = args.hops
        return w
    if long_run:
        return workloads.c3_voltage_search(n_controls=args.long_controls, seeds=args.seeds, hops=args.long_hops)
    return workloads.c3_voltage_search(n_controls=args.controls, seeds=args.seeds, hops=args.hops)


def config_of(args, w, n_gpus):
    lt = w["tables"]
    common = {"hops_per_member": int(args.hops), "N": int(lt.N), "P": int(lt.P),
              "l2": "flushed between timed steps (256 MiB write)", "rng": "Philox4x32-10 (seed, global member index)"}
    if args.workload == "c5":
        return {"workload": "C5 examples/scaling.py: uniform-random layout, 256 acceptors, 25 donors, 8 electrodes, "
                            f"{len(w['V'])} members with U(-150,150) voltages in total, {args.hops} hops per member per step",
                "members_total": int(len(w["V"])),
                "parallelism": f"ensemble-sharded x{n_gpus}, strong scaling (no data-path collective; final NCCL all_gather "
                               "of tallies)", **common}
    return {"workload": "C3 boolean_logic/voltage_search: reference layout 0 (30 acceptors, 3 donors), 8 electrodes, "
                        f"{args.controls} control vectors x 4 inputs x {args.seeds} seeds = {len(w['V'])} members per GPU, "
                        f"{args.hops} hops per member per step",
            "members_per_gpu": int(len(w["V"])),
            "parallelism": f"ensemble-sharded x{n_gpus} (no data-path collective; final NCCL all_gather of time f64 + "
                           "tallies int32)", **common}


# ---------------------------------------------------------------------------------------- CPU arms
def cpu_sample(w, seconds, variant=1, use_cache=True, semantics="go", nthreads=0):
    """Times the oracle on a bounded sample of the same workload on all host cores."""
    from oracle import oracle
    lt = w["tables"]
    cores = os.cpu_count() or 1
    nthreads = nthreads or cores

    def run(B, hops):
        V = w["V"][:: max(1, len(w["V"]) // B)][:B]
        E = lt.E_constant(V)
        t0 = time.perf_counter()
        if semantics == "go":
            r = oracle.go_ensemble(lt.N, lt.P, lt.nu, 1.0, lt.I_0, lt.R, lt.distances, E, lt.transitions_constant, V, hops,
                                   variant=variant, use_cache=use_cache, occupation0=w["occupation0"], seed0=1,
                                   nthreads=nthreads)
        else:
            r = oracle.py_ensemble(lt.N, lt.P, lt.nu, 1.0, lt.I_0, lt.R, lt.distances, E, lt.transitions_constant, V, hops,
                                   occupation0=w["occupation0"], seed0=1, nthreads=nthreads)
        dt = time.perf_counter() - t0
        return B * hops / dt, dt, r["threads"]

    hops = w["hops"]
    B0 = 2 * nthreads
    rate, dt, used = run(B0, min(hops, 2000))           # calibration
    B = int(max(B0, min(len(w["V"]), (rate * seconds) // hops // nthreads * nthreads)))
    rate, dt, used = run(B, hops)
    if dt < 0.6 * seconds and B < len(w["V"]):  # the state cache warms up: the calibration under-estimated the rate
        B = int(max(B, min(len(w["V"]), (rate * seconds) // hops // nthreads * nthreads)))
        rate, dt, used = run(B, hops)
    return dict(value=rate, seconds=dt, cores=used, members=B, hops=hops)


def _numba_worker(job):
    """One process = one core: the UNMODIFIED reference loop (kmc_dopant_networks.py:33-135) on a few members."""
    from oracle import numba_ref
    ref, seed_fn = numba_ref.load()
    lt_args, members, hops = job
    N, P, nu, kT, I_0, R, distances, tc = lt_args
    S = N + P

    def call(occ0, E, V, h):
        se = np.zeros(S); se[N:] = V
        return ref._simulate_discrete_record(N, P, nu, kT, I_0, R, 0.0, occ0.copy(), distances, E, se, tc, np.zeros((S, S)),
                                             np.zeros(S * S), np.zeros(P), h, False)
    call(*members[0], 50)  # JIT compilation: not timed
    done, t_run = 0, 0.0
    for occ0, E, V in members:
        t0 = time.perf_counter()
        call(occ0, E, V, hops)
        t_run += time.perf_counter() - t0
        done += hops
    return done, t_run


def cpu_numba_unmodified(w, seconds):
    """SURVEY 8(d) CPU baseline (1): the reference's own numba loop, one process per core via multiprocessing, JIT warm-up
    excluded.  Needs the reference tree (KMC_REFERENCE_ROOT or /root/reference): present in the build container, absent on
    the GPU boxes."""
    from oracle import numba_ref
    if not numba_ref.available():
        return {"unavailable": "reference tree absent on this box (KMC_REFERENCE_ROOT unset, /root/reference missing); measured "
                               "in the build container: profiles/r02/numba_unmodified_container.json"}
    try:
        import numba  # noqa: F401
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"numba not importable: {e}"}
    import multiprocessing as mp
    lt = w["tables"]
    cores = os.cpu_count() or 1
    hops = int(min(w["hops"], 20000))  # ~1.4e-5 s per hop at S = 38: a bounded sample
    per_core = max(1, int(seconds * 7e4 / hops))
    idx = np.linspace(0, len(w["V"]) - 1, cores * per_core).astype(np.int64)
    occ0 = np.asarray(w["occupation0"] if w["occupation0"] is not None else np.zeros(lt.N), dtype=bool)
    lt_args = (lt.N, lt.P, lt.nu, 1.0, lt.I_0, lt.R, np.ascontiguousarray(lt.distances), np.ascontiguousarray(lt.transitions_constant))
    jobs = []
    for c in range(cores):
        mem = [(occ0, np.ascontiguousarray(lt.E_constant(w["V"][i])), np.ascontiguousarray(w["V"][i])) for i in idx[c::cores]]
        jobs.append((lt_args, mem, hops))
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_numba_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return {"value": total / busy, "unit": UNIT, "cores": cores, "members": int(len(idx)), "hops": hops,
            "seconds_busiest_core": busy, "seconds_wall_incl_jit": wall,
            "what": "UNMODIFIED numba _simulate_discrete_record, one process per core (multiprocessing), JIT excluded"}


def reference_arm(args):
    """--impl reference: the reference's CPU path for this workload (the C restatement of the Go loop the
    reference's batched API runs: parallelSimulations -> simulateRecordPlus with its state cache,
    simulationWrapper.go:274-316) on all host cores; each step is a bounded sample of the workload AT THE SAME
    hops per member (the state cache warms as the reference's does)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    per_step = max(1.0, min(args.cpu_seconds, 100.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(w, per_step / 4)
    vals = []
    t_all = 0.0
    for _ in range(args.steps):
        s = cpu_sample(w, per_step)
        vals.append(s)
        t_all += s["seconds"]
    total_hops = sum(s["members"] * s["hops"] for s in vals)
    value = total_hops / t_all
    s0 = vals[-1]
    sample = f"{s0['members']} members x {s0['hops']} hops per step (members strided out of the full ensemble)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 rates / f64 time", "data": "synthetic",
            "config": config_of(args, w, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": s0["cores"], "kind": "port", "sample": sample,
                             "what": "C restatement of goSimulation simulateRecordPlus + state cache (Go toolchain "
                                     "unavailable), one trajectory per thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.workload == "c3" and not args.no_long:
        wl = workload(args, long_run=True)
        sl = cpu_sample(wl, min(20.0, args.cpu_seconds * 1.3))
        line["long_run"] = {"value": sl["value"], "unit": UNIT, "cores": sl["cores"],
                            "sample": f"{sl['members']} members x {sl['hops']} hops, {sl['seconds']:.1f} s",
                            "config": f"{len(wl['V'])} members x {args.long_hops} hops (voltage_search.py:49,53)"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------- our arm
class Leg:
    """One ensemble resident on this rank's GPU: device inputs for the `value` leg, pinned host buffers for `e2e`."""

    def __init__(self, torch, dist, lay, w, dev, world, member0):
        lt = w["tables"]
        self.torch, self.dist, self.lay, self.w, self.lt, self.world, self.member0 = torch, dist, lay, w, lt, world, member0
        B = self.B = len(w["V"])
        self.hops = int(w["hops"])

        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        self.V_h, self.kT_h = pin(w["V"]), pin(w["kT"])
        occ0 = w["occupation0"] if w["occupation0"] is not None else np.zeros(lt.N, dtype=bool)
        self.occ_h = pin(np.broadcast_to(occ0, (B, lt.N)).astype(np.uint8))
        self.basis_h = pin(lt.basis)
        self.V_d, self.kT_d, self.occ_d, self.basis_d = (x.to(dev) for x in (self.V_h, self.kT_h, self.occ_h, self.basis_h))
        self.time_d = torch.zeros(B, dtype=torch.float64, device=dev)
        self.eo_d = torch.zeros((B, lt.P), dtype=torch.int64, device=dev)
        self.eo32_d = torch.zeros((B, lt.P), dtype=torch.int32, device=dev)
        self.time_out = torch.zeros(B, dtype=torch.float64).pin_memory()
        self.eo_out = torch.zeros((B, lt.P), dtype=torch.int64).pin_memory()
        self.gather_t = [torch.empty_like(self.time_d) for _ in range(world)] if world > 1 else None
        self.gather_e = [torch.empty_like(self.eo32_d) for _ in range(world)] if world > 1 else None
        self.stream = torch.cuda.current_stream()
        self.h2d = self.V_h.numel() * 8 + self.kT_h.numel() * 8 + self.occ_h.numel() + self.basis_h.numel() * 8
        self.d2h = self.time_out.numel() * 8 + self.eo_out.numel() * 8

    def gather(self):
        """The path's only exchange: every rank ends up with all members' elapsed time (f64) and electrode tallies (int32)."""
        if self.world > 1:
            self.eo32_d.copy_(self.eo_d)
            self.dist.all_gather(self.gather_t, self.time_d)
            self.dist.all_gather(self.gather_e, self.eo32_d)

    def step_device(self, i):
        self.lay.run_device(self.B, self.hops, self.kT_d, self.V_d, self.time_d, self.eo_d, basis=self.basis_d,
                            occupation0=self.occ_d, seed=1000 + i, member_index0=self.member0, cuda_stream=self.stream.cuda_stream)
        self.gather()

    def step_e2e(self, i):
        """The public host API: pinned host buffers in, host results out (H2D + D2H inside the call); for N > 1 the results
        this call produced go back to the device and through the same collective as in step_device."""
        lay_run_host(self.lay, self.B, self.hops, self.kT_h.numpy(), self.V_h.numpy(), self.basis_h.numpy(), self.occ_h.numpy(),
                     self.time_out, self.eo_out, 2000 + i, self.member0, self.stream.cuda_stream)
        if self.world > 1:
            self.time_d.copy_(self.time_out, non_blocking=True)
            self.eo_d.copy_(self.eo_out, non_blocking=True)
            self.gather()


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from kmc_dn_b200.ensemble import Layout, launch_count, last_kernel

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = workload(args)
    lt = w["tables"]
    strong = args.workload == "c5"
    if strong:  # a fixed ensemble, contiguous blocks of members per rank (kmc_dn_b200/sharding.py)
        from kmc_dn_b200.sharding import shard_bounds
        B_total = len(w["V"])
        lo, hi = shard_bounds(B_total, world, rank)
        w["V"], w["kT"] = w["V"][lo:hi], w["kT"][lo:hi]
        member0 = lo
    B = len(w["V"])
    if strong:
        assert B * world == B_total, "strong-scaling workload: the member count must divide by the number of GPUs"
    else:
        member0 = rank * B  # global numbering: every rank holds a full C3 ensemble with its own Philox streams
    hops = args.hops
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R, device=local)
    leg = Leg(torch, dist, lay, w, dev, world, member0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = leg.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device(lg, steps, warmup, seed0=100):
        """CUDA events on the launching stream, L2 flushed between steps; returns ms over all steps."""
        for i in range(warmup):
            lg.step_device(i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            flush.fill_(i & 0xff)
            ev[i][0].record(stream)
            lg.step_device(seed0 + i)
            ev[i][1].record(stream)
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev)

    def timed_e2e(lg, steps, warmup):
        for i in range(warmup):
            lg.step_e2e(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            lg.step_e2e(10 + i)
        barrier()
        return (time.perf_counter() - t0) * 1e3

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    # ---- value: inputs resident in HBM
    sampler = ClockSampler(local); sampler.start()
    l0 = launch_count()
    t_wall0 = time.perf_counter()
    ms = timed_device(leg, args.steps, args.warmup)
    t_wall = time.perf_counter() - t_wall0
    launches = launch_count() - l0
    kernel = last_kernel().replace("kmc_", "").replace("_kernel", "")  # lanes | memo | wide
    clocks = sampler.stop()
    tsum = float(leg.time_d.sum().item())
    assert np.isfinite(tsum) and tsum > 0, "kernel produced no valid times"
    launches_timed = launches * args.steps // max(1, args.steps + args.warmup)

    # ---- e2e: the public host API (pinned host buffers in, host results out), copies inside the timed region
    e2e_ms = timed_e2e(leg, args.steps, min(2, args.warmup))
    ms, e2e_ms = max_over_ranks(ms, e2e_ms)
    hops_per_step_all = (float(B_total) if strong else float(B) * world) * hops
    value = hops_per_step_all * args.steps / (ms * 1e-3)
    e2e_value = hops_per_step_all * args.steps / (e2e_ms * 1e-3)

    # ---- long-run leg: the reference's real run length on a subset of the ensemble
    long_run = None
    if not strong and not args.no_long:
        wl = workload(args, long_run=True)
        lgl = Leg(torch, dist, lay, wl, dev, world, rank * len(wl["V"]))
        ksteps = max(1, min(2, args.steps))
        msl = timed_device(lgl, ksteps, 1, seed0=300)
        kernel_long = last_kernel()
        e2l = timed_e2e(lgl, ksteps, 1)
        msl, e2l = max_over_ranks(msl, e2l)
        hl = float(lgl.B) * world * lgl.hops
        long_run = {"value": hl * ksteps / (msl * 1e-3), "unit": UNIT, "ms_per_step": msl / ksteps, "steps": ksteps, "warmup": 1,
                    "e2e": {"value": hl * ksteps / (e2l * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(lgl.h2d),
                            "d2h_bytes_per_step": int(lgl.d2h), "ms_per_step": e2l / ksteps},
                    "kernel": kernel_long,
                    "config": {"members_per_gpu": int(lgl.B), "hops_per_member": int(lgl.hops),
                               "workload": f"C3 subset: {args.long_controls} control vectors x 4 inputs x {args.seeds} seeds, the "
                                           "reference's run length (voltage_search.py:49,53)"}}
        del lgl

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": config_of(args, w, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(leg.h2d), "d2h_bytes_per_step": int(leg.d2h),
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches_timed), "clocks": clocks, "wall_s_timed_region": t_wall}
        if long_run:
            line["long_run"] = long_run
        if world == 1 and args.ps_sims > 0 and not strong:
            line["e2e_parallelSimulations"] = parallel_simulations_leg(args, w, lt)
        stats = sample_statistics(lay, w, lt, member0, kernel)
        line["roofline"] = roofline(value / world, ms / args.steps, B, hops, lt, stats, lay, local, kernel)
        if world == 1 and not args.no_cpu_baseline:
            s = cpu_sample(w, args.cpu_seconds)
            s_nc = cpu_sample(w, args.cpu_seconds / 3, use_cache=False)
            s_py = cpu_sample(w, args.cpu_seconds / 3, semantics="py")
            other = {"go_loop_without_state_cache": s_nc["value"], "numba_loop_fp64_port": s_py["value"],
                     "note": "same port, cache off (wrapperSimulate) / numba semantics (python_simulation); hops/s on the same cores",
                     "numba_unmodified": cpu_numba_unmodified(w, args.cpu_seconds / 3)}
            if long_run:
                sl = cpu_sample(workload(args, long_run=True), args.cpu_seconds)
                other["long_run"] = {"value": sl["value"], "sample": f"{sl['members']} members x {sl['hops']} hops, {sl['seconds']:.1f} s",
                                     "what": "same port + state cache at the long-run leg's 1e6 hops per member"}
            line["cpu_baseline"] = {"value": s["value"], "unit": UNIT, "cores": s["cores"], "kind": "port",
                                    "sample": f"{s['members']} members x {s['hops']} hops (strided subset of the same "
                                              f"ensemble), {s['seconds']:.1f} s",
                                    "what": "C restatement of simulateRecordPlus + state cache (what parallelSimulations runs)",
                                    "other": other}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def lay_run_host(lay, B, hops, kT, V, basis, occ0, time_out, eo_out, seed, member0, cuda_stream):
    """Host-pointer call of the C ABI with caller-owned pinned buffers (what kmc_dn_b200.ensemble.Layout.run does,
    minus the per-call numpy allocations)."""
    import ctypes as C
    from kmc_dn_b200 import _lib
    a = _lib.EnsembleArgs()
    a.B, a.hops, a.prehops, a.mode, a.flags = B, hops, 0, 0, 0
    a.basis, a.electrode_v, a.kT, a.occupation0 = basis.ctypes.data, V.ctypes.data, kT.ctypes.data, occ0.ctypes.data
    a.seed, a.member_index0 = seed, member0
    a.time, a.electrode_occ = time_out.data_ptr(), eo_out.data_ptr()
    a.stream = cuda_stream
    if lay.lib.kmcb200_run_ensemble(lay._h, C.byref(a)):
        raise RuntimeError(_lib.last_error())


def parallel_simulations_leg(args, w, lt):
    """e2e through the reference's own batched plugin call: `parallelSimulations` with 14 GoSlices, every simulation
    carrying its own copy of the layout tables as parrallelSimulationBind.addSimulation builds them
    (goSimulation/parrallelSimulationBind.py:34-66); shape = what voltage_search.parallel_simulation produces (one
    entry per dn x test: distinct voltage vectors, no seed repeats).  Timed around the foreign call only (the slices
    are built once, as the reference's caller would hold them), H2D/D2H inside."""
    from kmc_dn_b200 import _lib
    lib = _lib.load()
    n = int(min(args.ps_sims, len(w["V"])))
    idx = (np.arange(n) * (len(w["V"]) // n)).astype(np.int64)
    N, P = lt.N, lt.P
    V = w["V"][idx]
    E = lt.E_constant(V)
    occ0 = np.asarray(w["occupation0"] if w["occupation0"] is not None else np.zeros(N), dtype=np.float64)

    def f(x):
        return np.ascontiguousarray(x, dtype=np.float64)
    arrs = dict(NSites=f(np.full(n, N)), NElectrodes=f(np.full(n, P)), nu=f(np.full(n, lt.nu)), kT=f(w["kT"][idx]),
                I_0=f(np.full(n, lt.I_0)), R=f(np.full(n, lt.R)), occupation=f(np.tile(occ0, n)),
                distances=f(np.tile(lt.distances.ravel(), n)), E_constant=f(E.ravel()),
                transitions_constant=f(np.tile(lt.transitions_constant.ravel(), n)), electrode_occupation=f(np.zeros(n * P)),
                hops=f(np.full(n, args.hops)), time=f(np.zeros(n)),
                site_energies=f(np.concatenate([np.zeros((n, N)), V], axis=1).ravel()))
    order = ["NSites", "NElectrodes", "nu", "kT", "I_0", "R", "occupation", "distances", "E_constant", "transitions_constant",
             "electrode_occupation", "hops", "time", "site_energies"]
    sl = [_lib.goslice(arrs[k]) for k in order]
    lib.parallelSimulations(*sl)  # warm-up: layout upload + table allocation
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.parallelSimulations(*sl)
    dt = (time.perf_counter() - t0) / reps
    from kmc_dn_b200.ensemble import last_kernel
    assert np.isfinite(arrs["time"]).all() and (arrs["time"] > 0).all()
    inb = sum(arrs[k].nbytes for k in order)
    return {"value": n * args.hops / dt, "unit": UNIT, "ms_per_call": dt * 1e3, "sims": n, "hops_per_sim": int(args.hops),
            "kernel": last_kernel(), "host_bytes_in_slices": int(inb),
            "what": "parallelSimulations (GoSlice ABI of simulationWrapper.go:274-316): per-simulation table copies compared / "
                    "hashed on the host, one ensemble launch per distinct layout, results written back into the slices"}


def sample_statistics(lay, w, lt, member0, kernel="memo", n_sample=4096):
    """One small DBG launch on a strided subset of the ensemble: cache hit rate and mean hole count, from which the
    algorithmic pair count A = n_h*(N-n_h) + N*P of SURVEY.md 8(d) follows.  For the thread-per-trajectory kernel the
    rate of state evaluations is taken from a launch of THAT kernel on a contiguous block of members (its tables are
    shared by the consecutive seeds of a voltage vector, which a strided subset would tear apart)."""
    n_sample = min(n_sample, len(w["V"]))
    idx = np.linspace(0, len(w["V"]) - 1, n_sample).astype(np.int64)
    hs = int(min(w["hops"], 100000))
    r = lay.run(hs, w["kT"][idx], w["V"][idx], basis=lt.basis, occupation0=w["occupation0"], seed=7,
                member_index0=member0, record=True, want_misses=True, kernel="warp" if lt.N <= 31 else None)
    nh = (r["avg_occupation"] / r["time"][:, None]).sum(1)
    A = float(np.mean(nh * (lt.N - nh) + lt.N * lt.P))
    if kernel == "lanes":
        r = lay.run(hs, w["kT"][:n_sample], w["V"][:n_sample], basis=lt.basis, occupation0=w["occupation0"], seed=7,
                    member_index0=member0, want_misses=True, kernel="lanes")
    return {"members": int(n_sample), "hops": hs, "miss_rate": float(r["misses"].mean() / hs), "mean_holes": float(nh.mean()),
            "pairs_per_hop_A": A, "pairs_per_hop_A_nominal": lt.N * (lt.N - 1) + 2 * lt.N * lt.P}


def roofline(hops_per_s_gpu, ms_per_step, B, hops, lt, stats, lay, device, kernel="memo"):
    """What bounds the hop loop (neither HBM nor tensor cores).

    Thread-per-trajectory kernel (kmc_lanes_kernel): a hop whose state is memoised is ONE table lookup -- a 32-byte
    sector at a place that depends on the trajectory's state, i.e. scattered over the threads of a warp.  The hardware
    serves scattered sectors at one L1TEX wavefront per clock and SM, whatever cache level they come from; the roofline is
    therefore `lookups/s`: achieved = hops/s (one algorithmic lookup per hop), peak = scattered 32-byte lookups per second
    measured on this device by kmcb200_measure_peak(what=3) (L2-resident table, 8 independent loads in flight per thread).
    Beside it: the issue-slot figure (warp-instructions per hop from the committed ncu capture x hops/s against the measured
    issue peak; null when this run's configuration is not the profiled one), SURVEY's MUFU.EX2 figure (the work of the
    MISS path: executed_frac = exp actually issued) and the HBM fraction (~0)."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ex2_peak = lay.lib.kmcb200_measure_peak(device, 0)
    issue_peak = lay.lib.kmcb200_measure_peak(device, 2)
    lookup_peak = lay.lib.kmcb200_measure_peak(device, 3)
    A = stats["pairs_per_hop_A"]
    alg = hops_per_s_gpu * A
    executed = hops_per_s_gpu * (stats["miss_rate"] * A)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_member = 8 * lt.P + 8 + lt.N + 8 + 8 * lt.P
    hbm_achieved = B * bytes_per_member / (ms_per_step * 1e-3) / 1e9
    # the committed ncu capture of the kernel that ran: this round's, else round 1's (warp-per-trajectory kernels: unchanged)
    ncu, same_cfg = {}, False
    for rnd in ("r02", "r01"):
        prof = os.path.join(ROOT, "profiles", f"ncu_{rnd}_{kernel}_kernel.json")
        if os.path.exists(prof):
            ncu = json.load(open(prof))
            # round 1's captures carry no configuration: they were taken on this bench's workload at 1e4 hops per member
            same_cfg = ((ncu.get("hops_per_member", 10000) == hops) and ncu.get("N", 30) == lt.N and ncu.get("P", 8) == lt.P
                        and (rnd == "r02" or kernel != "lanes"))
            break
    wih = ncu.get("warp_inst_per_hop") if same_cfg else None
    issue = {"achieved": hops_per_s_gpu * wih / 1e9 if wih else None, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
             "frac": hops_per_s_gpu * wih / issue_peak if wih else None, "warp_inst_per_hop": wih,
             "source": ncu.get("source") if same_cfg else "no ncu capture committed for this configuration",
             "peak_source": "dependent-free FFMA+LOP3 micro-kernel on this device (kmcb200_measure_peak 2); nominal 148 SM x 4/clk"}
    sfu = {"achieved": alg / 1e9, "peak": ex2_peak / 1e9, "unit": "Gexp/s", "frac": alg / ex2_peak, "executed_frac": executed / ex2_peak,
           "note": "SURVEY 8(d): hops/s x A (allowed pairs per hop) against the MUFU.EX2 peak measured on this device.  It is the "
                   "work of the MISS path: memoisation (like the reference's state cache) evaluates miss_rate of the hops, "
                   "executed_frac = exp actually issued"}
    hbm = {"achieved_gbs": hbm_achieved, "peak_gbs": hbm_peak, "frac": hbm_achieved / hbm_peak,
           "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}
    traffic = (B * hops * ncu["dram_bytes_per_hop"] if (same_cfg and "dram_bytes_per_hop" in ncu) else None)
    tnote = ("dram__bytes_read+write of the ncu capture per hop, scaled to this launch.  Algorithmic bytes per member = "
             f"{bytes_per_member} (inputs {8 * lt.P + 8 + lt.N} B, outputs {8 + 8 * lt.P} B) plus, for the memoised path, one 64-byte "
             "entry per distinct state of a run of members (DESIGN.md 3.0); the rest is table lines evicted from L2 and re-read")
    if kernel == "lanes":
        return {"bound": "l1tex-lookups", "kernel": "kmc_lanes_kernel", "achieved": hops_per_s_gpu / 1e9, "peak": lookup_peak / 1e9,
                "unit": "G lookups/s (scattered 32-byte sectors)", "frac": hops_per_s_gpu / lookup_peak,
                "peak_source": "kmcb200_measure_peak(what=3) on this device: independent 256-bit loads at pseudo-random sectors of an "
                               "L2-resident 32 MiB table; nominal 148 SM x 1 wavefront/clk x 1.965 GHz = 2.9e11/s",
                "algorithmic": "one state-table lookup (32 B) per hop: hops/s = lookups/s",
                "issue": issue, "sfu_algorithmic": sfu, "hbm": hbm, "traffic": traffic, "traffic_note": tnote, "sample": stats}
    issue.update({"bound": "issue", "kernel": f"kmc_{kernel}_kernel", "sfu_algorithmic": sfu, "hbm": hbm, "traffic": traffic,
                  "traffic_note": tnote, "sample": stats})
    return issue


if __name__ == "__main__":
    main()
