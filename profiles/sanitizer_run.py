#!/usr/bin/env python
"""Small invocations of every production kernel variant, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python profiles/sanitizer_run.py
    compute-sanitizer --tool racecheck python profiles/sanitizer_run.py

Covers hop_memo.cu (1 / 2 / 3 ranked events per acceptor, cache on / off, record + trace instantiation, second-level
table), hop_wide.cu (1 / 2 / 4 / 8 acceptors per lane, shared-memory and cp.async-ring sweeps, cache on / off) and
hop_lanes.cu (thread per trajectory: 1 / 2 / 3 ranked events, table on / off, runs of identical members, ragged
ensembles, trace instantiation) and its latency kernel for a few trajectories (kmc_solo_kernel).  `--lanes-only` restricts the
run to hop_lanes.cu, `--solo-only` to the latency kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from kmc_dn_b200.ensemble import Layout  # noqa: E402
from tests.util import synthetic_layout  # noqa: E402


def lanes():
    for N, P in [(5, 3), (10, 2), (16, 8), (30, 8), (31, 1), (25, 0)]:
        c = synthetic_layout(N, P, 11 + N, fill=0.6)
        lay = Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"])
        B, hops = 77, 250  # three warps, the last one ragged; runs of 4 identical members
        V = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) // 4)[:, None]
        E = np.tile(c["E_constant"], (B, 1))
        kw = dict(E_constant=E, occupation0=c["occupation"], seed=3, kernel="lanes", want_occupation=True)
        a = lay.run(hops, c["kT"], V, memo=True, **kw)
        b = lay.run(hops, c["kT"], V, memo=False, **kw)
        assert np.array_equal(a["time"], b["time"]) and np.array_equal(a["electrode_occupation"], b["electrode_occupation"])
        r = lay.run(hops, c["kT"], V[:40], E_constant=E[:40], occupation0=c["occupation"], seed=4, prehops=50, trace=True,
                    want_occupation=True, want_site_energies=True, want_misses=True, kernel="lanes")
        assert np.isfinite(r["time"]).any()
        # runs of 32 (the second warp of every block returns at once) and of 16 (every block splits); record outputs
        for rep in (32, 16):
            Vr = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) // rep)[:, None]
            lay.run(hops, c["kT"], Vr, **kw)
        lay.run(hops, c["kT"], V[:40], E_constant=E[:40], occupation0=c["occupation"], seed=5, record=True, kernel="lanes")
        lay.close()
        print(f"lanes N={N} P={P}: ok", flush=True)


def solo():
    """the latency kernel (hop_lanes.cu, kmc_solo_kernel: two warps per trajectory, state graph in shared memory): a table that
    keeps filling up, prehops, more members than CTAs"""
    for N, P in [(5, 3), (10, 2), (16, 8), (30, 8), (31, 1), (25, 0)]:
        c = synthetic_layout(N, P, 11 + N, fill=0.6)
        lay = Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"])
        B, hops = 5, 700
        V = np.tile(c["electrode_v"], (B, 1)) + np.arange(B)[:, None]
        E = np.tile(c["E_constant"], (B, 1))
        kw = dict(E_constant=E, occupation0=c["occupation"], seed=3, want_occupation=True, want_site_energies=True)
        for emax, prehops in ((None, 0), ("5", 100)):
            if emax:
                os.environ["KMCB200_SOLO_EMAX"] = emax
            a = lay.run(hops, c["kT"], V, kernel="solo", prehops=prehops, **kw)
            os.environ.pop("KMCB200_SOLO_EMAX", None)
            b = lay.run(hops, c["kT"], V, kernel="lanes", prehops=prehops, **kw)
            assert np.array_equal(a["time"], b["time"]) and np.array_equal(a["electrode_occupation"], b["electrode_occupation"])
        lay.close()
        print(f"solo N={N} P={P}: ok", flush=True)


def main():
    if "--solo-only" in sys.argv:
        return solo()
    lanes()
    solo()
    if "--lanes-only" in sys.argv:
        return
    cases = [(5, 3), (10, 2), (16, 8), (30, 8), (31, 1), (32, 8), (48, 8), (100, 5), (256, 8)]
    for N, P in cases:
        c = synthetic_layout(N, P, 11 + N, fill=0.85)
        lay = Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"])
        B = 6
        hops = 400 if N <= 64 else 120
        V = np.tile(c["electrode_v"], (B, 1)) + np.arange(B)[:, None]
        E = np.tile(c["E_constant"], (B, 1))
        kw = dict(E_constant=E, occupation0=c["occupation"], seed=3)
        a = lay.run(hops, c["kT"], V, memo=True, **kw)
        b = lay.run(hops, c["kT"], V, memo=False, **kw)
        assert np.array_equal(a["time"], b["time"]) and np.array_equal(a["electrode_occupation"], b["electrode_occupation"])
        r = lay.run(hops, c["kT"], V[:2], E_constant=E[:2], occupation0=c["occupation"], seed=4, prehops=50, record=True,
                    trace=True, want_occupation=True, want_site_energies=True, want_misses=True)
        assert np.isfinite(r["time"]).all()
        lay.close()
        print(f"N={N} P={P}: ok", flush=True)


if __name__ == "__main__":
    main()
