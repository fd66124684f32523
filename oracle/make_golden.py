"""Generates tests/golden/*.npz in the BUILD CONTAINER (needs /root/reference).

Run:  python oracle/make_golden.py

1. py_replay.npz   -- the UNMODIFIED numba loop (kmc_dopant_networks.py:33-135)
   replayed under numpy's MT19937 stream (numba's generator after
   np.random.seed(s) yields RandomState(s).random_sample(); each hop consumes
   u[2k] for the dwell time and u[2k+1] for the pick, SURVEY.md section 4 probe 5).
   Per case: inputs, the stream seed, the per-hop (from,to) trace obtained by
   calling the reference with hops=1 repeatedly (state arrays are mutated in
   place, RNG state persists), final occupation / electrode tallies / time.
2. fixtures_subset.npz -- a subset of the reference's own 400 `.kmc` golden
   fixtures (thesis_indrek/tests/*): pinned inputs plus the stored 5-run
   mean/stddev currents that validate_tests.py:80-135 accepts against.
3. electrostatics.npz -- electrodes, acceptor/donor positions and the stored
   eV_constant / comp_constant of fixtures (pins the FD-Laplace front end).
5. fixtures_all.npz -- ALL 400 fixtures in compact form (positions, electrode voltages, E_constant, scalars and
   the stored 5-run mean / stddev currents; distances and transitions_constant follow from the positions exactly
   as kmc_dopant_networks.py:657-695, 824-830 builds them, which is asserted here) for the full-size run of the
   reference's acceptance test (thesis_indrek/validate_tests.py:80-135, 299-350).
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import numba_ref  # noqa: E402
from kmc_dn_b200.fixtures import load_kmc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = numba_ref.REFERENCE_ROOT


def synthetic_case(N, P, seed, kT=1.0, I_0=100.0, a=0.25, M=0):
    """Uniform-random 2-D layout with electrodes spread on the boundary; E_constant is a
    synthetic smooth potential (the oracle-vs-reference check does not care where it came from)."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((N, 3)); acc[:, :2] = rng.random((N, 2))
    el = np.zeros((P, 4))
    for p in range(P):
        side = p % 4
        t = (p // 4 + 1) / (P // 4 + 2) if P > 4 else 0.5
        el[p, :2] = [(0.0, t), (1.0, t), (t, 0.0), (t, 1.0)][side]
        el[p, 3] = rng.uniform(-20, 20)
    pos = np.vstack([acc, el[:, :3]])
    S = N + P
    d = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
    R = N ** -0.5
    ab = a * R
    tc = np.exp(-2 * d / ab) - np.eye(S)
    w = np.exp(-4 * ((acc[:, None, :2] - el[None, :, :2]) ** 2).sum(-1))
    E = (w * el[None, :, 3]).sum(1) / np.maximum(w.sum(1), 1e-9)
    if M:
        don = np.zeros((M, 3)); don[:, :2] = rng.random((M, 2))
        E = E + I_0 * R * (1.0 / np.sqrt(((acc[:, None, :] - don[None, :, :]) ** 2).sum(-1))).sum(1)
    occ = np.zeros(N, dtype=bool); occ[rng.permutation(N)[: max(N - max(M, N // 3), 0)]] = True
    return dict(N=N, P=P, nu=1.0, kT=kT, I_0=I_0, R=R, distances=d, transitions_constant=tc,
                E_constant=E, electrode_v=el[:, 3].copy(), occupation=occ)


def fixture_case(path):
    d = load_kmc(path)
    return dict(N=int(d["N"]), P=int(d["P"]), nu=float(d["nu"]), kT=float(d["kT"]), I_0=float(d["I_0"]),
                R=float(d["R"]), distances=d["distances"], transitions_constant=d["transitions_constant"],
                E_constant=d["E_constant"], electrode_v=d["electrodes"][:, 3].copy(),
                occupation=d["occupation"].astype(bool))


def run_reference(ref, seed_fn, c, seed, hops, trace_hops):
    N, P = c["N"], c["P"]; S = N + P
    se = np.zeros(S); se[N:] = c["electrode_v"]
    occ = c["occupation"].copy(); eo = np.zeros(P, dtype=np.int64)
    transitions = np.zeros((S, S)); problist = np.zeros(S * S)
    seed_fn(seed)
    trace = np.zeros((trace_hops, 2), dtype=np.int32)
    time = 0.0
    base = dict(N_acceptors=N, N_electrodes=P, nu=c["nu"], kT=c["kT"], I_0=c["I_0"], R=c["R"], time=0.0,
                distances=c["distances"], E_constant=c["E_constant"], site_energies=se,
                transitions_constant=c["transitions_constant"], transitions=transitions, problist=problist)
    for k in range(trace_hops):
        before_occ = occ.copy(); before_eo = eo.copy()
        t, occ, eo, _, _ = ref._simulate_discrete_record(occupation=occ, electrode_occupation=eo, hops=1,
                                                         record=False, **base)
        time += t
        docc = occ.astype(int) - before_occ.astype(int); deo = eo - before_eo
        frm = np.where(docc < 0)[0]; to = np.where(docc > 0)[0]
        f = int(frm[0]) if frm.size else N + int(np.where(deo < 0)[0][0])
        g = int(to[0]) if to.size else N + int(np.where(deo > 0)[0][0])
        trace[k] = (f, g)
    rest = hops - trace_hops
    traffic = occ_time = None
    if rest > 0:
        t, occ, eo, traffic, occ_time = ref._simulate_discrete_record(occupation=occ, electrode_occupation=eo,
                                                                      hops=rest, record=True, **base)
        time += t
    return dict(trace=trace, time=time, occupation=occ.copy(), electrode_occupation=eo.copy(),
                site_energies=se.copy(), traffic_tail=traffic, occ_time_tail=occ_time, time_tail=t)


def layouts():
    """4. layouts.npz -- dopant layouts the reference's experiments ship as data and BASELINE.json's
    configs name (experiments/boolean_logic/random_layouts/*.npy: 100 layouts of 30 acceptors / 3 donors;
    the first four are kept)."""
    a = np.load(f"{REF}/experiments/boolean_logic/random_layouts/acceptor_layouts.npy")
    d = np.load(f"{REF}/experiments/boolean_logic/random_layouts/donor_layouts.npy")
    np.savez_compressed(os.path.join(OUT, "layouts.npz"), acceptor_layouts=a[:4], donor_layouts=d[:4])


def fixtures_all():
    sets = ["rnd_min_max", "rnd_min_max5M", "XOR_wide", "XOR_wide5M"]
    cols = {k: [] for k in ("acceptors", "donors", "electrodes", "E_constant", "occupation", "mean_currents",
                            "stddev_currents", "scalars")}
    names = []
    for setname in sets:
        for i in range(100):
            d = load_kmc(f"{REF}/thesis_indrek/tests/{setname}/test{i}.kmc")
            pos = np.vstack([d["acceptors"], d["electrodes"][:, :3]])
            dist = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1))
            tc = d["nu"] * np.exp(-2 * dist / d["ab"]) - np.eye(len(pos))
            assert np.abs(dist - d["distances"]).max() <= 1e-15 and np.abs(tc - d["transitions_constant"]).max() <= 1e-15
            names.append(f"{setname}/test{i}")
            for k in ("acceptors", "donors", "electrodes", "E_constant", "mean_currents", "stddev_currents"):
                cols[k].append(np.asarray(d[k], dtype=np.float64))
            cols["occupation"].append(np.asarray(d["occupation"], dtype=np.uint8))
            cols["scalars"].append(np.array([d["nu"], d["kT"], d["I_0"], d["R"], d["ab"]], dtype=np.float64))
    np.savez_compressed(os.path.join(OUT, "fixtures_all.npz"), names=np.array(names),
                        **{k: np.stack(v) for k, v in cols.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    fixtures_all()
    ref, seed_fn = numba_ref.load()
    cases = {
        "c1_basic_N10_P2": (synthetic_case(10, 2, 0), 11, 4000, 1500),
        "c2_grid_N16_P8": (synthetic_case(16, 8, 1, M=3), 12, 3000, 1000),
        "n5_p3_hot": (synthetic_case(5, 3, 2, kT=5.0, I_0=20.0), 13, 3000, 3000),
        "n31_p1": (synthetic_case(31, 1, 3, M=4), 14, 1500, 500),
        "fx_rnd_min_max_0": (fixture_case(f"{REF}/thesis_indrek/tests/rnd_min_max/test0.kmc"), 21, 3000, 1000),
        "fx_rnd_min_max_57": (fixture_case(f"{REF}/thesis_indrek/tests/rnd_min_max/test57.kmc"), 22, 3000, 1000),
        "fx_xor_wide_3": (fixture_case(f"{REF}/thesis_indrek/tests/XOR_wide/test3.kmc"), 23, 3000, 1000),
    }
    blob = {}
    for name, (c, seed, hops, th) in cases.items():
        r = run_reference(ref, seed_fn, c, seed, hops, th)
        for k, v in c.items():
            blob[f"{name}/{k}"] = np.asarray(v)
        blob[f"{name}/seed"] = np.asarray(seed); blob[f"{name}/hops"] = np.asarray(hops)
        for k, v in r.items():
            if v is not None:
                blob[f"{name}/ref_{k}"] = np.asarray(v)
        print(name, "time", r["time"], "eo", r["electrode_occupation"])
    np.savez_compressed(os.path.join(OUT, "py_replay.npz"), **blob)

    # --- subset of the reference's statistical fixtures
    blob = {}
    picks = [("rnd_min_max", [0, 1, 2, 57]), ("rnd_min_max5M", [0, 1, 2, 57]), ("XOR_wide", [0, 1, 2, 3]),
             ("XOR_wide5M", [0, 1, 2, 3])]
    keep = ["N", "M", "P", "nu", "kT", "I_0", "R", "ab", "mu", "res", "xdim", "ydim", "zdim", "distances",
            "transitions_constant", "E_constant", "eV_constant", "comp_constant", "electrodes", "acceptors",
            "donors", "occupation", "mean_currents", "stddev_currents", "expected_current", "time"]
    for setname, idx in picks:
        for i in idx:
            d = load_kmc(f"{REF}/thesis_indrek/tests/{setname}/test{i}.kmc")
            for k in keep:
                blob[f"{setname}/test{i}/{k}"] = np.asarray(d[k])
    np.savez_compressed(os.path.join(OUT, "fixtures_subset.npz"), **blob)

    # --- electrostatics pins: every 10th fixture of two sets
    blob = {}
    for setname in ["rnd_min_max", "XOR_wide"]:
        for i in range(0, 100, 5):
            d = load_kmc(f"{REF}/thesis_indrek/tests/{setname}/test{i}.kmc")
            for k in ["electrodes", "acceptors", "donors", "eV_constant", "comp_constant", "E_constant", "R", "I_0",
                      "mu", "res", "xdim", "ydim", "static_electrodes"]:
                blob[f"{setname}/test{i}/{k}"] = np.asarray(d[k])
    np.savez_compressed(os.path.join(OUT, "electrostatics.npz"), **blob)
    layouts()
    for f in sorted(glob.glob(os.path.join(OUT, "*.npz"))):
        print(f, os.path.getsize(f))


if __name__ == "__main__":
    main()
