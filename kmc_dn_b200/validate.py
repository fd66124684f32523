"""Acceptance harness over the reference's `.kmc` fixtures -- a matplotlib-free restatement of
thesis_indrek/validate_tests.py:80-135.

For every fixture: run the candidate 5 times (all of them as ONE ensemble launch), take per-electrode mean and
population standard deviation of the currents, and compare with the stored 5-run mean_currents/stddev_currents by
the Bhattacharyya distance D; D > 0.9 is "extreme" (validate_tests.py:134).

    python -m kmc_dn_b200.validate <dir-with-test*.kmc> [--hops 1000000] [--limit 20]
"""
import argparse
import glob
import math
import os

import numpy as np

from .fixtures import load_kmc


def calc_D(meanp, meanv, std_p, std_v):
    """validate_tests.py:80-87."""
    var_p, var_v = std_p ** 2, std_v ** 2
    if var_p == 0 or var_v == 0:
        return math.log(10000)
    return 0.25 * math.log(0.25 * (var_p / var_v + var_v / var_p + 2)) + 0.25 * ((meanp - meanv) ** 2) / (var_p + var_v)


def evaluate_fixture(d, hops, runs=5, seed=0, kernel=None):
    """Returns (D[P], currents[runs,P]) for one fixture dict (as returned by load_kmc).  kernel: None = the library's
    choice (5 members: warp per trajectory), "lanes" = the thread-per-trajectory kernel."""
    from .ensemble import Layout
    N, P = int(d["N"]), int(d["P"])
    lay = Layout(N, P, d["distances"], d["transitions_constant"], nu=float(d["nu"]), I_0=float(d["I_0"]), R=float(d["R"]))
    try:
        # the fixtures were produced by wrapperSimulateRecordPlus: all-empty start (generate_tests.py:51)
        r = lay.run(hops, float(d["kT"]), np.tile(d["electrodes"][:, 3], (runs, 1)),
                    E_constant=np.tile(d["E_constant"], (runs, 1)), seed=seed, kernel=kernel)
    finally:
        lay.close()
    cur = r["current"]
    mu, sd = cur.mean(0), cur.std(0)
    D = np.array([calc_D(d["mean_currents"][i], mu[i], d["stddev_currents"][i], sd[i]) for i in range(P)])
    return D, cur


def compact_fixtures(npz_path):
    """Iterates over tests/golden/fixtures_all.npz (all 400 fixtures of the reference in compact form, written by
    oracle/make_golden.py): yields (name, dict) with the fields evaluate_fixture needs; distances and
    transitions_constant are rebuilt from the positions as kmc_dopant_networks.py:657-695, 824-830 does."""
    z = np.load(npz_path)
    for i, name in enumerate(z["names"]):
        nu, kT, I_0, R, ab = (float(x) for x in z["scalars"][i])
        pos = np.vstack([z["acceptors"][i], z["electrodes"][i][:, :3]])
        dist = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1))
        tc = nu * np.exp(-2 * dist / ab) - np.eye(len(pos))
        yield str(name), dict(N=z["acceptors"].shape[1], P=z["electrodes"].shape[1], nu=nu, kT=kT, I_0=I_0, R=R,
                              distances=dist, transitions_constant=tc, electrodes=z["electrodes"][i],
                              E_constant=z["E_constant"][i], mean_currents=z["mean_currents"][i],
                              stddev_currents=z["stddev_currents"][i])


def acceptance_over_sets(npz_path, stride_5m=1, seed0=0, stride_1m=1, kernel=None):
    """The reference's acceptance run (validate_tests.py:299-350) over its four fixture sets at the fixtures' own run
    lengths (1e6 hops, 5e6 for the *5M sets).  Returns {set: dict(fixtures, pairs, D_mean, D_median, extreme)}."""
    per = {}
    for k, (name, d) in enumerate(compact_fixtures(npz_path)):
        setname, t = name.split("/")
        big = setname.endswith("5M")
        if int(t[4:]) % (stride_5m if big else stride_1m):
            continue
        D, _ = evaluate_fixture(d, 5_000_000 if big else 1_000_000, seed=seed0 + k, kernel=kernel)
        per.setdefault(setname, []).append(D)
    out = {}
    for setname, Ds in per.items():
        Ds = np.array(Ds)
        out[setname] = dict(fixtures=int(len(Ds)), pairs=int(Ds.size), D_mean=float(Ds.mean()), D_median=float(np.median(Ds)),
                            extreme=float((Ds > 0.9).mean()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("directory")
    ap.add_argument("--hops", type=int, default=None, help="default: 5e6 for '*5M' directories, else 1e6")
    ap.add_argument("--limit", type=int, default=0)
    a = ap.parse_args()
    files = sorted(glob.glob(os.path.join(a.directory, "test*.kmc")), key=lambda f: int("".join(c for c in os.path.basename(f) if c.isdigit())))
    if a.limit:
        files = files[:a.limit]
    hops = a.hops or (5_000_000 if a.directory.rstrip("/").endswith("5M") else 1_000_000)
    Ds = []
    for k, f in enumerate(files):
        D, _ = evaluate_fixture(load_kmc(f), hops, seed=k)
        Ds.append(D)
        flag = " EXTREME" if (D > 0.9).any() else ""
        print(f"{os.path.basename(f):14s} D mean {D.mean():.3f} max {D.max():.3f}{flag}")
    Ds = np.concatenate(Ds)
    print(f"{len(files)} fixtures, {len(Ds)} (fixture, electrode) pairs: D mean {Ds.mean():.3g}, sd {Ds.std():.3g}, "
          f"extreme (D > 0.9): {(Ds > 0.9).sum()}")


if __name__ == "__main__":
    main()
