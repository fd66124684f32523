"""Shared helpers for the test-suite (golden loaders, acceptance statistics)."""
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases(fname):
    """npz with keys '<case>/<field>' -> {case: {field: array}} (scalars unwrapped)."""
    z = np.load(os.path.join(GOLDEN, fname))
    out = {}
    for key in z.files:
        case, field = key.rsplit("/", 1)
        v = z[key]
        out.setdefault(case, {})[field] = v.item() if v.ndim == 0 else v
    return out


def site_energies_of(c):
    """site_energies[N:] = electrode voltages (kmc_dopant_networks.py:899); [:N] scratch."""
    se = np.zeros(c["N"] + c["P"])
    se[c["N"]:] = c["electrode_v"]
    return se


def calc_D(meanp, meanv, std_p, std_v):
    """Bhattacharyya distance, the reference's acceptance metric (thesis_indrek/validate_tests.py:80-87)."""
    var_p = std_p ** 2
    var_v = std_v ** 2
    if var_p == 0 or var_v == 0:
        return math.log(10000)
    part1 = 0.25 * math.log(0.25 * (var_p / var_v + var_v / var_p + 2))
    part2 = 0.25 * (((meanp - meanv) ** 2) / (var_p + var_v))
    return part1 + part2


def mean_std(x, axis=0):
    """population mean / std over runs (validate_tests.py:18-31: divides by len)."""
    x = np.asarray(x, dtype=np.float64)
    return x.mean(axis=axis), x.std(axis=axis)


def first_divergence(a, b):
    """index of the first differing hop between two (H,2) traces, or H."""
    n = min(len(a), len(b))
    neq = np.nonzero((a[:n] != b[:n]).any(axis=1))[0]
    return int(neq[0]) if neq.size else n


def go_stream(seed, hops):
    """Injected stream of the Go contract: Exp(1) float64 variates and float32 uniforms in [0,1)."""
    rng = np.random.default_rng(seed)
    e = rng.standard_exponential(hops)
    u = rng.random(hops, dtype=np.float32)
    return e, u


def synthetic_layout(N, P, seed, kT=1.0, I_0=100.0, a=0.25, fill=0.8):
    """Uniform-random 2-D layout with P electrodes on the boundary (shape of examples/scaling.py); E_constant is a
    smooth synthetic potential.  Same keys as the golden cases."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((N, 3)); acc[:, :2] = rng.random((N, 2))
    el = np.zeros((P, 4))
    for p in range(P):
        t = (p // 4 + 1) / (P // 4 + 2) if P > 4 else 0.5
        el[p, :2] = [(0.0, t), (1.0, t), (t, 0.0), (t, 1.0)][p % 4]
        el[p, 3] = rng.uniform(-30, 30)
    pos = np.vstack([acc, el[:, :3]])
    d = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
    R = N ** -0.5
    tc = np.exp(-2 * d / (a * R)) - np.eye(N + P)
    w = np.exp(-4 * ((acc[:, None, :2] - el[None, :, :2]) ** 2).sum(-1))
    E = (w * el[None, :, 3]).sum(1) / np.maximum(w.sum(1), 1e-9) + rng.normal(0, 3, N)
    occ = rng.random(N) < fill
    return dict(N=N, P=P, nu=1.0, kT=kT, I_0=I_0, R=R, distances=d, transitions_constant=tc, E_constant=E,
                electrode_v=el[:, 3].copy(), occupation=occ)
