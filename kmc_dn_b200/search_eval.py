"""Fitness evaluation of a whole generation of voltage genes in ONE ensemble launch -- the consumer
side of the hop loop in the reference's searches.

The reference evaluates a candidate by looping over the logic-table tests, re-solving the potential
with FEniCS and running one simulation per test (voltage_search.evaluate_error_corr,
voltage_search.py:111-136), or -- batched -- by queueing len(tests) simulations per candidate into
parallelSimulations (voltage_search.parallel_simulation / evaluate_error_corr_parallel, :138-185).
Here every (candidate, test, seed) triple is one member of one GPU ensemble; the potential comes from
the per-electrode basis (superposition), and the per-member currents are reduced to the reference's
error value on the host (a few kB).

The gene logic (mutation, cross-over, annealing schedule: dn_search.py) stays with the caller.
"""
import numpy as np


def perfect_correlation(tests):
    """10 for tests expecting True, 0 otherwise (voltage_search.py:40-46)."""
    return np.array([10 if t[1] else 0 for t in tests])


def error_corr(values, tests, corr_pow=1):
    """The reference's error for ONE candidate from its output currents `values[len(tests)]`
    (voltage_search.py:118-136 == :160-185): separation = highest_false - lowest_true with the reference's
    initial values (1, -1); if the separation is positive it is returned as is, else corr**corr_pow * separation,
    corr = max(0, corrcoef(perfect_correlation, values))."""
    values = np.asarray(values, dtype=np.float64)
    lowest_true, highest_false = 1, -1
    for v, t in zip(values, tests):
        if t[1]:
            lowest_true = min(lowest_true, v)
        else:
            highest_false = max(highest_false, v)
    with np.errstate(invalid="ignore", divide="ignore"):
        corr = np.corrcoef(perfect_correlation(tests), values)[0][1]
    separation = highest_false - lowest_true
    if corr < 0:
        corr = 0
    if separation > 0:
        return separation
    return (corr ** corr_pow) * separation


def error_diff(values, tests):
    """Separation only (voltage_search.evaluate_error_diff, voltage_search.py:92-109): highest_false - lowest_true with the
    reference's initial values (1, -1)."""
    lowest_true, highest_false = 1, -1
    for v, t in zip(np.asarray(values, dtype=np.float64), tests):
        if t[1]:
            lowest_true = min(lowest_true, v)
        else:
            highest_false = max(highest_false, v)
    return highest_false - lowest_true


def generation_members(controls, tests, n_electrodes, seeds=1, output_value=0.0):
    """Voltage matrix [G*T*seeds, P] of a generation: inputs on electrodes 0..len(test[0])-1
    (voltage_search.py:115-117), control genes on the following electrodes (init_random_voltages, :87-90),
    the output electrode (last) at `output_value`.  Member index = ((g*T + t)*seeds + s)."""
    controls = np.atleast_2d(np.asarray(controls, dtype=np.float64))
    G, C = controls.shape
    T = len(tests)
    n_in = len(tests[0][0])
    if n_in + C + 1 != n_electrodes:
        raise ValueError("electrodes = inputs + control genes + one output")
    V = np.full((G, T, n_electrodes), float(output_value))
    for t, test in enumerate(tests):
        V[:, t, :n_in] = np.asarray(test[0], dtype=np.float64)[None, :]
    V[:, :, n_in:n_in + C] = controls[:, None, :]
    return np.repeat(V.reshape(G * T, n_electrodes), seeds, axis=0)


def evaluate_generation(layout, basis, controls, tests, hops, kT=1.0, seeds=1, prehops=0, corr_pow=1, seed=0,
                        occupation0=None, output_electrode=None):
    """Run G candidates x T tests x `seeds` seeds as one ensemble on `layout` (kmc_dn_b200.ensemble.Layout) and
    return (errors[G], currents[G,T]) where currents are seed-averaged output-electrode currents
    (electrode_occupation/time, kmc_dopant_networks.py:618)."""
    P = layout.P
    out = P - 1 if output_electrode is None else output_electrode
    V = generation_members(controls, tests, P, seeds=seeds)
    r = layout.run(hops, kT, V, basis=basis, prehops=prehops, seed=seed, occupation0=occupation0)
    G, T = len(np.atleast_2d(controls)), len(tests)
    cur = r["current"][:, out].reshape(G, T, seeds).mean(axis=2)
    errors = np.array([error_corr(cur[g], tests, corr_pow) for g in range(G)])
    return errors, cur


# ---------------------------------------------------------------------------------------- a generation loop on top
def genes_of(controls, voltage_range):
    """Control voltages -> uint16 genes (voltage_search.getGenes, voltage_search.py:214-220)."""
    return np.uint16((np.asarray(controls, dtype=np.float64) + voltage_range) / voltage_range / 2 * 65535)


def controls_of(genes, voltage_range):
    """uint16 genes -> control voltages (voltage_search.getDnFromGenes, :223-227)."""
    return np.asarray(genes, dtype=np.float64) / 65535 * 2 * voltage_range - voltage_range


def genetic_search(layout, basis, tests, gen_size=32, generations=10, voltage_range=150.0, hops=100000, seeds=4, kT=1.0,
                   disparity=2.0, mut_rate=0.1, corr_pow=1, seed=0, occupation0=None, prehops=0, on_generation=None):
    """The reference's genetic voltage search (dn_search.genetic_search, dn_search.py:411-549, with voltage_search's gene
    coding) as a CONSUMER of the batched hop loop: every generation -- gen_size candidates x len(tests) tests x `seeds`
    seeds -- is ONE ensemble launch (evaluate_generation), where the reference runs gen_size * len(tests) simulations one
    after the other (or `parallel` dns at a time through parallelSimulations, voltage_search.py:138-157).

    Kept from the reference: uint16 genes per control electrode; the best `4 - gen_size % 2` candidates survive unchanged
    (:449); the others are bred from parents drawn by rank with the disparity weighting of :452-456 / :500-513; single-point
    cross-over; per-gene mutation with probability mut_rate.  Left out (gene logic outside the hot path):
    uniqueness forcing, uniqueness schedules, validation strategies, wall-clock budgets.
    Returns (best_error, best_controls, history[generation] = (best, mean) error)."""
    rng = np.random.default_rng(seed)
    n_in = len(tests[0][0])
    C = layout.P - n_in - 1
    genes = rng.integers(0, 65536, size=(gen_size, C), dtype=np.uint16)
    preserved = 4 - (gen_size % 2)
    n_cross = gen_size - preserved
    w = np.array([abs(disparity * ((1 - (i + 0.5) / n_cross) ** (disparity - 1))) for i in range(n_cross)])
    w = w + (n_cross - w.sum()) / n_cross  # (disparity_offset, :452-456)
    w = np.maximum(w, 0) / np.maximum(w, 0).sum()
    best = (np.inf, None)
    history = []
    for g in range(generations):
        errors, cur = evaluate_generation(layout, basis, controls_of(genes, voltage_range), tests, hops, kT=kT, seeds=seeds,
                                          prehops=prehops, corr_pow=corr_pow, seed=seed * 1000003 + g, occupation0=occupation0)
        order = np.argsort(np.where(np.isnan(errors), np.inf, errors))  # (undefined correlation: ranked last)
        if errors[order[0]] < best[0]:
            best = (float(errors[order[0]]), controls_of(genes[order[0]], voltage_range).copy())
        history.append((float(errors[order[0]]), float(np.nanmean(errors))))
        if on_generation:
            on_generation(g, errors, cur)
        if g == generations - 1:
            break
        nxt = [genes[order[i]].copy() for i in range(preserved)]
        ranked = genes[order[:n_cross]]
        while len(nxt) < gen_size:
            a, b = ranked[rng.choice(n_cross, p=w)], ranked[rng.choice(n_cross, p=w)]
            cut = int(rng.integers(1, C)) if C > 1 else 0
            child = np.concatenate([a[:cut], b[cut:]]).astype(np.uint16)
            mut = rng.random(C) < mut_rate
            child[mut] = rng.integers(0, 65536, size=int(mut.sum()), dtype=np.uint16)
            nxt.append(child)
        genes = np.array(nxt, dtype=np.uint16)
    return best[0], best[1], history
