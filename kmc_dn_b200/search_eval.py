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
