"""Generates tests/golden/search_eval.npz in the BUILD CONTAINER (needs /root/reference): golden (currents -> error) pairs of
the reference's fitness functions, computed by the UNMODIFIED reference code.

    python oracle/make_golden_search.py

`voltage_search.evaluate_error_corr` (voltage_search.py:111-136), `evaluate_error_corr_parallel` (:160-185) and
`evaluate_error_diff` (:92-109) are pure Python; they are imported from /root/reference with `fenics` and `matplotlib`
stubbed (module-level imports the fitness code never touches) and driven with a stand-in `dn` whose simulation call returns
prescribed output currents.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KMC_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def load_reference():
    _stub("fenics")
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    _stub("matplotlib.pyplot", ioff=lambda *a, **k: None, ion=lambda *a, **k: None)
    _stub("matplotlib.colors", LinearSegmentedColormap=object)
    for sub in ("animation", "gridspec", "patches", "cm", "ticker"):
        setattr(mpl, sub, _stub("matplotlib." + sub))
    _stub("mpl_toolkits"); _stub("mpl_toolkits.mplot3d", Axes3D=object)
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import voltage_search  # noqa: E402
    finally:
        os.chdir(cwd)
    return voltage_search


class FakeDn:
    """Stand-in for kmc_dn: the fitness functions set electrode voltages, call update_V and the simulation function, then read
    `current[output_electrode]` (sequential path) or `parrallel_results[i][2][output_electrode]` (batched path)."""

    def __init__(self, n_electrodes, currents):
        self.electrodes = np.zeros((n_electrodes, 4))
        self._currents = list(currents)
        self._k = 0
        self.current = np.zeros(n_electrodes)
        self.parrallel_results = [(1.0, np.zeros(n_electrodes), self._row(c, n_electrodes)) for c in currents]
        self.true_voltage = 50.0

    @staticmethod
    def _row(c, n):
        r = [0.0] * n
        r[n - 1] = float(c)
        return r

    def update_V(self):
        pass

    def go_simulation(self, **kw):
        self.current = np.array(self._row(self._currents[self._k % len(self._currents)], len(self.electrodes)))
        self._k += 1


def main():
    vs = load_reference()
    rng = np.random.default_rng(20261018)
    tables = {"xor": [((0, 0), False), ((0, 75), True), ((75, 0), True), ((75, 75), False)],
              "and": [((0, 0), False), ((0, 75), False), ((75, 0), False), ((75, 75), True)],
              "or": [((0, 0), False), ((0, 75), True), ((75, 0), True), ((75, 75), True)]}
    out = {}
    for name, tests in tables.items():
        vals, e_seq, e_par, e_diff, pows = [], [], [], [], []
        for k in range(40):
            scale = 10.0 ** rng.uniform(-6, -2)
            v = rng.normal(0, 1, len(tests)) * scale
            if k % 5 == 0:  # well separated candidates (positive separation)
                v = np.array([(-1.0 if t[1] else 1.0) for t in tests]) * scale * rng.uniform(0.5, 2) + rng.normal(0, 0.05, len(tests)) * scale
            cp = int(rng.integers(1, 4))
            s = object.__new__(vs.voltage_search)  # (the constructor draws random voltages and needs a full kmc_dn)
            s.tests = tests
            s.perfect_correlation = np.array([10 if t[1] else 0 for t in tests])
            s.corr_pow = cp
            s.simulation_func = "go_simulation"
            s.simulation_args = {}
            s.output_electrode = 7
            e_seq.append(vs.voltage_search.evaluate_error_corr(s, FakeDn(8, v)))
            e_par.append(vs.voltage_search.evaluate_error_corr_parallel(s, FakeDn(8, v)))
            e_diff.append(vs.voltage_search.evaluate_error_diff(s, FakeDn(8, v)))
            vals.append(v); pows.append(cp)
        out[f"{name}_values"] = np.array(vals)
        out[f"{name}_corr_pow"] = np.array(pows)
        out[f"{name}_error_corr"] = np.array(e_seq, dtype=np.float64)
        out[f"{name}_error_corr_parallel"] = np.array(e_par, dtype=np.float64)
        out[f"{name}_error_diff"] = np.array(e_diff, dtype=np.float64)
        out[f"{name}_tests_inputs"] = np.array([t[0] for t in tests], dtype=np.float64)
        out[f"{name}_tests_expected"] = np.array([t[1] for t in tests])
    path = os.path.join(ROOT, "tests", "golden", "search_eval.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("xor")})


if __name__ == "__main__":
    main()
