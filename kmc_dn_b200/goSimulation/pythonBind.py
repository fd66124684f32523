"""Drop-in for the reference's goSimulation/pythonBind.py: same public name, keyword
arguments, return tuples and error behaviour, bound to libkmcb200.so instead of the cgo
libSimulation.so.

Differences from the reference binding (goSimulation/pythonBind.py:49-90), none of them visible
to callers: numpy buffers are passed by pointer (the reference boxes every element into a
c_double, O(S^2) Python objects per call), and the library is loaded once per process
(the reference re-loads "./goSimulation/libSimulation.so" relative to the CWD on every call).
"""
import numpy as np

from .. import _lib
from .._lib import GoSlice, goslice  # noqa: F401  (GoSlice re-exported like the reference module)

_SYMBOLS = ("wrapperSimulate", "wrapperSimulateRecord", "wrapperSimulateRecordPlus", "wrapperSimulatePruned",
            "wrapperSimulateProbability")


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def callGoSimulation(N_acceptors, N_electrodes, nu, kT, I_0, R, time, occupation, distances, E_constant,
                     site_energies, transitions_constant, transitions, problist, electrode_occupation, hops,
                     record, goSpecificFunction, prune_threshold=0.0):
    """Signature and returns of goSimulation/pythonBind.py:49-90.

    Returns (time, occupation, electrode_occupation) or, when `record`,
    (time, occupation, electrode_occupation, traffic[S,S], average_occupation[N]).
    `occupation` comes back as the caller passed it (the exports never write it, :84)."""
    if goSpecificFunction not in _SYMBOLS:
        # the reference would raise AttributeError from getattr(lib, name) on an unknown export
        raise AttributeError(f"libkmcb200.so: undefined symbol: {goSpecificFunction}")
    if not isinstance(hops, (int, np.integer)):
        # ctypes rejects a float for c_int in the reference binding (pythonBind.py:65-68)
        raise TypeError("hops must be an int")
    lib = _lib.load()
    N = N_acceptors + N_electrodes
    d = _f64(distances); tc = _f64(transitions_constant)
    occ = _f64(occupation); Ec = _f64(E_constant); se = _f64(site_energies)
    eo = _f64(electrode_occupation)
    traffic = np.zeros(N * N); avg = np.zeros(N_acceptors)
    args = [int(N_acceptors), int(N_electrodes), float(nu), float(kT), float(I_0), float(R), float(time),
            goslice(occ), goslice(d), goslice(Ec), goslice(tc), goslice(eo), goslice(se), int(hops), bool(record),
            goslice(traffic), goslice(avg)]
    if goSpecificFunction == "wrapperSimulatePruned":
        args.insert(2, float(prune_threshold))
    t = getattr(lib, goSpecificFunction)(*args)
    r_eo = eo.astype(np.int64)
    r_occ = occ.astype(np.int64)
    if not record:
        return (t, r_occ, r_eo)
    return t, r_occ, r_eo, traffic.reshape(N, N), avg
