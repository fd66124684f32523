"""ctypes front end of the CPU parity oracle (oracle/kmc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by kmc_dn_b200/.
Every function restates a reference function; the file:line it follows is in
the C source next to it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmc_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with gcc (seconds)."""
    src = os.path.join(_HERE, "kmc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.kmc_oracle_py.restype = C.c_double
        _lib.kmc_oracle_go.restype = C.c_double
        _lib.kmc_oracle_go_ensemble.restype = C.c_int
        _lib.kmc_oracle_py_ensemble.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def py_simulate(N, P, nu, kT, I_0, R, occupation, distances, E_constant, site_energies,
                transitions_constant, electrode_occupation, hops, record=False, u=None, seed=0,
                trace=False):
    """numba semantics (kmc_dopant_networks.py:33-135).  Returns a dict; inputs are not mutated."""
    S = N + P
    occ = np.ascontiguousarray(np.asarray(occupation) > 0, dtype=np.uint8).copy()
    se = _c64(site_energies).copy()
    eo = np.ascontiguousarray(electrode_occupation, dtype=np.int64).copy()
    d = _c64(distances); tc = _c64(transitions_constant); E = _c64(E_constant)
    uu = None if u is None else _c64(u)
    if uu is not None:
        assert uu.size >= 2 * hops
    traffic = np.zeros((S, S)) if record else None
    occ_time = np.zeros(N) if record else None
    tr = np.zeros((hops, 2), dtype=np.int32) if trace else None
    t = lib().kmc_oracle_py(C.c_int(N), C.c_int(P), C.c_double(nu), C.c_double(kT), C.c_double(I_0),
                            C.c_double(R), _p(occ), _p(d), _p(E), _p(se), _p(tc), _p(eo),
                            C.c_int64(hops), C.c_int(int(record)), _p(uu), C.c_uint64(seed),
                            _p(traffic), _p(occ_time), _p(tr))
    return dict(time=t, occupation=occ.astype(bool), electrode_occupation=eo, site_energies=se,
                traffic=traffic, occ_time=occ_time, trace=tr)


def go_simulate(N, P, nu, kT, I_0, R, distances, E_constant, transitions_constant, site_energies,
                hops, variant=0, occupation=None, use_cache=False, record=False, cut=0.0,
                e=None, u=None, seed=0, trace=False):
    """Go semantics.  variant 0 = simulate (simulation.go:194-325), 1 = simulateRecordPlus (:327-432).
    occupation=None is the all-empty start of the single-run exports."""
    S = N + P
    d = _c64(distances); tc = _c64(transitions_constant); E = _c64(E_constant); se = _c64(site_energies)
    occ_in = None if occupation is None else _c64(np.asarray(occupation, dtype=np.float64))
    eo = np.zeros(P)
    ee = None if e is None else _c64(e)
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float32)
    if ee is not None:
        assert ee.size >= hops and uu is not None and uu.size >= hops
    traffic = np.zeros((S, S)) if record else None
    avg = np.zeros(N) if record else None
    occ_out = np.zeros(max(N, 1), dtype=np.uint8)
    tr = np.zeros((hops, 2), dtype=np.int32) if trace else None
    se_out = np.zeros(S, dtype=np.float32)
    miss = C.c_int64(0)
    t = lib().kmc_oracle_go(C.c_int(variant), C.c_int(N), C.c_int(P), C.c_double(nu), C.c_double(kT),
                            C.c_double(I_0), C.c_double(R), _p(occ_in), _p(d), _p(E), _p(tc), _p(eo), _p(se),
                            C.c_int64(hops), C.c_int(int(use_cache)), C.c_int(int(record)), C.c_double(cut),
                            _p(ee), _p(uu), C.c_uint64(seed), _p(traffic), _p(avg), _p(occ_out), _p(tr),
                            _p(se_out), C.byref(miss))
    return dict(time=t, occupation=occ_out[:N].astype(bool), electrode_occupation=eo, traffic=traffic,
                average_occupation=avg, trace=tr, site_energies=se_out, misses=miss.value)


def go_rates(N, P, nu, kT, I_0, R, occupation, distances, E_constant, transitions_constant, site_energies):
    """fp32 energies + dense rate matrix for a given occupation (simulation.go:226-234, :58-80)."""
    S = N + P
    occ = np.ascontiguousarray(np.asarray(occupation) > 0, dtype=np.uint8)
    se = np.zeros(S, dtype=np.float32); rates = np.zeros((S, S), dtype=np.float32)
    lib().kmc_oracle_go_rates(C.c_int(N), C.c_int(P), C.c_double(nu), C.c_double(kT), C.c_double(I_0),
                              C.c_double(R), _p(occ), _p(_c64(distances)), _p(_c64(E_constant)),
                              _p(_c64(transitions_constant)), _p(_c64(site_energies)), _p(se), _p(rates))
    return se, rates


def py_rates(N, P, nu, kT, I_0, R, occupation, distances, E_constant, transitions_constant, site_energies):
    """fp64 energies + rate matrix for a given occupation (kmc_dopant_networks.py:58-87)."""
    S = N + P
    occ = np.ascontiguousarray(np.asarray(occupation) > 0, dtype=np.uint8)
    se = np.zeros(S); rates = np.zeros((S, S))
    lib().kmc_oracle_py_rates(C.c_int(N), C.c_int(P), C.c_double(nu), C.c_double(kT), C.c_double(I_0),
                              C.c_double(R), _p(occ), _p(_c64(distances)), _p(_c64(E_constant)),
                              _p(_c64(transitions_constant)), _p(_c64(site_energies)), _p(se), _p(rates))
    return se, rates


def go_ensemble(N, P, nu, kT, I_0, R, distances, E_constant, transitions_constant, electrode_v, hops,
                variant=1, use_cache=True, occupation0=None, seed0=0, nthreads=0):
    """B members of one layout on host threads (shape of parallelSimulations,
    simulationWrapper.go:274-316).  E_constant [B,N], electrode_v [B,P], kT scalar or [B]."""
    E = _c64(E_constant); V = _c64(electrode_v)
    B = E.shape[0]
    kTa = _c64(np.broadcast_to(np.asarray(kT, dtype=np.float64), (B,)))
    occ0 = None if occupation0 is None else _c64(np.asarray(occupation0, dtype=np.float64))
    time = np.zeros(B); eo = np.zeros((B, P))
    used = lib().kmc_oracle_go_ensemble(C.c_int(variant), C.c_int(int(use_cache)), C.c_int(N), C.c_int(P),
                                        C.c_int64(B), C.c_double(nu), _p(kTa), C.c_double(I_0), C.c_double(R),
                                        _p(occ0), _p(_c64(distances)), _p(E), _p(_c64(transitions_constant)),
                                        _p(V), C.c_int64(hops), C.c_uint64(seed0), C.c_int(nthreads),
                                        _p(time), _p(eo))
    return dict(time=time, electrode_occupation=eo, threads=used)


def py_ensemble(N, P, nu, kT, I_0, R, distances, E_constant, transitions_constant, electrode_v, hops,
                occupation0=None, seed0=0, nthreads=0):
    E = _c64(E_constant); V = _c64(electrode_v)
    B = E.shape[0]
    kTa = _c64(np.broadcast_to(np.asarray(kT, dtype=np.float64), (B,)))
    occ0 = None if occupation0 is None else _c64(np.asarray(occupation0, dtype=np.float64))
    time = np.zeros(B); eo = np.zeros((B, P), dtype=np.int64)
    used = lib().kmc_oracle_py_ensemble(C.c_int(N), C.c_int(P), C.c_int64(B), C.c_double(nu), _p(kTa),
                                        C.c_double(I_0), C.c_double(R), _p(occ0), _p(_c64(distances)), _p(E),
                                        _p(_c64(transitions_constant)), _p(V), C.c_int64(hops),
                                        C.c_uint64(seed0), C.c_int(nthreads), _p(time), _p(eo))
    return dict(time=time, electrode_occupation=eo, threads=used)


def prob_simulate(N, P, nu, kT, I_0, R, distances, E_constant, transitions_constant, site_energies, hops, record=False):
    """Mean-field solver probSimulate (goSimulation/probabilitySimulation.go:53-157) behind
    wrapperSimulateProbability (simulationWrapper.go:218-233)."""
    S = N + P
    occ = np.zeros(max(N, 1)); eo = np.zeros(max(P, 1)); se = _c64(site_energies).copy()
    traffic = np.zeros((S, S)) if record else None
    avg = np.zeros(max(N, 1)) if record else None
    lib().kmc_oracle_prob.restype = C.c_double
    t = lib().kmc_oracle_prob(C.c_int(N), C.c_int(P), C.c_double(nu), C.c_double(kT), C.c_double(I_0), C.c_double(R),
                              _p(occ), _p(_c64(distances)), _p(_c64(E_constant)), _p(_c64(transitions_constant)), _p(eo),
                              _p(se), C.c_int64(hops), C.c_int(int(record)), _p(traffic), _p(avg))
    return dict(time=t, occupation=occ[:N], electrode_occupation=eo[:P], site_energies=se, traffic=traffic,
                average_occupation=None if avg is None else avg[:N])
