"""ctypes binding of libkmcb200.so (the C ABI declared in include/kmc_b200.h).

The library is the product: if it is missing or there is no CUDA device the calls
raise -- there is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libkmcb200.so")

MODE_FAST, MODE_GO_SIMULATE, MODE_GO_RECORDPLUS, MODE_PY, MODE_FAST_REFORDER, MODE_PROB = 0, 1, 2, 3, 4, 5
FLAG_DEVICE_PTRS = 1
FLAG_NO_MEMO = 2
FLAG_LANES = 4      # force the thread-per-trajectory kernel (hop_lanes.cu)
FLAG_NO_LANES = 8   # never use it
FLAG_SOLO = 16      # force the latency kernel (kmc_solo_kernel: a few trajectories, one warp each)
FLAG_NO_SOLO = 32   # never use it


class GoSlice(C.Structure):
    """Go slice header exactly as the reference declares it (goSimulation/pythonBind.py:28-30)."""
    _fields_ = [("data", C.POINTER(C.c_double)), ("len", C.c_longlong), ("cap", C.c_longlong)]


class EnsembleArgs(C.Structure):
    """kmcb200_ensemble_args (include/kmc_b200.h)."""
    _fields_ = [
        ("B", C.c_int64), ("hops", C.c_int64), ("prehops", C.c_int64), ("mode", C.c_int32), ("flags", C.c_int32),
        ("E_constant", C.c_void_p), ("basis", C.c_void_p), ("electrode_v", C.c_void_p), ("kT", C.c_void_p),
        ("occupation0", C.c_void_p), ("seed", C.c_uint64), ("member_index0", C.c_uint64),
        ("stream_e", C.c_void_p), ("stream_u", C.c_void_p), ("stream_u64", C.c_void_p),
        ("time", C.c_void_p), ("electrode_occ", C.c_void_p), ("occupation_out", C.c_void_p),
        ("site_energies_out", C.c_void_p), ("avg_occupation", C.c_void_p), ("traffic", C.c_void_p),
        ("trace", C.c_void_p), ("misses", C.c_void_p), ("prob_occupation", C.c_void_p),
        ("prob_electrode_occ", C.c_void_p), ("stream", C.c_void_p),
    ]


_SINGLE = [C.c_longlong, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
           GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, C.c_int, C.c_bool, GoSlice, GoSlice]
_PRUNED = [C.c_longlong, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
           GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, C.c_int, C.c_bool, GoSlice, GoSlice]

EXPORTS = ["wrapperSimulate", "wrapperSimulateRecord", "wrapperSimulateRecordPlus", "wrapperSimulatePruned",
           "wrapperSimulateProbability",
           "parallelSimulations", "kmcb200_device_count", "kmcb200_last_error", "kmcb200_version",
           "kmcb200_set_seed", "kmcb200_layout_create", "kmcb200_layout_destroy", "kmcb200_run_ensemble",
           "kmcb200_run_ensemble_multi",
           "kmcb200_probe_rates", "kmcb200_reduce_currents", "kmcb200_launch_count", "kmcb200_last_kernel", "kmcb200_sizeof_ensemble_args", "kmcb200_measure_peak"]

_lib = None


def load():
    """Load libkmcb200.so (built by `python -m kmc_dn_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is missing: build it with `python -m kmc_dn_b200.build` "
                           "(kmc_dn_b200 has no CPU fallback)")
    lib = C.CDLL(SO_PATH)
    for name in ("wrapperSimulate", "wrapperSimulateRecord", "wrapperSimulateRecordPlus", "wrapperSimulateProbability"):
        f = getattr(lib, name); f.argtypes = _SINGLE; f.restype = C.c_double
    lib.wrapperSimulatePruned.argtypes = _PRUNED
    lib.wrapperSimulatePruned.restype = C.c_double
    lib.parallelSimulations.argtypes = [GoSlice] * 14
    lib.parallelSimulations.restype = C.c_longlong
    lib.kmcb200_device_count.restype = C.c_int
    lib.kmcb200_last_error.restype = C.c_char_p
    lib.kmcb200_version.restype = C.c_char_p
    lib.kmcb200_set_seed.argtypes = [C.c_uint64]
    lib.kmcb200_layout_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                          C.c_double, C.c_double, C.c_double]
    lib.kmcb200_layout_create.restype = C.c_void_p
    lib.kmcb200_layout_destroy.argtypes = [C.c_void_p]
    lib.kmcb200_run_ensemble.argtypes = [C.c_void_p, C.POINTER(EnsembleArgs)]
    lib.kmcb200_run_ensemble.restype = C.c_int
    lib.kmcb200_run_ensemble_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(EnsembleArgs)]
    lib.kmcb200_run_ensemble_multi.restype = C.c_int
    lib.kmcb200_probe_rates.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p]
    lib.kmcb200_probe_rates.restype = C.c_int
    lib.kmcb200_reduce_currents.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
    lib.kmcb200_reduce_currents.restype = C.c_int
    lib.kmcb200_launch_count.restype = C.c_longlong
    lib.kmcb200_last_kernel.restype = C.c_char_p
    lib.kmcb200_measure_peak.argtypes = [C.c_int, C.c_int]
    lib.kmcb200_measure_peak.restype = C.c_double
    lib.kmcb200_sizeof_ensemble_args.restype = C.c_int
    if lib.kmcb200_sizeof_ensemble_args() != C.sizeof(EnsembleArgs):
        raise RuntimeError("kmcb200_ensemble_args layout mismatch between libkmcb200.so and kmc_dn_b200/_lib.py")
    _lib = lib
    return lib


def last_error():
    return load().kmcb200_last_error().decode()


def goslice(arr):
    """Zero-copy GoSlice over a contiguous float64 numpy array (the array must outlive the call)."""
    assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
    return GoSlice(arr.ctypes.data_as(C.POINTER(C.c_double)), arr.size, arr.size)
