"""Ensemble front end over the lean C ABI (kmcb200_layout_* / kmcb200_run_ensemble).

One `Layout` = one dopant layout (distances + transitions_constant tables) resident on one
GPU; `Layout.run(...)` executes B independent trajectories (seed x voltage vector x
temperature) of the reference's hop loop (goSimulation/simulation.go:194-432,
kmc_dopant_networks.py:33-135) in one launch.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (MODE_FAST, MODE_GO_SIMULATE, MODE_GO_RECORDPLUS, MODE_PY, MODE_FAST_REFORDER, MODE_PROB,  # noqa: F401
                   FLAG_DEVICE_PTRS, FLAG_NO_MEMO, FLAG_LANES, FLAG_NO_LANES, FLAG_SOLO, FLAG_NO_SOLO)


def _kernel_flags(kernel):
    """kernel: None = the library chooses; 'lanes' = thread-per-trajectory kernel (hop_lanes.cu); 'warp' = warp-per-
    trajectory kernels (hop_memo.cu / hop_wide.cu); 'solo' = latency kernel for a few trajectories (hop_lanes.cu)."""
    if kernel in (None, 'auto'):
        return 0
    return {'lanes': FLAG_LANES | FLAG_NO_SOLO, 'warp': FLAG_NO_LANES | FLAG_NO_SOLO, 'solo': FLAG_SOLO}[kernel]


def _host(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())  # torch tensor


class Layout:
    """Device-resident tables of one layout.

    distances, transitions_constant: (S,S) float64 as kmc_dn builds them
    (kmc_dopant_networks.py:657-695, 824-830); narrowed to float32 on upload exactly as the
    cgo wrappers do (goSimulation/simulationWrapper.go:37-56)."""

    def __init__(self, N, P, distances, transitions_constant, nu=1.0, I_0=100.0, R=1.0, prune_threshold=0.0,
                 device=0):
        self.lib = _lib.load()
        self.N, self.P, self.S = int(N), int(P), int(N) + int(P)
        d = _host(distances, np.float64); tc = _host(transitions_constant, np.float64)
        if d.shape != (self.S, self.S) or tc.shape != (self.S, self.S):
            raise ValueError("distances / transitions_constant must be (N+P, N+P)")
        self.device = device
        self.nu, self.I_0, self.R = float(nu), float(I_0), float(R)
        self._h = self.lib.kmcb200_layout_create(device, self.N, self.P, d.ctypes.data, tc.ctypes.data,
                                                 self.nu, self.I_0, self.R, float(prune_threshold))
        if not self._h:
            raise RuntimeError("kmcb200_layout_create: " + _lib.last_error())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.kmcb200_layout_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ host-buffer call
    def run(self, hops, kT, electrode_v, E_constant=None, basis=None, prehops=0, mode=MODE_FAST, occupation0=None,
            seed=0, member_index0=0, stream_e=None, stream_u=None, stream_u64=None, want_occupation=False,
            want_site_energies=False, record=False, trace=False, want_misses=False, memo=True, cuda_stream=None,
            kernel=None):
        """Run B members; every array is a host numpy array.  Returns a dict."""
        N, P, S = self.N, self.P, self.S
        V = _host(np.atleast_2d(electrode_v), np.float64)
        B = V.shape[0] if P > 0 else (np.atleast_2d(E_constant).shape[0])
        if P > 0 and V.shape != (B, P):
            raise ValueError("electrode_v must be (B,P)")
        kTa = _host(np.broadcast_to(np.asarray(kT, dtype=np.float64), (B,)), np.float64)
        Ec = None if E_constant is None else _host(np.atleast_2d(E_constant), np.float64)
        if Ec is not None and Ec.shape != (B, N):
            raise ValueError("E_constant must be (B,N)")
        bs = None if basis is None else _host(basis, np.float64)
        if bs is not None and bs.shape != (P + 1, N):
            raise ValueError("basis must be (P+1,N)")
        occ0 = None if occupation0 is None else _host(np.broadcast_to(np.asarray(occupation0) != 0, (B, N)), np.uint8)
        H = int(prehops) + int(hops)
        se_ = _host(stream_e, np.float64); su_ = _host(stream_u, np.float32); s64 = _host(stream_u64, np.float64)
        for arr, n in ((se_, B * H), (su_, B * H), (s64, 2 * B * H)):
            if arr is not None and arr.size != n:
                raise ValueError("injected stream has the wrong length")
        out = dict(time=np.zeros(B), electrode_occupation=np.zeros((B, P), dtype=np.int64))
        if want_occupation:
            out["occupation"] = np.zeros((B, N), dtype=np.uint8)
        if want_site_energies:
            out["site_energies"] = np.zeros((B, S))
        if record:
            out["avg_occupation"] = np.zeros((B, N))
            out["traffic"] = np.zeros((B, S, S))
        if trace:
            out["trace"] = np.zeros((B, int(hops), 2), dtype=np.int32)
        if want_misses:
            out["misses"] = np.zeros(B, dtype=np.int64)
        a = _lib.EnsembleArgs()
        a.B, a.hops, a.prehops, a.mode, a.flags = B, int(hops), int(prehops), int(mode), (0 if memo else FLAG_NO_MEMO)
        a.flags |= _kernel_flags(kernel)
        a.E_constant, a.basis, a.electrode_v, a.kT = _ptr(Ec), _ptr(bs), _ptr(V), _ptr(kTa)
        a.occupation0 = _ptr(occ0)
        a.misses = _ptr(out.get("misses"))
        a.seed, a.member_index0 = int(seed) & (2**64 - 1), int(member_index0)
        a.stream_e, a.stream_u, a.stream_u64 = _ptr(se_), _ptr(su_), _ptr(s64)
        a.time, a.electrode_occ = _ptr(out["time"]), _ptr(out["electrode_occupation"])
        a.occupation_out = _ptr(out.get("occupation"))
        a.site_energies_out = _ptr(out.get("site_energies"))
        a.avg_occupation, a.traffic, a.trace = _ptr(out.get("avg_occupation")), _ptr(out.get("traffic")), _ptr(out.get("trace"))
        a.stream = cuda_stream
        self._call(a)
        if "occupation" in out:
            out["occupation"] = out["occupation"].astype(bool)
        with np.errstate(divide="ignore", invalid="ignore"):
            out["current"] = out["electrode_occupation"] / out["time"][:, None]  # kmc_dopant_networks.py:618
        return out

    def _call(self, a):
        if self.lib.kmcb200_run_ensemble(self._h, C.byref(a)):
            raise RuntimeError("kmcb200_run_ensemble: " + _lib.last_error())

    # ------------------------------------------------------------------ device-pointer call (async)
    def run_device(self, B, hops, kT, electrode_v, time, electrode_occ, E_constant=None, basis=None, prehops=0,
                   mode=MODE_FAST, occupation0=None, seed=0, member_index0=0, occupation_out=None,
                   memo=True, cuda_stream=None, kernel=None):
        """Every array argument is a device tensor (torch) or raw device pointer holder with
        .data_ptr(); enqueues on `cuda_stream` and returns immediately."""
        a = _lib.EnsembleArgs()
        a.B, a.hops, a.prehops, a.mode = int(B), int(hops), int(prehops), int(mode)
        a.flags = FLAG_DEVICE_PTRS | (0 if memo else FLAG_NO_MEMO) | _kernel_flags(kernel)
        a.E_constant, a.basis, a.electrode_v, a.kT = _ptr(E_constant), _ptr(basis), _ptr(electrode_v), _ptr(kT)
        a.occupation0 = _ptr(occupation0)
        a.seed, a.member_index0 = int(seed) & (2**64 - 1), int(member_index0)
        a.time, a.electrode_occ, a.occupation_out = _ptr(time), _ptr(electrode_occ), _ptr(occupation_out)
        a.stream = cuda_stream
        if self.lib.kmcb200_run_ensemble(self._h, C.byref(a)):
            raise RuntimeError("kmcb200_run_ensemble: " + _lib.last_error())

    def reduce_currents_device(self, time, electrode_occ, group, out_sum, out_sumsq, out_count=None, cuda_stream=None):
        """Per-voltage-vector (sum, sum of squares, count) of the currents electrode_occ / time over `group` consecutive
        members, on the device (kmcb200_reduce_currents): the inputs are the device tensors of run_device, the outputs
        [B/group, P], [B/group, P], [B/group] float64 device tensors -- ready for an all_reduce / a small D2H copy."""
        B = int(time.numel()) if hasattr(time, "numel") else int(time.size)
        if self.lib.kmcb200_reduce_currents(self.device, _ptr(time), _ptr(electrode_occ), B, self.P, int(group), _ptr(out_sum),
                                            _ptr(out_sumsq), _ptr(out_count), cuda_stream):
            raise RuntimeError("kmcb200_reduce_currents: " + _lib.last_error())

    def run_prob(self, steps, kT, electrode_v, E_constant=None, basis=None, record=False):
        """Mean-field pre-screen (probSimulate, goSimulation/probabilitySimulation.go:53-157) for B members:
        `steps` relaxation steps from occupation 0.5.  Returns time[B], occupation[B,N] (fractional),
        electrode_occupation[B,P] (fractional), current, and with record: traffic, avg_occupation."""
        N, P, S = self.N, self.P, self.S
        V = _host(np.atleast_2d(electrode_v), np.float64)
        B = V.shape[0] if P > 0 else np.atleast_2d(E_constant).shape[0]
        kTa = _host(np.broadcast_to(np.asarray(kT, dtype=np.float64), (B,)), np.float64)
        Ec = None if E_constant is None else _host(np.atleast_2d(E_constant), np.float64)
        bs = None if basis is None else _host(basis, np.float64)
        out = dict(time=np.zeros(B), occupation=np.zeros((B, N)), electrode_occupation=np.zeros((B, P)),
                   site_energies=np.zeros((B, S)))
        if record:
            out["avg_occupation"] = np.zeros((B, N)); out["traffic"] = np.zeros((B, S, S))
        a = _lib.EnsembleArgs()
        a.B, a.hops, a.prehops, a.mode, a.flags = B, int(steps), 0, MODE_PROB, 0
        a.E_constant, a.basis, a.electrode_v, a.kT = _ptr(Ec), _ptr(bs), _ptr(V), _ptr(kTa)
        a.time, a.prob_occupation, a.prob_electrode_occ = _ptr(out["time"]), _ptr(out["occupation"]), _ptr(out["electrode_occupation"])
        a.site_energies_out = _ptr(out["site_energies"])
        a.avg_occupation, a.traffic = _ptr(out.get("avg_occupation")), _ptr(out.get("traffic"))
        if self.lib.kmcb200_run_ensemble(self._h, C.byref(a)):
            raise RuntimeError("kmcb200_run_ensemble: " + _lib.last_error())
        with np.errstate(divide="ignore", invalid="ignore"):
            out["current"] = out["electrode_occupation"] / out["time"][:, None]
        return out

    def probe_rates(self, E_constant, electrode_v, kT, occupation, site_energies=None):
        """fp32 energies and dense rate matrix of one state with the fast kernel's arithmetic."""
        N, P, S = self.N, self.P, self.S
        Ec = _host(E_constant, np.float64); V = _host(electrode_v, np.float64)
        occ = _host(np.asarray(occupation) != 0, np.uint8)
        se = np.zeros(S, dtype=np.float32) if site_energies is None else _host(site_energies, np.float32).copy()
        rates = np.zeros((S, S), dtype=np.float32)
        if self.lib.kmcb200_probe_rates(self._h, Ec.ctypes.data, V.ctypes.data if P else None, float(kT),
                                        occ.ctypes.data, se.ctypes.data, int(site_energies is not None),
                                        rates.ctypes.data):
            raise RuntimeError("kmcb200_probe_rates: " + _lib.last_error())
        return se, rates


def last_kernel():
    """Name of the hop kernel this thread's last run launched (kmcb200_last_kernel)."""
    return _lib.load().kmcb200_last_kernel().decode()


def launch_count():
    return int(_lib.load().kmcb200_launch_count())


class MultiLayout(Layout):
    """One layout replicated on several GPUs of the box, driven from ONE process: `run(...)` has Layout.run's
    signature and results (streams are numbered by global member index), the members are cut into contiguous
    blocks, one per device (kmcb200_run_ensemble_multi).  devices=None: all visible devices."""

    def __init__(self, N, P, distances, transitions_constant, nu=1.0, I_0=100.0, R=1.0, prune_threshold=0.0, devices=None):
        lib = _lib.load()
        if devices is None:
            devices = list(range(lib.kmcb200_device_count()))
        if not devices:
            raise RuntimeError("kmcb200: no CUDA device available (this library has no CPU fallback)")
        super().__init__(N, P, distances, transitions_constant, nu, I_0, R, prune_threshold, device=devices[0])
        self.devices = list(devices)
        self._others = [Layout(N, P, distances, transitions_constant, nu, I_0, R, prune_threshold, device=d) for d in devices[1:]]

    def _call(self, a):
        hs = [self._h] + [o._h for o in self._others]
        arr = (C.c_void_p * len(hs))(*hs)
        if self.lib.kmcb200_run_ensemble_multi(arr, len(hs), C.byref(a)):
            raise RuntimeError("kmcb200_run_ensemble_multi: " + _lib.last_error())

    def run_device(self, *args, **kw):
        raise RuntimeError("MultiLayout works on host buffers; use one Layout per device for device pointers")

    def close(self):
        for o in getattr(self, "_others", []):
            o.close()
        super().close()
