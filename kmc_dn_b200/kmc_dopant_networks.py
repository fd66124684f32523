"""Host class `kmc_dn`: same constructor, attributes and simulation entry points as the reference's
kmc_dopant_networks.kmc_dn (reference kmc_dopant_networks.py:165-1006), with the hop loop running on
the B200 through libkmcb200.so and the electrostatics done once per layout by superposition.

What is mirrored (reference file:line in each docstring): __init__/initialize/reset, go_simulation,
python_simulation, makeSimulation, update_V, calc_E_constant_V(_comp), calc_distances,
calc_transitions_constant, place_*_random, load_acceptors/donors, saveSelf/loadSelf, total_energy, dist.
What is new: `ensemble_simulation` (many voltage vectors x temperatures x seeds in one launch).
Not carried over: the unfinished Tsigankov stubs (:1011-1081, dead code in the reference).
"""
import os

import numpy as np

from . import fixtures
from .electrostatics import BasisPotentials
from .goSimulation.pythonBind import callGoSimulation


_LAYOUTS = None  # small LRU of device layouts keyed by content: python_simulation / prehops call the loop again and again
                 # with the same tables (kmc_dopant_networks.py:580-585) -- no cudaMalloc / upload per call


def _cached_layout(N, P, distances, transitions_constant, nu, I_0, R, keep=8):
    import collections
    import hashlib
    from .ensemble import Layout
    global _LAYOUTS
    if _LAYOUTS is None:
        _LAYOUTS = collections.OrderedDict()
    d = np.ascontiguousarray(distances, dtype=np.float64); tc = np.ascontiguousarray(transitions_constant, dtype=np.float64)
    h = hashlib.blake2b(digest_size=16)
    h.update(d.tobytes()); h.update(tc.tobytes())
    key = (int(N), int(P), float(nu), float(I_0), float(R), h.digest())
    lay = _LAYOUTS.pop(key, None)
    if lay is None:
        lay = Layout(N, P, d, tc, nu=nu, I_0=I_0, R=R)
        while len(_LAYOUTS) >= keep:
            _LAYOUTS.popitem(last=False)[1].close()
    _LAYOUTS[key] = lay  # most recently used last
    return lay


def _simulate_discrete_record(N_acceptors, N_electrodes, nu, kT, I_0, R, time, occupation, distances, E_constant,
                              site_energies, transitions_constant, transitions, problist, electrode_occupation,
                              hops, record=False, prehops=0):
    """GPU stand-in for the reference's numba loop of the same name (kmc_dopant_networks.py:33-135):
    same arguments, same 5-tuple result, fp64 numba arithmetic replayed op for op on the device
    (KMCB200_MODE_PY).  The random stream is numpy's global MT19937 -- the generator the numba loop
    draws from -- consumed as (dwell, pick) per hop, so `np.random.seed(s)` makes a run reproducible."""
    from .ensemble import MODE_PY
    hops = int(hops); prehops = int(prehops)
    N, P = int(N_acceptors), int(N_electrodes)
    u = np.random.random_sample(2 * (hops + prehops))
    lay = _cached_layout(N, P, distances, transitions_constant, nu, I_0, R)
    r = lay.run(hops, kT, np.asarray(site_energies, dtype=np.float64)[None, N:], prehops=prehops,
                E_constant=np.asarray(E_constant, dtype=np.float64)[None, :], mode=MODE_PY,
                occupation0=np.asarray(occupation)[None, :], stream_u64=u, want_occupation=True,
                want_site_energies=True, record=True)
    occupation[:] = r["occupation"][0]
    site_energies[:N] = r["site_energies"][0][:N]
    eo = np.asarray(electrode_occupation) + r["electrode_occupation"][0]  # the loop accumulates onto its input (:119,:123)
    return r["time"][0], occupation, eo, r["traffic"][0], r["avg_occupation"][0]


class kmc_dn():
    def __init__(self, N, M, xdim, ydim, zdim, mu=0, I_0=100, a=0.25, **kwargs):
        """Same signature and keyword arguments as the reference (kmc_dopant_networks.py:166-405):
        electrodes, static_electrodes (Px4: x, y, z, voltage), acceptors, donors, res,
        calc_E_constant ('calc_E_constant_V' | 'calc_E_constant_V_comp'), copy_from."""
        src = kwargs.get("copy_from")
        if src is not None:
            self.nu, self.kT, self.I_0 = src.nu, src.kT, src.I_0
        else:
            self.nu = 1
            self.kT = 1
            self.I_0 = I_0 * self.kT
        self.time = 0
        self.mu = mu
        self.N, self.M = N, M
        self.xdim, self.ydim, self.zdim = xdim, ydim, zdim
        if ydim == 0 and zdim == 0:
            self.dim = 1
            self.R = (N / xdim) ** (-1)
        elif zdim == 0:
            self.dim = 2
            self.R = (N / (xdim * ydim)) ** (-1 / 2)
        else:
            self.dim = 3
            self.R = (N / (xdim * ydim * zdim)) ** (-1 / 3)
        self.ab = a * self.R
        self.electrodes = kwargs["electrodes"].copy() if "electrodes" in kwargs else np.zeros((0, 4))
        self.P = self.electrodes.shape[0]
        if "acceptors" in kwargs:
            self.acceptors = kwargs["acceptors"].copy()
        if "donors" in kwargs:
            self.donors = kwargs["donors"].copy()
        self.static_electrodes = (kwargs["static_electrodes"].copy() if "static_electrodes" in kwargs
                                  else np.zeros((0, 4)))
        if "res" in kwargs:
            self.res = kwargs["res"]
        else:
            self.res = {1: xdim, 2: min([xdim, ydim]), 3: min([xdim, ydim, zdim])}[self.dim] / 100
        self.calc_E_constant = (self.calc_E_constant_V if kwargs.get("calc_E_constant") == "calc_E_constant_V"
                                else self.calc_E_constant_V_comp)
        self.initialize(dopant_placement=not hasattr(self, "acceptors"), charge_placement=not hasattr(self, "donors"))

    # ------------------------------------------------------------------ set-up (off the hot path)
    def initialize(self, dopant_placement=True, charge_placement=True, distances=True, V=True, E_constant=True):
        """kmc_dopant_networks.py:408-452: allocate state, place dopants/charges, tables, potential, energies."""
        S = self.N + self.P
        self.transitions = np.zeros((S, S))
        self.transitions_constant = np.zeros((S, S))
        self.distances = np.zeros((S, S))
        self.vectors = np.zeros((S, S, 3))
        self.site_energies = np.zeros((S,))
        self.problist = np.zeros(S ** 2)
        self.occupation = np.zeros(self.N, dtype=bool)
        self.electrode_occupation = np.zeros(self.P, dtype=int)
        if dopant_placement:
            self.place_dopants_random()
        if charge_placement:
            self.place_charges_random()
        if distances:
            self.calc_distances()
            self.calc_transitions_constant()
        if V:
            self.init_V()
        if E_constant:
            self.calc_E_constant()

    def reset(self):
        """kmc_dopant_networks.py:454-471: time, counters and electrode tallies restart; occupation is kept."""
        self.time = 0
        self.old_current = 0
        self.counter = 0
        self.electrode_occupation = np.zeros(self.P, dtype=int)

    def place_dopants_random(self):
        """kmc_dopant_networks.py:621-639 (same draw order: acceptors then donors)."""
        self.acceptors = np.random.rand(self.N, 3) * np.array([self.xdim, self.ydim, self.zdim])
        self.donors = np.random.rand(self.M, 3) * np.array([self.xdim, self.ydim, self.zdim])

    def place_charges_random(self):
        """kmc_dopant_networks.py:641-655: N-M holes by rejection sampling of sites."""
        self.occupation = np.zeros(self.N, dtype=bool)
        placed = 0
        while placed < self.N - self.M:
            trial = np.random.randint(self.N)
            if not self.occupation[trial]:
                self.occupation[trial] = True
                placed += 1

    def calc_distances(self):
        """kmc_dopant_networks.py:657-695: pairwise distances and unit vectors over acceptors then electrodes."""
        pos = np.vstack([self.acceptors[:, :3], self.electrodes[:, :3]])
        diff = pos[None, :, :] - pos[:, None, :]  # vectors[i,j] points from i to j
        self.distances = np.sqrt(diff[..., 0] ** 2 + diff[..., 1] ** 2 + diff[..., 2] ** 2)
        with np.errstate(invalid="ignore", divide="ignore"):
            self.vectors = np.where(self.distances[..., None] > 0, diff / self.distances[..., None], 0.0)

    def calc_transitions_constant(self):
        """kmc_dopant_networks.py:824-830."""
        self.transitions_constant = self.nu * np.exp(-2 * self.distances / self.ab)
        self.transitions_constant -= np.eye(self.transitions.shape[0])

    def update_electrodes(self, electrodes):
        """kmc_dopant_networks.py:701-704."""
        self.electrodes = electrodes
        self.P = self.electrodes.shape[0]
        self.initialize(dopant_placement=False, charge_placement=False)

    # ------------------------------------------------------------------ electrostatics by superposition
    def init_V(self):
        """Stands in for the FEniCS set-up + first solve (kmc_dopant_networks.py:706-796): solves ONE Laplace
        problem per electrode (plus the background) and keeps the basis potentials at the acceptors."""
        self.potentials = BasisPotentials(self.acceptors, self.electrodes, self.xdim, self.ydim, self.zdim,
                                          res=self.res, static_electrodes=self.static_electrodes)
        self.V = self._V_callable

    def _V_callable(self, x, y=0.0):
        """V(x[,y]) like the FEniCS function object the reference exposes (kmc_dopant_networks.py:795)."""
        sv = self.static_electrodes[:, 3] if self.static_electrodes.shape[0] else None
        return self.potentials.potential_at(x, y, self.electrodes[:, 3], mu=self.mu, static_v=sv)

    def _eV(self):
        sv = self.static_electrodes[:, 3] if self.static_electrodes.shape[0] else None
        return self.potentials.eV_constant(self.electrodes[:, 3], mu=self.mu, static_v=sv)

    def update_V(self):
        """kmc_dopant_networks.py:798-821: call after changing electrode voltages.  A mat-vec here, a FEM solve there."""
        self.calc_E_constant()

    def calc_E_constant_V(self):
        """kmc_dopant_networks.py:833-863."""
        self.eV_constant = self._eV()
        self.comp_constant = np.zeros((self.N,))
        self.E_constant = self.eV_constant
        self.site_energies[self.N:] = self.electrodes[:, 3]

    def calc_E_constant_V_comp(self):
        """kmc_dopant_networks.py:865-899."""
        self.eV_constant = self._eV()
        self.comp_constant = np.zeros((self.N,))
        for i in range(self.N):
            self.comp_constant[i] += self.I_0 * self.R * sum(1 / self.dist(self.acceptors[i], self.donors[k])
                                                             for k in range(self.M))
        self.E_constant = self.eV_constant + self.comp_constant
        self.site_energies[self.N:] = self.electrodes[:, 3]

    # ------------------------------------------------------------------ simulation entry points
    def go_simulation(self, hops=1E5, prehops=0, goSpecificFunction="wrapperSimulateRecord", record=False,
                      prune_threshold=0):
        """kmc_dopant_networks.py:473-511.  Sets time, occupation, electrode_occupation, current
        (+ traffic, average_occupation when record)."""
        self.makeSimulation(simulateFunction=callGoSimulation, preHopFunction=callGoSimulation, hops=hops,
                            prehops=prehops, goSpecificFunction=goSpecificFunction, record=record,
                            prune_threshold=prune_threshold)

    def python_simulation(self, hops=1E5, prehops=0, record=False):
        """kmc_dopant_networks.py:513-542 (fp64 numba semantics, replayed on the GPU)."""
        self.makeSimulation(simulateFunction=_simulate_discrete_record, preHopFunction=_simulate_discrete_record,
                            hops=hops, prehops=prehops, record=record)

    def makeSimulation(self, simulateFunction=None, preHopFunction=None, prehops=0, hops=1E5, record=False,
                       goSpecificFunction=None, prune_threshold=0.0):
        """kmc_dopant_networks.py:544-618, same argument dict handed to simulate_func."""
        if simulateFunction is not None:
            self.simulate_func = simulateFunction
        self.simulate_prehop = preHopFunction if preHopFunction is not None else _simulate_discrete_record
        self.reset()
        args = {"N_acceptors": self.N, "N_electrodes": self.P, "nu": self.nu, "kT": self.kT, "I_0": self.I_0,
                "R": self.R, "time": self.time, "occupation": self.occupation, "distances": self.distances,
                "E_constant": self.E_constant, "site_energies": self.site_energies,
                "transitions_constant": self.transitions_constant, "transitions": self.transitions,
                "problist": self.problist, "electrode_occupation": self.electrode_occupation, "record": False}
        if prehops != 0:  # always the fp64 loop, as in the reference (:580-585)
            args["hops"] = int(prehops)
            _, self.occupation, _, _, _ = _simulate_discrete_record(**args)
            self.reset()
            args["electrode_occupation"] = self.electrode_occupation
            args["occupation"] = self.occupation
        if goSpecificFunction is not None:
            args["goSpecificFunction"] = goSpecificFunction
            args["prune_threshold"] = prune_threshold
        args["hops"] = hops
        if record:
            args["record"] = True
            (self.time, self.occupation, self.electrode_occupation, self.traffic,
             occupations_in_time) = self.simulate_func(**args)
            self.average_occupation = [x / self.time for x in occupations_in_time]
        elif self.simulate_func == _simulate_discrete_record:
            self.time, self.occupation, self.electrode_occupation, _, _ = self.simulate_func(**args)
        else:
            self.time, self.occupation, self.electrode_occupation = self.simulate_func(**args)
        self.current = self.electrode_occupation / self.time

    def ensemble_simulation(self, voltages, hops=100000, prehops=0, kT=None, seeds=1, seed=0, device=0,
                            honour_occupation=True):
        """NEW (not in the reference): every row of `voltages` [B,P] (x `seeds` repeats) is one trajectory of this
        layout; E_constant comes from the basis potentials on the device.  Returns dict(time[B*seeds],
        electrode_occupation[B*seeds,P], current[B*seeds,P])."""
        from .ensemble import Layout
        V = np.repeat(np.atleast_2d(np.asarray(voltages, dtype=np.float64)), seeds, axis=0)
        kTs = np.full(len(V), float(self.kT)) if kT is None else np.repeat(np.broadcast_to(kT, (len(V) // seeds,)), seeds)
        sv = self.static_electrodes[:, 3] if self.static_electrodes.shape[0] else None
        comp = getattr(self, "comp_constant", None) if self.calc_E_constant == self.calc_E_constant_V_comp else None
        if comp is None or not np.any(comp):
            self.calc_E_constant()
            comp = self.comp_constant
        basis = self.potentials.kernel_basis(comp, mu=self.mu, static_v=sv)
        lay = Layout(self.N, self.P, self.distances, self.transitions_constant, nu=self.nu, I_0=self.I_0, R=self.R,
                     device=device)
        try:
            return lay.run(int(hops), kTs, V, basis=basis, prehops=int(prehops), seed=seed,
                           occupation0=self.occupation if honour_occupation else None)
        finally:
            lay.close()

    # ------------------------------------------------------------------ load / save
    def load_acceptors(self, acceptors):
        """kmc_dopant_networks.py:903-928 (recomputes R and sets ab = R)."""
        self.acceptors = acceptors
        self.N = self.acceptors.shape[0]
        if self.ydim == 0 and self.zdim == 0:
            self.R = (self.N / self.xdim) ** (-1)
        elif self.zdim == 0:
            self.R = (self.N / (self.xdim * self.ydim)) ** (-1 / 2)
        else:
            self.R = (self.N / (self.xdim * self.ydim * self.zdim)) ** (-1 / 3)
        self.ab = self.R
        self.initialize(V=False, dopant_placement=False)

    def load_donors(self, donors):
        """kmc_dopant_networks.py:930-941."""
        self.donors = donors
        self.M = self.donors.shape[0]
        self.initialize(V=False, dopant_placement=False)

    def saveSelf(self, fileName, rel_path=False):
        """kmc_dopant_networks.py:943-963: `.kmc` = pickle of every list/tuple/int/float/ndarray attribute."""
        if hasattr(self, "current"):
            self.expected_current = self.current
        path = os.path.join(os.path.dirname(__file__), fileName) if rel_path else fileName
        fixtures.save_kmc(path, {k: getattr(self, k) for k in dir(self) if not k.startswith("__")
                                 and isinstance(getattr(self, k), (list, tuple, int, float, np.ndarray))})

    def loadSelf(self, fileName, rel_path=False):
        """kmc_dopant_networks.py:965-978; reads files written by the reference too (restricted unpickler)."""
        path = os.path.join(os.path.dirname(__file__), fileName) if rel_path else fileName
        for key, val in fixtures.load_kmc(path).items():
            setattr(self, key, val)
        self.initialize(dopant_placement=False, charge_placement=False)

    # ------------------------------------------------------------------ misc
    def total_energy(self):
        """kmc_dopant_networks.py:982-1001: Coulomb sum over ionised acceptor pairs minus the electrostatic term."""
        ion = 1 - self.occupation.astype(np.float64)
        H = 0.0
        for i in range(self.N - 1):
            for j in range(i + 1, self.N):
                H += ion[i] * ion[j] / self.distances[i, j]
        H *= self.I_0 * self.R
        for i in range(self.N):
            H = H - ion[i] * self.eV_constant[i]
        return H

    @staticmethod
    def dist(ri, rj):
        """kmc_dopant_networks.py:1003-1006."""
        return np.sqrt((ri[0] - rj[0]) ** 2 + (ri[1] - rj[1]) ** 2 + (ri[2] - rj[2]) ** 2)
