"""Host-side set-up of a layout (the reference's off-hot-path producers) and the BASELINE.json
benchmark configurations C1..C5 as synthetic ensembles.

Layout tables follow kmc_dn.__init__/initialize (kmc_dopant_networks.py:166-452):
  R  = (N/(xdim*ydim))**-0.5 in 2-D (:350-358), ab = a*R (:361)
  distances[i,j] = euclidean distance over acceptors then electrodes (:657-695)
  transitions_constant = nu*exp(-2*distances/ab) - I (:824-830)
  comp_constant[i] = I_0*R*sum_k 1/|r_i - r_donor_k| (:892-894)
"""
import os

import numpy as np

from .electrostatics import BasisPotentials, comp_constant

GOLDEN_LAYOUTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                              "layouts.npz")


class LayoutTables:
    """Everything the hop loop needs for one dopant layout, built without FEniCS."""

    def __init__(self, acceptors, donors, electrodes, xdim=1.0, ydim=1.0, zdim=0.0, mu=0.0, I_0=100.0, a=0.25,
                 nu=1.0, kT=1.0, res=None, static_electrodes=None):
        self.acceptors = np.asarray(acceptors, dtype=np.float64).reshape(-1, 3)
        self.donors = np.asarray(donors, dtype=np.float64).reshape(-1, 3)
        self.electrodes = np.asarray(electrodes, dtype=np.float64).reshape(-1, 4)
        self.N, self.M, self.P = len(self.acceptors), len(self.donors), len(self.electrodes)
        self.S = self.N + self.P
        self.xdim, self.ydim, self.zdim, self.mu = xdim, ydim, zdim, mu
        self.nu, self.kT, self.I_0 = nu, kT, I_0 * kT  # kmc_dopant_networks.py:336-338
        if ydim == 0 and zdim == 0:
            self.R = (self.N / xdim) ** (-1)
        elif zdim == 0:
            self.R = (self.N / (xdim * ydim)) ** (-1 / 2)
        else:
            self.R = (self.N / (xdim * ydim * zdim)) ** (-1 / 3)
        self.ab = a * self.R
        pos = np.vstack([self.acceptors, self.electrodes[:, :3]])
        diff = pos[:, None, :] - pos[None, :, :]
        self.distances = np.sqrt(diff[..., 0] ** 2 + diff[..., 1] ** 2 + diff[..., 2] ** 2)
        self.transitions_constant = self.nu * np.exp(-2 * self.distances / self.ab) - np.eye(self.S)
        self.comp_constant = comp_constant(self.acceptors, self.donors, self.I_0, self.R) if self.M else np.zeros(self.N)
        self.potentials = BasisPotentials(self.acceptors, self.electrodes, xdim, ydim, zdim, res=res,
                                          static_electrodes=static_electrodes)
        sv = None if static_electrodes is None else np.asarray(static_electrodes, dtype=np.float64).reshape(-1, 4)[:, 3]
        self.basis = self.potentials.kernel_basis(self.comp_constant, mu=mu, static_v=sv)  # [P+1, N]

    def E_constant(self, electrode_v):
        """[(B,)N] = eV_constant + comp_constant (kmc_dopant_networks.py:896)."""
        return np.asarray(electrode_v, dtype=np.float64) @ self.basis[:self.P] + self.basis[self.P]

    def initial_occupation(self, rng):
        """N-M holes placed uniformly at random (place_charges_random, :641-655)."""
        occ = np.zeros(self.N, dtype=bool)
        occ[rng.permutation(self.N)[: self.N - self.M]] = True
        return occ


def electrodes8(xdim=1.0, ydim=1.0):
    """thesis_indrek/voltage_search_tests.py:19-30 (get8Electrodes); voltages filled in by the workload."""
    e = np.zeros((8, 4))
    e[0, :2] = [0, 3 * ydim / 4]; e[1, :2] = [xdim / 4, 0]; e[2, :2] = [xdim, ydim / 4]; e[3, :2] = [xdim, 3 * ydim / 4]
    e[4, :2] = [0, ydim / 4]; e[5, :2] = [3 * xdim / 4, 0]; e[6, :2] = [xdim / 4, ydim]; e[7, :2] = [3 * xdim / 4, ydim]
    return e


def _golden_layout(k=0):
    z = np.load(GOLDEN_LAYOUTS)
    return z["acceptor_layouts"][k], z["donor_layouts"][k]


def c1_basic(B=4096):
    """examples/basic.py:16-25: N=10, M=0, P=2 at (0,.5),(1,.5), V=(+10,-10); all-empty start."""
    rng = np.random.default_rng(0)
    acc = np.zeros((10, 3)); acc[:, :2] = rng.random((10, 2))
    el = np.zeros((2, 4)); el[0] = [0, 0.5, 0, 10]; el[1] = [1, 0.5, 0, -10]
    lt = LayoutTables(acc, np.zeros((0, 3)), el)
    V = np.tile(el[:, 3], (B, 1))
    return dict(name="C1 basic N=10 P=2", tables=lt, V=V, kT=np.ones(B), occupation0=None, prehops=0, hops=100000)


def c2_grid4x4(seeds=64):
    """experiments/grid4x4/grid4x4.py:26-39,59: 4x4 acceptor grid, 3x3 donors, 8 electrodes, IV sweep of
    electrode 0 over linspace(-100,100,100), others 0; prehops 1e3."""
    acc = np.zeros((16, 3)); don = np.zeros((9, 3))
    for i in range(4):
        for j in range(4):
            acc[4 * i + j, :2] = [(i + 1) / 5, (j + 1) / 5]
    for i in range(3):
        for j in range(3):
            don[3 * i + j, :2] = [i / 5 + 0.3, j / 5 + 0.3]
    el = np.zeros((8, 4))
    el[0, :2] = [0, 0.25]; el[1, :2] = [0, 0.75]; el[2, :2] = [1, 0.25]; el[3, :2] = [1, 0.75]
    el[4, :2] = [0.25, 0]; el[5, :2] = [0.75, 0]; el[6, :2] = [0.25, 1]; el[7, :2] = [0.75, 1]
    lt = LayoutTables(acc, don, el)
    sweep = np.linspace(-100, 100, 100)
    V = np.zeros((100 * seeds, 8)); V[:, 0] = np.repeat(sweep, seeds)
    occ = lt.initial_occupation(np.random.default_rng(1))
    return dict(name="C2 grid4x4 IV N=16 P=8", tables=lt, V=V, kT=np.ones(len(V)), occupation0=occ, prehops=1000,
                hops=100000)


def c3_voltage_search(n_controls=16384, seeds=16, hops=100000):
    """boolean_logic / voltage_search: reference layout 0 (30 acceptors / 3 donors), electrodes of
    voltage_search_tests.py:19-30, inputs on electrodes 0,1 in {0,75}^2 (:159), controls 2..6 ~ U(-150,150) (:76),
    electrode 7 = 0; n_controls x 4 inputs x seeds members."""
    acc, don = _golden_layout(0)
    lt = LayoutTables(acc, don, electrodes8())
    rng = np.random.default_rng(2026)
    ctrl = rng.uniform(-150, 150, size=(n_controls, 5))
    inputs = np.array([[0, 0], [0, 75], [75, 0], [75, 75]], dtype=np.float64)
    V = np.zeros((n_controls, 4, 8))
    V[:, :, 0:2] = inputs[None, :, :]
    V[:, :, 2:7] = ctrl[:, None, :]
    V = np.repeat(V.reshape(-1, 8), seeds, axis=0)  # member index = ((control*4 + input)*seeds + seed)
    occ = lt.initial_occupation(np.random.default_rng(3))
    return dict(name=f"C3 voltage_search N=30 P=8 B={len(V)}", tables=lt, V=V, kT=np.ones(len(V)), occupation0=occ,
                prehops=0, hops=hops)


def c4_temperature(n_T=64, seeds=1024):
    """experiments/temperature_dependence: same 30/3 layout, P=2 at (0,.5),(1,.5), V=(15,0),
    kT = logspace(0,1,n_T), prehops=1e5."""
    acc, don = _golden_layout(0)
    el = np.zeros((2, 4)); el[0] = [0, 0.5, 0, 15]; el[1] = [1, 0.5, 0, 0]
    lt = LayoutTables(acc, don, el)
    kT = np.repeat(np.logspace(0, 1, n_T), seeds)
    V = np.tile(el[:, 3], (len(kT), 1))
    occ = lt.initial_occupation(np.random.default_rng(4))
    return dict(name="C4 temperature N=30 P=2", tables=lt, V=V, kT=kT, occupation0=occ, prehops=100000, hops=100000)


def c5_scaling(N=256, M=25, B=8192):
    """examples/scaling.py: large uniform-random layout, 8 electrodes, U(-150,150) voltages."""
    rng = np.random.default_rng(0)
    acc = np.zeros((N, 3)); acc[:, :2] = rng.random((N, 2))
    don = np.zeros((M, 3)); don[:, :2] = rng.random((M, 2))
    lt = LayoutTables(acc, don, electrodes8())
    V = rng.uniform(-150, 150, size=(B, 8))
    occ = lt.initial_occupation(np.random.default_rng(5))
    return dict(name=f"C5 scaling N={N} P=8", tables=lt, V=V, kT=np.ones(B), occupation0=occ, prehops=0, hops=10000)
