"""Drop-in for the reference's goSimulation/parrallelSimulationBind.py (batched ensemble API).

Same class name (sic), methods and result layout: after runSimulation() every dn has
`dn.parrallel_results` extended by (time, electrode_occupation list, current list)
(parrallelSimulationBind.py:67-79)."""
import numpy as np

from .. import _lib
from .._lib import goslice


class parrallelSimulation():
    def __init__(self):
        self.N_acceptors = []; self.N_electrodes = []
        self.nu = []; self.kT = []; self.I_0 = []; self.R = []; self.time = []; self.hops = []
        self.occupation = []; self.electrode_occupation = []; self.E_constant = []; self.site_energies = []
        self.distances = []; self.transitions_constant = []
        self.dns = []

    def addSimulation(self, dn, hops):
        """Queue one simulation of `dn` in its current state (parrallelSimulationBind.py:37-48)."""
        self.N_acceptors.append(dn.N)
        self.N_electrodes.append(len(dn.electrode_occupation))
        for name in ("nu", "kT", "I_0", "R", "time"):
            getattr(self, name).append(float(getattr(dn, name)))
        self.hops.append(hops)
        for name in ("occupation", "electrode_occupation", "E_constant", "site_energies"):
            getattr(self, name).append(np.asarray(getattr(dn, name), dtype=np.float64).ravel().copy())
        for name in ("distances", "transitions_constant"):
            a = np.asarray(getattr(dn, name), dtype=np.float64)
            if a.ndim != 2 or a.shape[0] != a.shape[1]:
                raise Exception("None uniform array")  # parrallelSimulationBind.py:34-35
            getattr(self, name).append(a.ravel().copy())
        dn.parrallel_results = []
        self.dns.append(dn)

    def runSimulation(self):
        """One call into parallelSimulations (parrallelSimulationBind.py:50-79)."""
        if not self.dns:
            return
        lib = _lib.load()
        cat = lambda xs: np.ascontiguousarray(np.concatenate(xs)) if len(xs) else np.zeros(0)  # noqa: E731
        sc = lambda xs: np.ascontiguousarray(np.asarray(xs, dtype=np.float64))  # noqa: E731
        bufs = dict(N_acceptors=sc(self.N_acceptors), N_electrodes=sc(self.N_electrodes), nu=sc(self.nu), kT=sc(self.kT),
                    I_0=sc(self.I_0), R=sc(self.R), occupation=(cat(self.occupation) != 0).astype(np.float64),
                    distances=cat(self.distances), E_constant=cat(self.E_constant),
                    transitions_constant=cat(self.transitions_constant),
                    electrode_occupation=cat(self.electrode_occupation), hops=sc(self.hops), time=sc(self.time),
                    site_energies=cat(self.site_energies))
        order = ["N_acceptors", "N_electrodes", "nu", "kT", "I_0", "R", "occupation", "distances", "E_constant",
                 "transitions_constant", "electrode_occupation", "hops", "time", "site_energies"]
        lib.parallelSimulations(*[goslice(bufs[k]) for k in order])
        eo_all = bufs["electrode_occupation"]; times = bufs["time"]
        off = 0
        for i, dn in enumerate(self.dns):
            P = len(dn.electrodes)
            eo = eo_all[off:off + P].tolist()
            with np.errstate(divide="ignore", invalid="ignore"):
                current = (np.asarray(eo) / times[i]).tolist()
            dn.parrallel_results.append((float(times[i]), eo, current))
            off += P
