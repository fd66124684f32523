"""Electrostatics front end: per-electrode BASIS potentials at the acceptor sites, once per layout.

The reference solves Laplace's equation with FEniCS for every voltage vector
(kmc_dopant_networks.py:706-821) and samples it at the acceptors (:865-899).  The problem is linear
in the electrode voltages, so

    eV_constant[i] = mu * phi_bg(r_i) + sum_p V_p * phi_p(r_i)          (superposition)

with phi_p the solution for electrode p at 1 and everything else at 0.  `BasisPotentials` computes
phi once per layout; E_constant for a whole ensemble of voltage vectors is then the mat-vec the
hop kernel fuses into its trajectory set-up (kmcb200_ensemble_args.basis).

Discretisation = the reference's: P1 elements on DOLFIN's RectangleMesh(nx, ny) with
nx = int(xdim // res) (:774-777), every boundary node Dirichlet (:697-699, :783-785), boundary values
from the rule of :741-763 (first matching electrode wins, else mu).  On that right-triangulated
uniform grid the P1 stiffness matrix IS the 5-point finite-difference Laplacian, and V(x,y) is the
P1 interpolant on cells split along the bottom-left -> top-right ("right") diagonal.  Stored
eV_constant of the reference's fixtures is reproduced to ~1e-11 (tests/test_electrostatics.py).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def _boundary_owner_2d(xs, ys, electrodes, static_electrodes, xdim):
    """owner[ix,iy] for boundary nodes: index into concat(electrodes, static_electrodes) or -1 (mu).
    Mirrors the C++ expression string built at kmc_dopant_networks.py:741-763."""
    surplus = xdim / 10
    nxp, nyp = len(xs), len(ys)
    owner = np.full((nxp, nyp), -2, dtype=np.int64)  # -2 = interior
    allel = np.vstack([electrodes[:, :4], static_electrodes[:, :4]]) if len(static_electrodes) else electrodes[:, :4]
    for ix in range(nxp):
        for iy in range(nyp):
            if not (ix == 0 or iy == 0 or ix == nxp - 1 or iy == nyp - 1):
                continue
            x, y = xs[ix], ys[iy]
            own = -1
            for k in range(allel.shape[0]):
                ex, ey = allel[k, 0], allel[k, 1]
                if ex == 0 or ex == xdim:
                    hit = (x == ex) and (y >= ey - surplus) and (y <= ey + surplus)
                else:
                    hit = (x >= ex - surplus) and (x <= ex + surplus) and (y == ey)
                if hit:
                    own = k
                    break
            owner[ix, iy] = own
    return owner


class BasisPotentials:
    """phi[(P + n_static + 1), N]: rows 0..P-1 electrodes, then static electrodes, last row the
    background (non-electrode boundary at 1)."""

    def __init__(self, acceptors, electrodes, xdim, ydim=0.0, zdim=0.0, res=None, static_electrodes=None):
        acceptors = np.asarray(acceptors, dtype=np.float64)
        electrodes = np.asarray(electrodes, dtype=np.float64).reshape(-1, 4)
        static_electrodes = (np.zeros((0, 4)) if static_electrodes is None
                             else np.asarray(static_electrodes, dtype=np.float64).reshape(-1, 4))
        self.P, self.n_static, self.N = electrodes.shape[0], static_electrodes.shape[0], acceptors.shape[0]
        if ydim == 0 and zdim == 0:
            self.dim = 1
        elif zdim == 0:
            self.dim = 2
        else:
            raise NotImplementedError("3-D electrostatics: the reference's own 3-D path is unfinished "
                                      "(kmc_dopant_networks.py:852-858 indexes an undefined array)")
        if res is None:
            res = (xdim if self.dim == 1 else min(xdim, ydim)) / 100  # kmc_dopant_networks.py:387-390
        self.res = res
        if self.dim == 1:
            self.phi = self._solve_1d(acceptors, electrodes, static_electrodes, xdim)
        else:
            self.phi = self._solve_2d(acceptors, electrodes, static_electrodes, xdim, ydim, res)

    # -- 1-D: only the two end points are boundary nodes, so V is the straight line between them
    def _solve_1d(self, acceptors, electrodes, static_electrodes, xdim):
        allel = np.vstack([electrodes, static_electrodes])
        K = allel.shape[0]
        phi = np.zeros((K + 1, self.N))
        ends = []
        for xb in (0.0, float(xdim)):
            own = K  # background
            for k in range(K):
                if allel[k, 0] == xb:  # 'x[0] == e_x ? e : ...' (:736-739), first match wins
                    own = k
                    break
            ends.append(own)
        t = acceptors[:, 0] / xdim
        phi[ends[0]] += 1.0 - t
        phi[ends[1]] += t
        self._grid = ("1d", ends, xdim, K)
        return phi

    def _solve_2d(self, acceptors, electrodes, static_electrodes, xdim, ydim, res):
        nx, ny = int(xdim // res), int(ydim // res)
        xs = np.array([0.0 + (xdim - 0.0) * i / nx for i in range(nx + 1)])
        ys = np.array([0.0 + (ydim - 0.0) * j / ny for j in range(ny + 1)])
        hx, hy = xdim / nx, ydim / ny
        owner = _boundary_owner_2d(xs, ys, electrodes, static_electrodes, xdim)
        K = self.P + self.n_static
        interior = owner == -2
        idx = -np.ones(owner.shape, dtype=np.int64)
        idx[interior] = np.arange(interior.sum())
        n_int = int(interior.sum())
        # 5-point Laplacian on interior nodes; Dirichlet neighbours go to the right-hand sides
        rows, cols, vals = [], [], []
        rhs = np.zeros((n_int, K + 1))
        wx, wy = 1.0 / hx ** 2, 1.0 / hy ** 2
        ii, jj = np.nonzero(interior)
        for ix, iy in zip(ii, jj):
            r = idx[ix, iy]
            rows.append(r); cols.append(r); vals.append(2 * wx + 2 * wy)
            for dx, dy, w in ((-1, 0, wx), (1, 0, wx), (0, -1, wy), (0, 1, wy)):
                jx, jy = ix + dx, iy + dy
                if interior[jx, jy]:
                    rows.append(r); cols.append(idx[jx, jy]); vals.append(-w)
                else:
                    own = owner[jx, jy]
                    rhs[r, K if own == -1 else own] += w
        A = sp.csc_matrix((vals, (rows, cols)), shape=(n_int, n_int))
        sol = spla.splu(A).solve(rhs)
        # full nodal fields
        U = np.zeros((K + 1, nx + 1, ny + 1))
        for k in range(K + 1):
            U[k][interior] = sol[:, k]
            U[k][owner == (-1 if k == K else k)] = 1.0
        self._grid = ("2d", xs, ys, hx, hy, U)
        phi = np.zeros((K + 1, self.N))
        for a in range(self.N):
            phi[:, a] = self.basis_at(acceptors[a, 0], acceptors[a, 1])
        return phi

    def basis_at(self, x, y=0.0):
        """All basis potentials at an arbitrary point: P1 interpolation, cells split along the "right"
        diagonal (bottom-left -> top-right, DOLFIN's RectangleMesh default)."""
        if self._grid[0] == "1d":
            _, ends, xdim, K = self._grid
            out = np.zeros(K + 1)
            out[ends[0]] += 1.0 - x / xdim
            out[ends[1]] += x / xdim
            return out
        _, xs, ys, hx, hy, U = self._grid
        nx, ny = len(xs) - 1, len(ys) - 1
        cx = min(max(int(np.floor(x / hx)), 0), nx - 1)
        cy = min(max(int(np.floor(y / hy)), 0), ny - 1)
        s = (x - xs[cx]) / hx
        t = (y - ys[cy]) / hy
        v00, v10, v01, v11 = U[:, cx, cy], U[:, cx + 1, cy], U[:, cx, cy + 1], U[:, cx + 1, cy + 1]
        if s >= t:
            return v00 + s * (v10 - v00) + t * (v11 - v10)
        return v00 + t * (v01 - v00) + s * (v11 - v01)

    def potential_at(self, x, y, electrode_v, mu=0.0, static_v=None):
        """V(x,y) for given electrode voltages -- what the reference's FEniCS function object returns."""
        b = self.basis_at(x, y)
        v = float(np.dot(np.asarray(electrode_v, dtype=np.float64), b[:self.P]) + mu * b[-1])
        if self.n_static:
            v += float(np.dot(np.asarray(static_v, dtype=np.float64), b[self.P:self.P + self.n_static]))
        return v

    # ------------------------------------------------------------------
    def eV_constant(self, electrode_v, mu=0.0, static_v=None):
        """eV_constant[(B,)N] for electrode voltages [(B,)P] (kmc_dopant_networks.py:876-882)."""
        V = np.asarray(electrode_v, dtype=np.float64)
        out = V @ self.phi[:self.P] + mu * self.phi[-1]
        if self.n_static:
            out = out + np.asarray(static_v, dtype=np.float64) @ self.phi[self.P:self.P + self.n_static]
        return out

    def kernel_basis(self, comp_constant=None, mu=0.0, static_v=None):
        """[P+1,N] array for kmcb200_ensemble_args.basis: rows 0..P-1 = phi_p, row P = everything that
        does not depend on the swept voltages (compensation term, background, static electrodes)."""
        const = mu * self.phi[-1]
        if self.n_static:
            const = const + np.asarray(static_v, dtype=np.float64) @ self.phi[self.P:self.P + self.n_static]
        if comp_constant is not None:
            const = const + np.asarray(comp_constant, dtype=np.float64)
        return np.vstack([self.phi[:self.P], const[None, :]])


def comp_constant(acceptors, donors, I_0, R):
    """I_0*R*sum_k 1/|r_i - r_donor_k|  (kmc_dopant_networks.py:892-894)."""
    acceptors = np.asarray(acceptors, dtype=np.float64); donors = np.asarray(donors, dtype=np.float64)
    out = np.zeros(acceptors.shape[0])
    for i in range(acceptors.shape[0]):
        out[i] = I_0 * R * sum(1 / np.sqrt(((acceptors[i] - donors[k]) ** 2).sum()) for k in range(donors.shape[0]))
    return out
