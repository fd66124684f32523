#!/usr/bin/env python
"""Round-2 study (CPU, numpy): how much of a state's total rate lies in pairs beyond a static threshold on the transition constant, on
states visited by the C5 layout (256 acceptors, kT = 1, +-150 V) -- i.e. whether the UNPRUNED sweep could use neighbour lists.  Prints, per
threshold: pairs kept, an energy-aware upper bound on the dropped mass, and the true dropped mass, as fractions of the total rate."""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from kmc_dn_b200 import workloads
w = workloads.c5_scaling(); lt = w["tables"]; N,P = lt.N, lt.P; S=N+P
tc = np.asarray(lt.transitions_constant, dtype=np.float64); d = np.asarray(lt.distances, dtype=np.float64)
I0R = lt.I_0*lt.R
rng = np.random.default_rng(1)
dd = d[:N,:N].copy(); np.fill_diagonal(dd, np.inf)
ths=(1e-6,1e-8,1e-10,1e-12)
out=[]
for mem in range(6):
    V = w["V"][mem*997]; kT = w["kT"][0]
    Ec = np.asarray(lt.E_constant(V), dtype=np.float64)
    occ = np.asarray(w["occupation0"]).astype(bool).copy()
    for step in range(150):
        e = np.concatenate([Ec[:N] - I0R*(1.0/dd[:, ~occ]).sum(axis=1), V])
        occm = np.concatenate([occ, np.ones(P,bool)]); empm = np.concatenate([~occ, np.ones(P,bool)])
        shape = np.outer(occm, empm); shape[N:,N:]=False; np.fill_diagonal(shape,False)
        dE = e[None,:]-e[:,None]
        dEc = dE.copy(); dEc[:N,:N] += I0R/dd
        R = np.where(shape, tc*np.exp(-np.maximum(dEc,0)/kT), 0.0)
        tot = R.sum()
        row=[tot/tc.max()]
        emin = np.where(empm, e, np.inf).min()   # lowest target energy
        emax = np.where(occm, e, -np.inf).max()  # highest source energy
        for th in ths:
            near = tc > th*tc.max()
            far_src = (tc*(~near)*empm[None,:]).sum(axis=1)        # per source: far mass to currently-possible targets (could be static w/o empm)
            far_src_static = (tc*(~near)).sum(axis=1)
            b2 = (far_src_static*occm*np.exp(-np.maximum(emin - e,0)/kT)).sum()
            # per-target bound too: far_dst_j * exp(-(e_j - emax))
            far_dst_static = (tc*(~near)).sum(axis=0)
            b3 = (far_dst_static*empm*np.exp(-np.maximum(e - emax,0)/kT)).sum()
            row += [int((shape&near).sum()), min(b2,b3)/tot, (R*(~near)).sum()/tot]
        out.append(row)
        if step>=0:
            p=(R/tot).ravel(); k=rng.choice(len(p),p=p); i,j=divmod(k,S)
            if i<N: occ[i]=False
            if j<N: occ[j]=True
s=np.array(out)
np.set_printoptions(linewidth=220, precision=3)
print("cols: tot/maxtc, [near pairs, energy-aware bound/tot, TRUE dropped/tot] for th", ths)
for q in (50,90,99,100):
    print(q, np.percentile(s,q,axis=0))
