"""The drop-in boundary exercised with the reference's LITERAL calling convention:
the shared library sits at ./goSimulation/libSimulation.so relative to the CWD and is re-loaded on every call
(goSimulation/pythonBind.py:64, parrallelSimulationBind.py:59), every array is boxed element by element into c_double
objects (pythonBind.py:6-25), the scalar arguments are FIVE c_double (nu, kT, I_0, R, time) and a 32-bit c_int `hops`
(pythonBind.py:65-72), GoSlices travel by value.  The caller below is a restatement of that sequence (the reference tree
is not on the GPU box); it runs in a fresh interpreter whose CWD is a scratch directory holding a copy of libkmcb200.so
under the reference's file name."""
import json
import os
import shutil
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CALLER = textwrap.dedent('''
    import json, sys
    from ctypes import *
    import numpy as np

    class GoSlice(Structure):                                   # pythonBind.py:28-30
        _fields_ = [("data", POINTER(c_double)), ("len", c_longlong), ("cap", c_longlong)]

    def flattenDouble(arr2):                                    # pythonBind.py:6-15
        arr = []
        for row in arr2:
            for ele in row:
                arr.append(c_double(np.float64(ele).item()))
        return (c_double * len(arr))(*arr), len(arr2[0]), len(arr)

    def getGoSlice(arr):                                        # pythonBind.py:17-25
        rArr = [c_double(ele) for ele in arr]
        return GoSlice((c_double * len(rArr))(*rArr), len(rArr), len(rArr))

    def values(s):
        return [s.data[i] for i in range(s.len)]

    def single(c, hops, func, record=False, prune_threshold=0.0):   # callGoSimulation, pythonBind.py:49-90
        N = c["N"] + c["P"]
        d, _, s = flattenDouble(c["distances"]); tc, _, ts = flattenDouble(c["transitions_constant"])
        d = GoSlice(d, s, s); tc = GoSlice(tc, ts, ts)
        occ = getGoSlice(c["occupation"]); Ec = getGoSlice(c["E_constant"]); se = getGoSlice(c["site_energies"])
        traffic = getGoSlice(np.zeros(N * N)); avg = getGoSlice(np.zeros(c["N"])); eo = getGoSlice(np.zeros(c["P"]))
        lib = cdll.LoadLibrary("./goSimulation/libSimulation.so")
        f = getattr(lib, func)
        f.argtypes = [c_longlong, c_longlong, c_double, c_double, c_double, c_double, c_double,
                      GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, GoSlice, c_int, c_bool, GoSlice, GoSlice]
        f.restype = c_double
        args = [c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], 0.0, occ, d, Ec, tc, eo, se, hops, record, traffic, avg]
        if func == "wrapperSimulatePruned":
            f.argtypes = [c_longlong, c_longlong, c_double] + f.argtypes[2:]
            args = args[:2] + [prune_threshold] + args[2:]
        t = f(*args)
        return dict(time=t, eo=values(eo), occupation=values(occ), traffic=values(traffic), avg=values(avg))

    def batched(cs, hops):                                      # parrallelSimulationBind.py:34-79
        L = dict(N=[], P=[], nu=[], kT=[], I_0=[], R=[], time=[], occupation=[], distances=[], E_constant=[],
                 transitions_constant=[], electrode_occupation=[], site_energies=[], hops=[])
        for c in cs:
            L["N"].append(c["N"]); L["P"].append(c["P"])
            for k in ("nu", "kT", "I_0", "R"):
                L[k].append(c[k])
            L["time"].append(0.0); L["hops"].append(hops)
            L["occupation"].extend(1 if o else 0 for o in c["occupation"])
            L["electrode_occupation"].extend([0.0] * c["P"])
            L["E_constant"].extend(c["E_constant"]); L["site_energies"].extend(c["site_energies"])
            for k in ("distances", "transitions_constant"):
                L[k].extend(e for row in c[k] for e in row)
        go = {k: getGoSlice(v) for k, v in L.items()}
        lib = cdll.LoadLibrary("./goSimulation/libSimulation.so")
        lib.parallelSimulations.argtypes = [GoSlice] * 14
        lib.parallelSimulations.restype = c_longlong
        done = lib.parallelSimulations(go["N"], go["P"], go["nu"], go["kT"], go["I_0"], go["R"], go["occupation"], go["distances"],
                                       go["E_constant"], go["transitions_constant"], go["electrode_occupation"], go["hops"],
                                       go["time"], go["site_energies"])
        return dict(done=done, time=values(go["time"]), eo=values(go["electrode_occupation"]))

    job = json.load(open(sys.argv[1]))
    c = job["case"]
    out = dict(single=[single(c, job["hops"], "wrapperSimulate") for _ in range(job["calls"])],
               record=single(c, job["hops"], "wrapperSimulateRecord", record=True),
               recordplus=single(c, job["hops"], "wrapperSimulateRecordPlus"),
               pruned=single(c, job["hops"], "wrapperSimulatePruned", prune_threshold=1e-7),
               batched=batched([c] * job["calls"], job["hops"]))
    json.dump(out, open(sys.argv[2], "w"))
''')


def test_reference_calling_convention(tmp_path, golden_py):
    from kmc_dn_b200 import _lib
    from kmc_dn_b200.goSimulation.pythonBind import callGoSimulation
    from tests.util import site_energies_of
    c = golden_py["fx_rnd_min_max_0"]
    N, P = c["N"], c["P"]
    S = N + P
    (tmp_path / "goSimulation").mkdir()
    shutil.copy(_lib.SO_PATH, tmp_path / "goSimulation" / "libSimulation.so")
    case = dict(N=N, P=P, nu=float(c["nu"]), kT=float(c["kT"]), I_0=float(c["I_0"]), R=float(c["R"]),
                distances=c["distances"].tolist(), transitions_constant=c["transitions_constant"].tolist(),
                occupation=[bool(o) for o in c["occupation"]], E_constant=c["E_constant"].tolist(),
                site_energies=site_energies_of(c).tolist())
    hops, calls = 200000, 6
    (tmp_path / "caller.py").write_text(CALLER)
    (tmp_path / "job.json").write_text(json.dumps(dict(case=case, hops=hops, calls=calls)))
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}  # nothing of this repo on the caller's path
    subprocess.run([sys.executable, "caller.py", "job.json", "out.json"], cwd=tmp_path, check=True, env=env, timeout=600)
    out = json.loads((tmp_path / "out.json").read_text())

    def currents(rs):
        return np.array([np.array(r["eo"]) / r["time"] for r in rs])
    lit = currents(out["single"])
    assert np.isfinite(lit).all() and all(r["time"] > 0 for r in out["single"])
    # the exports start from the all-empty state and never write the occupation back (pythonBind.py:84, simulationWrapper.go:90)
    assert out["single"][0]["occupation"] == [float(o) for o in c["occupation"]]
    # against this repo's mirror binding (numpy buffers by pointer, library loaded once): same currents within statistics
    mir = []
    for _ in range(calls):
        t, _, eo = callGoSimulation(N, P, c["nu"], c["kT"], c["I_0"], c["R"], 0.0, c["occupation"].astype(float), c["distances"],
                                    c["E_constant"], site_energies_of(c), c["transitions_constant"], None, None, np.zeros(P), hops,
                                    False, "wrapperSimulate")
        mir.append(eo / t)
    mir = np.array(mir)
    z = np.abs(lit.mean(0) - mir.mean(0)) / np.sqrt(lit.var(0) / calls + mir.var(0) / calls + 1e-300)
    assert (z < 6).all(), z
    # record: antisymmetric traffic whose row sums reproduce the electrode tallies; occupied times within [0, time]
    r = out["record"]
    tr = np.array(r["traffic"]).reshape(S, S)
    np.testing.assert_array_equal(tr, -tr.T)
    np.testing.assert_array_equal(tr[:, N:].sum(0), np.array(r["eo"]))
    assert (np.array(r["avg"]) >= 0).all() and (np.array(r["avg"]) <= r["time"] * (1 + 1e-9)).all()
    # RecordPlus ignores `record` (simulationWrapper.go:164-165); Pruned takes the extra leading double
    for k in ("recordplus", "pruned"):
        assert out[k]["time"] > 0 and np.isfinite(out[k]["eo"]).all()
    zp = np.abs(np.array(out["pruned"]["eo"]) / out["pruned"]["time"] - lit.mean(0)) / (lit.std(0) + 1e-300)
    assert (zp < 8).all(), zp
    # the batched export through boxed slices: returns 0, honours the input occupation, writes time and tallies in place
    b = out["batched"]
    assert b["done"] == 0 and len(b["time"]) == calls and all(t > 0 for t in b["time"])
    cb = np.array(b["eo"]).reshape(calls, P) / np.array(b["time"])[:, None]
    zb = np.abs(cb.mean(0) - lit.mean(0)) / np.sqrt(cb.var(0) / calls + lit.var(0) / calls + 1e-300)
    assert (zb < 6).all(), zb
