"""CPU suite, part 2: electrostatics front end and the kmc_dn host class set-up against the reference's
stored fixture fields (inputs pinned to fp64 round-off), plus the benchmark workload builders."""
import os

import numpy as np
import pytest

from tests.util import load_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def estat():
    return load_cases("electrostatics.npz")


def test_basis_potentials_reproduce_fenics_eV_constant(estat):
    """Superposition of per-electrode FD/P1 solutions == the FEniCS solve stored in the reference's fixtures
    (kmc_dopant_networks.py:706-899).  40 fixtures, |error| <= 2e-11 on values of magnitude ~100."""
    from kmc_dn_b200.electrostatics import BasisPotentials, comp_constant
    assert len(estat) == 40
    by_layout = {}
    for name, c in estat.items():
        key = (c["acceptors"].tobytes(), c["electrodes"][:, :3].tobytes())
        if key not in by_layout:
            by_layout[key] = BasisPotentials(c["acceptors"], c["electrodes"], float(c["xdim"]), float(c["ydim"]), 0.0,
                                             res=float(c["res"]))
        bp = by_layout[key]
        ev = bp.eV_constant(c["electrodes"][:, 3], mu=float(c["mu"]))
        np.testing.assert_allclose(ev, c["eV_constant"], rtol=0, atol=2e-11, err_msg=name)
        cc = comp_constant(c["acceptors"], c["donors"], float(c["I_0"]), float(c["R"]))
        np.testing.assert_allclose(cc, c["comp_constant"], rtol=1e-14, err_msg=name)
        basis = bp.kernel_basis(cc, mu=float(c["mu"]))
        np.testing.assert_allclose(c["electrodes"][:, 3] @ basis[:-1] + basis[-1], c["E_constant"], rtol=0, atol=2e-11)
    assert len(by_layout) < len(estat)  # XOR fixtures share layouts: one factorisation serves many voltage vectors


def test_basis_potentials_partition_of_unity_and_boundary_rule(estat):
    from kmc_dn_b200.electrostatics import BasisPotentials
    c = estat["rnd_min_max/test0"]
    bp = BasisPotentials(c["acceptors"], c["electrodes"], 1.0, 1.0, 0.0, res=0.01)
    np.testing.assert_allclose(bp.phi.sum(0), 1.0, atol=1e-12)  # all boundaries at 1 -> V == 1 everywhere
    V = c["electrodes"][:, 3]
    for e in c["electrodes"]:  # on an electrode's centre the potential is the electrode voltage
        assert bp.potential_at(e[0], e[1], V) == pytest.approx(e[3], abs=1e-12)
    assert bp.potential_at(0.0, 0.5, V, mu=3.0) == pytest.approx(3.0)  # bare boundary sits at mu (:763)
    # electrode modelled as point +- xdim/10 (:742-752): 0.34 is inside electrode 0's segment, 0.36 is not
    assert bp.potential_at(0.0, 0.34, V) == pytest.approx(c["electrodes"][0, 3])
    assert abs(bp.potential_at(0.0, 0.36, V)) < abs(c["electrodes"][0, 3])


def test_one_dimensional_potential_is_linear():
    """dim 1 (validation/set/set.py set-up): only the two end points are Dirichlet nodes (:736-739, :772)."""
    from kmc_dn_b200.electrostatics import BasisPotentials
    acc = np.array([[0.5, 0, 0], [0.2, 0, 0]])
    el = np.array([[0.0, 0, 0, 4.0], [1.0, 0, 0, -2.0]])
    bp = BasisPotentials(acc, el, 1.0)
    np.testing.assert_allclose(bp.eV_constant(el[:, 3]), [1.0, 2.8])
    with pytest.raises(NotImplementedError):
        BasisPotentials(acc, el, 1.0, 1.0, 1.0)


def test_kmc_dn_setup_matches_fixture_fields(fixtures_subset):
    """kmc_dn built from a fixture's acceptors/donors/electrodes reproduces the stored distances,
    transitions_constant, comp_constant, eV_constant, E_constant and R/ab."""
    from kmc_dn_b200.kmc_dopant_networks import kmc_dn
    for name in ("rnd_min_max/test1", "XOR_wide/test2"):
        f = fixtures_subset[name]
        dn = kmc_dn(int(f["N"]), int(f["M"]), 1, 1, 0, electrodes=f["electrodes"], acceptors=f["acceptors"],
                    donors=f["donors"])
        assert dn.R == f["R"] and dn.ab == f["ab"] and dn.P == 8 and dn.dim == 2 and dn.res == f["res"]
        np.testing.assert_array_equal(dn.distances, f["distances"])
        np.testing.assert_allclose(dn.transitions_constant, f["transitions_constant"], rtol=0, atol=2e-16)
        np.testing.assert_allclose(dn.comp_constant, f["comp_constant"], rtol=1e-14)
        np.testing.assert_allclose(dn.eV_constant, f["eV_constant"], rtol=0, atol=2e-11)
        np.testing.assert_allclose(dn.E_constant, f["E_constant"], rtol=0, atol=2e-11)
        np.testing.assert_array_equal(dn.site_energies[dn.N:], f["electrodes"][:, 3])
        # update_V after a voltage change == superposition
        dn.electrodes[:, 3] *= 0.5
        dn.update_V()
        np.testing.assert_allclose(dn.eV_constant, 0.5 * f["eV_constant"], rtol=0, atol=2e-11)
        assert dn.V(dn.acceptors[3, 0], dn.acceptors[3, 1]) == pytest.approx(dn.eV_constant[3], abs=1e-12)


def test_kmc_dn_random_placement_and_persistence(tmp_path):
    from kmc_dn_b200.kmc_dopant_networks import kmc_dn
    np.random.seed(3)
    el = np.zeros((2, 4)); el[0] = [0, 0.5, 0, 10]; el[1] = [1, 0.5, 0, -10]
    dn = kmc_dn(10, 2, 1, 1, 0, electrodes=el)
    assert dn.occupation.sum() == 8 and dn.acceptors.shape == (10, 3) and (dn.acceptors[:, 2] == 0).all()
    assert dn.vectors.shape == (12, 12, 3)
    np.testing.assert_allclose(np.linalg.norm(dn.vectors[0, 1]), 1.0)
    np.testing.assert_allclose(dn.vectors[0, 1], -dn.vectors[1, 0])
    H = dn.total_energy()
    assert np.isfinite(H)
    dn.current = np.array([1.0, -1.0])
    p = tmp_path / "x.kmc"
    dn.saveSelf(str(p))
    dn2 = kmc_dn(10, 2, 1, 1, 0, electrodes=el)
    dn2.loadSelf(str(p))
    np.testing.assert_array_equal(dn2.acceptors, dn.acceptors)
    np.testing.assert_array_equal(dn2.E_constant, dn.E_constant)
    np.testing.assert_array_equal(dn2.expected_current, [1.0, -1.0])
    assert not dn2.occupation.any()  # like the reference: loadSelf re-initialises, occupation restarts empty (:978, :434)
    dn3 = kmc_dn(10, 2, 1, 1, 0, electrodes=el, copy_from=dn)
    assert dn3.I_0 == dn.I_0
    dn.load_donors(np.random.rand(3, 3) * [1, 1, 0])
    assert dn.M == 3 and dn.comp_constant.shape == (10,) and dn.occupation.sum() == 7


def test_workload_builders_are_deterministic_and_shaped():
    from kmc_dn_b200 import workloads
    w = workloads.c3_voltage_search(n_controls=8, seeds=2, hops=100)
    lt = w["tables"]
    assert (lt.N, lt.M, lt.P) == (30, 3, 8) and w["V"].shape == (64, 8)
    assert (w["V"][:, 7] == 0).all() and set(np.unique(w["V"][:, :2])) == {0.0, 75.0}
    assert (w["V"][0] == w["V"][1]).all() and not (w["V"][0] == w["V"][2]).all()  # seeds are the fastest index
    assert np.abs(w["V"][:, 2:7]).max() <= 150 and w["occupation0"].sum() == 27
    w2 = workloads.c3_voltage_search(n_controls=8, seeds=2, hops=100)
    np.testing.assert_array_equal(w["V"], w2["V"])
    np.testing.assert_allclose(lt.E_constant(w["V"][:3]), w["V"][:3] @ lt.basis[:8] + lt.basis[8])
    for f, shape in ((workloads.c1_basic, (10, 2)), (workloads.c2_grid4x4, (16, 8)), (workloads.c4_temperature, (30, 2))):
        w = f()
        assert (w["tables"].N, w["tables"].P) == shape and len(w["V"]) == len(w["kT"])
    w = workloads.c5_scaling(N=64, M=6, B=16)
    assert w["tables"].S == 72


def test_search_error_function_matches_reference_formula():
    """voltage_search.py:118-136 on hand-checked inputs."""
    from kmc_dn_b200.search_eval import error_corr, generation_members, perfect_correlation
    tests = [((0, 0), False), ((0, 75), True), ((75, 0), True), ((75, 75), False)]  # XOR (voltage_search_tests.py:159)
    np.testing.assert_array_equal(perfect_correlation(tests), [0, 10, 10, 0])
    # separated: highest_false (max(-1, 0.1, 0.2)) - lowest_true (min(1, 0.5, 0.6)) = -0.3, corr = 0.9734...
    vals = [0.1, 0.5, 0.6, 0.2]
    corr = np.corrcoef([0, 10, 10, 0], vals)[0][1]
    assert error_corr(vals, tests) == pytest.approx(corr * (0.2 - 0.5))
    assert error_corr(vals, tests, corr_pow=3) == pytest.approx(corr ** 3 * (0.2 - 0.5))
    # not separated: positive separation is returned as is
    assert error_corr([0.4, 0.1, 0.6, 0.2], tests) == pytest.approx(0.4 - 0.1)
    # anti-correlated: corr clamps to 0
    assert error_corr([0.5, -0.2, -0.1, 0.6], tests) == pytest.approx(0.6 - (-0.2))
    assert error_corr([0.0, -0.2, -0.1, -0.3], tests) == pytest.approx(0.0 - (-0.2))
    # currents far below the initial bounds (1, -1): highest_false stays -1 only if every false current < -1
    assert error_corr([-2.0, 3.0, 4.0, -5.0], tests) == pytest.approx(1.0 * (-1 - 1) * np.corrcoef([0, 10, 10, 0], [-2, 3, 4, -5])[0][1])
    V = generation_members(np.array([[1., 2, 3, 4, 5], [6, 7, 8, 9, 10]]), tests, 8, seeds=3)
    assert V.shape == (2 * 4 * 3, 8)
    np.testing.assert_array_equal(V[0], [0, 0, 1, 2, 3, 4, 5, 0])
    np.testing.assert_array_equal(V[3], [0, 75, 1, 2, 3, 4, 5, 0])   # seeds are the fastest index
    np.testing.assert_array_equal(V[-1], [75, 75, 6, 7, 8, 9, 10, 0])
    with pytest.raises(ValueError):
        generation_members(np.zeros((1, 4)), tests, 8)


def test_search_error_functions_match_the_unmodified_reference():
    """error_corr / error_diff against golden (currents -> error) pairs computed by the UNMODIFIED reference functions
    voltage_search.evaluate_error_corr, evaluate_error_corr_parallel and evaluate_error_diff (voltage_search.py:92-185),
    imported in the build container by oracle/make_golden_search.py: 3 logic tables x 40 candidates, corr_pow 1..3."""
    import os
    from kmc_dn_b200.search_eval import error_corr, error_diff
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "search_eval.npz"))
    n = 0
    for name in ("xor", "and", "or"):
        tests = [(tuple(i), bool(e)) for i, e in zip(z[f"{name}_tests_inputs"], z[f"{name}_tests_expected"])]
        for v, cp, e_seq, e_par, e_dif in zip(z[f"{name}_values"], z[f"{name}_corr_pow"], z[f"{name}_error_corr"],
                                              z[f"{name}_error_corr_parallel"], z[f"{name}_error_diff"]):
            assert e_seq == e_par  # the reference's two code paths agree with each other
            assert error_corr(v, tests, corr_pow=int(cp)) == e_seq, (name, v, cp)
            assert error_diff(v, tests) == e_dif
            n += 1
    assert n == 120


def test_genetic_search_gene_coding():
    """uint16 gene coding of voltage_search.getGenes / getDnFromGenes (voltage_search.py:214-227)."""
    from kmc_dn_b200.search_eval import controls_of, genes_of
    v = np.array([-150.0, -75.0, 0.0, 149.0, 150.0])
    g = genes_of(v, 150.0)
    assert g.dtype == np.uint16 and g[0] == 0 and g[-1] == 65535
    np.testing.assert_allclose(controls_of(g, 150.0), v, atol=300 / 65535)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line carrying the keys of
    the bench contract; it needs no GPU.  Tiny sample here -- the driver's run uses the defaults."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "0.5", "--controls", "64", "--hops", "2000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "hops/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_bench_refuses_to_run_without_gpu():
    """Our arm has no CPU path: without a CUDA device bench.py exits with an error instead of printing a number."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_kernel_choice_argument():
    """`kernel=` of Layout.run maps onto the C ABI's KMCB200_FLAG_LANES / _NO_LANES / _SOLO / _NO_SOLO; unknown names are refused."""
    from kmc_dn_b200 import _lib
    from kmc_dn_b200.ensemble import _kernel_flags
    assert _kernel_flags(None) == 0 and _kernel_flags("auto") == 0
    assert _lib.FLAG_LANES == 4 and _lib.FLAG_NO_LANES == 8 and _lib.FLAG_SOLO == 16 and _lib.FLAG_NO_SOLO == 32  # include/kmc_b200.h
    assert _kernel_flags("lanes") == _lib.FLAG_LANES | _lib.FLAG_NO_SOLO and _kernel_flags("warp") == _lib.FLAG_NO_LANES | _lib.FLAG_NO_SOLO
    assert _kernel_flags("solo") == _lib.FLAG_SOLO
    with pytest.raises(KeyError):
        _kernel_flags("fastest")
    header = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    assert "KMCB200_FLAG_LANES = 4" in header and "KMCB200_FLAG_NO_LANES = 8" in header


def test_committed_ncu_summaries_feed_the_bench_roofline():
    """bench.py's roofline reads warp-instructions per hop of the kernel that ran from profiles/ncu_r01_<kernel>_kernel.json."""
    import json
    for kernel in ("lanes", "memo", "wide"):
        j = json.load(open(os.path.join(ROOT, "profiles", f"ncu_r01_{kernel}_kernel.json")))
        assert j["warp_inst_per_hop"] > 1 and kernel in j["kernel"], kernel
        assert 0 < j["issue_active_pct"] <= 100
