"""GPU parity suite (-m gpu): the CUDA hop loops, called through the C ABI, against the CPU oracle
on the same seeded inputs, against the committed golden vectors of the unmodified reference, and
against the reference's own statistical fixtures.  Nothing here reads /root/reference."""
import numpy as np
import pytest

from tests.util import calc_D, first_divergence, go_stream, site_energies_of, synthetic_layout

pytestmark = pytest.mark.gpu


def _layout(c, prune=0.0):
    from kmc_dn_b200.ensemble import Layout
    return Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"],
                  prune_threshold=prune)


def _fixture_case(f):
    return dict(N=int(f["N"]), P=int(f["P"]), nu=float(f["nu"]), kT=float(f["kT"]), I_0=float(f["I_0"]), R=float(f["R"]),
                distances=f["distances"], transitions_constant=f["transitions_constant"], E_constant=f["E_constant"],
                electrode_v=f["electrodes"][:, 3].copy(), occupation=f["occupation"].astype(bool))


# ------------------------------------------------------------------ check 1: deterministic replay
def test_replay_py_mode_reproduces_unmodified_numba_reference(golden_py):
    """MODE_PY under numpy's MT19937 stream reproduces the hop sequence, occupations and electrode
    tallies of the UNMODIFIED _simulate_discrete_record bit-exactly (integers), time to 1e-12 (CUDA's
    fp64 exp/log are not glibc's, so the last bits of the rates may differ)."""
    from kmc_dn_b200.ensemble import MODE_PY
    for name, c in golden_py.items():
        hops = int(c["hops"])
        u = np.random.RandomState(int(c["seed"])).random_sample(2 * hops)
        lay = _layout(c)
        r = lay.run(hops, c["kT"], c["electrode_v"][None, :], E_constant=c["E_constant"][None, :], mode=MODE_PY,
                    occupation0=c["occupation"][None, :], stream_u64=u, want_occupation=True, want_site_energies=True,
                    trace=True)
        th = c["ref_trace"].shape[0]
        assert first_divergence(r["trace"][0], c["ref_trace"]) == th, name
        assert (r["occupation"][0] == c["ref_occupation"]).all(), name
        assert (r["electrode_occupation"][0] == c["ref_electrode_occupation"]).all(), name
        assert r["time"][0] == pytest.approx(c["ref_time"], rel=1e-12), name
        np.testing.assert_allclose(r["site_energies"][0], c["ref_site_energies"], rtol=1e-13, atol=1e-12)
        lay.close()


def test_replay_py_mode_full_trace_and_record_vs_oracle(golden_py):
    from oracle import oracle
    from kmc_dn_b200.ensemble import MODE_PY
    for name in ("c1_basic_N10_P2", "fx_rnd_min_max_57"):
        c = golden_py[name]
        hops = 2000
        B = 3
        us = [np.random.RandomState(100 + b).random_sample(2 * hops) for b in range(B)]
        lay = _layout(c)
        r = lay.run(hops, c["kT"], np.tile(c["electrode_v"], (B, 1)), E_constant=np.tile(c["E_constant"], (B, 1)),
                    mode=MODE_PY, occupation0=c["occupation"], stream_u64=np.stack(us), want_occupation=True,
                    record=True, trace=True)
        for b in range(B):
            o = oracle.py_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["occupation"], c["distances"],
                                   c["E_constant"], site_energies_of(c), c["transitions_constant"],
                                   np.zeros(c["P"], dtype=np.int64), hops, record=True, u=us[b], trace=True)
            assert first_divergence(r["trace"][b], o["trace"]) == hops, (name, b)
            assert (r["occupation"][b] == o["occupation"]).all()
            assert (r["electrode_occupation"][b] == o["electrode_occupation"]).all()
            np.testing.assert_array_equal(r["traffic"][b], o["traffic"])
            np.testing.assert_allclose(r["avg_occupation"][b], o["occ_time"], rtol=1e-11)
            assert r["time"][b] == pytest.approx(o["time"], rel=1e-12)
        lay.close()


@pytest.mark.parametrize("variant", [0, 1])
def test_replay_go_modes_bit_exact_vs_oracle(golden_py, variant):
    """MODE_GO_SIMULATE / MODE_GO_RECORDPLUS vs the C restatement of simulate / simulateRecordPlus under
    the same injected (Exp, float32-uniform) stream: identical hop sequence, occupation, tallies, and
    bit-identical fp64 time and fp32 site energies."""
    from oracle import oracle
    from kmc_dn_b200.ensemble import MODE_GO_SIMULATE, MODE_GO_RECORDPLUS
    mode = MODE_GO_RECORDPLUS if variant else MODE_GO_SIMULATE
    for name, c in golden_py.items():
        hops = 3000
        B = 2
        streams = [go_stream(1000 + 7 * b, hops) for b in range(B)]
        occ0 = [None, c["occupation"]]
        lay = _layout(c)
        r = lay.run(hops, c["kT"], np.tile(c["electrode_v"], (B, 1)), E_constant=np.tile(c["E_constant"], (B, 1)),
                    mode=mode, occupation0=np.stack([np.zeros(c["N"], bool), c["occupation"]]),
                    stream_e=np.stack([s[0] for s in streams]), stream_u=np.stack([s[1] for s in streams]),
                    want_occupation=True, want_site_energies=True, record=(variant == 0), trace=True)
        for b in range(B):
            o = oracle.go_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                                   c["transitions_constant"], site_energies_of(c), hops, variant=variant,
                                   occupation=occ0[b], e=streams[b][0], u=streams[b][1], trace=True,
                                   record=(variant == 0))
            assert first_divergence(r["trace"][b], o["trace"]) == hops, (name, b)
            assert (r["occupation"][b] == o["occupation"]).all(), (name, b)
            assert (r["electrode_occupation"][b] == o["electrode_occupation"]).all(), (name, b)
            assert r["time"][b] == o["time"], (name, b)
            np.testing.assert_array_equal(r["site_energies"][b].astype(np.float32), o["site_energies"])
            if variant == 0:
                np.testing.assert_array_equal(r["traffic"][b], o["traffic"])
                np.testing.assert_array_equal(r["avg_occupation"][b], o["average_occupation"])
        lay.close()


def test_replay_go_pruned_list(golden_py):
    """prune threshold (simulation.go:200-215, wrapperSimulatePruned)."""
    from oracle import oracle
    from kmc_dn_b200.ensemble import MODE_GO_SIMULATE
    c = golden_py["fx_xor_wide_3"]
    hops = 2000
    e, u = go_stream(77, hops)
    lay = _layout(c, prune=1e-5)
    r = lay.run(hops, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], mode=MODE_GO_SIMULATE,
                stream_e=e, stream_u=u, trace=True)
    o = oracle.go_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                           c["transitions_constant"], site_energies_of(c), hops, variant=0, cut=1e-5, e=e, u=u, trace=True)
    assert first_divergence(r["trace"][0], o["trace"]) == hops
    assert r["time"][0] == o["time"]
    lay.close()


def test_reference_order_kernel_replays_oracle_until_a_rounding_tie(golden_py):
    """MODE_FAST_REFORDER = production arithmetic (fp32 rates via ex2.approx, fp64 energies / prefix / time) with
    the reference's row-major event order, so under an injected stream it follows the fp32 Go restatement hop
    for hop; it may part only where u*total falls within rounding distance of a list boundary."""
    from oracle import oracle
    from kmc_dn_b200.ensemble import MODE_FAST_REFORDER as MODE_FAST
    agree = []
    for name, c in golden_py.items():
        hops = 4000
        e, u = go_stream(4242, hops)
        lay = _layout(c)
        r = lay.run(hops, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], mode=MODE_FAST,
                    occupation0=c["occupation"][None], stream_e=e, stream_u=u, trace=True, want_occupation=True)
        o = oracle.go_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                               c["transitions_constant"], site_energies_of(c), hops, variant=1,
                               occupation=c["occupation"], e=e, u=u, trace=True)
        k = first_divergence(r["trace"][0], o["trace"])
        agree.append((name, k))
        lay.close()
    # per-hop tie probability is ~(#boundaries)*2^-23; most cases run thousands of hops in lock-step
    assert min(k for _, k in agree) >= 50, agree
    assert np.median([k for _, k in agree]) >= 1000, agree


# ------------------------------------------------------------------ check 2: energies and rates
def test_fast_kernel_energies_and_rates_vs_oracle(golden_py, fixtures_subset):
    """north_star check 2: site energies within 1e-6 (relative to the energy scale) of the Go restatement; rate matrices
    within 1e-6 relative for EVERY entry when evaluated at identical fp32 energies -- the production arithmetic follows the
    reference's roundings of the exponent (float32(-dE/kT), simulation.go:73; kmc_device.cuh boltz), also for kT != 1 where
    that is a correctly rounded division; end to end (device energies) within the reference's own fp32 summation noise:
    the device energies are exact sums, the reference accumulates in fp32 (a few ulp of the largest energy, over kT)."""
    from oracle import oracle
    cases = dict(golden_py)
    for k in ("rnd_min_max/test1", "XOR_wide5M/test2"):
        cases[k] = _fixture_case(fixtures_subset[k])
    rng = np.random.default_rng(0)
    for name, c in cases.items():
        lay = _layout(c)
        for trial in range(6):
            occ = c["occupation"] if trial == 0 else rng.random(c["N"]) < rng.uniform(0.2, 0.95)
            kT = c["kT"] * (1.0, 1.0, 1.0, 1.0, 0.37, 2.9)[trial]  # (the last two: the division is not exact)
            se_o, r_o = oracle.go_rates(c["N"], c["P"], c["nu"], kT, c["I_0"], c["R"], occ, c["distances"],
                                        c["E_constant"], c["transitions_constant"], site_energies_of(c))
            se_d, r_d = lay.probe_rates(c["E_constant"], c["electrode_v"], kT, occ)
            scale = max(np.abs(c["E_constant"]).max(), 1.0)
            np.testing.assert_allclose(se_d, se_o, rtol=1e-6, atol=1e-6 * scale, err_msg=name)
            _, r_same = lay.probe_rates(c["E_constant"], c["electrode_v"], kT, occ, site_energies=se_o)
            assert ((r_same > 0) == (r_o > 0)).all() or np.abs(r_o[(r_same > 0) != (r_o > 0)]).max() < 1e-37
            live = r_o > 1e-30  # (everything above the range where float32 runs out of exponent)
            np.testing.assert_allclose(r_same[live], r_o[live], rtol=1e-6, err_msg=f"{name} kT={kT}")
            live = r_o > 1e-9 * r_o.max()
            # end to end (device energies): a few float32 ulps of the largest energy, over kT, in the exponent
            tol = 6 * float(np.spacing(np.float32(np.abs(se_o).max()))) / kT + 2e-6
            np.testing.assert_allclose(r_d[live], r_o[live], rtol=tol, err_msg=name)
        lay.close()


def test_fast_kernel_incremental_energies_do_not_drift(golden_py):
    """After 2e5 hops the incrementally updated energies equal a from-scratch evaluation of the final state."""
    c = golden_py["fx_rnd_min_max_0"]
    lay = _layout(c)
    r = lay.run(200000, c["kT"], np.tile(c["electrode_v"], (4, 1)), E_constant=np.tile(c["E_constant"], (4, 1)),
                seed=3, want_occupation=True, want_site_energies=True)
    for b in range(4):
        se, _ = lay.probe_rates(c["E_constant"], c["electrode_v"], c["kT"], r["occupation"][b])
        np.testing.assert_array_equal(r["site_energies"][b].astype(np.float32), se)
    lay.close()


def test_fast_kernel_one_hop_event_distribution_matches_oracle_rates(golden_py, fixtures_subset):
    """Order-independent check of the production pick: from one fixed state, 2^18 members take ONE hop each;
    the empirical distribution over (from,to) must match rate_ij / sum(rate) of the oracle (chi-square), and
    no disallowed pair may ever be chosen."""
    from oracle import oracle
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"],
             "n5_p3_hot": golden_py["n5_p3_hot"], "XOR_wide/test3": _fixture_case(fixtures_subset["XOR_wide/test3"]),
             # the general kernel: 2, 4 and 8 acceptor slots per lane, pair table in shared / global memory
             "N48_P8": synthetic_layout(48, 8, 1), "N100_P5": synthetic_layout(100, 5, 2, kT=3.0, I_0=30.0),
             "N256_P8": synthetic_layout(256, 8, 3, kT=4.0, I_0=20.0, fill=0.9)}
    B = 1 << 18
    for name, c in cases.items():
        S = c["N"] + c["P"]
        _, r_o = oracle.go_rates(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["occupation"], c["distances"],
                                 c["E_constant"], c["transitions_constant"], site_energies_of(c))
        p = r_o.astype(np.float64).ravel(); p /= p.sum()
        lay = _layout(c)
        r = lay.run(1, c["kT"], np.tile(c["electrode_v"], (B, 1)), E_constant=np.tile(c["E_constant"], (B, 1)),
                    occupation0=c["occupation"], seed=77, trace=True, kernel="warp")  # (lanes kernel: tests/test_gpu_lanes.py)
        lay.close()
        ev = r["trace"][:, 0, 0].astype(np.int64) * S + r["trace"][:, 0, 1]
        cnt = np.bincount(ev, minlength=S * S).astype(np.float64)
        assert cnt[p == 0].sum() == 0, name  # structurally forbidden pairs are never picked
        big = p * B >= 20
        obs = np.append(cnt[big], cnt[~big].sum()); exp = np.append(p[big] * B, p[~big].sum() * B)
        keep = exp > 0
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        dof = keep.sum() - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof) + 5, (name, chi2, dof)
        # the dwell times are Exp(total): mean of time over members = 1/total
        total = r_o.astype(np.float64).sum()
        assert r["time"].mean() * total == pytest.approx(1.0, abs=5 / np.sqrt(B))


# ------------------------------------------------------------------ check 3: statistics
def _five_run_D(f, cur):
    mu, sd = cur.mean(0), cur.std(0)
    return np.array([calc_D(f["mean_currents"][i], mu[i], f["stddev_currents"][i], sd[i]) for i in range(len(mu))])


def test_fast_kernel_currents_pass_reference_acceptance(fixtures_subset):
    """The reference's own acceptance test (thesis_indrek/validate_tests.py:80-135): 5 runs per fixture,
    per-electrode Bhattacharyya distance against the stored 5-run mean/stddev; D > 0.9 is 'extreme'."""
    Ds = []
    rel = []
    for name, f in fixtures_subset.items():
        c = _fixture_case(f)
        hops = 5_000_000 if "5M" in name else 1_000_000  # same lengths as the fixtures (generate_tests.py:169-171);
        # shorter runs would bias the small currents: the all-empty start injects ~N-M holes once
        lay = _layout(c)
        # fixtures were generated by wrapperSimulateRecordPlus => all-empty start (generate_tests.py:51)
        r = lay.run(hops, c["kT"], np.tile(c["electrode_v"], (5, 1)), E_constant=np.tile(c["E_constant"], (5, 1)),
                    seed=11, member_index0=0)
        cur = r["current"]
        Ds.append(_five_run_D(f, cur))
        ref = np.asarray(f["mean_currents"]); big = np.abs(ref) > 0.05 * np.abs(ref).max()
        rel.append(np.abs(cur.mean(0)[big] - ref[big]) / np.abs(ref[big]))
        lay.close()
    Ds = np.concatenate(Ds); rel = np.concatenate(rel)
    assert Ds.mean() < 0.9, Ds.mean()
    assert (Ds > 0.9).mean() < 0.25, (Ds > 0.9).mean()
    assert np.median(rel) < 0.02 and rel.max() < 0.15, (np.median(rel), rel.max())


def test_reference_acceptance_over_the_fixture_sets():
    """The reference's acceptance run (thesis_indrek/validate_tests.py:299-350) at the fixtures' own run lengths:
    every 2nd fixture of rnd_min_max and XOR_wide (1e6 hops x 5 runs each) and every 10th of the 5e6-hop sets,
    per-electrode Bhattacharyya distance against the stored 5-run statistics (which the reference produced with
    wrapperSimulateRecordPlus, generate_tests.py:45-63).  The run over ALL 400 fixtures is profiles/run_acceptance.py
    -> profiles/acceptance_r01.json: D_mean 0.198 / 0.143 / 0.373 / 0.217, extreme 1.9 / 0.9 / 6.6 / 2.0 % on
    rnd_min_max / rnd_min_max5M / XOR_wide / XOR_wide5M.  Calibration: the CPU oracle (restated Go loop, its own RNG)
    scores D_mean 0.204 / extreme 1.8 % on rnd_min_max and D_mean 0.384 / extreme 6.4 % on XOR_wide against the same
    fixtures; two 5-run samples of one distribution cannot do much better."""
    import os
    from kmc_dn_b200.validate import acceptance_over_sets
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures_all.npz")
    res = acceptance_over_sets(path, stride_5m=10, stride_1m=2)
    assert res["rnd_min_max"]["fixtures"] == 50 and res["XOR_wide"]["fixtures"] == 50 and res["XOR_wide5M"]["fixtures"] == 10
    for setname, r in res.items():
        assert r["D_mean"] < 0.7, (setname, r)
        assert r["extreme"] < 0.2, (setname, r)
    assert res["rnd_min_max"]["D_median"] < 0.2 and res["XOR_wide"]["D_median"] < 0.25, res


def test_fast_kernel_ensemble_mean_within_confidence_interval_of_oracle(golden_py):
    """Ensemble-averaged currents: 256 GPU members vs 256 oracle members (simulateRecordPlus semantics, own
    RNG).  |mean_gpu - mean_cpu| <= 4.5 * sqrt(se_gpu^2 + se_cpu^2) per electrode (two-sample z, ~1e-5 false
    alarm per electrode)."""
    from oracle import oracle
    for name in ("fx_rnd_min_max_0", "c2_grid_N16_P8", "n5_p3_hot", "N48_P8", "N100_P5"):
        c = golden_py[name] if name in golden_py else \
            {"N48_P8": synthetic_layout(48, 8, 1), "N100_P5": synthetic_layout(100, 5, 2, kT=3.0, I_0=30.0)}[name]
        B, hops = (256, 20000) if c["N"] <= 32 else (192, 6000)
        E = np.tile(c["E_constant"], (B, 1)); V = np.tile(c["electrode_v"], (B, 1))
        lay = _layout(c)
        g = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], prehops=2000, seed=5)
        o = oracle.go_ensemble(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], E,
                               c["transitions_constant"], V, hops + 2000, variant=1, use_cache=c["N"] <= 64,
                               occupation0=c["occupation"], seed0=99)  # (getKey drops acceptors beyond 64: no cache there)
        # compare net carrier counts per unit time; the oracle has no prehops, so use equal total lengths
        g2 = lay.run(hops + 2000, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=6)
        cg = g2["electrode_occupation"] / g2["time"][:, None]
        co = o["electrode_occupation"] / o["time"][:, None]
        ok = np.isfinite(cg).all(1) & np.isfinite(co).all(1)
        cg, co = cg[ok], co[ok]
        z = np.abs(cg.mean(0) - co.mean(0)) / np.sqrt(cg.var(0) / len(cg) + co.var(0) / len(co) + 1e-300)
        assert (z < 4.5).all(), (name, z)
        assert np.isfinite(g["time"]).all()
        lay.close()


def test_fast_kernel_results_do_not_depend_on_batching(golden_py):
    """Member m draws from Philox stream (seed, member_index0+m): splitting an ensemble (as ranks do) is invisible."""
    c = golden_py["fx_xor_wide_3"]
    B = 40
    E = np.tile(c["E_constant"], (B, 1)); V = np.tile(c["electrode_v"], (B, 1)) + np.arange(B)[:, None]
    lay = _layout(c)
    a = lay.run(3000, c["kT"], V, E_constant=E, seed=9)
    b1 = lay.run(3000, c["kT"], V[:13], E_constant=E[:13], seed=9, member_index0=0)
    b2 = lay.run(3000, c["kT"], V[13:], E_constant=E[13:], seed=9, member_index0=13)
    np.testing.assert_array_equal(a["time"], np.concatenate([b1["time"], b2["time"]]))
    np.testing.assert_array_equal(a["electrode_occupation"], np.concatenate([b1["electrode_occupation"], b2["electrode_occupation"]]))
    lay.close()


def test_state_memoisation_is_transparent(golden_py, fixtures_subset):
    """The per-warp state cache memoises a pure function of the occupation: with it disabled
    (KMCB200_FLAG_NO_MEMO) every trajectory is bit-identical -- hop sequence, time, tallies, occupation --
    and the cache really is used (rate structures evaluated on a small fraction of the hops)."""
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "n5_p3_hot": golden_py["n5_p3_hot"],
             "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"],  # two ranked events per acceptor (three for N <= 10)
             "c1_basic_N10_P2": golden_py["c1_basic_N10_P2"], "XOR_wide/test1": _fixture_case(fixtures_subset["XOR_wide/test1"]),
             # hop_wide.cu: multi-word masks, 1 / 2 / 4 / 8 acceptors per lane
             "N32_P8": synthetic_layout(32, 8, 4), "N48_P8": synthetic_layout(48, 8, 1),
             "N100_P5": synthetic_layout(100, 5, 2, kT=3.0, I_0=30.0), "N256_P8": synthetic_layout(256, 8, 3, fill=0.9)}
    for name, c in cases.items():
        B, hops = 24, 6000
        V = np.tile(c["electrode_v"], (B, 1)) + np.linspace(0, 5, B)[:, None]
        E = np.tile(c["E_constant"], (B, 1))
        lay = _layout(c)
        kw = dict(E_constant=E, occupation0=c["occupation"], prehops=500, seed=21, trace=True, want_occupation=True,
                  want_site_energies=True, want_misses=True, record=True)
        a = lay.run(hops, c["kT"], V, memo=True, **kw)
        b = lay.run(hops, c["kT"], V, memo=False, **kw)
        lay.close()
        np.testing.assert_array_equal(a["trace"], b["trace"])
        np.testing.assert_array_equal(a["time"], b["time"])
        np.testing.assert_array_equal(a["electrode_occupation"], b["electrode_occupation"])
        np.testing.assert_array_equal(a["occupation"], b["occupation"])
        np.testing.assert_array_equal(a["site_energies"], b["site_energies"])
        np.testing.assert_array_equal(a["avg_occupation"], b["avg_occupation"])
        assert (b["misses"] == hops + 500).all(), name
        assert (a["misses"] < b["misses"]).all() and a["misses"].mean() < 0.9 * (hops + 500), (name, a["misses"].mean())


def test_wide_memo_kernel_agrees_with_general_kernel(monkeypatch):
    """hop_wide.cu (memoised, top events first) and hop_fast.cu (every hop a full sweep, lane-major list) sample the
    same Markov chain: ensemble means of time and electrode currents agree (two-sample z) on the scaling layout."""
    c = synthetic_layout(256, 8, 3, fill=0.9)
    B, hops = 512, 4000
    V = np.tile(c["electrode_v"], (B, 1)); E = np.tile(c["E_constant"], (B, 1))
    lay = _layout(c)
    a = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=31, want_misses=True)
    monkeypatch.setenv("KMCB200_NO_MEMO_KERNEL", "1")
    b = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=32)
    monkeypatch.delenv("KMCB200_NO_MEMO_KERNEL")
    lay.close()
    assert a["misses"].mean() < 0.5 * hops, a["misses"].mean()  # the cache is used
    za = np.abs(a["time"].mean() - b["time"].mean()) / np.sqrt(a["time"].var() / B + b["time"].var() / B)
    ca = a["electrode_occupation"] / a["time"][:, None]; cb = b["electrode_occupation"] / b["time"][:, None]
    z = np.abs(ca.mean(0) - cb.mean(0)) / np.sqrt(ca.var(0) / B + cb.var(0) / B + 1e-300)
    assert za < 4.5 and (z < 4.5).all(), (za, z)


def test_random_layouts_memoisation_and_trace_consistency():
    """40 random layouts (N = 1..70, P = 0..9, random filling / temperature / interaction strength): memoised and
    unmemoised runs are bit-identical, every traced hop is allowed, tallies follow from the trace."""
    rng = np.random.default_rng(2026)
    for it in range(40):
        N = int(rng.choice([1, 2, 3, 7, 10, 11, 15, 16, 17, 24, 25, 30, 31, 32, 33, 40, 64, 65, 70]))
        P = int(rng.integers(0, 10))
        if P == 0 and N < 2:
            P = 1
        c = synthetic_layout(N, P, 100 + it, kT=float(rng.choice([0.5, 1.0, 4.0])), I_0=float(rng.choice([0.0, 30.0, 100.0])),
                             fill=float(rng.uniform(0.1, 0.9)))
        if P == 0 and (c["occupation"].all() or not c["occupation"].any()):
            c["occupation"][0] = not c["occupation"][0]
        B, hops = 3, 1500
        V = np.tile(c["electrode_v"], (B, 1)) + rng.normal(0, 5, (B, P))
        E = np.tile(c["E_constant"], (B, 1))
        lay = _layout(c)
        kw = dict(E_constant=E, occupation0=c["occupation"], seed=it, trace=True, want_occupation=True, record=True)
        a = lay.run(hops, c["kT"], V, memo=True, **kw)
        b = lay.run(hops, c["kT"], V, memo=False, **kw)
        lay.close()
        tag = (it, N, P)
        np.testing.assert_array_equal(a["trace"], b["trace"], err_msg=str(tag))
        np.testing.assert_array_equal(a["time"], b["time"], err_msg=str(tag))
        np.testing.assert_array_equal(a["avg_occupation"], b["avg_occupation"], err_msg=str(tag))
        for m in range(B):
            if not np.isfinite(a["time"][m]):
                continue  # dead state reached (closed system that ran out of moves)
            occ = c["occupation"].copy()
            eo = np.zeros(P, dtype=np.int64)
            for f, t in a["trace"][m]:
                assert f != t and not (f >= N and t >= N), tag
                if f < N:
                    assert occ[f], tag
                    occ[f] = False
                else:
                    eo[f - N] -= 1
                if t < N:
                    assert not occ[t], tag
                    occ[t] = True
                else:
                    eo[t - N] += 1
            np.testing.assert_array_equal(a["occupation"][m].astype(bool), occ, err_msg=str(tag))
            np.testing.assert_array_equal(a["electrode_occupation"][m], eo, err_msg=str(tag))


def test_multi_device_entry_equals_single_device(fixtures_subset):
    """kmcb200_run_ensemble_multi (one process, one host thread per layout copy) returns exactly what one device
    returns: the members are cut into contiguous blocks and streams are numbered by global member index.  With one
    GPU the copies live on the same device (the sharding, offsets and threading are what is under test); with
    several they are spread over all of them."""
    from kmc_dn_b200 import _lib
    from kmc_dn_b200.ensemble import Layout, MultiLayout
    c = _fixture_case(fixtures_subset["XOR_wide/test0"])
    ndev = _lib.load().kmcb200_device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    rng = np.random.default_rng(4)
    B = 1003  # ragged
    V = np.tile(c["electrode_v"], (B, 1)) + rng.normal(0, 20, (B, c["P"]))
    E = np.tile(c["E_constant"], (B, 1)) + rng.normal(0, 1, (B, c["N"]))
    kT = rng.uniform(0.5, 2.0, B)
    one = Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"])
    many = MultiLayout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"],
                       devices=devices)
    kw = dict(E_constant=E, occupation0=c["occupation"], prehops=100, seed=5, member_index0=77, want_occupation=True,
              want_site_energies=True, record=True, trace=True)
    a = one.run(700, kT, V, **kw)
    b = many.run(700, kT, V, **kw)
    for k in ("time", "electrode_occupation", "occupation", "site_energies", "avg_occupation", "traffic", "trace"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    a = one.run(3000, kT, V, E_constant=E, seed=6)
    b = many.run(3000, kT, V, E_constant=E, seed=6)
    np.testing.assert_array_equal(a["time"], b["time"])
    np.testing.assert_array_equal(a["electrode_occupation"], b["electrode_occupation"])
    one.close(); many.close()


def test_superposition_matvec_equals_explicit_E_constant(fixtures_subset):
    """E_constant[m,:] = basis[P,:] + V[m,:] @ basis[:P,:] on device == passing E_constant explicitly."""
    f = fixtures_subset["XOR_wide/test0"]
    c = _fixture_case(f)
    rng = np.random.default_rng(1)
    basis = rng.normal(size=(c["P"] + 1, c["N"]))
    V = rng.uniform(-100, 100, size=(6, c["P"]))
    E = basis[c["P"]][None, :] + V @ basis[:c["P"]]
    lay = _layout(c)
    a = lay.run(5000, c["kT"], V, E_constant=E, seed=2)
    b = lay.run(5000, c["kT"], V, basis=basis, seed=2)
    np.testing.assert_array_equal(a["electrode_occupation"], b["electrode_occupation"])
    np.testing.assert_allclose(a["time"], b["time"], rtol=1e-6)
    lay.close()


# ------------------------------------------------------------------ edge cases
def test_edge_cases(golden_py):
    from kmc_dn_b200.ensemble import Layout
    c = golden_py["c1_basic_N10_P2"]
    lay = _layout(c)
    # hops = 0: time 0, tallies 0, current = eo/0 (kmc_dopant_networks.py:618; used on purpose by the reference)
    r = lay.run(0, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], want_occupation=True,
                occupation0=c["occupation"][None])
    assert r["time"][0] == 0.0 and not r["electrode_occupation"].any()
    assert (r["occupation"][0] == c["occupation"]).all()
    # ragged tail: B not a multiple of the warps per CTA
    r = lay.run(100, c["kT"], np.tile(c["electrode_v"], (37, 1)), E_constant=np.tile(c["E_constant"], (37, 1)), seed=1)
    assert (r["time"] > 0).all() and len(set(r["time"])) == 37
    lay.close()
    # no electrodes, closed system (Boltzmann set-up of kmc_dopant_networks_utils.py:546-636): hops only between acceptors
    N = 6
    rng = np.random.default_rng(5)
    pos = rng.random((N, 2)); d = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1))
    tc = np.exp(-2 * d / 0.3) - np.eye(N)
    lay = Layout(N, 0, d, tc, I_0=2.0, R=N ** -0.5)
    occ = np.array([1, 1, 1, 0, 0, 0], bool)
    r = lay.run(5000, 1.0, np.zeros((2, 0)), E_constant=np.zeros((2, N)), occupation0=occ, want_occupation=True, seed=4)
    assert (r["occupation"].sum(1) == 3).all() and (r["time"] > 0).all()
    # fully occupied, no electrodes: no transition possible -> dead state, time = +inf (e/0 in simulation.go:297)
    r = lay.run(10, 1.0, np.zeros((1, 0)), E_constant=np.zeros((1, N)), occupation0=np.ones(N, bool))
    assert np.isinf(r["time"][0])
    lay.close()
    # S == 32 exactly (one full row slot), S == 64 (two full slots), and the 4- / 8-slot general kernel
    for N, P in ((30, 2), (31, 1), (32, 8), (56, 8), (64, 3), (128, 8), (200, 8)):
        pos = rng.random((N + P, 2)); d = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1))
        tc = np.exp(-2 * d / (0.25 * N ** -0.5)) - np.eye(N + P)
        lay = Layout(N, P, d, tc, I_0=50.0, R=N ** -0.5)
        r = lay.run(2000, 1.0, rng.uniform(-30, 30, (3, P)), E_constant=rng.normal(0, 10, (3, N)), seed=8,
                    want_occupation=True)
        assert np.isfinite(r["time"]).all() and (r["time"] > 0).all()
        assert r["electrode_occupation"].sum(1).tolist() == (-r["occupation"].sum(1)).tolist()  # hole conservation
        lay.close()


def test_record_tallies_of_fast_kernel(golden_py):
    """traffic is the antisymmetric net count and average_occupation the pre-hop occupied time (simulation.go:309-317)."""
    for c in (golden_py["n5_p3_hot"], synthetic_layout(48, 8, 1), synthetic_layout(200, 8, 6, fill=0.9)):
        lay = _layout(c)
        hops = 20000
        r = lay.run(hops, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], record=True, trace=True,
                    occupation0=c["occupation"], want_occupation=True, want_site_energies=True, seed=12)
        N, S = c["N"], c["N"] + c["P"]
        tr = np.zeros((S, S))
        np.add.at(tr, (r["trace"][0][:, 0], r["trace"][0][:, 1]), 1.0)
        np.testing.assert_array_equal(r["traffic"][0], tr - tr.T)
        net_in = -(tr - tr.T).sum(1)[N:]  # holes arriving at each electrode
        np.testing.assert_array_equal(r["electrode_occupation"][0], net_in.astype(np.int64))
        assert (r["avg_occupation"][0] <= r["time"][0] * (1 + 1e-6)).all() and (r["avg_occupation"][0] >= 0).all()
        # replaying the trace on the host: every hop allowed (source occupied / target empty), final mask as reported,
        # final energies = from-scratch energies of that mask
        occ = c["occupation"].copy()
        for f, t in r["trace"][0]:
            if f < N:
                assert occ[f]; occ[f] = False
            if t < N:
                assert not occ[t]; occ[t] = True
        np.testing.assert_array_equal(r["occupation"][0].astype(bool), occ)
        se, _ = lay.probe_rates(c["E_constant"], c["electrode_v"], c["kT"], occ)
        np.testing.assert_array_equal(r["site_energies"][0].astype(np.float32), se)
        lay.close()


# ------------------------------------------------------------------ the drop-in exports
def test_libsimulation_exports_through_goslices(fixtures_subset):
    """wrapperSimulateRecordPlus / wrapperSimulate / wrapperSimulatePruned / parallelSimulations called exactly as
    goSimulation/pythonBind.py:49-90 and parrallelSimulationBind.py:50-66 call them."""
    from kmc_dn_b200.goSimulation.pythonBind import callGoSimulation
    from kmc_dn_b200.goSimulation.parrallelSimulationBind import parrallelSimulation
    f = fixtures_subset["rnd_min_max/test2"]
    c = _fixture_case(f)
    N, P = c["N"], c["P"]
    base = dict(N_acceptors=N, N_electrodes=P, nu=c["nu"], kT=c["kT"], I_0=c["I_0"], R=c["R"], time=0.0,
                occupation=c["occupation"], distances=c["distances"], E_constant=c["E_constant"],
                site_energies=site_energies_of(c), transitions_constant=c["transitions_constant"],
                transitions=np.zeros((N + P, N + P)), problist=np.zeros((N + P) ** 2),
                electrode_occupation=np.zeros(P, dtype=int), hops=1000000)
    curs = []
    for _ in range(5):
        t, occ, eo = callGoSimulation(record=False, goSpecificFunction="wrapperSimulateRecordPlus", **base)
        assert occ.shape == (N,) and eo.shape == (P,) and t > 0
        curs.append(eo / t)
    D = _five_run_D(f, np.array(curs))
    assert D.mean() < 0.9, D
    t, occ, eo, traffic, avg = callGoSimulation(record=True, goSpecificFunction="wrapperSimulate", **base)
    assert traffic.shape == (N + P, N + P) and np.allclose(traffic, -traffic.T) and avg.shape == (N,)
    assert np.abs(eo / t - np.asarray(f["mean_currents"])).max() < 0.2 * np.abs(f["mean_currents"]).max()
    t, occ, eo = callGoSimulation(record=False, goSpecificFunction="wrapperSimulatePruned", prune_threshold=1e-7, **base)
    assert np.abs(eo / t - np.asarray(f["mean_currents"])).max() < 0.2 * np.abs(f["mean_currents"]).max()

    # batched: 6 simulations, two different layouts interleaved, occupation honoured
    class DN:  # the attributes parrallelSimulation.addSimulation reads (parrallelSimulationBind.py:37-48)
        pass
    dns = []
    for k in range(6):
        ff = fixtures_subset["rnd_min_max/test2" if k % 2 == 0 else "XOR_wide/test1"]
        cc = _fixture_case(ff)
        dn = DN()
        dn.N = cc["N"]; dn.nu = cc["nu"]; dn.kT = cc["kT"]; dn.I_0 = cc["I_0"]; dn.R = cc["R"]; dn.time = 0.0
        dn.occupation = cc["occupation"]; dn.electrode_occupation = np.zeros(cc["P"], dtype=int)
        dn.E_constant = cc["E_constant"]; dn.site_energies = site_energies_of(cc); dn.distances = cc["distances"]
        dn.transitions_constant = cc["transitions_constant"]; dn.electrodes = ff["electrodes"]; dn.ref = ff
        dns.append(dn)
    par = parrallelSimulation()
    for dn in dns:
        par.addSimulation(dn, 1000000)
    par.runSimulation()
    for dn in dns:
        t, eo, cur = dn.parrallel_results[0]
        ref = np.asarray(dn.ref["mean_currents"])
        assert t > 0 and np.abs(np.asarray(cur) - ref).max() < 0.2 * np.abs(ref).max()


# ------------------------------------------------------------------ the reference's known-answer physics checks
def test_single_electron_transistor_closed_form():
    """validation/set/set.py:12-72: one acceptor midway between source and drain in 1-D, ab = 1e5*R (distance
    irrelevant), gate energy U_G = U_S/2.  Drain current vs the closed form
    I = G_sg>*G_gd>/(G_sg>+G_gd>) - G_sg<*G_gd</(G_sg<+G_gd<),  G = nu*min(1, exp(dU/kT))  (set.py:24-29)."""
    from kmc_dn_b200.electrostatics import BasisPotentials
    from kmc_dn_b200.ensemble import Layout

    def exp_tresh(x):
        return np.where(x <= 0, np.exp(np.minimum(x, 0)), 1.0)

    def analytical(U_S, U_G, U_D, kT):
        a = exp_tresh((U_S - U_G) / kT); c = exp_tresh((U_G - U_S) / kT)
        d = exp_tresh((U_G - U_D) / kT); b = exp_tresh((U_D - U_G) / kT)
        return a * d / (a + d) - c * b / (c + b)

    acc = np.array([[0.5, 0.0, 0.0]])
    el = np.array([[0.0, 0, 0, 10.0], [1.0, 0, 0, 0.0]])
    pos = np.vstack([acc, el[:, :3]])
    dist = np.abs(pos[:, None, 0] - pos[None, :, 0])
    R = 1.0
    tc = np.exp(-2 * dist / (100000 * R)) - np.eye(3)
    phi = BasisPotentials(acc, el, 1.0)
    bias = np.linspace(-10, 10, 21)
    seeds = 16
    V = np.zeros((len(bias) * seeds, 2)); V[:, 0] = np.repeat(bias, seeds)
    Ec = phi.eV_constant(V)
    np.testing.assert_allclose(Ec[:, 0], V[:, 0] / 2)  # U_G = U_S/2 (plotting.py:29)
    lay = Layout(1, 2, dist, tc, nu=1.0, I_0=100.0, R=R)
    r = lay.run(100000, 1.0, V, E_constant=Ec, seed=3)
    lay.close()
    cur = r["current"][:, 1].reshape(len(bias), seeds)
    mean, sem = cur.mean(1), cur.std(1) / np.sqrt(seeds)
    ref = analytical(bias, bias / 2, 0.0, 1.0)
    assert np.all(np.abs(mean - ref) < 5 * sem + 2e-3), np.abs(mean - ref).max()
    assert np.abs(mean - ref).max() < 0.01
    np.testing.assert_allclose(r["current"][:, 0], -r["current"][:, 1], atol=2e-5)  # what enters at S leaves at D


def test_boltzmann_statistics_without_electrodes():
    """kmc_dopant_networks_utils.py:546-636 / validation/boltzmann_validation.py: no electrodes, fixed carrier
    number; time-weighted occupancy must follow exp(-H/kT)/Z with H = kmc_dn.total_energy
    (kmc_dopant_networks.py:982-1001).  Checked on the per-site occupancies the record tallies deliver."""
    import itertools
    from kmc_dn_b200.ensemble import Layout
    rng = np.random.default_rng(11)
    N, n_holes = 6, 3
    pos = rng.random((N, 2))
    d = np.sqrt(((pos[:, None] - pos[None]) ** 2).sum(-1))
    R = N ** -0.5
    I_0 = 3.0
    tc = np.exp(-2 * d / (0.5 * R)) - np.eye(N)
    eV = rng.uniform(-1, 1, N)
    # exact Boltzmann site occupancies
    w, occs = [], []
    for holes in itertools.combinations(range(N), n_holes):
        occ = np.zeros(N); occ[list(holes)] = 1
        ion = 1 - occ
        H = 0.0
        for i in range(N - 1):
            for j in range(i + 1, N):
                H += ion[i] * ion[j] / d[i, j]
        H = H * I_0 * R - (ion * eV).sum()
        w.append(np.exp(-H)); occs.append(occ)
    w = np.array(w) / np.sum(w)
    exact = (w[:, None] * np.array(occs)).sum(0)
    B = 64
    occ0 = np.zeros(N, bool); occ0[:n_holes] = True
    lay = Layout(N, 0, d, tc, nu=1.0, I_0=I_0, R=R)
    r = lay.run(200000, 1.0, np.zeros((B, 0)), E_constant=np.tile(eV, (B, 1)), occupation0=occ0, prehops=2000, seed=5,
                record=True)
    lay.close()
    frac = r["avg_occupation"] / r["time"][:, None]
    mean, sem = frac.mean(0), frac.std(0) / np.sqrt(B)
    assert np.all(np.abs(mean - exact) < 5 * sem + 1e-3), (mean, exact)
    assert abs(mean.sum() - n_holes) < 1e-6


def test_generation_fitness_in_one_launch(fixtures_subset):
    """a whole generation (candidates x logic-table tests x seeds) evaluated as one ensemble equals
    evaluating every (candidate, test) separately, and the reduction is the reference's error function."""
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout
    from kmc_dn_b200.search_eval import error_corr, evaluate_generation, generation_members
    w = workloads.c3_voltage_search(n_controls=4, seeds=1, hops=1000)
    lt = w["tables"]
    tests = [((0, 0), False), ((0, 75), True), ((75, 0), True), ((75, 75), False)]
    rng = np.random.default_rng(4)
    controls = rng.uniform(-150, 150, (6, 5))
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    errs, cur = evaluate_generation(lay, lt.basis, controls, tests, hops=20000, seeds=4, seed=9, occupation0=w["occupation0"])
    assert errs.shape == (6,) and cur.shape == (6, 4) and np.isfinite(errs).all()
    V = generation_members(controls, tests, 8, seeds=4)
    r = lay.run(20000, 1.0, V, E_constant=lt.E_constant(V), seed=9, occupation0=w["occupation0"])
    lay.close()
    np.testing.assert_allclose(r["current"][:, 7].reshape(6, 4, 4).mean(2), cur, rtol=1e-6, atol=1e-12)
    for g in range(6):
        assert errs[g] == pytest.approx(error_corr(cur[g], tests))


def test_pruned_production_currents_vs_oracle(fixtures_subset):
    """The production kernels with a pruned transition list (simulation.go:200-215, wrapperSimulatePruned,
    validate_tests.py:323: pairs with tc <= 1e-7 max(tc) dropped) against the CPU restatement of the Go loop run with the
    SAME cut: two-sample z per electrode current and for the elapsed time -- on a 30-acceptor fixture (narrow kernels,
    both of them) and on the 256-acceptor layout of examples/scaling.py (wide kernel), with and without pruning."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout

    def oracle_runs(N, P, nu, kT, I_0, R, d, tc, E, V, hops, cut, n, occ0):
        def one(k):
            se = np.zeros(N + P); se[N:] = V
            o = oracle.go_simulate(N, P, nu, kT, I_0, R, d, E, tc, se, hops, variant=1, occupation=occ0, use_cache=False,
                                   cut=cut, seed=1000 + k)
            return np.append(o["electrode_occupation"] / o["time"], o["time"])
        with ThreadPoolExecutor(max_workers=16) as ex:  # (ctypes releases the GIL)
            return np.array(list(ex.map(one, range(n))))

    def check(tag, g, o):
        cg = np.column_stack([g["electrode_occupation"] / g["time"][:, None], g["time"]])
        z = np.abs(cg.mean(0) - o.mean(0)) / np.sqrt(cg.var(0) / len(cg) + o.var(0) / len(o) + 1e-300)
        assert (z < 5).all(), (tag, z)

    c = _fixture_case(fixtures_subset["rnd_min_max/test2"])
    for cut in (0.0, 1e-7, 1e-3):
        hops, n = 20000, 48
        o = oracle_runs(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["transitions_constant"],
                        c["E_constant"], c["electrode_v"], hops, cut, n, c["occupation"])
        lay = _layout(c, prune=cut)
        for kernel in ("warp", "lanes"):
            g = lay.run(hops, c["kT"], np.tile(c["electrode_v"], (256, 1)), E_constant=np.tile(c["E_constant"], (256, 1)),
                        occupation0=c["occupation"], seed=17, kernel=kernel)
            check((cut, kernel), g, o)
        lay.close()
    w = workloads.c5_scaling(N=256, M=25, B=64)
    lt = w["tables"]
    V = w["V"][0]
    E = lt.E_constant(V)
    for cut in (0.0, 1e-7):
        hops, n = 3000, 32
        o = oracle_runs(lt.N, lt.P, lt.nu, 1.0, lt.I_0, lt.R, lt.distances, lt.transitions_constant, E, V, hops, cut, n,
                        w["occupation0"])
        lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R, prune_threshold=cut)
        g = lay.run(hops, 1.0, np.tile(V, (256, 1)), E_constant=np.tile(E, (256, 1)), occupation0=w["occupation0"], seed=23)
        lay.close()
        check((256, cut), g, o)


def test_device_reduction_of_currents(fixtures_subset):
    """(sum x, sum x^2, n) per voltage vector on the device == the host reduction of the per-member currents
    (kmc_dopant_networks.py:618; consumers voltage_search.py:160-185, validate_tests.py:80-135)."""
    import torch
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout
    w = workloads.c3_voltage_search(n_controls=8, seeds=16, hops=2000)
    lt = w["tables"]
    B, P, g = len(w["V"]), lt.P, 16
    dev = torch.device("cuda", 0)
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    V = torch.from_numpy(w["V"]).to(dev); kT = torch.from_numpy(w["kT"]).to(dev); basis = torch.from_numpy(lt.basis).to(dev)
    t = torch.zeros(B, dtype=torch.float64, device=dev); eo = torch.zeros((B, P), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lay.run_device(B, 2000, kT, V, t, eo, basis=basis, seed=3, cuda_stream=st)
    s1 = torch.zeros((B // g, P), dtype=torch.float64, device=dev); s2 = torch.zeros_like(s1)
    n = torch.zeros(B // g, dtype=torch.float64, device=dev)
    lay.reduce_currents_device(t, eo, g, s1, s2, n, cuda_stream=st)
    torch.cuda.synchronize()
    lay.close()
    cur = (eo.double() / t[:, None]).cpu().numpy().reshape(B // g, g, P)
    np.testing.assert_allclose(s1.cpu().numpy(), cur.sum(1), rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(s2.cpu().numpy(), (cur ** 2).sum(1), rtol=1e-12, atol=1e-30)
    assert (n.cpu().numpy() == g).all()


def test_genetic_search_runs_generations_as_single_launches(fixtures_subset):
    """SURVEY 8f-2 as a consumer: the reference's genetic voltage search (dn_search.py:411-549) on top of the batched hop
    loop -- every generation is ONE ensemble launch; the fitness of a generation is the reference's error function (pinned
    to the unmodified reference by tests/golden/search_eval.npz) of the seed-averaged output currents."""
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout, launch_count
    from kmc_dn_b200.search_eval import error_corr, genetic_search
    w = workloads.c3_voltage_search(n_controls=4, seeds=1, hops=1000)
    lt = w["tables"]
    tests = [((0, 0), False), ((0, 75), True), ((75, 0), True), ((75, 75), False)]
    lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R)
    seen = []

    def check(g, errors, cur):
        seen.append(g)
        # (a candidate whose four currents have no variance has an undefined correlation: nan, as in the reference)
        np.testing.assert_allclose(errors, [error_corr(cur[k], tests) for k in range(len(errors))], rtol=1e-12, equal_nan=True)
    l0 = launch_count()
    best, controls, hist = genetic_search(lay, lt.basis, tests, gen_size=16, generations=4, hops=5000, seeds=4, seed=3,
                                          occupation0=w["occupation0"], on_generation=check)
    assert launch_count() - l0 == 4 and seen == [0, 1, 2, 3]  # one kernel launch per generation
    lay.close()
    assert np.isfinite(best) and controls.shape == (5,) and np.abs(controls).max() <= 150
    assert best == pytest.approx(min(h[0] for h in hist))


def test_mean_field_prescreen_vs_oracle(golden_py, fixtures_subset):
    """probSimulate (probabilitySimulation.go:53-157) on the GPU vs its C restatement: time, fractional
    occupations, electrode tallies, acceptor energies, traffic and occupied time to fp64 rounding; and through the
    wrapperSimulateProbability export exactly as dn_search's strategy 0 calls it (dn_search.py:49-52)."""
    from oracle import oracle
    from kmc_dn_b200.goSimulation.pythonBind import callGoSimulation
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"],
             "n5_p3_hot": golden_py["n5_p3_hot"], "N48_P8": synthetic_layout(48, 8, 1, kT=2.0, I_0=30.0)}
    # The reference accumulates pair by pair; the kernel sums rows/columns per lane and reduces over the warp.  One
    # step therefore agrees to fp64 rounding; the relaxation map then amplifies that (the step limiter is
    # non-smooth), so long runs are compared at 1e-5.
    for name, c in cases.items():
        lay = _layout(c)
        V = np.stack([c["electrode_v"], c["electrode_v"] * 0.5])
        E = np.stack([c["E_constant"], c["E_constant"]])
        for steps, tol in ((1, 1e-12), (10, 1e-10), (300, 1e-5)):
            o = oracle.prob_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                                     c["transitions_constant"], site_energies_of(c), steps, record=True)
            r = lay.run_prob(steps, c["kT"], V, E_constant=E, record=True)
            scale_eo = np.abs(o["electrode_occupation"]).max() + 1e-300
            assert r["time"][0] == pytest.approx(o["time"], rel=tol), (name, steps)
            np.testing.assert_allclose(r["occupation"][0], o["occupation"], rtol=0, atol=10 * tol, err_msg=name)
            np.testing.assert_allclose(r["electrode_occupation"][0], o["electrode_occupation"], rtol=0, atol=100 * tol * scale_eo)
            np.testing.assert_allclose(r["site_energies"][0], o["site_energies"], rtol=10 * tol, atol=1e3 * tol, err_msg=name)
            np.testing.assert_allclose(r["avg_occupation"][0], o["average_occupation"], rtol=0, atol=100 * tol * o["time"])
            np.testing.assert_allclose(r["traffic"][0], o["traffic"], rtol=0, atol=100 * tol * np.abs(o["traffic"]).max())
        # member 1 (halved electrode energies) against its own oracle run: members are independent
        se1 = site_energies_of(c); se1[c["N"]:] *= 0.5
        o1 = oracle.prob_simulate(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                                  c["transitions_constant"], se1, 300)
        assert r["time"][1] == pytest.approx(o1["time"], rel=1e-5), name
        np.testing.assert_allclose(r["occupation"][1], o1["occupation"], rtol=0, atol=1e-4, err_msg=name)
        lay.close()
    c = golden_py["fx_rnd_min_max_0"]
    N, P = c["N"], c["P"]
    o = oracle.prob_simulate(N, P, c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], c["E_constant"],
                             c["transitions_constant"], site_energies_of(c), 1000)
    t, occ, eo = callGoSimulation(N_acceptors=N, N_electrodes=P, nu=c["nu"], kT=c["kT"], I_0=c["I_0"], R=c["R"], time=0.0,
                                  occupation=c["occupation"], distances=c["distances"], E_constant=c["E_constant"],
                                  site_energies=site_energies_of(c), transitions_constant=c["transitions_constant"],
                                  transitions=np.zeros((N + P, N + P)), problist=np.zeros((N + P) ** 2),
                                  electrode_occupation=np.zeros(P, dtype=int), hops=1000, record=False,
                                  goSpecificFunction="wrapperSimulateProbability")
    assert t == pytest.approx(o["time"], rel=1e-4)
    # the reference binding truncates the (fractional) results to int (pythonBind.py:83-84)
    assert np.abs(eo - o["electrode_occupation"].astype(np.int64)).max() <= 1
    assert np.abs(occ - o["occupation"].astype(np.int64)).max() <= 1


def test_host_class_simulation_entry_points(fixtures_subset):
    """kmc_dn.go_simulation / python_simulation / ensemble_simulation (mirror of kmc_dopant_networks.py:473-618):
    python_simulation replays the numba loop under numpy's MT19937 stream, so after np.random.seed(s) it must
    equal the oracle's run on RandomState(s); go_simulation must reproduce the fixture's currents."""
    from oracle import oracle
    from kmc_dn_b200.kmc_dopant_networks import kmc_dn
    f = fixtures_subset["rnd_min_max/test1"]
    dn = kmc_dn(int(f["N"]), int(f["M"]), 1, 1, 0, electrodes=f["electrodes"], acceptors=f["acceptors"], donors=f["donors"])
    np.testing.assert_allclose(dn.E_constant, f["E_constant"], atol=2e-11)
    # --- python_simulation == numba semantics under the same stream
    dn.occupation = f["occupation"].astype(bool).copy()
    occ0 = dn.occupation.copy()
    hops = 1500
    np.random.seed(77)
    dn.python_simulation(hops=hops, record=True)
    u = np.random.RandomState(77).random_sample(2 * hops)
    se = np.zeros(dn.N + dn.P); se[dn.N:] = dn.electrodes[:, 3]
    o = oracle.py_simulate(dn.N, dn.P, dn.nu, dn.kT, dn.I_0, dn.R, occ0, dn.distances, dn.E_constant, se,
                           dn.transitions_constant, np.zeros(dn.P, dtype=np.int64), hops, record=True, u=u)
    assert (dn.occupation == o["occupation"]).all()
    assert (dn.electrode_occupation == o["electrode_occupation"]).all()
    assert dn.time == pytest.approx(o["time"], rel=1e-12)
    np.testing.assert_array_equal(dn.traffic, o["traffic"])
    np.testing.assert_allclose(np.asarray(dn.average_occupation) * dn.time, o["occ_time"], rtol=1e-10)
    np.testing.assert_allclose(dn.current, o["electrode_occupation"] / o["time"], rtol=1e-12)
    # prehops run first, on the same stream (kmc_dopant_networks.py:580-585)
    dn.occupation = occ0.copy()
    np.random.seed(78)
    dn.python_simulation(hops=500, prehops=300)
    u = np.random.RandomState(78).random_sample(2 * 800)
    a = oracle.py_simulate(dn.N, dn.P, dn.nu, dn.kT, dn.I_0, dn.R, occ0, dn.distances, dn.E_constant, se,
                           dn.transitions_constant, np.zeros(dn.P, dtype=np.int64), 300, u=u[:600])
    b = oracle.py_simulate(dn.N, dn.P, dn.nu, dn.kT, dn.I_0, dn.R, a["occupation"], dn.distances, dn.E_constant, se,
                           dn.transitions_constant, np.zeros(dn.P, dtype=np.int64), 500, u=u[600:])
    assert (dn.occupation == b["occupation"]).all() and (dn.electrode_occupation == b["electrode_occupation"]).all()
    assert dn.time == pytest.approx(b["time"], rel=1e-12)
    # --- go_simulation: the default export, statistics of the fixture
    ref = np.asarray(f["mean_currents"]); big = np.abs(ref) > 0.05 * np.abs(ref).max()
    curs = []
    for _ in range(3):
        dn.go_simulation(hops=1000000)
        assert dn.electrode_occupation.shape == (8,) and dn.time > 0
        curs.append(dn.current.copy())
    np.testing.assert_allclose(np.mean(curs, 0)[big], ref[big], rtol=0.05)
    with pytest.raises(TypeError):
        dn.go_simulation(hops=1E5)  # like the reference binding, ctypes' c_int rejects a float
    dn.go_simulation(hops=200000, record=True, goSpecificFunction="wrapperSimulate")
    assert np.asarray(dn.traffic).shape == (38, 38) and len(dn.average_occupation) == 30
    # --- ensemble_simulation: IV sweep of electrode 0 as one launch; zero bias everywhere -> zero mean current
    V = np.zeros((8, 8)); V[:, 0] = np.linspace(-100, 100, 8)
    r = dn.ensemble_simulation(V, hops=50000, prehops=5000, seeds=4, seed=1)
    assert r["current"].shape == (32, 8) and np.isfinite(r["current"]).all()
    cur0 = r["current"].reshape(8, 4, 8).mean(1)[:, 0]
    assert cur0[0] > 0 > cur0[-1] or cur0[0] < 0 < cur0[-1]  # the swept electrode's current changes sign with its bias


def test_wide_sparse_sweep_is_bit_identical():
    """A layout built with a prune threshold (simulation.go:200-215) has a pair table that is mostly exact zeros; hop_wide.cu then
    walks the non-zero pairs only (a miss costs O(neighbours)).  Trace, time, tallies, occupation and energies must be
    bit-identical with the dense sweep over every (occupied, empty) pair (KMCB200_WIDE_SPARSE=0) -- 256 acceptors (8 per lane)
    and 100 (4 per lane), cache on and off, thresholds that leave ~5 % and ~1 % of the pairs."""
    import os
    from kmc_dn_b200 import workloads
    from kmc_dn_b200.ensemble import Layout, last_kernel

    for N, M, cut in ((256, 25, 1e-7), (256, 25, 1e-4), (100, 10, 1e-7), (256, 25, 0.0)):
        w = workloads.c5_scaling(N=N, M=M, B=24)
        lt = w["tables"]
        lay = Layout(lt.N, lt.P, lt.distances, lt.transitions_constant, nu=lt.nu, I_0=lt.I_0, R=lt.R, prune_threshold=cut)
        res = {}
        for sparse in ("1", "0"):
            for memo in (True, False):
                os.environ["KMCB200_WIDE_SPARSE"] = sparse
                try:
                    res[sparse, memo] = lay.run(600, w["kT"][:24], w["V"][:24], basis=lt.basis, occupation0=w["occupation0"], seed=5,
                                                memo=memo, trace=True, want_occupation=True, want_site_energies=True)
                finally:
                    os.environ.pop("KMCB200_WIDE_SPARSE", None)
                assert last_kernel() == "kmc_wide_kernel"
        ref = res["0", False]
        assert np.isfinite(ref["time"]).all() and (ref["time"] > 0).all()
        for key, r in res.items():
            for k in ("time", "electrode_occupation", "occupation", "trace", "site_energies"):
                np.testing.assert_array_equal(r[k], ref[k], err_msg=f"N={N} cut={cut} sparse,memo={key}: {k}")
        lay.close()
