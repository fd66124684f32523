"""GPU parity suite of the thread-per-trajectory kernel (csrc/hop_lanes.cu, `kernel="lanes"`), through the C ABI:
the state table must be invisible (bit-identical with every hop evaluated from scratch), every traced hop must be
allowed, the one-hop event distribution must match the oracle's rate matrix, and currents must agree with the
warp-per-trajectory kernel, the CPU oracle and the reference's fixtures.  Nothing here reads /root/reference."""
import numpy as np
import pytest

from tests.util import calc_D, site_energies_of, synthetic_layout

pytestmark = pytest.mark.gpu


def _layout(c):
    from kmc_dn_b200.ensemble import Layout
    return Layout(c["N"], c["P"], c["distances"], c["transitions_constant"], nu=c["nu"], I_0=c["I_0"], R=c["R"])


def _fixture_case(f):
    return dict(N=int(f["N"]), P=int(f["P"]), nu=float(f["nu"]), kT=float(f["kT"]), I_0=float(f["I_0"]), R=float(f["R"]),
                distances=f["distances"], transitions_constant=f["transitions_constant"], E_constant=f["E_constant"],
                electrode_v=f["electrodes"][:, 3].copy(), occupation=f["occupation"].astype(bool))


def _check_trace(c, r, occupation0):
    """every traced hop is allowed (simulation.go:40-55); final occupation and tallies follow from the trace"""
    N, P = c["N"], c["P"]
    for m in range(len(r["time"])):
        if not np.isfinite(r["time"][m]):
            continue
        occ = occupation0.copy()
        eo = np.zeros(P, dtype=np.int64)
        for f, t in r["trace"][m]:
            assert f != t and not (f >= N and t >= N), (m, f, t)
            if f < N:
                assert occ[f], (m, f, t)
                occ[f] = False
            else:
                eo[f - N] -= 1
            if t < N:
                assert not occ[t], (m, f, t)
                occ[t] = True
            else:
                eo[t - N] += 1
        np.testing.assert_array_equal(r["occupation"][m].astype(bool), occ)
        np.testing.assert_array_equal(r["electrode_occupation"][m], eo)


@pytest.mark.parametrize("runs", ["of16", "distinct", "identical", "of4_ragged", "of5"])
def test_lanes_table_is_transparent(golden_py, fixtures_subset, runs):
    """The table memoises a pure function of (parameters, occupation): with it disabled (memo=False: every hop of
    every trajectory is evaluated from scratch) trace, time, tallies, occupation and energies are bit-identical --
    for runs of 16 seeds per voltage vector (shared tables), all-distinct members (one table each), 32 identical
    members, a ragged ensemble with runs of 4, and runs of 5 (runs may start anywhere in a warp's 32 members)."""
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "n5_p3_hot": golden_py["n5_p3_hot"],
             "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"], "c1_basic_N10_P2": golden_py["c1_basic_N10_P2"],
             "XOR_wide/test1": _fixture_case(fixtures_subset["XOR_wide/test1"]),
             "N31_P5": synthetic_layout(31, 5, 7), "N25_P0": synthetic_layout(25, 0, 8, fill=0.5)}
    for name, c in cases.items():
        B, hops, P = {"of16": 96, "distinct": 64, "identical": 64, "of4_ragged": 77, "of5": 80}[runs], 3000, c["P"]
        rep = {"of16": 16, "distinct": 1, "identical": B, "of4_ragged": 4, "of5": 5}[runs]
        V = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) // rep)[:, None] * 0.37
        E = np.tile(c["E_constant"], (B, 1)) + (np.arange(B) // rep)[:, None] * 0.11
        kT = c["kT"] * (1.0 + 0.05 * ((np.arange(B) // rep) % 3))
        lay = _layout(c)
        kw = dict(E_constant=E, occupation0=c["occupation"], prehops=300, seed=21, trace=True, want_occupation=True,
                  want_site_energies=True, want_misses=True, kernel="lanes")
        a = lay.run(hops, kT, V, memo=True, **kw)
        b = lay.run(hops, kT, V, memo=False, **kw)
        lay.close()
        np.testing.assert_array_equal(a["trace"], b["trace"], err_msg=name)
        np.testing.assert_array_equal(a["time"], b["time"], err_msg=name)
        np.testing.assert_array_equal(a["electrode_occupation"], b["electrode_occupation"], err_msg=name)
        np.testing.assert_array_equal(a["occupation"], b["occupation"], err_msg=name)
        np.testing.assert_array_equal(a["site_energies"], b["site_energies"], err_msg=name)
        alive = np.isfinite(a["time"])
        assert alive.any(), name
        # the table really is used: fewer evaluations than hops
        assert (a["misses"] <= b["misses"]).all(), name
        if name in ("fx_rnd_min_max_0", "n5_p3_hot", "c2_grid_N16_P8"):
            assert a["misses"][alive].mean() < 0.7 * b["misses"][alive].mean(), (name, a["misses"].mean(), b["misses"].mean())


def test_lanes_trace_is_valid_on_random_layouts():
    """30 random layouts (N = 1..31, P = 0..9, random filling / temperature / interaction strength), ragged ensembles
    with runs of identical members: every traced hop is allowed, tallies and the final occupation follow from the trace,
    and site energies equal a from-scratch evaluation of the final state."""
    rng = np.random.default_rng(2027)
    for it in range(30):
        N = int(rng.choice([1, 2, 3, 7, 10, 11, 15, 16, 17, 24, 25, 30, 31]))
        P = int(rng.integers(0, 10))
        if P == 0 and N < 2:
            P = 1
        c = synthetic_layout(N, P, 300 + it, kT=float(rng.choice([0.5, 1.0, 4.0])), I_0=float(rng.choice([0.0, 30.0, 100.0])),
                             fill=float(rng.uniform(0.1, 0.9)))
        if P == 0 and (c["occupation"].all() or not c["occupation"].any()):
            c["occupation"][0] = not c["occupation"][0]
        B, hops = int(rng.integers(1, 70)), 600
        rep = int(rng.choice([1, 2, 8]))
        V = np.tile(c["electrode_v"], (B, 1)) + rng.normal(0, 5, ((B + rep - 1) // rep, P)).repeat(rep, axis=0)[:B]
        E = np.tile(c["E_constant"], (B, 1))
        lay = _layout(c)
        r = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=it, trace=True, want_occupation=True,
                    want_site_energies=True, kernel="lanes")
        _check_trace(c, r, c["occupation"])
        for m in range(min(B, 3)):
            if np.isfinite(r["time"][m]):
                se, _ = lay.probe_rates(c["E_constant"], V[m], c["kT"], r["occupation"][m])
                np.testing.assert_array_equal(r["site_energies"][m].astype(np.float32), se, err_msg=str((it, N, P)))
        lay.close()


def test_lanes_one_hop_event_distribution_matches_oracle_rates(golden_py, fixtures_subset):
    """Order-independent check of the pick: from one fixed state, 2^18 members take ONE hop each; the empirical
    distribution over (from,to) must match rate_ij / sum(rate) of the oracle (chi-square), no disallowed pair may
    ever be chosen, and the mean dwell time is 1 / total rate."""
    from oracle import oracle
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"],
             "n5_p3_hot": golden_py["n5_p3_hot"], "XOR_wide/test3": _fixture_case(fixtures_subset["XOR_wide/test3"]),
             "c1_basic_N10_P2": golden_py["c1_basic_N10_P2"]}
    B = 1 << 18
    for name, c in cases.items():
        S = c["N"] + c["P"]
        _, r_o = oracle.go_rates(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["occupation"], c["distances"],
                                 c["E_constant"], c["transitions_constant"], site_energies_of(c))
        p = r_o.astype(np.float64).ravel(); p /= p.sum()
        lay = _layout(c)
        r = lay.run(1, c["kT"], np.tile(c["electrode_v"], (B, 1)), E_constant=np.tile(c["E_constant"], (B, 1)),
                    occupation0=c["occupation"], seed=78, trace=True, kernel="lanes")
        lay.close()
        ev = r["trace"][:, 0, 0].astype(np.int64) * S + r["trace"][:, 0, 1]
        cnt = np.bincount(ev, minlength=S * S).astype(np.float64)
        assert cnt[p == 0].sum() == 0, name
        big = p * B >= 20
        obs = np.append(cnt[big], cnt[~big].sum()); exp = np.append(p[big] * B, p[~big].sum() * B)
        keep = exp > 0
        chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
        dof = keep.sum() - 1
        assert chi2 < dof + 5 * np.sqrt(2 * dof) + 5, (name, chi2, dof)
        total = r_o.astype(np.float64).sum()
        assert r["time"].mean() * total == pytest.approx(1.0, abs=5 / np.sqrt(B))


def test_lanes_currents_agree_with_warp_kernel_and_oracle(golden_py):
    """Ensemble-averaged currents and elapsed time: 2048 members on the lanes kernel vs 2048 on the warp-per-trajectory
    kernel (other seeds), and 256 oracle members (simulateRecordPlus semantics, own RNG); two-sample z per electrode
    < 4.5."""
    from oracle import oracle
    for name in ("fx_rnd_min_max_0", "c2_grid_N16_P8", "n5_p3_hot"):
        c = golden_py[name]
        B, hops = 2048, 20000
        E = np.tile(c["E_constant"], (B, 1)); V = np.tile(c["electrode_v"], (B, 1))
        lay = _layout(c)
        g = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=5, kernel="lanes")
        w = lay.run(hops, c["kT"], V, E_constant=E, occupation0=c["occupation"], seed=6, kernel="warp")
        lay.close()
        Bo = 256
        o = oracle.go_ensemble(c["N"], c["P"], c["nu"], c["kT"], c["I_0"], c["R"], c["distances"], E[:Bo],
                               c["transitions_constant"], V[:Bo], hops, variant=1, occupation0=c["occupation"], seed0=99)
        cg = g["electrode_occupation"] / g["time"][:, None]
        cw = w["electrode_occupation"] / w["time"][:, None]
        co = o["electrode_occupation"] / o["time"][:, None]
        for other in (cw, co):
            z = np.abs(cg.mean(0) - other.mean(0)) / np.sqrt(cg.var(0) / len(cg) + other.var(0) / len(other) + 1e-300)
            assert (z < 4.5).all(), (name, z)
        zt = abs(g["time"].mean() - w["time"].mean()) / np.sqrt(g["time"].var() / B + w["time"].var() / B)
        assert zt < 4.5, (name, zt)


def test_lanes_results_do_not_depend_on_batching_or_kernel_geometry(golden_py):
    """Member m draws from Philox stream (seed, member_index0+m) and the table is invisible: splitting an ensemble (as
    ranks do) anywhere -- also inside a run of identical members -- changes nothing."""
    c = golden_py["fx_xor_wide_3"]
    B = 100
    E = np.tile(c["E_constant"], (B, 1)); V = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) // 8)[:, None]
    lay = _layout(c)
    kw = dict(seed=9, kernel="lanes")
    a = lay.run(3000, c["kT"], V, E_constant=E, **kw)
    b1 = lay.run(3000, c["kT"], V[:13], E_constant=E[:13], member_index0=0, **kw)
    b2 = lay.run(3000, c["kT"], V[13:], E_constant=E[13:], member_index0=13, **kw)
    np.testing.assert_array_equal(a["time"], np.concatenate([b1["time"], b2["time"]]))
    np.testing.assert_array_equal(a["electrode_occupation"], np.concatenate([b1["electrode_occupation"], b2["electrode_occupation"]]))
    # An ensemble that leaves warp slots empty gives every block of 32 members two warps and splits it where a run starts at
    # member 16 (hop_lanes.cu, "halves"): blocks that split (runs of 8 or 16), blocks that do not (runs of 32, of 24: member
    # 16 is inside a run), a ragged last block -- bit-identical with one warp per block (KMCB200_LANES_HALVES=0).
    import os
    for rep, Bh in ((8, 100), (16, 112), (32, 96), (24, 120), (1, 49), (16, 17)):
        Vh = np.tile(c["electrode_v"], (Bh, 1)) + (np.arange(Bh) // rep)[:, None]
        Eh = np.tile(c["E_constant"], (Bh, 1))
        res = []
        for h in ("1", "0"):
            os.environ["KMCB200_LANES_HALVES"] = h
            try:
                res.append(lay.run(3000, c["kT"], Vh, E_constant=Eh, **kw))
            finally:
                os.environ.pop("KMCB200_LANES_HALVES", None)
        for k in res[0]:
            if isinstance(res[0][k], np.ndarray):
                np.testing.assert_array_equal(res[0][k], res[1][k], err_msg=f"runs of {rep}, {Bh} members: {k}")
        assert np.isfinite(res[0]["time"]).all() and (res[0]["time"] > 0).all()
    lay.close()


def test_lanes_superposition_and_prehops(fixtures_subset):
    """basis + electrode voltages (the on-device mat-vec) give exactly what explicit E_constant gives; prehops restart
    the tallies but keep the occupation: a run of prehops + hops equals the tail of ... the same stream."""
    f = fixtures_subset["rnd_min_max/test2"]
    c = _fixture_case(f)
    N, P, B = c["N"], c["P"], 48
    rng = np.random.default_rng(1)
    # (values on a binary grid: the fp64 sums are exact in any order, so host and device agree to the bit)
    basis = np.round(np.vstack([rng.normal(0, 0.3, (P, N)), c["E_constant"][None] * 0.5]) * 64) / 64
    V = np.round(np.tile(c["electrode_v"], (B, 1)) * 4) / 4 + (np.arange(B) // 16)[:, None] * 2.0
    E = basis[P][None] + V @ basis[:P]
    lay = _layout(c)
    a = lay.run(4000, c["kT"], V, basis=basis, prehops=1000, seed=4, want_occupation=True, kernel="lanes")
    b = lay.run(4000, c["kT"], V, E_constant=E, prehops=1000, seed=4, want_occupation=True, kernel="lanes")
    # prehops: the occupation after 1000 + 4000 hops is that of a plain 5000-hop run of the same stream
    d = lay.run(5000, c["kT"], V, E_constant=E, seed=4, want_occupation=True, trace=True, kernel="lanes")
    lay.close()
    np.testing.assert_array_equal(a["occupation"], b["occupation"])
    np.testing.assert_array_equal(a["electrode_occupation"], b["electrode_occupation"])
    np.testing.assert_array_equal(a["time"], b["time"])
    np.testing.assert_array_equal(b["occupation"], d["occupation"])
    tail = d["trace"][:, 1000:]
    eo = np.zeros((B, P), dtype=np.int64)
    for e in range(P):
        eo[:, e] = (tail[:, :, 1] == N + e).sum(1) - (tail[:, :, 0] == N + e).sum(1)
    np.testing.assert_array_equal(b["electrode_occupation"], eo)


def test_lanes_currents_pass_reference_acceptance(fixtures_subset):
    """The reference's own acceptance test (thesis_indrek/validate_tests.py:80-135) on the lanes kernel: 5 runs per
    fixture at the fixture's own run length, per-electrode Bhattacharyya distance against the stored 5-run mean/stddev."""
    Ds, rel = [], []
    for name, f in fixtures_subset.items():
        c = _fixture_case(f)
        hops = 5_000_000 if "5M" in name else 1_000_000
        lay = _layout(c)
        r = lay.run(hops, c["kT"], np.tile(c["electrode_v"], (5, 1)), E_constant=np.tile(c["E_constant"], (5, 1)),
                    seed=12, kernel="lanes")
        lay.close()
        cur = r["current"]
        mu, sd = cur.mean(0), cur.std(0)
        Ds.append(np.array([calc_D(f["mean_currents"][i], mu[i], f["stddev_currents"][i], sd[i]) for i in range(len(mu))]))
        ref = np.asarray(f["mean_currents"]); big = np.abs(ref) > 0.05 * np.abs(ref).max()
        rel.append(np.abs(mu[big] - ref[big]) / np.abs(ref[big]))
    Ds = np.concatenate(Ds); rel = np.concatenate(rel)
    assert Ds.mean() < 0.9, Ds.mean()
    assert (Ds > 0.9).mean() < 0.25, (Ds > 0.9).mean()
    assert np.median(rel) < 0.02 and rel.max() < 0.15, (np.median(rel), rel.max())


def test_lanes_edge_cases(golden_py):
    """hops = 0, a single member, a dead state (closed system with nothing to do), layouts the kernel cannot take are refused."""
    c = golden_py["fx_rnd_min_max_0"]
    lay = _layout(c)
    r = lay.run(0, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], occupation0=c["occupation"],
                want_occupation=True, kernel="lanes")
    assert r["time"][0] == 0.0 and (r["electrode_occupation"] == 0).all()
    np.testing.assert_array_equal(r["occupation"][0], c["occupation"])
    r = lay.run(1000, c["kT"], c["electrode_v"][None], E_constant=c["E_constant"][None], seed=3, kernel="lanes")
    assert np.isfinite(r["time"][0]) and r["time"][0] > 0
    lay.close()
    big = synthetic_layout(40, 4, 3)  # more than 31 acceptors: one mask word is not enough
    lay = _layout(big)
    with pytest.raises(RuntimeError):
        lay.run(10, big["kT"], big["electrode_v"][None], E_constant=big["E_constant"][None], kernel="lanes")
    lay.close()
    d = synthetic_layout(6, 0, 5)
    d["occupation"][:] = True  # full, no electrodes: no transition possible
    lay = _layout(d)
    r = lay.run(50, d["kT"], np.zeros((3, 0)), E_constant=np.tile(d["E_constant"], (3, 1)), occupation0=d["occupation"],
                want_occupation=True, kernel="lanes")
    lay.close()
    assert np.isinf(r["time"]).all() and r["occupation"].all()


def test_kernel_selection(golden_py):
    """The library picks the thread-per-trajectory kernel for large ensembles of small layouts and the warp-per-
    trajectory kernels otherwise (kmcb200_last_kernel); results of the two agree statistically (tested above)."""
    from kmc_dn_b200.ensemble import last_kernel
    c = golden_py["fx_rnd_min_max_0"]
    lay = _layout(c)
    for B, want in ((70000, "kmc_lanes_kernel"), (13000, "kmc_lanes_kernel"), (3000, "kmc_memo_kernel"), (1000, "kmc_solo_kernel"), (100, "kmc_solo_kernel")):
        r = lay.run(50, c["kT"], np.tile(c["electrode_v"], (B, 1)), E_constant=np.tile(c["E_constant"], (B, 1)), seed=1)
        assert last_kernel() == want, (B, last_kernel())
        assert np.isfinite(r["time"]).all()
    lay.run(50, c["kT"], np.tile(c["electrode_v"], (70000, 1)), E_constant=np.tile(c["E_constant"], (70000, 1)), seed=1,
            kernel="warp")
    assert last_kernel() == "kmc_memo_kernel"
    r = lay.run(50, c["kT"], np.tile(c["electrode_v"], (70000, 1)), E_constant=np.tile(c["E_constant"], (70000, 1)), seed=1,
                record=True)  # a large record=True ensemble stays on the thread-per-trajectory kernel
    assert last_kernel() == "kmc_lanes_kernel"
    assert np.isfinite(r["avg_occupation"]).all() and (r["traffic"] == -r["traffic"].transpose(0, 2, 1)).all()
    lay.close()
    d = synthetic_layout(40, 4, 3)
    lay = _layout(d)
    lay.run(20, d["kT"], np.tile(d["electrode_v"], (70000, 1)), E_constant=np.tile(d["E_constant"], (70000, 1)), seed=1)
    assert last_kernel() == "kmc_wide_kernel"
    lay.close()


def test_lanes_record_outputs(golden_py, fixtures_subset):
    """record=True on the thread-per-trajectory kernel (simulation.go:309-317): traffic is the antisymmetrised histogram of
    the traced hops, the occupied times add up to (holes x time), both are bit-identical with the table off, restart at the
    prehops boundary, and agree with the warp-per-trajectory kernel's tallies within statistical error."""
    for name, c in (("fx_rnd_min_max_0", golden_py["fx_rnd_min_max_0"]), ("c2_grid_N16_P8", golden_py["c2_grid_N16_P8"]),
                    ("XOR_wide/test1", _fixture_case(fixtures_subset["XOR_wide/test1"]))):
        N, P = c["N"], c["P"]
        S = N + P
        B, hops, pre = 48, 4000, 500
        V = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) // 16)[:, None] * 0.5
        E = np.tile(c["E_constant"], (B, 1))
        lay = _layout(c)
        kw = dict(E_constant=E, occupation0=c["occupation"], prehops=pre, seed=33, record=True, trace=True, want_occupation=True,
                  kernel="lanes")
        a = lay.run(hops, c["kT"], V, memo=True, **kw)
        b = lay.run(hops, c["kT"], V, memo=False, **kw)
        for k in ("traffic", "avg_occupation", "time", "trace", "electrode_occupation"):
            np.testing.assert_array_equal(a[k], b[k], err_msg=f"{name} {k}")
        # traffic from the trace
        tr = np.zeros((B, S, S))
        for m in range(B):
            np.add.at(tr[m], (a["trace"][m, :, 0], a["trace"][m, :, 1]), 1.0)
        np.testing.assert_array_equal(a["traffic"], tr - tr.transpose(0, 2, 1), err_msg=name)
        # occupied time: sum over sites = integral of the number of holes; every site within [0, time]
        ao = a["avg_occupation"]
        assert (ao >= -1e-9 * a["time"][:, None]).all() and (ao <= a["time"][:, None] * (1 + 1e-9)).all(), name
        # against the warp-per-trajectory kernel (other streams): mean occupation per site
        Bs = 512
        Vs = np.tile(c["electrode_v"], (Bs, 1)); Es = np.tile(c["E_constant"], (Bs, 1))
        g = lay.run(6000, c["kT"], Vs, E_constant=Es, occupation0=c["occupation"], prehops=pre, seed=5, record=True, kernel="lanes")
        w = lay.run(6000, c["kT"], Vs, E_constant=Es, occupation0=c["occupation"], prehops=pre, seed=6, record=True, kernel="warp")
        lay.close()
        og, ow = g["avg_occupation"] / g["time"][:, None], w["avg_occupation"] / w["time"][:, None]
        z = np.abs(og.mean(0) - ow.mean(0)) / np.sqrt(og.var(0) / Bs + ow.var(0) / Bs + 1e-12)
        assert (z < 5).all(), (name, z)
        # number of holes: sum of the occupied times / time stays within [0, N] and matches between the kernels
        assert abs(og.sum(1).mean() - ow.sum(1).mean()) < 0.05 * N, name
        tg, tw = g["traffic"].mean(0), w["traffic"].mean(0)
        sd = np.sqrt(g["traffic"].var(0) / Bs + w["traffic"].var(0) / Bs) + 1e-9
        assert (np.abs(tg - tw) < 5.5 * sd + 0.02).all(), name


def test_lanes_injected_stream(golden_py):
    """Injected variates (simulation.go:297-299: e ~ Exp(1) for the dwell time, u ~ U[0,1) for the pick) on the thread-per-
    trajectory kernel: bit-identical with the table off, every hop allowed, and the elapsed time is sum e_k / total rate of
    the state before hop k (oracle rates, fp32: 1e-5)."""
    from oracle import oracle
    c = golden_py["fx_rnd_min_max_0"]
    N, P = c["N"], c["P"]
    B, hops = 6, 60
    rng = np.random.default_rng(8)
    e = rng.exponential(size=(B, hops)); u = rng.random((B, hops)).astype(np.float32)
    V = np.tile(c["electrode_v"], (B, 1)); E = np.tile(c["E_constant"], (B, 1))
    lay = _layout(c)
    kw = dict(E_constant=E, occupation0=c["occupation"], stream_e=e, stream_u=u, trace=True, want_occupation=True, kernel="lanes")
    a = lay.run(hops, c["kT"], V, memo=True, **kw)
    b = lay.run(hops, c["kT"], V, memo=False, **kw)
    lay.close()
    np.testing.assert_array_equal(a["trace"], b["trace"])
    np.testing.assert_array_equal(a["time"], b["time"])
    _check_trace(c, a, c["occupation"].copy())
    for m in range(B):
        occ = c["occupation"].copy()
        t = 0.0
        for k, (f, to) in enumerate(a["trace"][m]):
            se = site_energies_of(c)
            _, rates = oracle.go_rates(N, P, c["nu"], c["kT"], c["I_0"], c["R"], occ, c["distances"], c["E_constant"],
                                       c["transitions_constant"], se)
            t += e[m, k] / rates.astype(np.float64).sum()
            if f < N:
                occ[f] = False
            if to < N:
                occ[to] = True
        assert a["time"][m] == pytest.approx(t, rel=2e-5), (m, a["time"][m], t)


def test_solo_kernel_is_bit_identical_with_lanes_kernel(golden_py, fixtures_subset):
    """The latency kernel (a few trajectories, one warp each, the visited states as a graph in shared memory: hop_lanes.cu,
    kmc_solo_kernel) builds its entries with the thread-per-trajectory kernel's evaluation, draws the same variates and
    resolves the tail with the same exact pick: time, tallies, occupation and energies are bit-identical -- on layouts
    of 5 to 31 acceptors, 0 to 8 electrodes, with prehops, with a table of 3 or 40 entries that keeps being dropped, on 24
    random layouts / fillings / temperatures, on a dead state, for hops = 0, and for more members than CTAs."""
    import os
    from kmc_dn_b200.ensemble import last_kernel
    cases = {"fx_rnd_min_max_0": golden_py["fx_rnd_min_max_0"], "n5_p3_hot": golden_py["n5_p3_hot"],
             "c2_grid_N16_P8": golden_py["c2_grid_N16_P8"], "c1_basic_N10_P2": golden_py["c1_basic_N10_P2"],
             "XOR_wide/test1": _fixture_case(fixtures_subset["XOR_wide/test1"]),
             "N31_P5": synthetic_layout(31, 5, 7), "N25_P0": synthetic_layout(25, 0, 8, fill=0.5)}
    for name, c in cases.items():
        B, P = 6, c["P"]
        V = np.tile(c["electrode_v"], (B, 1)) + np.arange(B)[:, None] * (1.0 if P else 0.0)
        E = np.tile(c["E_constant"], (B, 1))
        lay = _layout(c)
        for hops, prehops, emax in ((4000, 0, None), (700, 300, None), (1500, 0, "3"), (1500, 100, "40"), (0, 0, None)):
            kw = dict(E_constant=E, occupation0=c["occupation"], seed=21, prehops=prehops, want_occupation=True,
                      want_site_energies=True)
            ref = lay.run(hops, c["kT"], V, kernel="lanes", **kw)
            assert last_kernel() == "kmc_lanes_kernel"
            if emax:
                os.environ["KMCB200_SOLO_EMAX"] = emax
            try:
                got = lay.run(hops, c["kT"], V, kernel="solo", **kw)
            finally:
                os.environ.pop("KMCB200_SOLO_EMAX", None)
            assert last_kernel() == "kmc_solo_kernel"
            for k in ("time", "electrode_occupation", "occupation", "site_energies"):
                np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{name} hops={hops} prehops={prehops} emax={emax}: {k}")
        lay.close()
    # random layouts, fillings and temperatures
    rng = np.random.default_rng(77)
    for t in range(24):
        N, P = int(rng.integers(3, 32)), int(rng.integers(0, 9))
        c = synthetic_layout(N, P, 100 + t, fill=float(rng.uniform(0.15, 0.9)))
        B = 3
        V = np.tile(c["electrode_v"], (B, 1)) + rng.normal(0, 5, size=(B, P))
        E = np.tile(c["E_constant"], (B, 1))
        kT = c["kT"] * float(rng.choice([0.3, 1.0, 4.0]))
        lay = _layout(c)
        kw = dict(E_constant=E, occupation0=c["occupation"], seed=int(rng.integers(1 << 30)), prehops=int(rng.choice([0, 130])),
                  want_occupation=True, want_site_energies=True)
        ref = lay.run(2500, kT, V, kernel="lanes", **kw)
        if t % 2:
            os.environ["KMCB200_SOLO_EMAX"] = "7"
        try:
            got = lay.run(2500, kT, V, kernel="solo", **kw)
        finally:
            os.environ.pop("KMCB200_SOLO_EMAX", None)
        lay.close()
        for k in ("time", "electrode_occupation", "occupation", "site_energies"):
            np.testing.assert_array_equal(got[k], ref[k], err_msg=f"random layout {t} (N={N}, P={P}): {k}")
    # more members than CTAs (two per SM): the CTAs loop
    c = golden_py["fx_rnd_min_max_0"]
    B = 700
    V = np.tile(c["electrode_v"], (B, 1)) + (np.arange(B) % 50)[:, None]
    E = np.tile(c["E_constant"], (B, 1))
    lay = _layout(c)
    ref = lay.run(600, c["kT"], V, E_constant=E, seed=4, kernel="lanes")
    got = lay.run(600, c["kT"], V, E_constant=E, seed=4, kernel="solo")
    lay.close()
    np.testing.assert_array_equal(got["time"], ref["time"])
    np.testing.assert_array_equal(got["electrode_occupation"], ref["electrode_occupation"])
    # a dead state
    d = synthetic_layout(6, 0, 5)
    d["occupation"][:] = True
    lay = _layout(d)
    r = lay.run(50, d["kT"], np.zeros((3, 0)), E_constant=np.tile(d["E_constant"], (3, 1)), occupation0=d["occupation"],
                want_occupation=True, kernel="solo")
    lay.close()
    assert np.isinf(r["time"]).all() and r["occupation"].all()
