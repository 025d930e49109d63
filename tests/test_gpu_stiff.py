"""The Rosenbrock-23 path (odl_*_ros23_kernel, solver='auto') on the stiff host-virus variant
(BASELINE.json config 4: rates spanning 6 orders of magnitude) against scipy odeint (LSODA -> BDF)."""
import numpy as np
import pytest

from oracle import odelib_oracle as orc
from tests.helpers import STATES, SUMS, demo_df, device_model, golden, obs_tables_from_oracle, oracle_rhs

pytestmark = pytest.mark.gpu


def stiff_thetas(n, seed=0):
    """two_i with tau ~ lognorm(0.5, 1e4), lam ~ lognorm(0.5, 1e-2); mu, phi, beta around (0.5, 1e-7, 50)."""
    rng = np.random.default_rng(seed)
    th = np.empty((n, 5))
    th[:, 0] = 0.5 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 1] = 1e-7 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 2] = 50.0 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 3] = 1e-2 * np.exp(0.5 * rng.standard_normal(n))
    th[:, 4] = 1e4 * np.exp(0.5 * rng.standard_normal(n))
    return th


def test_ros23_matches_lsoda_on_stiff_systems():
    dm, tab = device_model("two_i")
    theta = stiff_thetas(32)
    out = dm.sweep(theta, solver="ros23", return_pred=True, max_steps=2000000)
    assert np.all(out["status"] == 0)
    rhs = oracle_rhs("two_i")
    for k in range(len(theta)):
        vec, chi, r2 = orc.solve_unit(rhs, theta[k], tab, 1e-11, 1e-11, mxstep=500000)
        # order-2 method at rtol=atol=1.49e-8: global error of a few 1e-7 .. 1e-6 relative
        np.testing.assert_allclose(out["pred"][k], vec, rtol=2e-5, atol=1e-3)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=2e-4)


def test_ros23_tight_tolerance_converges_to_lsoda():
    dm, tab = device_model("two_i")
    theta = stiff_thetas(4, seed=1)
    out = dm.sweep(theta, solver="ros23", rtol=1e-11, atol=1e-11, return_pred=True, max_steps=5000000)
    assert np.all(out["status"] == 0)
    rhs = oracle_rhs("two_i")
    for k in range(len(theta)):
        vec, chi, _ = orc.solve_unit(rhs, theta[k], tab, 1e-13, 1e-13, mxstep=500000)
        np.testing.assert_allclose(out["pred"][k], vec, rtol=2e-7, atol=1e-4)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=2e-6)


def test_ros23_agrees_with_dopri5_on_nonstiff_systems():
    g = golden("two_i")
    dm, _ = device_model("two_i")
    theta = g["theta"][:8]
    a = dm.sweep(theta, solver="dopri5", rtol=1e-10, atol=1e-10, return_pred=True)
    b = dm.sweep(theta, solver="ros23", rtol=1e-10, atol=1e-10, return_pred=True, max_steps=5000000)
    ok = (a["status"] == 0) & (b["status"] == 0) & np.all(a["pred"] > 1.0, axis=1)
    assert ok.sum() >= 5
    np.testing.assert_allclose(b["pred"][ok], a["pred"][ok], rtol=5e-6)


def test_radau5_matches_lsoda_on_stiff_systems():
    dm, tab = device_model("two_i")
    theta = stiff_thetas(32, seed=3)
    out = dm.sweep(theta, solver="radau5", return_pred=True)
    assert np.all(out["status"] == 0) and out["nsteps"].max() < 3000
    rhs = oracle_rhs("two_i")
    for k in range(len(theta)):
        vec, chi, r2 = orc.solve_unit(rhs, theta[k], tab, 1e-11, 1e-11, mxstep=500000)
        np.testing.assert_allclose(out["pred"][k], vec, rtol=1e-6, atol=1e-3)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=2e-5)
    tight = dm.sweep(theta[:4], solver="radau5", rtol=1e-12, atol=1e-12, return_pred=True)
    for k in range(4):
        vec, chi, _ = orc.solve_unit(rhs, theta[k], tab, 1e-13, 1e-13, mxstep=500000)
        # RADAU5 works at rtol' = 0.1 rtol^(2/3) internally (1e-9 for 1e-12)
        np.testing.assert_allclose(tight["pred"][k], vec, rtol=5e-7, atol=1e-4)
        np.testing.assert_allclose(tight["chi"][k], chi, rtol=5e-6)


def test_auto_routes_stiff_systems_and_keeps_the_rest():
    dm, tab = device_model("two_i")
    g = golden("two_i")
    theta = np.vstack([stiff_thetas(24, seed=2), g["theta"][:24]])
    plain = dm.sweep(theta, solver="dopri5", stiff_check=True, max_steps=2000000)
    assert (plain["status"][:24] == 4).sum() >= 12            # DOPRI5 alone flags (most of) the stiff block
    from odelib_b200 import _capi
    auto0 = dm.sweep(theta, solver="auto", max_steps=2000000, auto_flags=_capi.AUTO_NO_HANDOVER)
    assert np.all(auto0["status"][:24] == 0)
    # stiff pass from t0 (AUTO_NO_HANDOVER): every system was finished by exactly one of the two steppers (DOPRI5, or the
    # BDF stiff pass): same kernel, same numbers -- a solve does not depend on which systems share its warp or which pass
    # it ran in
    bdf = dm.sweep(theta, solver="bdf", max_steps=2000000)
    dop = dm.sweep(theta, solver="dopri5", max_steps=2000000)
    same_bdf = (auto0["chi"] == bdf["chi"]) | (np.isnan(auto0["chi"]) & np.isnan(bdf["chi"]))
    same_dop = (auto0["chi"] == dop["chi"]) | (np.isnan(auto0["chi"]) & np.isnan(dop["chi"]))
    assert np.all(same_bdf | same_dop)
    assert same_bdf[:24].sum() >= 12 and same_dop[24:].sum() >= 20
    # the default: the stiff pass CONTINUES from where the DOPRI5 pass stopped (t, y and the observation columns staged so
    # far travel with the row).  Rows DOPRI5 finished are untouched; the others agree with the BDF solve from t0 to
    # solver accuracy and take fewer BDF steps
    auto = dm.sweep(theta, solver="auto", max_steps=2000000)
    assert np.all(auto["status"][:24] == 0)
    assert np.array_equal(auto["chi"][same_dop], auto0["chi"][same_dop])
    handed = same_bdf & ~same_dop
    np.testing.assert_allclose(auto["chi"][handed], auto0["chi"][handed], rtol=5e-3)   # (BDF's chi at the default tolerance: ~1e-4)
    assert auto["nsteps"][handed].sum() < auto0["nsteps"][handed].sum()
    # far fewer steps than the explicit method needs on the stiff block
    assert np.median(auto["nsteps"][:24]) < 0.2 * np.median(dop["nsteps"][:24])
    # the Radau5 stiff pass is still selectable and agrees with the BDF one to solver accuracy
    rad = dm.sweep(theta, solver="auto", tail_solver="radau5", max_steps=2000000)
    np.testing.assert_allclose(rad["chi"][:24], auto["chi"][:24], rtol=2e-5)


def test_bdf_matches_lsoda_on_stiff_systems():
    """The stiff pass of the auto sweep: variable-order BDF, the method family of LSODA's stiff branch."""
    dm, tab = device_model("two_i")
    theta = stiff_thetas(32, seed=3)
    out = dm.sweep(theta, solver="bdf", return_pred=True)
    assert np.all(out["status"] == 0) and out["nsteps"].max() < 6000
    rhs = oracle_rhs("two_i")
    for k in range(len(theta)):
        vec, chi, r2 = orc.solve_unit(rhs, theta[k], tab, 1e-11, 1e-11, mxstep=500000)
        np.testing.assert_allclose(out["pred"][k], vec, rtol=5e-6, atol=1e-3)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=1e-4)
    again = dm.sweep(theta[::-1].copy(), solver="bdf")
    assert np.array_equal(again["chi"][::-1], out["chi"], equal_nan=True)      # lane placement does not matter
    tight = dm.sweep(theta[:4], solver="bdf", rtol=1e-11, atol=1e-11, return_pred=True, max_steps=2000000)
    for k in range(4):
        vec, chi, _ = orc.solve_unit(rhs, theta[k], tab, 1e-13, 1e-13, mxstep=500000)
        np.testing.assert_allclose(tight["pred"][k], vec, rtol=2e-7, atol=1e-4)


def test_mcmc_on_the_stiff_variant_ros23_and_auto():
    """Config 4: chains on synthetic stiff data; ROS23, BDF and auto (per solve: DOPRI5 until Hairer's test calls the solve
    stiff, then the same solve on BDF) make the same decisions on host streams."""
    from odelib_b200 import demo_models
    from odelib_b200.engine import DeviceModel
    center = np.array([0.5, 1e-7, 50.0, 1e-2, 1e4])
    tab0 = orc.build_tables(demo_df("two_i"), STATES["two_i"], SUMS["two_i"], 1000, {"S": 5236900})
    # synthetic data: the reference odeint at the centre + log-normal noise (seed 0) on the demo's time points
    vec, _, _ = orc.solve_unit(orc.two_i, center, tab0, 1e-12, 1e-12, mxstep=500000)
    rng = np.random.default_rng(0)
    df = demo_df("two_i").sort_values(by=["organism", "time"]).reset_index(drop=True)
    df["abundance"] = vec * np.exp(0.2 * rng.standard_normal(len(vec)))
    df["log_sigma"] = 0.2
    tab = orc.build_tables(df, STATES["two_i"], SUMS["two_i"], 1000, {"S": 5236900, "V": 10981000})
    f, n, P, groups = demo_models.MODELS["two_i"]
    dm = DeviceModel(f, n, P, groups)
    dm.set_data(obs_tables_from_oracle(tab), tab.y0)
    C, nits = 8, 40
    starts = center * np.exp(0.02 * rng.standard_normal((C, 5)))
    z = 0.05 * rng.standard_normal((C, nits - 1, 5))
    u = rng.random((C, nits - 1))
    a = dm.mcmc(starts, nits=nits, rng_mode="host", z=z, u=u, solver="ros23", trace=True, max_steps=2000000)
    b = dm.mcmc(starts, nits=nits, rng_mode="host", z=z, u=u, solver="auto", trace=True, max_steps=2000000)
    assert np.isfinite(a["chinew"]).all() and a["fail_count"].sum() == 0
    # auto = DOPRI5 until Hairer's test calls a proposal stiff, then BDF for that solve: same chi to solver accuracy
    # (BDF at the default tolerance: chi within 1e-4 of the tight solve, see test_bdf_sweep_matches_odeint)
    np.testing.assert_allclose(b["chinew"], a["chinew"], rtol=2e-4)
    assert (a["accepted"] != b["accepted"]).sum() <= 2 and b["fail_count"].sum() == 0
    # every solve here is stiff: the solves auto hands to BDF are the BDF kernel's solves, bit for bit
    d = dm.mcmc(starts, nits=nits, rng_mode="host", z=z, u=u, solver="bdf", trace=True, max_steps=2000000)
    same = b["chinew"] == d["chinew"]                             # (the rest DOPRI5 finished itself, stiff or not)
    assert same.mean() > 0.5, same.mean()
    np.testing.assert_allclose(b["chinew"], d["chinew"], rtol=2e-4)
    ref = orc.mh_chain(orc.two_i, starts[0], tab, 5, nits=nits, z=z[0], u=u[0], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(a["chinew"][0], ref["chinew"], rtol=5e-4)
    c = dm.mcmc(starts, nits=nits, rng_mode="host", z=z, u=u, solver="radau5", trace=True)
    np.testing.assert_allclose(c["chinew"][0], ref["chinew"], rtol=2e-5)
    assert c["step_count"].sum() < 0.5 * a["step_count"].sum()


def test_facade_picks_the_bdf_kernel_for_a_stiff_posterior():
    """ModelFramework(solver='auto') (the default): chains whose starts the capped DOPRI5 pass cannot finish run on the
    BDF kernel -- the drop-in needs no hint for config 4's stiff variant; the demo's posterior stays on DOPRI5."""
    import scipy.stats
    import odelib_b200 as ODElib
    from odelib_b200 import demo_models
    center = dict(mu=0.5, phi=1e-7, beta=50.0, lam=1e-2, tau=1e4)
    pobj = {p: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": 0.5, "scale": v}, init_value=v)
            for p, v in center.items()}
    m = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES["two_i"],
                              state_names=demo_models.STATE_NAMES["two_i"], dataframe=demo_df("two_i"),
                              state_summations={"H": ["S", "I1", "I2"]}, S=5236900, **pobj)
    post = m.MCMC(chain_inits=[dict(center)] * 4, iterations_per_chain=60, print_report=False)
    assert m._last_solver == "bdf" and len(post) == 4 * 29 and np.isfinite(post["chi"]).all()
    assert m._last_mcmc["fail_count"].sum() == 0
    g = golden("two_i")
    pobj2 = {p: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=sc)
             for p, (s, sc) in demo_models.PRIORS["two_i"].items()}
    m2 = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES["two_i"],
                               state_names=demo_models.STATE_NAMES["two_i"], dataframe=demo_df("two_i"),
                               state_summations={"H": ["S", "I1", "I2"]}, S=5236900, **pobj2)
    th = dict(zip(m2.get_pnames(), g["chain_def_s0_theta0"]))
    m2.MCMC(chain_inits=[th] * 2, iterations_per_chain=40, print_report=False)
    assert m2._last_solver == "dopri5"


def test_chains_that_exhaust_the_explicit_budget_are_rerun_on_bdf():
    """solver='auto' on a non-stiff start: DOPRI5 with a bounded step budget per solve; a chain that ever exhausts it is
    re-run whole on the BDF kernel with the same random streams (chain_ids keep its Philox key).  Every chain is then
    either the plain DOPRI5 chain or the plain BDF chain, bit for bit.  (In the first run a failing chain stops at its
    first failed solve -- stop_failed -- which the chains that never fail must not notice.)"""
    import scipy.stats
    import odelib_b200 as ODElib
    from odelib_b200 import demo_models
    pobj = {p: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=sc)
            for p, (s, sc) in demo_models.PRIORS["two_i"].items()}
    m = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES["two_i"],
                              state_names=demo_models.STATE_NAMES["two_i"], dataframe=demo_df("two_i"),
                              state_summations={"H": ["S", "I1", "I2"]}, S=5236900, **pobj)
    m.EXPLICIT_STEP_BUDGET = 230                                 # low enough that some chains of this set exhaust it
    m.STEP_BUDGET_FACTOR = 0                                     # (no headroom over the probe's step counts)
    rng = np.random.default_rng(3)
    center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
    starts = center * np.exp(0.3 * rng.standard_normal((48, 5)))
    dm = m._device()
    probe = dm.sweep(starts, solver="dopri5", max_steps=230)
    starts = starts[probe["status"] == 0][:32]                  # all starts themselves are within the budget
    assert len(starts) >= 16
    C, nits = len(starts), 60
    raw = m._run_chains([s for s in starts], list(range(C)), nits, nits // 2, [], rng="philox", return_raw=True)
    assert m._last_solver == "dopri5" and 0 < m._last_rerun < C
    assert raw["fail_count"].sum() == 0
    key = dict(nits=nits, rng_mode="philox", seed=int(m.random_seed), chain_ids=np.arange(C), trace=True)
    plain = dm.mcmc(starts, max_steps=230, **key)
    bdf = dm.mcmc(starts, solver="bdf", max_steps=2000000, **key)
    bad = plain["fail_count"] > 0
    assert bad.sum() == m._last_rerun
    assert np.array_equal(raw["samples"][~bad], plain["samples"][~bad])
    assert np.array_equal(raw["samples"][bad], bdf["samples"][bad])
    assert np.array_equal(raw["summaries"][bad], bdf["summaries"][bad])
    assert np.array_equal(raw["best_chi"][bad], bdf["best_chi"][bad], equal_nan=True)
    # stop_failed: the same chains are marked, the others are untouched, a marked chain is the plain chain up to its stop
    for K in (1, 8):
        stop = dm.mcmc(starts, max_steps=230, stop_failed=True, speculate=K, **key)
        assert np.array_equal(stop["fail_count"] > 0, bad)
        assert np.array_equal(stop["samples"][~bad], plain["samples"][~bad])
        assert np.array_equal(stop["summaries"][~bad], plain["summaries"][~bad])
        first = np.argmax(np.isnan(plain["chinew"]), axis=1)
        for c in np.flatnonzero(bad):
            assert np.array_equal(stop["chinew"][c, :first[c] + 1], plain["chinew"][c, :first[c] + 1], equal_nan=True)
            # (K lanes per chain: the round that holds the first failure may consume further failed proposals -- all of
            # them rejected -- before the chain stops)
            assert 1 <= stop["fail_count"][c] <= K


def test_two_stepper_chain_kernel_hands_single_solves_to_bdf():
    """odl_mcmc(solver=auto) with an explicit budget: per solve, DOPRI5 within the budget, else the same solve on BDF
    (LSODA's method switch, per solve).  A chain whose solves all stay within the budget is the plain DOPRI5 chain (the
    two kernels are compiled separately: equal to rounding, decisions identical); in the others every solve the plain
    kernel gave up on has a finite chi, and up to the first of them the chain is the plain chain."""
    dm, _ = device_model("two_i")
    rng = np.random.default_rng(3)
    center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
    starts = center * np.exp(0.3 * rng.standard_normal((48, 5)))
    probe = dm.sweep(starts, solver="dopri5", max_steps=230)
    starts = starts[probe["status"] == 0][:32]
    C, nits = len(starts), 60
    key = dict(nits=nits, rng_mode="philox", seed=11, chain_ids=np.arange(C), trace=True)
    plain = dm.mcmc(starts, max_steps=230, **key)
    both = dm.mcmc(starts, solver="auto", explicit_budget=230, max_steps=2000000, **key)
    bad = plain["fail_count"] > 0
    assert 0 < bad.sum() < C and both["fail_count"].sum() == 0
    # (separately compiled kernels: a step accepted in one may be rejected in the other at a rounding's distance from
    # err = 1, after which the two solves differ at solver accuracy -- 1e-7 at the default tolerance -- not at rounding)
    assert (both["accepted"][~bad] != plain["accepted"][~bad]).sum() <= 1
    same = (both["accepted"] == plain["accepted"]).all(axis=1) & ~bad
    np.testing.assert_allclose(both["chinew"][same], plain["chinew"][same], rtol=2e-6)
    np.testing.assert_allclose(both["samples"][same], plain["samples"][same], rtol=2e-6)
    gave_up = np.isnan(plain["chinew"]) & bad[:, None]
    first = np.argmax(gave_up, axis=1)
    assert np.isfinite(both["chinew"][bad, first[bad]]).all()
    for c in np.flatnonzero(bad):
        assert (both["accepted"][c, :first[c]] != plain["accepted"][c, :first[c]]).sum() <= 1
        if np.array_equal(both["accepted"][c, :first[c]], plain["accepted"][c, :first[c]]):
            np.testing.assert_allclose(both["chinew"][c, :first[c]], plain["chinew"][c, :first[c]], rtol=2e-6)
