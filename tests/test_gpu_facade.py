"""The drop-in surface on the GPU: these read like a session with the reference (the demo notebook) and
are checked against vectors recorded from the unmodified reference."""
import numpy as np
import pandas as pd
import pytest

from tests.helpers import golden
from tests.test_facade_host import make_array_model, make_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["zero_i", "one_i", "two_i"])
def test_integrate_and_fitstats_like_the_reference(name):
    g = golden(name)
    m = make_model(name, rtol=1e-12, atol=1e-12)
    th = g["theta"][0]
    m.set_parameters(**dict(zip(m.get_pnames(), th)))
    pred = m.integrate(predict_obs=True, as_dataframe=False)
    assert list(pred) == list(g["obs_order"])
    vec = np.concatenate([pred[s] for s in pred])
    np.testing.assert_allclose(vec, g["pred_tight"][0], rtol=2e-9)
    assert m.get_chi(pred) == pytest.approx(g["chi_tight"][0], rel=1e-9)
    fs = m.get_fitstats()
    assert fs["R^2"] == pytest.approx(g["r2_tight"][0], rel=1e-8, abs=1e-9)
    assert fs["AIC"] == pytest.approx(2 * g["chi_tight"][0] + 2 * int(g["pnum"]), rel=1e-9)
    full = m.integrate()
    assert list(full.columns) == m.get_snames() + ["time"] and len(full) == len(m.times)
    raw = m.integrate(as_dataframe=False, sum_subpopulations=False)
    assert raw.shape == (len(m.times), len(m._snames))
    long = m.integrate(predict_obs=True)
    assert len(long) == 37 and list(long.columns) == ["time", "abundance"]
    res = m.get_residuals()
    assert len(res) == 37


def test_fit_survey_frame_and_chi_consistency():
    m = make_model("zero_i")
    np.random.seed(0)
    sv = m.fit_survey(samples=2000, cpu_cores=8)
    assert list(sv.columns) == ["mu", "phi", "beta", "chi"] and len(sv) == 2000
    again = m.sweep(sv[["mu", "phi", "beta"]].to_numpy())
    assert np.array_equal(again["chi"], sv["chi"].to_numpy(), equal_nan=True)
    assert (sv["chi"] < 666).sum() > 20          # ~9.5 % of zero_i prior draws pass sd=6


def test_metropolis_hastings_reproduces_the_reference_chain():
    """Same start, same seed: the GPU chain IS the reference chain (its random numbers are regenerated)."""
    from odelib_b200.Statistics import Samplers
    name = "zero_i"
    g = golden(name)
    pre = "chain_tight_s0_"
    m = make_model(name, rtol=1e-13, atol=1e-13)
    m.set_parameters(**dict(zip(m.get_pnames(), g[pre + "theta0"])))
    m.random_seed = 0
    df = Samplers.MetropolisHastings(m, nits=int(g[pre + "nits"]), print_progress=False)
    assert list(df.columns) == m.get_pnames() + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"]
    kept = g[pre + "kept"]
    assert len(df) == len(kept)
    np.testing.assert_array_equal(df["iteration"].to_numpy(), kept[:, -2])
    np.testing.assert_array_equal(df["acceptance_ratio"].to_numpy(), kept[:, -1])     # identical decisions
    np.testing.assert_allclose(df[m.get_pnames()].to_numpy(), kept[:, :3], rtol=1e-12)
    np.testing.assert_allclose(df["chi"].to_numpy(), kept[:, 3], rtol=1e-8)


def test_hundreds_of_chains_keep_the_reference_streams():
    """MCMC's default rng regenerates every chain's own numpy stream (seed = chain index, Framework.py:1015) with
    the library's host code, so chain 7 of a 300-chain call is the chain a single seeded run produces."""
    from odelib_b200.Statistics import Samplers
    m = make_model("two_i")
    th = m.get_parameters(as_dict=True)
    post = m.MCMC(chain_inits=[th] * 300, iterations_per_chain=60, print_report=False)
    for c in (7, 299):
        m.set_parameters(**th)
        m.random_seed = c
        one = Samplers.MetropolisHastings(m, nits=60, burnin=30, print_progress=False)
        mine = post[post["chain#"] == c].drop(columns="chain#").reset_index(drop=True)
        assert len(one) == len(mine) and len(one) > 0
        np.testing.assert_array_equal(mine.to_numpy(), one.reset_index(drop=True).to_numpy())


def test_mcmc_demo_call_shapes_and_report(capsys):
    m = make_model("two_i")
    np.random.seed(3)
    post = m.MCMC(chain_inits=8, iterations_per_chain=200, cpu_cores=8, fitsurvey_samples=4000, sd_fitdistance=6.0)
    cols = m.get_pnames() + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio", "chain#"]
    assert list(post.columns) == cols
    assert len(post) == 8 * 99 and set(post["chain#"]) == set(range(8))       # nits-1-burnin rows per chain
    assert post["iteration"].min() == 101 and post["iteration"].max() == 199
    assert "Fitting Report" in capsys.readouterr().out
    assert m.rhat is not None and len(m.rhat) == 5
    # explicit starts + static parameter: column reports the prior scale (reference quirk A13)
    starts = [dict(zip(m.get_pnames(), post[m.get_pnames()].iloc[i])) for i in (0, 150)]
    p2 = m.MCMC(chain_inits=starts, iterations_per_chain=100, static_parameters=["tau"], print_report=False)
    assert np.all(p2["tau"] == 1) and len(p2) == 2 * 49
    with pytest.raises(ValueError):
        m.MCMC(chain_inits=2, iterations_per_chain=50, fitsurvey_samples=50, sd_fitdistance=0.01, print_report=False)


def test_chain_start_selection_on_the_device_equals_the_pandas_filter():
    """f1 (Framework.py:993-1016): threshold filter + ordered compaction + gather on the device pick exactly the rows
    `fitsurvey[fitsurvey['chi'] < cutchi].sample(n, replace=True)` picks under the same numpy seed."""
    import torch
    m = make_model("zero_i")
    dm = m._device()
    np.random.seed(4)
    ps = m._lhs_samples(5000)[m.get_pnames()]
    theta = np.ascontiguousarray(ps.to_numpy(dtype=np.float64))
    res = m.sweep(theta)
    survey = ps.reset_index(drop=True).copy()
    survey["chi"] = res["chi"]
    survey = survey.dropna()
    cut = 37 * 6.0 ** 2 / 2
    np.random.seed(11)
    ref = survey[survey["chi"] < cut].sample(16, replace=True)[m.get_pnames()].to_numpy()
    th_dev = torch.from_numpy(theta).cuda()
    chi_dev = torch.from_numpy(res["chi"]).cuda()
    index, count = dm.select_below(chi_dev, cut)
    assert count == int((survey["chi"] < cut).sum()) and count > 50
    assert np.array_equal(index[:count].cpu().numpy(), np.flatnonzero(res["chi"] < cut))      # ascending, NaN excluded
    np.random.seed(11)
    got = dm.gather_rows(th_dev, np.random.choice(count, size=16, replace=True), index=index).cpu().numpy()
    assert np.array_equal(got, ref)
    # ragged sizes around the 1024-row tiles; nothing / everything selected
    for n in (1, 1023, 1025, 4097):
        idx, c = dm.select_below(chi_dev[:n].contiguous(), cut)
        assert np.array_equal(idx[:c].cpu().numpy(), np.flatnonzero(res["chi"][:n] < cut))
    assert dm.select_below(chi_dev, -1.0)[1] == 0
    assert dm.select_below(chi_dev, np.inf)[1] == int(np.isfinite(res["chi"]).sum())


def test_report_and_best_parameters_come_from_device_reductions(capsys):
    """f2 (Framework.py:11-17, :725-731, :1047-1060): pooled log-moments and the best kept row from the kernel equal
    rawstats / idxmin over the posterior frame; posterior='summary' returns them without building the frame."""
    from odelib_b200.Framework import PosteriorSummary, rawstats
    m = make_model("two_i")
    np.random.seed(3)
    post = m.MCMC(chain_inits=8, iterations_per_chain=200, fitsurvey_samples=4000, sd_fitdistance=6.0)
    capsys.readouterr()
    s = m.posterior_summary()
    for p in m.get_pnames():
        med, sd = rawstats(post[p])
        assert s.stats[p][0] == pytest.approx(med, rel=1e-12) and s.stats[p][1] == pytest.approx(sd, rel=1e-9)
    row = post.loc[post["chi"].idxmin()]
    assert s.best_chi == row["chi"] and all(s.best[p] == row[p] for p in m.get_pnames())
    assert all(float(m.get_parameters(as_dict=True)[p]) == row[p] for p in m.get_pnames())        # set_best_params ran
    assert s.n_rows == len(post) and s.n_chains == 8
    m2 = make_model("two_i")
    np.random.seed(3)
    summ = m2.MCMC(chain_inits=8, iterations_per_chain=200, fitsurvey_samples=4000, sd_fitdistance=6.0,
                   posterior="summary")
    assert isinstance(summ, PosteriorSummary) and "Fitting Report" in capsys.readouterr().out
    assert summ.best == s.best and summ.stats == s.stats and summ.rhat == s.rhat
    assert m2._last_mcmc["samples"] is None                     # no sample rows were produced at all


def test_explore_equilibriums_returns_final_states():
    """f3 (Framework.py:819-854): the final state of every LHS sample, from a two-point output grid."""
    from scipy.integrate import odeint
    from oracle import odelib_oracle as orc
    m = make_model("one_i")
    np.random.seed(5)
    eq = m.explore_equilibriums(samples=64)
    assert list(eq.columns) == m.get_snames(after_summation=False) + m.get_pnames() and len(eq) == 64
    y0 = list(m.get_inits())
    ok = 0
    for k in (0, 17, 63):
        th = eq[m.get_pnames()].iloc[k].to_numpy()
        ref = odeint(orc.one_i, y0, [m.times[0], m.times[-1]], args=(list(th),), rtol=1e-12, atol=1e-12, mxstep=500000)[-1]
        got = eq[m.get_snames(after_summation=False)].iloc[k].to_numpy()
        if np.all(np.isfinite(got)):
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1.0)
            ok += 1
    assert ok >= 2
    full = m.integrate()                                          # the full output grid is back in place
    assert len(full) == len(m.times)


def test_gradient_batched_and_continuation():
    """f3 (Framework.py:1063-1127, fixed): a 1-D parameter scan, batched when the runs are independent, a
    continuation seeded by the previous end point otherwise -- both against scipy on the same recipe."""
    from scipy.integrate import odeint
    from oracle import odelib_oracle as orc
    m = make_model("one_i")
    names = m.get_snames(after_summation=False)
    th0 = m._current_theta()
    k = m.get_pnames().index("beta")
    betas = np.linspace(10.0, 40.0, 7)
    ends = m.gradient("beta", betas, seed_equilibrium=False, aggregate_enpoints=True, print_status=False)
    assert list(ends.columns) == names + ["beta"] and len(ends) == 7
    np.testing.assert_array_equal(ends["beta"].to_numpy(), betas)
    y0 = np.asarray(m.get_inits(), dtype=np.float64)
    for i in (0, 3, 6):
        th = th0.copy(); th[k] = betas[i]
        ref = odeint(orc.one_i, y0, [m.times[0], m.times[-1]], args=(list(th),), rtol=1e-12, atol=1e-12, mxstep=500000)[-1]
        np.testing.assert_allclose(ends[names].iloc[i].to_numpy(), ref, rtol=2e-5, atol=1.0)
    # every grid row of every run, and the same end points
    full = m.gradient("beta", betas[:3], seed_equilibrium=False, print_status=False)
    assert len(full) == 3 * len(m.times)
    np.testing.assert_allclose(full[names].iloc[len(m.times) - 1].to_numpy(), ends[names].iloc[0].to_numpy(), rtol=1e-6, atol=1e-3)
    np.testing.assert_array_equal(full[names].iloc[len(m.times)].to_numpy(), y0)
    # continuation: run i starts from run i-1's final state floored at 0.001
    cont = m.gradient("beta", betas[:4], seed_equilibrium=True, aggregate_enpoints=True, print_status=False)
    init = y0.copy()
    for i in range(4):
        th = th0.copy(); th[k] = betas[i]
        ref = odeint(orc.one_i, init, [m.times[0], m.times[-1]], args=(list(th),), rtol=1e-12, atol=1e-12, mxstep=500000)[-1]
        np.testing.assert_allclose(cont[names].iloc[i].to_numpy(), ref, rtol=2e-5, atol=1.0)
        init = np.clip(cont[names].iloc[i].to_numpy(), 0.001, None)     # like for like: seed scipy as the scan was seeded
    np.testing.assert_array_equal(m._current_theta(), th0)        # the scanned parameter is restored
    assert len(m.integrate()) == len(m.times)
    assert len(m.gradient("beta", [], print_status=False)) == 0
    with pytest.raises(ValueError):
        m.gradient("nope", betas)


def test_array_valued_parameters_on_the_device():
    """f4: an array-valued parameter reaches the RHS as an array.  integrate / get_chi with arrays run in the reference
    (its value for this model, recorded from the unmodified reference, is below); its sample_lhs (Samplers.py:45) and
    chain (Framework.py:99) branches for arrays raise, so surveys and chains are checked against the equivalent scalar
    model instead."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "array_param.npz"))   # tests/golden/make_array_param.py
    m = make_array_model()
    np.testing.assert_array_equal(m._current_theta(), g["theta"])
    pred = m.integrate(predict_obs=True, as_dataframe=False)
    assert m.get_chi(pred) == pytest.approx(float(g["chi"]), rel=2e-6)        # reference at odeint's default tolerance
    for s_ in ("S", "V"):
        ok = g["pred_" + s_] > 1.0
        np.testing.assert_allclose(pred[s_][ok], g["pred_" + s_][ok], rtol=5e-6)
    assert list(g["raises"]) == ["ValueError", "TypeError"]                   # what the reference does with arrays there
    z = make_model("zero_i")                                      # phi = phi[0] + phi[1]: the same arithmetic
    rows = np.exp(np.random.default_rng(0).normal(size=(64, 4)) * 0.3) * m._current_theta()
    a = m.sweep(rows)
    b = z.sweep(np.column_stack([rows[:, 0], rows[:, 1] + rows[:, 2], rows[:, 3]]))
    np.testing.assert_allclose(a["chi"], b["chi"], rtol=1e-12)
    np.random.seed(4)
    sv = m.fit_survey(samples=200)
    assert list(sv.columns) == ["mu", "phi[0]", "phi[1]", "beta", "chi"] and len(sv) == 200
    post = m.MCMC(chain_inits=[{"phi": [0.7e-8, 0.65e-8], "beta": 19.0}] * 3, iterations_per_chain=80,
                  static_parameters=["mu"], print_report=True)
    assert list(post.columns) == ["mu", "phi[0]", "phi[1]", "beta", "chi", "rsquared", "aic", "iteration",
                                  "acceptance_ratio", "chain#"]
    assert len(post) == 3 * 39 and np.all(post["mu"] == 1e-8)     # static column: the prior's scale (quirk A13)
    assert post["phi[0]"].nunique() > 1 and post["phi[1]"].nunique() > 1
    assert np.all(post["aic"] == 2 * post["chi"] + 2 * 4)
    assert np.shape(m.parameters["phi"].val) == (2,)              # best row written back with the shape kept
    best = post.loc[post["chi"].idxmin()]
    np.testing.assert_array_equal(m.parameters["phi"].val, [best["phi[0]"], best["phi[1]"]])
    # a structural zero stays zero along the chain and in the survey
    s0 = make_array_model(phi=(1.35e-8, 0.0))
    p0 = s0.MCMC(chain_inits=[{}] * 2, iterations_per_chain=40, print_report=False)
    assert np.all(p0["phi[1]"] == 0.0) and p0["phi[0]"].nunique() > 1
    assert np.all(p0["aic"] == 2 * p0["chi"] + 2 * 3)


def test_device_latin_hypercube_sampling_of_the_priors():
    """odl_sample_lhs (Samplers.py:6-51 on the device): every column visits every stratum exactly once, the k-th
    smallest value of a column lies between the prior's ppf at k/n and (k+1)/n, columns are independent, the sample is
    reproducible through numpy's seed, and fit_survey uses it for large surveys with the same frame layout."""
    import scipy.stats
    m = make_model("two_i")
    dm = m._device()
    n = 40000
    table = m._prior_table()
    assert [t[0] for t in table] == ["lognorm"] * 5
    th = dm.sample_lhs(table, n, seed=123).cpu().numpy()
    again = dm.sample_lhs(table, n, seed=123).cpu().numpy()
    other = dm.sample_lhs(table, n, seed=124).cpu().numpy()
    assert np.array_equal(th, again) and not np.array_equal(th, other)
    edges = np.arange(n + 1) / n
    for j, (kind, s_, loc, scale) in enumerate(table):
        lo = scipy.stats.lognorm.ppf(edges[:-1], s=s_, loc=loc, scale=scale)
        hi = scipy.stats.lognorm.ppf(edges[1:], s=s_, loc=loc, scale=scale)
        col = np.sort(th[:, j])
        assert np.all(col >= lo * (1 - 1e-12)) and np.all(col <= hi * (1 + 1e-12))       # one point per stratum, right ppf
    c = np.corrcoef(np.log(th).T)
    assert np.abs(c - np.eye(5)).max() < 0.03
    mixed = dm.sample_lhs([("const", 2.5, 0, 0), ("norm", 0, 1.0, 2.0), ("uniform", 0, -1.0, 4.0), ("lognorm", 1.0, 0, 3.0),
                           ("const", 7.0, 0, 0)], 1000, seed=5).cpu().numpy()
    assert np.all(mixed[:, 0] == 2.5) and np.all(mixed[:, 4] == 7.0)
    assert -1.0 <= mixed[:, 2].min() < -0.99 and 2.99 < mixed[:, 2].max() <= 3.0
    assert abs(mixed[:, 1].mean() - 1.0) < 0.01 and abs(mixed[:, 1].std() - 2.0) < 0.02
    np.random.seed(8)
    sv = m.fit_survey(samples=70000)                             # >= DEVICE_SAMPLING_FROM: sampled on the device
    np.random.seed(8)
    sv2 = m.fit_survey(samples=70000)
    assert list(sv.columns) == m.get_pnames() + ["chi"] and len(sv) == 70000 and sv.equals(sv2)
    host = m.sweep(np.ascontiguousarray(sv[m.get_pnames()].to_numpy()))
    assert np.array_equal(host["chi"], sv["chi"].to_numpy(), equal_nan=True)
    with pytest.raises(NotImplementedError):
        m.parameters["mu"].dist = scipy.stats.gamma
        m.fit_survey(samples=10, sampler="device")


def test_edge_sizes_of_the_device_side_helpers():
    """Empty and one-row inputs of odl_select_below / odl_gather_rows / odl_sample_lhs, and bad arguments."""
    import torch
    from odelib_b200 import _capi
    m = make_model("zero_i")
    dm = m._device()
    empty = torch.empty(0, dtype=torch.float64, device="cuda")
    idx, c = dm.select_below(empty, 1.0)
    assert c == 0 and idx.numel() == 0
    one = torch.tensor([0.5], dtype=torch.float64, device="cuda")
    assert dm.select_below(one, 1.0)[1] == 1 and dm.select_below(one, 0.5)[1] == 0          # strict <
    nan = torch.tensor([float("nan"), 2.0, float("inf"), -1.0], dtype=torch.float64, device="cuda")
    idx, c = dm.select_below(nan, float("inf"))
    assert c == 2 and idx[:2].cpu().tolist() == [1, 3]
    src = torch.arange(12, dtype=torch.float64, device="cuda").reshape(4, 3)
    assert dm.gather_rows(src, np.array([], dtype=np.int64)).shape == (0, 3)
    got = dm.gather_rows(src, np.array([3, 3, 0]))
    assert got.cpu().tolist() == [[9, 10, 11], [9, 10, 11], [0, 1, 2]]
    table = m._prior_table()
    assert dm.sample_lhs(table, 0).shape == (0, 3)
    one_row = dm.sample_lhs(table, 1, seed=1).cpu().numpy()
    assert one_row.shape == (1, 3) and np.all(np.isfinite(one_row)) and np.all(one_row > 0)
    three = dm.sample_lhs(table, 3, seed=1).cpu().numpy()
    assert len(set(np.argsort(three[:, 0]))) == 3
    with pytest.raises(_capi.OdlError):
        dm.sample_lhs([(7, 1.0, 0.0, 1.0)] * 3, 4)                # unknown prior kind


def test_mcmc_use_priors_is_opt_in_and_changes_the_target(capsys):
    """MCMC(use_priors=True): the posterior ratio (device prior log-densities + Hastings term) instead of the
    reference's likelihood-only ratio; the default stays the reference's chain."""
    m = make_model("two_i")
    g = golden("two_i")
    th = dict(zip(m.get_pnames(), g["chain_def_s0_theta0"]))
    a = m.MCMC(chain_inits=[th] * 4, iterations_per_chain=120, print_report=False, rng="philox")
    b = m.MCMC(chain_inits=[th] * 4, iterations_per_chain=120, print_report=False, rng="philox")
    c = m.MCMC(chain_inits=[th] * 4, iterations_per_chain=120, print_report=False, rng="philox", use_priors=True)
    assert a.equals(b) and not a.equals(c)
    assert list(c.columns) == list(a.columns) and len(c) == len(a)
    # static parameters drop out of the prior product (they never move)
    d = m.MCMC(chain_inits=[th] * 2, iterations_per_chain=60, print_report=False, rng="philox", use_priors=True,
               static_parameters=["tau"])
    assert np.all(d["tau"] == 1)


def test_large_reference_runs_take_their_streams_from_the_device_generator(monkeypatch):
    """Beyond HOST_STREAM_DOUBLES the facade regenerates the reference's numpy streams on the device
    (odl_reference_streams_device) instead of switching to Philox: the same chains as with host-generated streams --
    decisions identical, samples to rounding (the gaussians agree to 1 ulp)."""
    m = make_model("two_i")
    g = golden("two_i")
    th = dict(zip(m.get_pnames(), g["chain_def_s0_theta0"]))
    a = m.MCMC(chain_inits=[th] * 6, iterations_per_chain=120, print_report=False)
    monkeypatch.setattr(type(m), "HOST_STREAM_DOUBLES", 100)
    b = m.MCMC(chain_inits=[th] * 6, iterations_per_chain=120, print_report=False)
    assert len(a) == len(b) == 6 * 59
    assert np.array_equal(a["iteration"].to_numpy(), b["iteration"].to_numpy())
    assert np.array_equal(a["acceptance_ratio"].to_numpy(), b["acceptance_ratio"].to_numpy())   # same decisions
    np.testing.assert_allclose(b[m.get_pnames()].to_numpy(), a[m.get_pnames()].to_numpy(), rtol=1e-12)
    np.testing.assert_allclose(b["chi"].to_numpy(), a["chi"].to_numpy(), rtol=1e-9)
