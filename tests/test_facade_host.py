"""Host-side logic of the ModelFramework facade (no GPU): the tables it derives equal the reference's
(golden vectors), names / errors / return shapes follow ODElib/Framework.py."""
import numpy as np
import pandas as pd
import pytest
import scipy.stats

import odelib_b200 as ODElib
from odelib_b200 import demo_models
from odelib_b200.Statistics import Samplers, stats
from oracle import odelib_oracle as orc
from tests.helpers import PRIORS, STATES, SUMS, TSTEPS, demo_df, golden


def make_model(name, **kw):
    pri = {n: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=sc)
           for n, s, sc in PRIORS[name]}
    extra = {} if name == "zero_i" else {"S": 5236900}
    return ODElib.ModelFramework(ODE=demo_models.MODELS[name][0], parameter_names=[p[0] for p in PRIORS[name]],
                                 state_names=STATES[name], dataframe=demo_df(name), state_summations=SUMS[name],
                                 t_steps=TSTEPS[name], **pri, **extra, **kw)


def split_phi(y, t, ps):
    """zero_i with the adsorption rate given as an ARRAY-valued parameter of two summands (f4: the reference hands
    arrays straight to the RHS, Framework.py:656)."""
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    a = phi[0] + phi[1]
    return np.array([mu * S - a * S * V, beta * a * S * V - a * S * V])


def make_array_model(phi=(0.7e-8, 0.65e-8), **kw):
    LN = scipy.stats.lognorm
    return ODElib.ModelFramework(
        ODE=split_phi, parameter_names=["mu", "phi", "beta"], state_names=["S", "V"], dataframe=demo_df("zero_i"),
        mu=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 3, "scale": 1e-8}, init_value=1e-6),
        phi=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 3, "scale": 1e-8}, init_value=list(phi)),
        beta=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 1, "scale": 25}, init_value=19.4),
        t_steps=288, **kw)


def test_array_valued_parameters_flat_layout():
    """f4: one slot per element; parameter_names and get_parameters keep the reference's shapes."""
    from odelib_b200.tracer import trace
    m = make_array_model()
    assert m.get_pnames() == ["mu", "phi", "beta"]
    assert m._flat_names == ("mu", "phi[0]", "phi[1]", "beta") and m._flat_owner() == ["mu", "phi", "phi", "beta"]
    np.testing.assert_array_equal(m._current_theta(), [1e-6, 0.7e-8, 0.65e-8, 19.4])
    ps = m.get_parameters()[0]
    assert np.shape(ps[1]) == (2,) and np.shape(ps[0]) == ()
    assert m._pnum == 4                                           # AIC counts non-zero elements (Framework.py:260-263)
    # the fixture recorded from the unmodified reference (tests/golden/make_array_param.py) is what scipy gives on the
    # facade's own tables: same y0, grid, observation picks and chi
    import os
    from scipy.integrate import odeint
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "array_param.npz"))
    np.testing.assert_array_equal(np.asarray(m.get_inits(), float), g["y0"])
    mod = odeint(split_phi, m.get_inits(), m.times, args=(m.get_parameters()[0],))
    pred = {s_: mod[:, i][m._pred_tindex[s_]] for i, s_ in enumerate(m.get_snames()) if s_ in m._pred_tindex}
    np.testing.assert_array_equal(pred["S"], g["pred_S"])
    assert m.get_chi(pred) == float(g["chi"])
    # the RHS traced through the slot adapter == the user's function on arrays
    tm = trace(m._device_ode(), 2, 4)
    y = np.array([5.0e6, 1.0e7])
    np.testing.assert_allclose(tm.evaluate(tm.outputs, y, 0.0, m._current_theta()),
                               split_phi(y, 0.0, [1e-6, np.array([0.7e-8, 0.65e-8]), 19.4]), rtol=1e-15)
    # element names address single elements; theta vectors round-trip with the shapes kept
    m.set_parameters(**{"phi[1]": 1e-9})
    np.testing.assert_array_equal(m.parameters["phi"].val, [0.7e-8, 1e-9])
    m._set_theta(np.array([1.0, 2.0, 3.0, 4.0]))
    assert np.shape(m.parameters["phi"].val) == (2,) and float(m.parameters["beta"].val) == 4.0
    with pytest.raises(Exception, match="unknown parameter"):
        m.set_parameters(**{"phi[2]": 1.0})
    # LHS: one stratified column per NON-ZERO element, zeros are structural (Samplers.py:26-32)
    z = make_array_model(phi=(1.3e-8, 0.0))
    np.random.seed(2)
    sv = z._lhs_samples(50)
    assert list(sv.columns) == ["mu", "phi[0]", "phi[1]", "beta"] and np.all(sv["phi[1]"] == 0.0)
    u = scipy.stats.lognorm.cdf(sv["phi[0]"], s=3, scale=1e-8)
    assert sorted(np.floor(u * 50).astype(int)) == list(range(50))
    assert z._prior_table()[2] == ("const", 0.0, 0.0, 0.0) and z._prior_table()[1][0] == "lognorm" and z._pnum == 3
    # the reference chain draws one increment per element, one prior rvs per parameter object, one uniform
    walking = [m.parameters[p] for p in m.get_pnames()]
    zz, uu = Samplers.reference_streams_batch([5], walking, 30)
    rs = np.random.RandomState(5)
    for i in range(30):
        for j in range(4):
            assert zz[0, i, j] == rs.normal(0, 0.05)
        for _ in range(3):
            rs.standard_normal()
        assert uu[0, i] == rs.rand()


@pytest.mark.parametrize("name", ["zero_i", "one_i", "two_i"])
def test_ctor_tables_equal_reference(name):
    g = golden(name)
    m = make_model(name)
    assert np.array_equal(m.times, g["times"])
    order = [s for s in m.get_snames(after_summation=True) if s in m._pred_tindex]
    assert order == list(g["obs_order"])
    for s in order:
        assert np.array_equal(m._pred_tindex[s], g["tindex_" + s])
    assert np.array_equal(np.concatenate([m._obs_logabundance[s] for s in order]), g["ln_obs"])
    assert np.array_equal(np.concatenate([m._obs_logsigma[s] for s in order]), g["log_sigma"])
    assert np.array_equal(np.asarray(m.get_inits(), float), g["y0"])
    assert m._pnum == int(g["pnum"]) and m._samples == 37
    groups = m._observe_groups()
    assert [tuple(x) for x in groups] == [tuple(x) for x in demo_models.MODELS[name][3]]


def test_names_errors_and_parameter_semantics():
    m = make_model("two_i")
    assert m.get_pnames() == ["mu", "phi", "beta", "lam", "tau"]
    assert m.get_snames() == ["H", "V"] and m.get_snames(after_summation=False) == ["S", "I1", "I2", "V"]
    with pytest.raises(Exception, match="unknown parameter"):
        m.set_parameters(nope=1.0)
    with pytest.raises(Exception, match="unknown state"):
        m.set_inits(nope=1.0)
    with pytest.raises(ValueError):
        ODElib.ModelFramework(demo_models.two_i, ["mu", "phi", "beta", "lam", "tau"], ["S", "I1", "I2", "V"],
                              state_summations={"H": ["S", "I1"], "G": ["I1", "V"]})
    with pytest.raises(ValueError):
        ODElib.parameter()
    m.set_parameters(mu=2e-8)
    assert m.get_parameters()[0][0] == 2e-8 and m.get_parameters(as_dict=True)["mu"] == 2e-8
    c = m.copy(overwrite={"mu": 3e-8, "S": 10.0})
    assert c.get_parameters()[0][0] == 3e-8 and m.get_parameters()[0][0] == 2e-8
    assert c.get_inits()[0] == 10.0 and m.get_inits()[0] == 5236900
    assert "Current State Summations" in repr(m)
    p = ODElib.parameter(scipy.stats.lognorm, {"s": 1, "scale": 2.0}, init_value=1.5)
    np.random.seed(0)
    p.rwalk()
    assert float(p.val) == pytest.approx(1.5 * np.exp(0.05 * 1.764052345967664), rel=1e-15)
    assert ODElib.parameter(init_value=3.0).pdf() == 1.0


def test_host_stats_match_reference_formulas():
    rng = np.random.default_rng(0)
    O = rng.normal(10, 1, 37); S = rng.uniform(0.05, 0.6, 37)
    with np.errstate(all="ignore"):
        C = np.log(rng.normal(3e4, 2e4, 37))        # some negative -> NaN
        S[5] = 0.0
        ref = orc.chi(O, C, S)
    assert stats.chi(O, C, S) == pytest.approx(float(ref), rel=1e-14)
    assert stats.chi(O, np.full(37, np.nan), S) is np.ma.masked
    assert stats.AIC(10.0, 3) == 26.0
    Cd = {"a": np.array([1.0, 2.0, np.nan]), "b": np.array([3.0, 4.0])}
    Od = {"a": np.array([1.5, 2.5, 3.0]), "b": np.array([2.0, 5.0])}
    sst = 3 * np.var(Od["a"]) + 2 * np.var(Od["b"])
    assert stats.Rsqrd(Cd, Od) == pytest.approx(1 - (0.25 + 0.25 + 1 + 1) / sst)


def test_reference_streams_equal_golden_and_lhs_is_latin():
    g = golden("two_i")
    m = make_model("two_i")
    walking = [m.parameters[p] for p in m.get_pnames()]
    z, u = Samplers.reference_streams(0, walking, int(g["chain_def_s0_nits"]) - 1)
    assert np.array_equal(z, g["chain_def_s0_z"]) and np.array_equal(u, g["chain_def_s0_u"])
    # odd count per iteration (one parameter without prior) takes the scalar path
    walking[0] = ODElib.parameter(init_value=1.0)
    z2, u2 = Samplers.reference_streams(3, walking, 20)
    rs = np.random.RandomState(3)
    for i in range(20):
        for j in range(5):
            assert z2[i, j] == rs.normal(0, 0.05)
        for _ in range(4):
            rs.standard_normal()
        assert u2[i] == rs.rand()
    # many chains at once: the library's MT19937 / polar-gauss restatement is numpy's legacy stream bit for bit,
    # for both the even case and the odd one (gauss cache carried across the uniform), and for any 32-bit seed
    even = [m.parameters[p] for p in m.get_pnames()]
    for wk in (even, walking, even[:3]):
        seeds = [0, 1, 7, 12345, 2 ** 32 - 1]
        zb, ub = Samplers.reference_streams_batch(seeds, wk, 300)
        for c, sd in enumerate(seeds):
            zr, ur = Samplers.reference_streams(sd, wk, 300)
            assert np.array_equal(zb[c], zr) and np.array_equal(ub[c], ur)
    zb, ub = Samplers.reference_streams_batch(range(600), even, 400)          # enough work for the threaded split
    for c in (0, 299, 599):
        zr, ur = Samplers.reference_streams(c, even, 400)
        assert np.array_equal(zb[c], zr) and np.array_equal(ub[c], ur)
    np.random.seed(1)
    d = Samplers.lhs(3, 50)
    assert d.shape == (50, 3)
    for j in range(3):
        assert sorted(np.floor(d[:, j] * 50).astype(int)) == list(range(50))
    np.random.seed(2)
    df = m._lhs_samples(64)
    assert list(df.columns) == m.get_pnames() and len(df) == 64 and (df > 0).all().all()


def test_replicate_dataframe_format():
    rows = []
    for org in ("S", "V"):
        for t in (0.0, 1.0, 2.0):
            for r in range(3):
                rows.append({"organism": org, "time": t, "abundance": 100.0 * (1 + r) * (1 + t), "replicate": r})
    pri = {n: ODElib.parameter(init_value=v) for n, v in (("mu", 1e-8), ("phi", 1e-8), ("beta", 20.0))}
    m = ODElib.ModelFramework(demo_models.zero_i, ["mu", "phi", "beta"], ["S", "V"], dataframe=pd.DataFrame(rows),
                              t_steps=21, **pri)
    assert m._samples == 6
    np.testing.assert_allclose(m._obs_logabundance["S"], [np.log([100., 200, 300]).mean() + np.log(1 + t) for t in (0, 1, 2)])
    assert list(m._pred_tindex["V"]) == [0, 10, 20]


def test_pooled_log_stats_equal_rawstats_over_the_frame():
    """f2: Welford summaries per chain -> pooled log-mean / log-std (ddof=1) == what rawstats takes from the
    concatenated posterior frame (Framework.py:11-17)."""
    from odelib_b200.Framework import _rawstats_from_logmoments, rawstats
    from odelib_b200.rhat import pooled_log_stats
    rng = np.random.default_rng(0)
    C, n, P = 7, 41, 3
    x = np.exp(rng.normal(size=(C, n, P)) * [0.1, 1.0, 3.0] + [0.0, -18.0, 3.0])
    summ = np.zeros((C, 1 + 2 * P))
    for c in range(C):                                           # the kernel's recurrence
        for i in range(n):
            summ[c, 0] += 1.0
            lx = np.log(x[c, i])
            d = lx - summ[c, 1:1 + P]
            summ[c, 1:1 + P] += d / summ[c, 0]
            summ[c, 1 + P:] += d * (lx - summ[c, 1:1 + P])
    N, mean, std = pooled_log_stats(summ, P)
    assert N == C * n
    for q in range(P):
        med, sd = rawstats(pd.Series(x[:, :, q].ravel()))
        m2, s2 = _rawstats_from_logmoments(mean[q], std[q])
        assert m2 == pytest.approx(med, rel=1e-12) and s2 == pytest.approx(sd, rel=1e-10)


def test_effective_sample_size_from_chain_summaries():
    """f2: n_eff = m n var+ / B from the per-chain (count, mean, M2) rows alone; ~m n for independent draws, far
    below it for chains stuck at different levels, never above the number of kept rows."""
    from odelib_b200.rhat import ess_from_summaries, rhat_from_summaries
    rng = np.random.default_rng(1)
    m, n, P = 64, 200, 2
    x = rng.normal(size=(m, n, P))
    x[:, :, 1] += 3.0 * rng.normal(size=(m, 1))                   # second parameter: chains disagree
    summ = np.concatenate([np.full((m, 1), float(n)), x.mean(axis=1), ((x - x.mean(axis=1, keepdims=True)) ** 2).sum(axis=1)], axis=1)
    ess = ess_from_summaries(summ, P)
    W = x.var(axis=1, ddof=1).mean(axis=0)
    B = n * x.mean(axis=1).var(axis=0, ddof=1)
    np.testing.assert_allclose(ess, np.minimum(m * n, m * n * ((n - 1) / n * W + B / n) / B), rtol=1e-12)
    assert ess[0] > 0.5 * m * n and ess[1] < 0.02 * m * n and np.all(ess <= m * n)
    assert rhat_from_summaries(summ, P)[1] > 2.0
    same = np.concatenate([np.full((3, 1), 5.0), np.ones((3, 1)), np.ones((3, 1))], axis=1)   # B == 0: capped
    assert ess_from_summaries(same, 1)[0] == 15.0


def test_chain_start_picks_are_what_dataframe_sample_draws():
    """f1: `good.sample(n, replace=True)` (Framework.py:1012) == rows `np.random.choice(len(good), n, replace=True)`
    of `good` under the same global numpy seed -- the device path needs only len(good) from the survey."""
    good = pd.DataFrame({"a": np.arange(100.0, 163.0), "chi": np.linspace(1, 2, 63)}, index=np.arange(5, 68))
    for seed in (0, 1, 12345):
        np.random.seed(seed)
        ref = good.sample(17, replace=True)["a"].to_numpy()
        np.random.seed(seed)
        picks = np.random.choice(len(good), size=17, replace=True)
        assert np.array_equal(good["a"].to_numpy()[picks], ref)


class _RecordingDeviceModel:
    """Stands in for engine.DeviceModel: records what the facade uploads (no GPU, no compile)."""
    instances = 0

    def __init__(self, ode, n_state, n_param, groups, device=None, y0_from_param=False, cache_dir=None):
        type(self).instances += 1
        self.n_state, self.n_param, self.y0_from_param = n_state, n_param, y0_from_param
        self.loaded_y0 = self.loaded_grid = self.loaded_map = None
        self.uploads = 0

    def set_data(self, tables, y0, y0_from_param=None):
        self.loaded_y0, self.loaded_obs = np.array(y0), tables.ln_obs.copy()
        self.loaded_map = None if y0_from_param is None else np.array(y0_from_param)
        self.uploads += 1

    def set_grid(self, times, y0, y0_from_param=None):
        self.loaded_grid = np.array(times)


def test_copies_share_the_compiled_model_but_not_its_loaded_tables(monkeypatch):
    """a copy with other initial states / data must not leave ITS tables behind for the original."""
    import odelib_b200.Framework as F
    monkeypatch.setattr(F, "DeviceModel", _RecordingDeviceModel)
    _RecordingDeviceModel.instances = 0
    a = make_model("two_i")
    dm = a._device()
    n0 = dm.uploads
    assert a._device() is dm and dm.uploads == n0                      # nothing changed: nothing re-uploaded
    b = a.copy(overwrite={"S": 1.0e6})
    assert b._device() is dm and _RecordingDeviceModel.instances == 1   # shared compiled model
    assert dm.loaded_y0[0] == 1.0e6
    a._device()                                                         # the original loads its own states again
    assert dm.loaded_y0[0] == 5236900 and dm.uploads == n0 + 2
    # a copy given other data: its observation rows are what is loaded when IT computes, the original's when it does
    df2 = demo_df("two_i").copy()
    df2["abundance"] = df2["abundance"] * 2.0
    c = a.copy()
    c.reset_dataframe(df2)
    c._device()
    obs_c = dm.loaded_obs.copy()
    a._device()
    assert not np.array_equal(obs_c, dm.loaded_obs)
    np.testing.assert_allclose(obs_c, dm.loaded_obs + np.log(2.0), rtol=1e-14)
    # interleaved calls never reuse the other instance's tables
    for m_, s0 in ((b, 1.0e6), (a, 5236900), (b, 1.0e6)):
        m_._device()
        assert dm.loaded_y0[0] == s0


def test_state0_parameters_host_semantics(monkeypatch):
    """'<state>0' parameters: the map reaches the device tables; set_best_params copies the best row's values into the
    initial states (Framework.py:725-731) exactly as the unmodified reference did (tests/golden/make_state0.py)."""
    import odelib_b200.Framework as F
    monkeypatch.setattr(F, "DeviceModel", _RecordingDeviceModel)
    g = golden("state0")
    LN = scipy.stats.lognorm
    pri = [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25), ("S0", 0.3, 5.0e6), ("V0", 0.3, 1.1e7)]
    pobj = {n: ODElib.parameter(stats_gen=LN, hyperparameters={"s": s, "scale": sc}, init_value=v)
            for (n, s, sc), v in zip(pri, g["start"])}
    m = ODElib.ModelFramework(ODE=lambda y, t, ps: demo_models.zero_i(y, t, ps), parameter_names=[p[0] for p in pri],
                              state_names=["S", "V"], dataframe=demo_df("zero_i"), t_steps=288, **pobj)
    np.testing.assert_array_equal(m._y0_map(), [3, 4])
    dm = m._device()
    assert dm.y0_from_param and np.array_equal(dm.loaded_map, [3, 4])
    np.testing.assert_array_equal(dm.loaded_y0, g["y0"])              # istates from the data, not S0 / V0
    cols = [p[0] for p in pri] + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"]
    post = pd.DataFrame(g["chain_walk_kept"], columns=cols)
    post["chain#"] = 0
    m.set_best_params(post)
    np.testing.assert_array_equal(np.asarray(m.get_inits(), float), g["best_inits"])
    np.testing.assert_array_equal(m._current_theta(), g["best_theta"])
