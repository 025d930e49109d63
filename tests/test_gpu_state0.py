"""'<state>0' parameters on the GPU path against vectors recorded from the unmodified reference
(tests/golden/make_state0.py): such a parameter sets the state's initial value for the solves of MCMC PROPOSALS only
(Samplers.py:110-114, restored on reject :139-143); integrate, the survey seam and the chain's a-priori solve start
from istates (Framework.py:647-650, :41-48; Samplers.py:88); set_best_params copies the best row's values into the
initial states (Framework.py:730-731).  The fixture's S0 / V0 differ from the t == 0 data on purpose."""
import numpy as np
import pandas as pd
import pytest
import scipy.stats

import odelib_b200 as ODElib
from odelib_b200 import demo_models
from odelib_b200.engine import DeviceModel
from oracle import odelib_oracle as orc
from tests.helpers import demo_df, golden, obs_tables_from_oracle, oracle_tables

pytestmark = pytest.mark.gpu
PRI = [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25), ("S0", 0.3, 5.0e6), ("V0", 0.3, 1.1e7)]
Y0MAP = np.array([3, 4], np.int32)


def rhs5(y, t, ps):                                                # zero_i; ps[3], ps[4] = S0, V0 are not rates
    return demo_models.zero_i(y, t, ps)


def orc_rhs5(y, t, ps):
    return orc.zero_i(y, t, ps[:3])


def device_model5(**kw):
    tab = oracle_tables("zero_i")
    dm = DeviceModel(rhs5, 2, 5, None, y0_from_param=True, **kw)
    dm.set_data(obs_tables_from_oracle(tab), tab.y0, Y0MAP)
    dm.set_grid(tab.times, tab.y0, Y0MAP)
    return dm, tab


def facade_model(g, **kw):
    pobj = {n: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=v)
            for (n, s, sc), v in zip(PRI, g["start"])}
    return ODElib.ModelFramework(ODE=rhs5, parameter_names=[p[0] for p in PRI], state_names=["S", "V"],
                                 dataframe=demo_df("zero_i"), t_steps=288, **pobj, **kw)


def _replay(chi0, chinew, u):
    acc = np.zeros(len(u), bool); c = chi0
    for k in range(len(u)):
        with np.errstate(all="ignore"):
            if np.exp(c - chinew[k]) > u[k]:
                acc[k] = True; c = chinew[k]
    return acc


def test_sweep_and_trajectory_start_from_istates():
    """The survey seam and integrate ignore S0 / V0 (they differ from istates in every row of the fixture)."""
    g = golden("state0")
    dm, tab = device_model5()
    out = dm.sweep(g["theta"], return_pred=True)
    np.testing.assert_allclose(out["pred"], g["pred_def"], rtol=5e-6)
    np.testing.assert_allclose(out["chi"], g["chi_def"], rtol=2e-5)
    np.testing.assert_allclose(out["chi"], g["fit_worker_chi"], rtol=2e-5)
    auto = dm.sweep(g["theta"], solver="auto")                    # the ordering key does not read S0 / V0 either
    np.testing.assert_array_equal(auto["chi"], out["chi"])
    traj, status, _ = dm.trajectory(g["theta"][1:2])
    assert status[0] == 0
    np.testing.assert_array_equal(traj[0, 0], tab.y0)
    traj2, _, _ = dm.trajectory(g["theta"][1:2], y0=np.array([1.0e6, 2.0e6]))   # explicit inits win (integrate(inits=))
    np.testing.assert_array_equal(traj2[0, 0], [1.0e6, 2.0e6])


@pytest.mark.parametrize("tag,walk", [("walk", [0, 1, 2, 3, 4]), ("staticV0", [0, 1, 2, 3])])
@pytest.mark.parametrize("spec", [1, 8])
def test_chain_matches_reference(tag, walk, spec):
    """Reference streams in, reference chain out: a-priori chi from istates, every proposal solved from ITS S0 / V0 (a
    static V0 included), rejected proposals leave no trace in the next one."""
    g = golden("state0")
    pre = f"chain_{tag}_"
    nits = int(g[pre + "nits"])
    dm, tab = device_model5()
    z, u = g[pre + "z"], g[pre + "u"]
    out = dm.mcmc(g[pre + "theta0"][None], nits=nits, rng_mode="host", z=z[None], u=u[None], walk=walk, trace=True,
                  pnum=int(g["pnum"]), speculate=spec, rtol=1e-11, atol=1e-11, max_steps=2000000)
    # the oracle chain at tight tolerance on the same streams is the arbiter of the decisions (the golden chain ran at
    # scipy's default tolerance); it equals the reference chain decision by decision (test_oracle_golden.py)
    wmask = np.isin(np.arange(5), walk)
    ref = orc.mh_chain(orc_rhs5, g[pre + "theta0"], tab, int(g["pnum"]), nits=nits, walk=wmask, z=z, u=u,
                       rtol=1e-12, atol=1e-12, y0_from_param={0: 3, 1: 4})
    assert np.array_equal(ref["accepted"], g[pre + "accepted"])
    assert np.array_equal(out["accepted"][0].astype(bool), ref["accepted"])
    np.testing.assert_allclose(out["chinew"][0], ref["chinew"], rtol=1e-8)
    np.testing.assert_allclose(out["chinew"][0], g[pre + "chinew"], rtol=2e-5)
    P = 5
    np.testing.assert_allclose(out["samples"][0][:, :P], ref["kept"][:, :P], rtol=1e-12)
    np.testing.assert_allclose(out["samples"][0][:, P], ref["kept"][:, P], rtol=1e-8)
    np.testing.assert_array_equal(out["samples"][0][:, P + 3:], ref["kept"][:, P + 3:])
    # a-priori chi: istates (the data's t == 0 rows), NOT the S0 / V0 of the starting point
    forced = dm.mcmc(g[pre + "theta0"][None], nits=nits, rng_mode="forced", forced=g[pre + "proposals"][None],
                     u=u[None], walk=walk, trace=True, pnum=int(g["pnum"]), speculate=spec)
    assert np.array_equal(forced["accepted"][0].astype(bool), _replay(float(g[pre + "chi0"]), forced["chinew"][0], u))
    first_acc = int(np.flatnonzero(g[pre + "accepted"])[0])
    assert first_acc > 0                                          # iterations before it compare against the a-priori chi
    np.testing.assert_allclose(forced["chinew"][0], g[pre + "chinew"], rtol=2e-5)


def test_facade_chain_and_best_params():
    """MetropolisHastings through the facade = the reference chain (its own numpy streams regenerated by the library);
    the model is left at the chain's last point with the initial states following S0 / V0; set_best_params."""
    g = golden("state0")
    m = facade_model(g, rtol=1e-11, atol=1e-11)
    m.random_seed = 3
    pre = "chain_walk_"
    np.testing.assert_allclose(m.get_chi(m.integrate(predict_obs=True, as_dataframe=False)), g["chi_def"][0], rtol=2e-5)
    frame = ODElib.Statistics.Samplers.MetropolisHastings(m, nits=int(g[pre + "nits"]), print_progress=False)
    kept = g[pre + "kept"]
    np.testing.assert_array_equal(frame["iteration"].to_numpy(), kept[:, 8].astype(int))
    np.testing.assert_array_equal(frame["acceptance_ratio"].to_numpy(), kept[:, 9])        # same decisions
    np.testing.assert_allclose(frame[[p[0] for p in PRI]].to_numpy(), kept[:, :5], rtol=1e-12)
    np.testing.assert_allclose(frame["chi"].to_numpy(), kept[:, 5], rtol=2e-5)
    np.testing.assert_allclose(m._current_theta(), g[pre + "final_theta"], rtol=1e-12)
    np.testing.assert_allclose(np.asarray(m.get_inits(), float), g[pre + "final_inits"], rtol=1e-12)
    post = pd.DataFrame(kept, columns=[p[0] for p in PRI] + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"])
    post["chain#"] = 0
    m.set_best_params(post)
    np.testing.assert_array_equal(np.asarray(m.get_inits(), float), g["best_inits"])
    # integrate now starts from the best row's S0 / V0 -- through istates, as in the reference
    mod = m.integrate(as_dataframe=False, sum_subpopulations=False)
    np.testing.assert_array_equal(mod[0], g["best_inits"])
    # a survey row's chi does not depend on its S0 / V0 columns (Framework.py:41-48)
    th = np.tile(m._current_theta(), (4, 1))
    th[:, 3] *= [1.0, 0.5, 2.0, 1.0]
    th[:, 4] *= [1.0, 1.0, 1.0, 3.0]
    chi = m.sweep(th)["chi"]
    assert np.all(chi == chi[0])
