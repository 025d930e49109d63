#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference does not import as-is here, so three harness shims are
installed first -- none of them touches reference source or reference arithmetic:
  1. stub ``matplotlib`` / ``matplotlib.pyplot`` modules (Framework.py:6, plotting only);
  2. stub ``pyDOE2.lhs`` (Samplers.py:3; not exercised by anything recorded here);
  3. ``pd.Series.iteritems = pd.Series.items`` (Framework.py:247, removed in pandas >= 2).

What is recorded, per demo model (zero_i / one_i / two_i, Demo_InfectionStates.ipynb:60-128):
  * ``theta[K,P]``      parameter sets: posterior-like points + seeded prior draws
  * ``pred_*[K,n_obs]`` reference predictions at the observation rows (chi's concatenation order)
  * ``chi_*``, ``r2_*`` ``ModelFramework.get_chi`` / ``get_Rsqrd``
    ``*_def``  : odeint at scipy's default tolerance (what the reference really runs)
    ``*_tight``: odeint at rtol=atol=1e-13 (module-level name ``ODElib.Framework.odeint`` rebound
                 to a functools.partial -- the call site Framework.py:656 is untouched)
  * tables: ``times``, ``tindex_<organism>``, ``ln_obs``, ``log_sigma`` (concatenation order), ``y0``
  * chains: ``Samplers.MetropolisHastings`` with the random streams (np.random.normal /
    np.random.rand wrappers), every proposal, chinew, decision, and the returned frame,
    at default and tight tolerance.
"""
import contextlib
import functools
import io
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    pyd = types.ModuleType("pyDOE2")

    def lhs(n, samples=None):
        samples = samples or n
        u = np.random.rand(samples, n)
        edges = np.linspace(0, 1, samples + 1)
        pts = u * (edges[1:] - edges[:-1])[:, None] + edges[:-1][:, None]
        out = np.empty_like(pts)
        for j in range(n):
            out[:, j] = pts[np.random.permutation(samples), j]
        return out

    pyd.lhs = lhs
    sys.modules["pyDOE2"] = pyd
    pd.Series.iteritems = pd.Series.items
    sys.path.insert(0, "/root/reference")


install_shims()
warnings.filterwarnings("ignore")
import ODElib  # noqa: E402
import ODElib.Framework as RF  # noqa: E402
from ODElib.Statistics import Samplers as RS  # noqa: E402
from scipy.integrate import odeint as _odeint  # noqa: E402


# --- the demo models, as in the notebook (:60-128) -------------------------------------------
def two_i(y, t, ps):
    mu, phi, beta, lam, tau = ps[0], ps[1], ps[2], ps[3], ps[4]
    S, I1, I2, V = y[0], y[1], y[2], y[3]
    dSdt = mu * S - phi * S * V
    dI1dt = phi * S * V - tau * I1
    dI2dt = tau * I1 - lam * I2
    dVdt = beta * lam * I2 - phi * S * V
    return np.array([dSdt, dI1dt, dI2dt, dVdt])


def one_i(y, t, ps):
    mu, phi, beta, lam = ps[0], ps[1], ps[2], ps[3]
    S, I1, V = y[0], y[1], y[2]
    dSdt = mu * S - phi * S * V
    dI1dt = phi * S * V - lam * I1
    dVdt = beta * lam * I1 - phi * S * V
    return np.array([dSdt, dI1dt, dVdt])


def zero_i(y, t, ps):
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    dSdt = mu * S - phi * S * V
    dVdt = beta * phi * S * V - phi * S * V
    return np.array([dSdt, dVdt])


LN = scipy.stats.lognorm
PRIORS = {  # Demo_InfectionStates.ipynb:885-891, :8575-8578, :17472-17476
    "zero_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25)],
    "one_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1)],
    "two_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1), ("tau", 2, 1)],
}
# posterior-like points: notebook rows (:2297-2307) and fitting-report medians (:8613-8619, :15120-15128)
POSTERIOR = {
    "zero_i": [(1.480838e-08, 1.364223e-08, 19.386877), (1.364139e-08, 1.352514e-08, 19.442711),
               (4.594495e-06, 1.334745e-08, 19.110142), (1.36e-8, 1.35e-8, 19.44)],
    "one_i": [(1.238e-08, 3.550e-08, 19.40, 1.835)],
    "two_i": [(7.475e-09, 1.069e-07, 19.73, 1.934, 2.799)],
}


def make_model(name):
    df = pd.read_csv(os.path.join(HERE, "demodata.csv"))
    P = ODElib.parameter
    pri = {n: P(stats_gen=LN, hyperparameters={"s": s, "scale": sc}, init_value=sc) for n, s, sc in PRIORS[name]}
    if name == "zero_i":
        df = df.replace({"virus": "V", "host": "S"})
        return ODElib.ModelFramework(ODE=zero_i, parameter_names=["mu", "phi", "beta"], state_names=["S", "V"],
                                     dataframe=df, t_steps=288, **pri)
    df = df.replace({"virus": "V", "host": "H"})
    if name == "one_i":
        return ODElib.ModelFramework(ODE=one_i, parameter_names=["mu", "phi", "beta", "lam"],
                                     state_names=["S", "I1", "V"], dataframe=df,
                                     state_summations={"H": ["S", "I1"]}, S=5236900, **pri)
    return ODElib.ModelFramework(ODE=two_i, parameter_names=["mu", "phi", "beta", "lam", "tau"],
                                 state_names=["S", "I1", "I2", "V"], dataframe=df,
                                 state_summations={"H": ["S", "I1", "I2"]}, S=5236900, **pri)


@contextlib.contextmanager
def tolerance(tol):
    """Rebind the module-level name the reference calls at Framework.py:656."""
    if tol is None:
        yield
        return
    old = RF.odeint
    RF.odeint = functools.partial(_odeint, rtol=tol, atol=tol, mxstep=200000)
    try:
        yield
    finally:
        RF.odeint = old


def solve(model, theta):
    pred = model.integrate(parameters=(list(theta),), predict_obs=True, as_dataframe=False)
    chi = model.get_chi(pred)
    r2 = model.get_Rsqrd(pred)
    vec = np.concatenate([pred[s] for s in pred])
    return vec, (np.nan if chi is np.ma.masked else float(chi)), float(r2)


def record_chain(model, theta0, seed, nits, tol):
    """Run the reference sampler, spying on its RNG calls and on get_chi."""
    m = model.copy(overwrite=dict(zip(model.get_pnames(), theta0)))
    m.random_seed = seed
    zs, us, thetas, chis = [], [], [], []
    real_normal, real_rand = np.random.normal, np.random.rand
    real_chi = m.get_chi

    def normal(*a, **k):
        v = real_normal(*a, **k)
        zs.append(float(v))
        return v

    def rand(*a, **k):
        v = real_rand(*a, **k)
        us.append(float(v))
        return v

    def get_chi(mod_dict):
        c = real_chi(mod_dict)
        thetas.append([float(m.parameters[p].val) for p in m.get_pnames()])
        chis.append(np.nan if c is np.ma.masked else float(c))
        return c

    np.random.normal, np.random.rand, m.get_chi = normal, rand, get_chi
    try:
        with tolerance(tol), contextlib.redirect_stdout(io.StringIO()):
            frame = RS.MetropolisHastings(m, nits=nits, burnin=int(nits / 2), print_progress=False)
    finally:
        np.random.normal, np.random.rand = real_normal, real_rand
    P = len(m.get_pnames())
    n_iter = nits - 1
    z = np.array(zs).reshape(n_iter, P)
    u = np.array(us)
    props = np.array(thetas[1:])     # first get_chi call is the a-priori solve (Samplers.py:89)
    chinew = np.array(chis[1:])
    cols = m.get_pnames() + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"]
    kept = frame[cols].to_numpy(dtype=float)
    # decisions from the running acceptance ratio are only visible after burn-in; recompute exactly
    # as the reference does (Samplers.py:124-127) from its own chi values
    acc = np.zeros(n_iter, bool)
    chi_cur = chis[0]
    for k in range(n_iter):
        with np.errstate(all="ignore"):
            a = np.exp(np.log(np.exp(chi_cur - chinew[k])))
        if a > u[k]:
            acc[k] = True
            chi_cur = chinew[k]
    # consistency: running acceptance of the reference frame must equal ours
    it = kept[:, P + 3].astype(int)
    assert np.allclose(kept[:, P + 4], np.cumsum(acc)[it - 1] / it, rtol=0, atol=1e-15)
    return dict(z=z, u=u, proposals=props, chinew=chinew, accepted=acc, kept=kept,
                chi0=np.array(chis[0]), theta0=np.array(theta0, float))


def main():
    rng = np.random.default_rng(20261018)
    for name in ("zero_i", "one_i", "two_i"):
        model = make_model(name)
        pri = PRIORS[name]
        K_prior = 28
        draws = np.column_stack([sc * np.exp(s * rng.standard_normal(K_prior)) for _, s, sc in pri])
        theta = np.vstack([np.array(POSTERIOR[name], float), draws])
        out = {"theta": theta, "times": model.times, "y0": np.asarray(model.get_inits(), float),
               "pnum": np.array(model._pnum)}
        order = [s for s in model.get_snames(after_summation=True) if s in model._pred_tindex]
        out["obs_order"] = np.array(order)
        for s in order:
            out["tindex_" + s] = model._pred_tindex[s]
        out["ln_obs"] = np.concatenate([model._obs_logabundance[s] for s in order])
        out["log_sigma"] = np.concatenate([model._obs_logsigma[s] for s in order])
        out["cutchi6"] = np.array(float(model.get_chi(
            {s: np.exp(model._obs_logabundance[s] + 6.0 * model._obs_logsigma[s]) for s in model._obs_logabundance})))
        for tag, tol in (("def", None), ("tight", 1e-13)):
            res = []
            with tolerance(tol):
                for th in theta:
                    res.append(solve(model, th))
            out["pred_" + tag] = np.array([r[0] for r in res])
            out["chi_" + tag] = np.array([r[1] for r in res])
            out["r2_" + tag] = np.array([r[2] for r in res])
        start = POSTERIOR[name][-1]
        for tag, tol, nits, seeds in (("def", None, 400, (0, 1)), ("tight", 1e-13, 200, (0,))):
            for seed in seeds:
                ch = record_chain(model, start, seed, nits, tol)
                for k, v in ch.items():
                    out[f"chain_{tag}_s{seed}_{k}"] = v
                out[f"chain_{tag}_s{seed}_nits"] = np.array(nits)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, {k: np.shape(v) for k, v in out.items() if k.startswith(("theta", "pred", "chain_def_s0_kept"))})


if __name__ == "__main__":
    main()
