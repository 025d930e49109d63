"""Run the UNMODIFIED reference (oracle/_ref, installed by oracle/make_ref.py)  --  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.  The reference does not import as-is in this
image; three shims are installed first, none of which touches reference source or arithmetic:
  1. stub ``matplotlib`` / ``matplotlib.pyplot`` (Framework.py:6 -- plotting only, not installed here);
  2. stub ``pyDOE2.lhs`` (Samplers.py:3, :33 -- not installed here): the classic Latin-hypercube design, one uniform
     point per stratum and dimension, strata permuted per dimension, from numpy's global RandomState as pyDOE2 draws;
  3. ``pd.Series.iteritems = pd.Series.items`` (Framework.py:247, :276 -- removed in pandas >= 2).
"""
import contextlib
import io
import os
import sys
import time
import types
import warnings

import numpy as np
import pandas as pd
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_mod = None


def available():
    return os.path.exists(os.path.join(REF, "ODElib", "Framework.py"))


def load():
    """-> the reference's ``ODElib`` module (imported from oracle/_ref under the shims)."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError("oracle/_ref is absent: run `python oracle/make_ref.py` in the build container")
    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if "pyDOE2" not in sys.modules:
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
    if not hasattr(pd.Series, "iteritems"):
        pd.Series.iteritems = pd.Series.items
    if REF not in sys.path:
        sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")                  # as the demo notebook does (Demo_InfectionStates.ipynb:29-31)
    import ODElib
    assert os.path.realpath(ODElib.__file__).startswith(os.path.realpath(REF)), ODElib.__file__
    _mod = ODElib
    return ODElib


# the demo models as the notebook writes them (Demo_InfectionStates.ipynb:60-128); module level: Pool pickles them
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
    return np.array([mu * S - phi * S * V, phi * S * V - lam * I1, beta * lam * I1 - phi * S * V])


def zero_i(y, t, ps):
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    return np.array([mu * S - phi * S * V, beta * phi * S * V - phi * S * V])


PRIORS = {  # Demo_InfectionStates.ipynb:885-891, :8575-8578, :17472-17476 (lognorm s, scale)
    "zero_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25)],
    "one_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1)],
    "two_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1), ("tau", 2, 1)],
}


def demo_model(name, demodata_csv, init=None):
    """The notebook's ModelFramework for ``name`` on demo/demodata.csv (:843, :8581-8595, :17479-17499)."""
    ODElib = load()
    df = pd.read_csv(demodata_csv)
    pri = {n: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc},
                               init_value=(sc if init is None else init[i]))
           for i, (n, s, sc) in enumerate(PRIORS[name])}
    names = [p[0] for p in PRIORS[name]]
    with contextlib.redirect_stdout(io.StringIO()):
        if name == "zero_i":
            return ODElib.ModelFramework(ODE=zero_i, parameter_names=names, state_names=["S", "V"],
                                         dataframe=df.replace({"virus": "V", "host": "S"}), t_steps=288, **pri)
        df = df.replace({"virus": "V", "host": "H"})
        if name == "one_i":
            return ODElib.ModelFramework(ODE=one_i, parameter_names=names, state_names=["S", "I1", "V"], dataframe=df,
                                         state_summations={"H": ["S", "I1"]}, S=5236900, **pri)
        return ODElib.ModelFramework(ODE=two_i, parameter_names=names, state_names=["S", "I1", "I2", "V"], dataframe=df,
                                     state_summations={"H": ["S", "I1", "I2"]}, S=5236900, **pri)


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


def time_fit_survey(model, samples, cores, seed=0):
    """``ModelFramework.fit_survey(samples, cpu_cores)`` (Framework.py:800-816: LHS of the priors, round-robin chunks,
    multiprocessing.Pool fan-out of _Fit_worker, concat) -> (solves/s, seconds, frame)."""
    np.random.seed(seed)
    t0 = time.perf_counter()
    frame = _quiet(model.fit_survey, samples=int(samples), cpu_cores=int(cores))
    dt = time.perf_counter() - t0
    return samples / dt, dt, frame


def time_single_chain(model, nits, seed=0):
    """``Samplers.MetropolisHastings(model, nits)`` on a copy (Samplers.py:53-174) -> (chain-steps/s, seconds, frame)."""
    from ODElib.Statistics import Samplers
    m = model.copy()
    m.random_seed = seed
    t0 = time.perf_counter()
    frame = _quiet(Samplers.MetropolisHastings, m, nits=int(nits), print_progress=False)
    dt = time.perf_counter() - t0
    return (nits - 1) / dt, dt, frame


def time_mcmc(model, starts, nits, cores):
    """``ModelFramework.MCMC(chain_inits=[dict...], iterations_per_chain, cpu_cores)`` (Framework.py:946-1061) with
    explicit chain starts (no survey) -> (chain-steps/s over all chains, seconds, frame)."""
    m = model.copy()
    t0 = time.perf_counter()
    frame = _quiet(m.MCMC, chain_inits=list(starts), iterations_per_chain=int(nits), cpu_cores=int(cores),
                   print_report=False)
    dt = time.perf_counter() - t0
    return len(starts) * (nits - 1) / dt, dt, frame
