"""The synthetic workloads of BASELINE.json configs 3-5 as ``ModelFramework`` instances.

None of them exists in the reference (its only data set is demo/demodata.csv): they generalise the demo's infection
models and are built through the reference-facing surface -- a user RHS in Python, parameter objects with lognorm
priors, a DataFrame of observations -- so that everything measured on them goes through the same tracer / NVRTC /
C-ABI path as a user's model.  The observations are synthetic: the model's own trajectory at the centre of the priors
(integrated on the GPU at tight tolerance through ``ModelFramework.integrate``) with log-normal noise, on the demo's
observation times.  ``spec_only=True`` returns the RHS and sizes without touching a GPU (used by the build to put the
NVRTC cubins of these models into the cache).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.stats

from . import demo_models

DEMO_TIMES = [0.0, 0.2, 0.3, 0.5, 0.7, 0.9, 1.0, 1.2, 1.3, 1.5, 1.7, 1.8, 2.0, 2.2, 2.3, 2.5, 2.8, 3.0]


def _lognorm(center, s):
    from .Framework import parameter
    return parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": float(center)}, init_value=float(center))


def _with_synthetic_data(make, organisms, sigma, seed):
    """make(df) -> ModelFramework.  Observations = the model's own predictions at its current parameters x log-normal
    noise (sigma in log space), one block of rows per observed organism on the demo's time points."""
    rows = [{"organism": o, "time": t, "abundance": 1.0, "log_sigma": sigma} for o in organisms for t in DEMO_TIMES]
    m = make(pd.DataFrame(rows))
    tight = m.rtol, m.atol
    m.rtol = m.atol = 1e-11
    pred = m.integrate(predict_obs=True, as_dataframe=False)
    m.rtol, m.atol = tight
    rng = np.random.default_rng(seed)
    data = []
    for o in organisms:
        vals = np.maximum(np.asarray(pred[o], dtype=float), 1e-3) * np.exp(sigma * rng.standard_normal(len(DEMO_TIMES)))
        data += [{"organism": o, "time": t, "abundance": float(v), "log_sigma": sigma} for t, v in zip(DEMO_TIMES, vals)]
    m.reset_dataframe(pd.DataFrame(data))
    return m


def nclass(N, device=None, spec_only=False, seed=None):
    """Config 3: S, I1..IN, V with N latent infected classes (N = 1: the demo's one_i on the demo data); H = S + sum I."""
    from .Framework import ModelFramework
    if N == 1:
        rhs, n, P, groups = demo_models.MODELS["one_i"]
        center = np.array([1.238e-08, 3.550e-08, 19.40, 1.835])
        if spec_only:
            return rhs, n, P, groups
        import os
        df = pd.read_csv(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "demodata.csv"))
        pri = {p: _lognorm(c, s) for p, c, s in zip(demo_models.PARAMETER_NAMES["one_i"], center, (3, 3, 1, 2))}
        m = ModelFramework(ODE=rhs, parameter_names=demo_models.PARAMETER_NAMES["one_i"], state_names=demo_models.STATE_NAMES["one_i"],
                           dataframe=df.replace({"virus": "V", "host": "H"}), state_summations={"H": ["S", "I1"]}, S=5236900,
                           device=device, **pri)
        return m, center
    rhs = demo_models.n_class(N)
    names = ["S"] + [f"I{k}" for k in range(1, N + 1)] + ["V"]
    groups = [tuple(range(N + 1)), (N + 1,)]
    if spec_only:
        return rhs, N + 2, 5, groups
    center = np.array([0.3, 1.0e-7, 20.0, 2.0, 2.8 * N / 2])
    pn = ["mu", "phi", "beta", "lam", "tau"]
    inits = {"S": 5236900.0, "V": 10981000.0}

    def make(df):
        pri = {p: _lognorm(c, 0.5) for p, c in zip(pn, center)}
        return ModelFramework(ODE=rhs, parameter_names=pn, state_names=names, dataframe=df, state_summations={"H": names[:-1]},
                              device=device, **inits, **pri)
    return _with_synthetic_data(make, ["H", "V"], 0.2, N if seed is None else seed), center


def network(H=5, V=5, device=None, spec_only=False, seed=1):
    """Config 5: H hosts x V viruses, H + H V + V states (35), H + H V + 2 V parameters (40); observables
    H_i = S_i + sum_j I_ij and V_j."""
    from .Framework import ModelFramework
    rhs, n, P, groups = demo_models.network(H, V)
    if spec_only:
        return rhs, n, P, groups
    names = [f"S{i}" for i in range(H)] + [f"I{i}{j}" for i in range(H) for j in range(V)] + [f"V{j}" for j in range(V)]
    sums = {f"H{i}": [f"S{i}"] + [f"I{i}{j}" for j in range(V)] for i in range(H)}
    rng = np.random.default_rng(seed)
    center = np.concatenate([0.3 * np.exp(0.2 * rng.standard_normal(H)), 2e-8 * np.exp(0.5 * rng.standard_normal(H * V)),
                             20 * np.exp(0.1 * rng.standard_normal(V)), 2.0 * np.exp(0.2 * rng.standard_normal(V))])
    pn = [f"mu{i}" for i in range(H)] + [f"phi{i}{j}" for i in range(H) for j in range(V)] + [f"beta{j}" for j in range(V)] + \
         [f"lam{j}" for j in range(V)]
    inits = {f"S{i}": 1e6 * (1 + i) for i in range(H)}
    inits.update({f"V{j}": 2e6 * (1 + j) for j in range(V)})

    def make(df):
        pri = {p: _lognorm(c, 0.5) for p, c in zip(pn, center)}
        return ModelFramework(ODE=rhs, parameter_names=pn, state_names=names, dataframe=df, state_summations=sums, device=device,
                              **inits, **pri)
    return _with_synthetic_data(make, [f"H{i}" for i in range(H)] + [f"V{j}" for j in range(V)], 0.2, seed), center


def stiff_thetas(n, seed=0):
    """Config 4: two_i with tau ~ lognorm(0.5, 1e4), lam ~ lognorm(0.5, 1e-2) -- Jacobian spectrum over six orders of
    magnitude -- and mu, phi, beta around (0.5, 1e-7, 50)."""
    rng = np.random.default_rng(seed)
    th = np.empty((n, 5))
    th[:, 0] = 0.5 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 1] = 1e-7 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 2] = 50.0 * np.exp(0.2 * rng.standard_normal(n))
    th[:, 3] = 1e-2 * np.exp(0.5 * rng.standard_normal(n))
    th[:, 4] = 1e4 * np.exp(0.5 * rng.standard_normal(n))
    return th


def stiff(device=None):
    """Config 4: the demo's two_i model on the demo data with the stiff priors above."""
    import os
    from .Framework import ModelFramework
    rhs, n, P, groups = demo_models.MODELS["two_i"]
    center = np.array([0.5, 1e-7, 50.0, 1e-2, 1e4])
    df = pd.read_csv(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "demodata.csv"))
    pri = {p: _lognorm(c, s) for p, c, s in zip(demo_models.PARAMETER_NAMES["two_i"], center, (0.2, 0.2, 0.2, 0.5, 0.5))}
    m = ModelFramework(ODE=rhs, parameter_names=demo_models.PARAMETER_NAMES["two_i"], state_names=demo_models.STATE_NAMES["two_i"],
                       dataframe=df.replace({"virus": "V", "host": "H"}), state_summations={"H": ["S", "I1", "I2"]}, S=5236900,
                       device=device, **pri)
    return m, center


# models whose cubins the build puts into the cache beside the demo models' (name -> spec_only tuple)
def cache_specs():
    return {"n_class_4": nclass(4, spec_only=True), "n_class_10": nclass(10, spec_only=True), "network_5x5": network(spec_only=True)}
