"""Shared test helpers: demo workloads built both for the oracle and for the product."""
import os

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Demo_InfectionStates.ipynb:885-891, :8575-8578, :17472-17476  (lognorm s, scale)
PRIORS = {
    "zero_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25)],
    "one_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1)],
    "two_i": [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 20), ("lam", 2, 0.1), ("tau", 2, 1)],
}
STATES = {"zero_i": ["S", "V"], "one_i": ["S", "I1", "V"], "two_i": ["S", "I1", "I2", "V"]}
SUMS = {"zero_i": None, "one_i": {"H": ["S", "I1"]}, "two_i": {"H": ["S", "I1", "I2"]}}
TSTEPS = {"zero_i": 288, "one_i": 1000, "two_i": 1000}


def demo_df(name):
    df = pd.read_csv(os.path.join(GOLDEN, "demodata.csv"))
    return df.replace({"virus": "V", "host": "S" if name == "zero_i" else "H"})


def golden(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


def oracle_tables(name):
    from oracle import odelib_oracle as orc
    inits = None if name == "zero_i" else {"S": 5236900}
    return orc.build_tables(demo_df(name), STATES[name], SUMS[name], TSTEPS[name], inits)


def oracle_rhs(name):
    from oracle import odelib_oracle as orc
    return getattr(orc, name)


def obs_tables_from_oracle(tab):
    """oracle Tables -> product ObsTables (same content the facade derives itself)."""
    from odelib_b200.engine import ObsTables
    cols = tab.out_columns()
    columns = [(cols[s], tab.tindex[s], tab.ln_obs[s], tab.log_sigma[s]) for s in tab.obs_order]
    return ObsTables(tab.times, columns)


def device_model(name, **kw):
    """DeviceModel of a demo workload with the demo data loaded."""
    from odelib_b200 import demo_models
    from odelib_b200.engine import DeviceModel
    f, n, P, groups = demo_models.MODELS[name]
    dm = DeviceModel(f, n, P, groups, **kw)
    tab = oracle_tables(name)
    dm.set_data(obs_tables_from_oracle(tab), tab.y0)
    dm.set_grid(tab.times, tab.y0)
    return dm, tab


def prior_draws(name, n, seed=0):
    rng = np.random.default_rng(seed)
    return np.column_stack([sc * np.exp(s * rng.standard_normal(n)) for _, s, sc in PRIORS[name]])


DEMO_TIMES = [0.0, 0.2, 0.3, 0.5, 0.7, 0.9, 1.0, 1.2, 1.3, 1.5, 1.7, 1.8, 2.0, 2.2, 2.3, 2.5, 2.8, 3.0]


def synthetic_problem(rhs, state_names, sums, center, y0, organisms, seed=0, sigma=0.2, t_steps=1000, device_kw=None):
    """Synthetic data set: the oracle's odeint at `center` + log-normal noise on the demo's
    time points, one block of rows per observed organism.  Returns (DeviceModel with the data loaded, oracle tables)."""
    import pandas as pd
    from odelib_b200.engine import DeviceModel
    from oracle import odelib_oracle as orc
    rows = [{"organism": o, "time": t, "abundance": 1.0, "log_sigma": sigma} for o in organisms for t in DEMO_TIMES]
    df0 = pd.DataFrame(rows)
    inits = dict(zip(state_names, y0))
    tab0 = orc.build_tables(df0, state_names, sums, t_steps, inits)
    tab0.y0 = np.asarray(y0, float)
    vec, _, _ = orc.solve_unit(rhs, center, tab0, 1e-12, 1e-12, y0=np.asarray(y0, float), mxstep=500000)
    rng = np.random.default_rng(seed)
    df = df0.sort_values(by=["organism", "time"], kind="stable").reset_index(drop=True)
    # rows of tab0 are ordered like chi concatenates them: by out-column order, which is not the alphabetical
    # organism order in general -> assign per organism
    off = 0
    for o in tab0.obs_order:
        nrow = len(tab0.tindex[o])
        df.loc[df["organism"] == o, "abundance"] = np.maximum(vec[off:off + nrow], 1e-3) * np.exp(sigma * rng.standard_normal(nrow))
        off += nrow
    tab = orc.build_tables(df, state_names, sums, t_steps, inits)
    tab.y0 = np.asarray(y0, float)
    n, P = len(state_names), len(center)
    groups = [tab.sum_index.get(i, (i,)) for i in tab.keep] if tab.sum_index else [(i,) for i in range(n)]
    dm = DeviceModel(rhs, n, P, groups, **(device_kw or {}))
    dm.set_data(obs_tables_from_oracle(tab), tab.y0)
    return dm, tab
