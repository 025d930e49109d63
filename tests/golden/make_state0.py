#!/usr/bin/env python
"""Generate tests/golden/state0.npz by running the UNMODIFIED reference (/root/reference) on a model whose
parameters include '<state>0' names.  Build container only; shims as make_golden.py.

What the reference does with such parameters (and what the vectors pin):
  * ``integrate`` / ``_Fit_worker`` / the chain's a-priori solve start from ``istates`` -- the '<state>0' VALUE IS
    IGNORED there (Framework.py:647-650, :41-48; Samplers.py:88);
  * the proposal loop copies every '<state>0' parameter (static ones too) into the initial states
    (Samplers.py:110-114), the reject branch restores the walking ones (:139-143);
  * ``set_best_params`` copies the best row's '<state>0' values into the initial states (Framework.py:730-731).
The fixture's S0 / V0 start values differ from the t == 0 data on purpose.

    python tests/golden/make_state0.py
"""
import contextlib
import io
import os

import numpy as np
import pandas as pd

import make_golden as G   # installs the shims, imports the reference

HERE = os.path.dirname(os.path.abspath(__file__))
ODElib, RS, LN = G.ODElib, G.RS, G.LN

PNAMES = ["mu", "phi", "beta", "S0", "V0"]
START = (1.36e-8, 1.35e-8, 19.44, 5.2e6, 1.1e7)          # S0, V0 != the data's t == 0 rows (5236900, 10981000)
PRI = [("mu", 3, 1e-8), ("phi", 3, 1e-8), ("beta", 1, 25), ("S0", 0.3, 5.0e6), ("V0", 0.3, 1.1e7)]


def zero_i_s0(y, t, ps):
    mu, phi, beta = ps[0], ps[1], ps[2]                   # ps[3], ps[4] = S0, V0: initial values, not rates
    S, V = y[0], y[1]
    return np.array([mu * S - phi * S * V, beta * phi * S * V - phi * S * V])


def make_model():
    df = pd.read_csv(os.path.join(HERE, "demodata.csv")).replace({"virus": "V", "host": "S"})
    P = ODElib.parameter
    pri = {n: P(stats_gen=LN, hyperparameters={"s": s, "scale": sc}, init_value=v) for (n, s, sc), v in zip(PRI, START)}
    return ODElib.ModelFramework(ODE=zero_i_s0, parameter_names=PNAMES, state_names=["S", "V"], dataframe=df,
                                 t_steps=288, **pri)


def record_chain(model, theta0, seed, nits, static=()):
    """As make_golden.record_chain, with static parameters and the initial states seen by every solve."""
    m = model.copy(overwrite=dict(zip(model.get_pnames(), theta0)))
    m.random_seed = seed
    zs, us, thetas, chis, y0s = [], [], [], [], []
    real_normal, real_rand, real_chi = np.random.normal, np.random.rand, m.get_chi

    def normal(*a, **k):
        v = real_normal(*a, **k); zs.append(float(v)); return v

    def rand(*a, **k):
        v = real_rand(*a, **k); us.append(float(v)); return v

    def get_chi(mod_dict):
        c = real_chi(mod_dict)
        thetas.append([float(m.parameters[p].val) for p in m.get_pnames()])
        y0s.append([float(v) for v in m.get_inits()])
        chis.append(np.nan if c is np.ma.masked else float(c))
        return c

    np.random.normal, np.random.rand, m.get_chi = normal, rand, get_chi
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            frame = RS.MetropolisHastings(m, nits=nits, burnin=int(nits / 2), static_parameters=set(static),
                                          print_progress=False)
    finally:
        np.random.normal, np.random.rand = real_normal, real_rand
    n_walk = len(PNAMES) - len(static)
    n_iter = nits - 1
    cols = m.get_pnames() + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"]
    chinew = np.array(chis[1:])
    acc = np.zeros(n_iter, bool)
    chi_cur = chis[0]
    u = np.array(us)
    for k in range(n_iter):
        with np.errstate(all="ignore"):
            a = np.exp(np.log(np.exp(chi_cur - chinew[k])))
        if a > u[k]:
            acc[k] = True
            chi_cur = chinew[k]
    return dict(z=np.array(zs).reshape(n_iter, n_walk), u=u, proposals=np.array(thetas[1:]), chinew=chinew,
                accepted=acc, kept=frame[cols].to_numpy(dtype=float), chi0=np.array(chis[0]),
                y0_apriori=np.array(y0s[0]), y0_solves=np.array(y0s[1:]), theta0=np.array(theta0, float),
                final_inits=np.asarray(m.get_inits(), float),
                final_theta=np.array([float(m.parameters[p].val) for p in m.get_pnames()]))


def main():
    model = make_model()
    out = {"start": np.array(START), "y0": np.asarray(model.get_inits(), float), "times": model.times,
           "pnum": np.array(model._pnum)}
    # integrate / the survey seam at parameter sets whose S0, V0 differ from istates: the values are ignored
    rng = np.random.default_rng(7)
    theta = np.array(START) * np.exp(0.2 * rng.standard_normal((6, 5)))
    theta[0] = START
    res = [G.solve(model, th) for th in theta]
    out["theta"] = theta
    out["pred_def"] = np.array([r[0] for r in res])
    out["chi_def"] = np.array([r[1] for r in res])
    out["r2_def"] = np.array([r[2] for r in res])
    with contextlib.redirect_stdout(io.StringIO()):
        fw = G.RF._Fit_worker(model.copy(), [tuple(th) for th in theta])
    out["fit_worker_chi"] = fw["chi"].to_numpy(dtype=float)
    for tag, static in (("walk", ()), ("staticV0", ("V0",))):
        ch = record_chain(model, START, 3, 300, static)
        for k, v in ch.items():
            out[f"chain_{tag}_{k}"] = v
        out[f"chain_{tag}_nits"] = np.array(300)
    # set_best_params: the best row's S0 / V0 become the initial states (Framework.py:725-731)
    kept = out["chain_walk_kept"]
    post = pd.DataFrame(kept, columns=PNAMES + ["chi", "rsquared", "aic", "iteration", "acceptance_ratio"])
    post["chain#"] = 0
    m2 = model.copy()
    m2.set_best_params(post)
    out["best_inits"] = np.asarray(m2.get_inits(), float)
    out["best_theta"] = np.array([float(m2.parameters[p].val) for p in PNAMES])
    path = os.path.join(HERE, "state0.npz")
    np.savez_compressed(path, **out)
    print("->", path, "chi_def", out["chi_def"][:3], "accept rate", out["chain_walk_accepted"].mean(),
          "y0 apriori", out["chain_walk_y0_apriori"], "first solve y0", out["chain_walk_y0_solves"][0])


if __name__ == "__main__":
    main()
