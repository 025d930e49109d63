#!/usr/bin/env python
"""Generate tests/golden/array_param.npz by running the UNMODIFIED reference (/root/reference) on a model with an
ARRAY-valued parameter.  Build container only; same three harness shims as make_golden.py.

The reference hands parameter arrays straight to the RHS (Framework.py:656): ctor, integrate and get_chi work with
them; its sample_lhs (Samplers.py:45) and MetropolisHastings (Framework.py:99) branches raise, which is recorded too.

    python tests/golden/make_array_param.py
"""
import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
mpl = types.ModuleType("matplotlib")
plt = types.ModuleType("matplotlib.pyplot")
mpl.pyplot = plt
sys.modules["matplotlib"] = mpl
sys.modules["matplotlib.pyplot"] = plt
pyd = types.ModuleType("pyDOE2")
pyd.lhs = lambda n, samples=None: np.random.rand(samples or n, n)
sys.modules["pyDOE2"] = pyd
pd.Series.iteritems = pd.Series.items
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
import ODElib  # noqa: E402
from ODElib.Statistics import Samplers  # noqa: E402


def split_phi(y, t, ps):                                          # zero_i with phi = phi[0] + phi[1]
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    a = phi[0] + phi[1]
    return np.array([mu * S - a * S * V, beta * a * S * V - a * S * V])


df = pd.read_csv(os.path.join(HERE, "demodata.csv"))
df["organism"] = df["organism"].map({"virus": "V", "host": "S"})
LN = scipy.stats.lognorm
PHI = [0.7e-8, 0.65e-8]
m = ODElib.ModelFramework(
    ODE=split_phi, parameter_names=["mu", "phi", "beta"], state_names=["S", "V"], dataframe=df,
    mu=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 3, "scale": 1e-8}, init_value=1e-6),
    phi=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 3, "scale": 1e-8}, init_value=PHI),
    beta=ODElib.parameter(stats_gen=LN, hyperparameters={"s": 1, "scale": 25}, init_value=19.4),
    t_end=3, t_steps=288)
pred = m.integrate(predict_obs=True, as_dataframe=False)
chi = float(m.get_chi(pred))
raised = {}
for name, call in (("MetropolisHastings", lambda: Samplers.MetropolisHastings(m, nits=20)), ("fit_survey", lambda: m.fit_survey(samples=5))):
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            call()
        raised[name] = ""
    except Exception as exc:  # noqa: BLE001 - the point is to record what the reference does
        raised[name] = type(exc).__name__
np.savez(os.path.join(HERE, "array_param.npz"), theta=np.array([1e-6, PHI[0], PHI[1], 19.4]), chi=chi,
         pred_S=pred["S"], pred_V=pred["V"], y0=np.asarray(m.get_inits(), float),
         raises=np.array([raised["MetropolisHastings"], raised["fit_survey"]]))
print("chi", chi, "raises", raised)
