"""Beyond the three demo models: the N-class infection chain (BASELINE config 3, n = N+2 states) and the
5-host x 5-virus network (config 5, 35 states / 40 parameters), traced from Python like any user model and
checked against the oracle (scipy odeint) on synthetic data of the demo's shape."""
import numpy as np
import pytest

from odelib_b200 import demo_models
from oracle import odelib_oracle as orc
from tests.helpers import synthetic_problem

pytestmark = pytest.mark.gpu


def nclass_problem(N):
    names = ["S"] + [f"I{k}" for k in range(1, N + 1)] + ["V"]
    sums = {"H": ["S"] + [f"I{k}" for k in range(1, N + 1)]}
    center = np.array([0.3, 1.0e-7, 20.0, 2.0, 2.8 * N / 2])
    y0 = [5236900.0] + [0.0] * N + [10981000.0]
    return demo_models.n_class(N), names, sums, center, y0, ["H", "V"]


@pytest.mark.parametrize("N", [3, 6, 10])
def test_n_class_chain_sweep_and_mcmc(N):
    rhs, names, sums, center, y0, orgs = nclass_problem(N)
    dm, tab = synthetic_problem(rhs, names, sums, center, y0, orgs, seed=N)
    rng = np.random.default_rng(N)
    theta = center * np.exp(0.1 * rng.standard_normal((24, 5)))
    out = dm.sweep(theta, rtol=1e-11, atol=1e-11, return_pred=True, max_steps=2000000)
    assert np.all(out["status"] == 0)
    for k in range(0, 24, 4):
        vec, chi, r2 = orc.solve_unit(rhs, theta[k], tab, 1e-12, 1e-12, mxstep=500000)
        np.testing.assert_allclose(out["pred"][k], vec, rtol=2e-8, atol=1e-6)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=5e-8)
    # chains on host streams against the oracle chain
    nits, C = 30, 4
    z = 0.05 * rng.standard_normal((C, nits - 1, 5))
    u = rng.random((C, nits - 1))
    res = dm.mcmc(theta[:C], nits=nits, rng_mode="host", z=z, u=u, rtol=1e-11, atol=1e-11, trace=True, max_steps=2000000)
    ref = orc.mh_chain(rhs, theta[0], tab, 5, nits=nits, z=z[0], u=u[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(res["chinew"][0], ref["chinew"], rtol=5e-8)
    assert np.array_equal(res["accepted"][0].astype(bool), ref["accepted"])
    # the stiff steppers compile and agree for this state count as well
    rad = dm.sweep(theta[:4], solver="radau5", rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(rad["chi"], out["chi"][:4], rtol=1e-5)


def test_five_by_five_network():
    rhs, n, P, groups = demo_models.network(5, 5)
    H, V = 5, 5
    names = [f"S{i}" for i in range(H)] + [f"I{i}{j}" for i in range(H) for j in range(V)] + [f"V{j}" for j in range(V)]
    sums = {f"H{i}": [f"S{i}"] + [f"I{i}{j}" for j in range(V)] for i in range(H)}
    rng = np.random.default_rng(1)
    center = np.concatenate([0.3 * np.exp(0.2 * rng.standard_normal(H)), 2e-8 * np.exp(0.5 * rng.standard_normal(H * V)),
                             20 * np.exp(0.1 * rng.standard_normal(V)), 2.0 * np.exp(0.2 * rng.standard_normal(V))])
    y0 = [1e6 * (1 + i) for i in range(H)] + [0.0] * (H * V) + [2e6 * (1 + j) for j in range(V)]
    orgs = [f"H{i}" for i in range(H)] + [f"V{j}" for j in range(V)]
    dm, tab = synthetic_problem(rhs, names, sums, center, y0, orgs, seed=1)
    assert dm.n_state == 35 and dm.n_param == 40 and dm.n_obs == 180
    theta = center * np.exp(0.05 * rng.standard_normal((40, P)))
    out = dm.sweep(theta, rtol=1e-10, atol=1e-10, return_pred=True, max_steps=2000000)
    assert np.all(out["status"] == 0)
    for k in (0, 13, 39):
        vec, chi, r2 = orc.solve_unit(rhs, theta[k], tab, 1e-12, 1e-12, mxstep=500000)
        np.testing.assert_allclose(out["pred"][k], vec, rtol=1e-7, atol=1e-5)
        np.testing.assert_allclose(out["chi"][k], chi, rtol=1e-6)
    res = dm.mcmc(theta[:32], nits=24, seed=3)
    assert np.isfinite(res["samples"]).all() and res["samples"].shape == (32, 11, P + 5)
    assert res["fail_count"].sum() == 0
