"""Beyond the three demo models: the N-class infection chain (BASELINE config 3, n = N+2 states) and the
5-host x 5-virus network (config 5, 35 states / 40 parameters), traced from Python like any user model and
checked against the oracle (scipy odeint) on synthetic data of the demo's shape."""
import numpy as np
import pytest

from odelib_b200 import engine

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


@pytest.mark.parametrize("which", ["n_class_10", "network_5x5"])
def test_cooperative_mapping_agrees_with_thread_per_system(which):
    """n > 8: the default kernels spread a system over several lanes (state slices in registers, RHS from a shared row);
    asking for speculate=1 selects the thread-per-system kernel (arrays in local memory).  Same algorithm, different
    summation order in the error norm: chi agrees to solver accuracy, decisions agree, the AUTO sweep agrees."""
    rng = np.random.default_rng(7)
    if which == "n_class_10":
        rhs, names, sums, center, y0, orgs = nclass_problem(10)
        P = 5
    else:
        rhs, n, P, groups = demo_models.network(5, 5)
        H, V = 5, 5
        names = [f"S{i}" for i in range(H)] + [f"I{i}{j}" for i in range(H) for j in range(V)] + [f"V{j}" for j in range(V)]
        sums = {f"H{i}": [f"S{i}"] + [f"I{i}{j}" for j in range(V)] for i in range(H)}
        r1 = np.random.default_rng(1)
        center = np.concatenate([0.3 * np.exp(0.2 * r1.standard_normal(H)), 2e-8 * np.exp(0.5 * r1.standard_normal(H * V)),
                                 20 * np.exp(0.1 * r1.standard_normal(V)), 2.0 * np.exp(0.2 * r1.standard_normal(V))])
        y0 = [1e6 * (1 + i) for i in range(H)] + [0.0] * (H * V) + [2e6 * (1 + j) for j in range(V)]
        orgs = [f"H{i}" for i in range(H)] + [f"V{j}" for j in range(V)]
    dm, tab = synthetic_problem(rhs, names, sums, center, y0, orgs, seed=1)
    assert dm.kernel_info("sweep_coop")["regs"] > 0
    theta = center * np.exp(0.05 * rng.standard_normal((37, P)))          # ragged: 37 systems do not fill the groups
    coop = dm.sweep(theta, rtol=1e-10, atol=1e-10, max_steps=2000000)
    tps = dm.sweep(theta, rtol=1e-10, atol=1e-10, max_steps=2000000, stiff_check=True)   # stiff_check -> thread-per-system
    assert np.all(coop["status"] == 0) and np.all(tps["status"] == 0)
    np.testing.assert_allclose(coop["chi"], tps["chi"], rtol=1e-7)
    # AUTO's bulk pass is the cooperative kernel (rows it finishes within its step cap carry exactly its numbers)
    auto = dm.sweep(theta, solver="auto", max_steps=2000000)
    plain = dm.sweep(theta, max_steps=engine.AUTO_CAP)
    fin = plain["status"] == 0
    assert fin.sum() >= 30 and np.all(auto["status"] == 0)
    assert np.array_equal(auto["chi"][fin], plain["chi"][fin])
    C, nits = 9, 40
    a = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True)
    b = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True, speculate=1)
    np.testing.assert_allclose(a["chinew"], b["chinew"], rtol=2e-5)
    assert (a["accepted"] != b["accepted"]).sum() <= 1
    same = (a["accepted"] == b["accepted"]).all(axis=1)
    np.testing.assert_allclose(a["samples"][same][:, :, :P], b["samples"][same][:, :, :P], rtol=1e-12)
    assert np.array_equal(a["best_iteration"][same], b["best_iteration"][same])
    np.testing.assert_allclose(a["summaries"][same], b["summaries"][same], rtol=1e-9, atol=1e-12)
    seg = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True, segments=3)
    assert np.array_equal(seg["samples"], a["samples"]) and np.array_equal(seg["chain_state"], a["chain_state"])
    # posterior-ratio chains (prior log-densities in the kernel): both mappings make the same decisions
    prior = [("lognorm", 0.3, 0.0, float(c_)) for c_ in center]
    pa = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True, prior=prior)
    pb = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True, prior=prior, speculate=1)
    assert (pa["accepted"] != a["accepted"]).sum() > 0 and (pa["accepted"] != pb["accepted"]).sum() <= 1
    np.testing.assert_allclose(pa["chain_state"][:, 5][(pa["accepted"] == pb["accepted"]).all(axis=1)],
                               pb["chain_state"][:, 5][(pa["accepted"] == pb["accepted"]).all(axis=1)], rtol=1e-10)
    # prefetching width of the cooperative kernel (speculate = -K: K groups of lanes per chain): the same chain, bit for bit
    for K in (1, 2, 4):
        k = dm.mcmc(theta[:C], nits=nits, seed=4, trace=True, speculate=-K)
        for key in ("samples", "chinew", "accepted", "summaries", "chain_state", "theta", "step_count", "fail_count", "best_theta"):
            assert np.array_equal(a[key], k[key], equal_nan=True), (K, key)


def test_sliced_cooperative_rhs_agrees_with_the_full_one():
    """The per-lane right-hand side of the cooperative kernels (tracer.slice_plan; an option) against the
    default (every lane evaluates the whole traced RHS): same arithmetic per output, another component-to-lane layout,
    so the group-reduced error norms sum in another order -- chi to solver accuracy, decisions identical."""
    from odelib_b200.engine import DeviceModel
    rhs, names, sums, center, y0, orgs = nclass_problem(10)
    dm, tab = synthetic_problem(rhs, names, sums, center, y0, orgs, seed=10)
    ds = DeviceModel(rhs, 12, 5, dm.groups, sliced_rhs=True)
    ds.set_data(dm.tables, tab.y0)
    rng = np.random.default_rng(3)
    theta = center * np.exp(0.1 * rng.standard_normal((64, 5)))
    a = dm.sweep(theta, rtol=1e-10, atol=1e-10, max_steps=2000000)
    b = ds.sweep(theta, rtol=1e-10, atol=1e-10, max_steps=2000000)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    np.testing.assert_allclose(b["chi"], a["chi"], rtol=1e-7)
    ra = dm.mcmc(theta[:16], nits=40, seed=2, trace=True, rtol=1e-10, atol=1e-10)
    rb = ds.mcmc(theta[:16], nits=40, seed=2, trace=True, rtol=1e-10, atol=1e-10)
    assert np.array_equal(ra["accepted"], rb["accepted"])
    np.testing.assert_allclose(rb["chinew"], ra["chinew"], rtol=1e-6)
