"""The oracle restatement (oracle/odelib_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import odelib_oracle as orc
from tests.helpers import golden, oracle_rhs, oracle_tables

MODELS = ["zero_i", "one_i", "two_i"]


@pytest.mark.parametrize("name", MODELS)
def test_tables_match_reference(name):
    g = golden(name)
    tab = oracle_tables(name)
    assert np.array_equal(tab.times, g["times"])
    assert list(tab.obs_order) == list(g["obs_order"])
    for s in tab.obs_order:
        assert np.array_equal(tab.tindex[s], g["tindex_" + s])
    assert np.array_equal(np.concatenate([tab.ln_obs[s] for s in tab.obs_order]), g["ln_obs"])
    assert np.array_equal(np.concatenate([tab.log_sigma[s] for s in tab.obs_order]), g["log_sigma"])
    assert np.array_equal(tab.y0, g["y0"])
    assert orc.cutchi(tab, 6.0) == pytest.approx(float(g["cutchi6"]), rel=1e-14)
    assert orc.cutchi(tab, 6.0) == pytest.approx(tab.n_obs * 36 / 2, rel=1e-12)  # n_obs*sd^2/2


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("tag,tol", [("def", None), ("tight", 1e-13)])
def test_solve_unit_matches_reference(name, tag, tol):
    g = golden(name)
    tab, rhs = oracle_tables(name), oracle_rhs(name)
    for k, th in enumerate(g["theta"]):
        vec, chi, r2 = orc.solve_unit(rhs, th, tab, tol, tol, mxstep=200000 if tol else 0)
        # same scipy, same call: bitwise here; 1e-12 leaves room for a different libm/SIMD build
        np.testing.assert_allclose(vec, g["pred_" + tag][k], rtol=1e-12, atol=0, equal_nan=True)
        np.testing.assert_allclose(chi, g["chi_" + tag][k], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(r2, g["r2_" + tag][k], rtol=1e-12, equal_nan=True)


def test_notebook_known_answers():
    """Weak known-answer vectors printed in the demo notebook (Demo_InfectionStates.ipynb:2297-2307)."""
    tab, rhs = oracle_tables("zero_i"), oracle_rhs("zero_i")
    rows = [((1.480838e-08, 1.364223e-08, 19.386877), 108.070809),
            ((1.364139e-08, 1.352514e-08, 19.442711), 108.023903),
            ((4.594495e-06, 1.334745e-08, 19.110142), 109.682589)]
    for th, chi_printed in rows:
        _, chi, _ = orc.solve_unit(rhs, th, tab)
        assert chi == pytest.approx(chi_printed, rel=2e-7)   # 7 printed digits


@pytest.mark.parametrize("name", MODELS)
def test_streams_regenerated_without_running_reference(name):
    g = golden(name)
    for seed in (0, 1):
        z, u = orc.reference_streams(seed, g["theta"].shape[1], int(g[f"chain_def_s{seed}_nits"]) - 1)
        assert np.array_equal(z, g[f"chain_def_s{seed}_z"])
        assert np.array_equal(u, g[f"chain_def_s{seed}_u"])


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("tag,tol,seed", [("def", None, 0), ("def", None, 1), ("tight", 1e-13, 0)])
def test_mh_chain_matches_reference(name, tag, tol, seed):
    g = golden(name)
    tab, rhs = oracle_tables(name), oracle_rhs(name)
    pre = f"chain_{tag}_s{seed}_"
    nits = int(g[pre + "nits"])
    out = orc.mh_chain(rhs, g[pre + "theta0"], tab, int(g["pnum"]), nits=nits, seed=seed, rtol=tol, atol=tol)
    assert np.array_equal(out["accepted"], g[pre + "accepted"])
    np.testing.assert_allclose(out["proposals"], g[pre + "proposals"], rtol=1e-13)
    np.testing.assert_allclose(out["chinew"], g[pre + "chinew"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(out["kept"], g[pre + "kept"], rtol=1e-12)


def test_chi_masks_invalid_terms():
    """stats.py:41: non-finite terms vanish; everything invalid -> masked."""
    O = np.log(np.array([10.0, 20.0, 30.0]))
    S = np.array([0.1, 0.2, 0.0])
    with np.errstate(all="ignore"):
        C = np.log(np.array([11.0, -1.0, 30.0]))
        c = orc.chi(O, C, S)
    assert float(c) == pytest.approx((np.log(10 / 11)) ** 2 / (2 * 0.01))
    with np.errstate(all="ignore"):
        assert orc.chi(O, np.log(np.array([-1.0, -1.0, -1.0])), S) is np.ma.masked


def test_rhat_definition():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 500, 3))
    r = orc.rhat(x)
    assert np.all(np.abs(r - 1) < 0.02)
    x[0] += 3.0
    assert np.all(orc.rhat(x) > 1.2)


# ---- '<state>0' parameters (tests/golden/make_state0.py; Samplers.py:110-114, :139-143, Framework.py:730-731) ----
STATE0_MAP = {0: 3, 1: 4}          # state S <- parameter S0 (slot 3), V <- V0 (slot 4)


def state0_rhs(y, t, ps):
    return orc.zero_i(y, t, ps[:3])


def test_state0_values_are_ignored_outside_the_proposal_loop():
    """integrate / _Fit_worker start from istates whatever S0, V0 hold (Framework.py:647-650, :41-48)."""
    g = golden("state0")
    tab = oracle_tables("zero_i")
    assert np.array_equal(tab.y0, g["y0"])
    for k, th in enumerate(g["theta"]):
        vec, chi, r2 = orc.solve_unit(state0_rhs, th, tab)
        np.testing.assert_allclose(vec, g["pred_def"][k], rtol=1e-12)
        np.testing.assert_allclose(chi, g["chi_def"][k], rtol=1e-12)
        np.testing.assert_allclose(chi, g["fit_worker_chi"][k], rtol=1e-12)
        np.testing.assert_allclose(r2, g["r2_def"][k], rtol=1e-12)


@pytest.mark.parametrize("tag,walk", [("walk", [1, 1, 1, 1, 1]), ("staticV0", [1, 1, 1, 1, 0])])
def test_state0_chain_matches_reference(tag, walk):
    """A-priori solve from istates, every proposal from its own S0 / V0 (a static V0 included), restored on reject."""
    g = golden("state0")
    tab = oracle_tables("zero_i")
    pre = f"chain_{tag}_"
    nits = int(g[pre + "nits"])
    z, u = orc.reference_streams(3, sum(walk), nits - 1)
    assert np.array_equal(z, g[pre + "z"]) and np.array_equal(u, g[pre + "u"])
    out = orc.mh_chain(state0_rhs, g[pre + "theta0"], tab, int(g["pnum"]), nits=nits, walk=np.array(walk, bool), z=z, u=u,
                       y0_from_param=STATE0_MAP)
    assert np.array_equal(g[pre + "y0_apriori"], tab.y0)                       # the reference's own a-priori solve
    np.testing.assert_array_equal(g[pre + "y0_solves"], g[pre + "proposals"][:, 3:5])
    assert np.array_equal(out["accepted"], g[pre + "accepted"]) and out["accepted"].sum() > 5
    np.testing.assert_allclose(out["proposals"], g[pre + "proposals"], rtol=1e-13)
    np.testing.assert_allclose(out["chinew"], g[pre + "chinew"], rtol=1e-12)
    kept = g[pre + "kept"].copy()
    if tag == "staticV0":                                                      # quirk A13: static column = prior scale
        assert np.all(kept[:, 4] == 1.1e7)
        kept[:, 4] = out["kept"][:, 4]
    np.testing.assert_allclose(out["kept"], kept, rtol=1e-12)
    # without the a-priori exception the chain is a different chain: the restatement must not apply S0 there
    _, chi_wrong, _ = orc.solve_unit(state0_rhs, g[pre + "theta0"], tab, y0=g[pre + "theta0"][3:5])
    assert abs(chi_wrong - float(g[pre + "chi0"])) > 1e-3
