"""GPU parity of the device-resident Metropolis-Hastings kernel (odl_mcmc) with the reference chain
(golden vectors recorded from the unmodified Samplers.MetropolisHastings) and with the oracle."""
import numpy as np
import pytest

from odelib_b200 import engine
from oracle import odelib_oracle as orc
from tests.helpers import device_model, golden, oracle_rhs

pytestmark = pytest.mark.gpu
MODELS = ["zero_i", "one_i", "two_i"]


def _near_tie(chi_cur, chinew, u, delta):
    with np.errstate(all="ignore"):
        return np.abs((chi_cur - chinew) - np.log(u)) < delta


def _replay(chi0, chinew, u):
    """Decisions and running chi implied by a chinew sequence (Samplers.py:124-127)."""
    acc = np.zeros(len(u), bool); cur = np.empty(len(u)); c = chi0
    for k in range(len(u)):
        cur[k] = c
        with np.errstate(all="ignore"):
            if np.exp(c - chinew[k]) > u[k]:
                acc[k] = True; c = chinew[k]
    return acc, cur


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("tag,tol,delta,rtol_chi", [("def", None, 1e-4, 1e-6), ("tight", 1e-13, 1e-8, 1e-9)])
def test_teacher_forced_decisions_match_reference(name, tag, tol, delta, rtol_chi):
    """Reference proposals fed to the GPU: chinew within tolerance, decisions identical except near-ties
    |(chi - chinew) - ln u| < delta.  The recorded chains hold NO near-tie (smallest decision margin 7.7e-4 at the
    default tolerance, where chinew differs by <= 2.6e-7 relative / 4.5e-5 absolute from the reference's -- measured,
    tools/tolerance_probe.py), so every decision and every kept sample must be the reference's."""
    g = golden(name)
    pre = f"chain_{tag}_s0_"
    nits = int(g[pre + "nits"])
    dm, _ = device_model(name)
    out = dm.mcmc(g[pre + "theta0"][None, :], nits=nits, rng_mode="forced", forced=g[pre + "proposals"][None],
                  u=g[pre + "u"][None], rtol=tol, atol=tol, trace=True, pnum=int(g["pnum"]), max_steps=2000000)
    ref_acc, ref_cur = _replay(float(g[pre + "chi0"]), g[pre + "chinew"], g[pre + "u"])
    assert np.array_equal(ref_acc, g[pre + "accepted"])
    fin = np.isfinite(g[pre + "chinew"])
    np.testing.assert_allclose(out["chinew"][0][fin], g[pre + "chinew"][fin], rtol=rtol_chi, atol=1e-7)
    tie = _near_tie(ref_cur, g[pre + "chinew"], g[pre + "u"], delta)
    assert tie.sum() == 0
    first_tie = np.flatnonzero(tie)[0] if tie.any() else len(tie)
    assert np.array_equal(out["accepted"][0][:first_tie].astype(bool), ref_acc[:first_tie])
    if not tie.any():
        # no near-tie: the whole chain (kept rows: theta, iteration, acceptance ratio) is the reference's
        kept = g[pre + "kept"]; P = dm.n_param
        np.testing.assert_array_equal(out["samples"][0][:, :P], kept[:, :P])          # bit-identical samples
        np.testing.assert_array_equal(out["samples"][0][:, P + 3:], kept[:, P + 3:])  # iteration, acceptance
        np.testing.assert_allclose(out["samples"][0][:, P:P + 3], kept[:, P:P + 3], rtol=rtol_chi, atol=1e-6)


@pytest.mark.parametrize("name", MODELS)
def test_free_running_host_streams_match_reference(name):
    """Host z/u streams (the reference's own random numbers): the device forms exp(log(theta)+z) itself."""
    g = golden(name)
    pre = "chain_tight_s0_"
    nits = int(g[pre + "nits"])
    dm, _ = device_model(name)
    z, u = orc.reference_streams(0, dm.n_param, nits - 1)
    assert np.array_equal(z, g[pre + "z"])
    out = dm.mcmc(g[pre + "theta0"][None, :], nits=nits, rng_mode="host", z=z[None], u=u[None], rtol=1e-13,
                  atol=1e-13, trace=True, pnum=int(g["pnum"]), max_steps=2000000)
    ref_acc, ref_cur = _replay(float(g[pre + "chi0"]), g[pre + "chinew"], g[pre + "u"])
    tie = _near_tie(ref_cur, g[pre + "chinew"], g[pre + "u"], 1e-7)
    first_tie = np.flatnonzero(tie)[0] if tie.any() else len(tie)
    assert first_tie > 50
    assert np.array_equal(out["accepted"][0][:first_tie].astype(bool), ref_acc[:first_tie])
    if not tie.any():
        kept = g[pre + "kept"]; P = dm.n_param
        # device exp/log differ from numpy's by <= 1 ulp per step: a few ulp after a random walk
        np.testing.assert_allclose(out["samples"][0][:, :P], kept[:, :P], rtol=1e-12)
        np.testing.assert_allclose(out["samples"][0][:, P], kept[:, P], rtol=1e-8)
        np.testing.assert_array_equal(out["samples"][0][:, P + 3:], kept[:, P + 3:])


def test_philox_chains_match_oracle_with_recomputed_streams():
    """Device Philox streams recomputed on the host and fed to the oracle chain: same decisions."""
    name = "zero_i"
    g = golden(name)
    dm, tab = device_model(name)
    nits, C, seed = 120, 6, 1234
    theta0 = np.tile(g["chain_def_s0_theta0"], (C, 1)) * np.exp(0.01 * np.arange(C))[:, None]
    out = dm.mcmc(theta0, nits=nits, rng_mode="philox", seed=seed, chain_offset=10, rtol=1e-12, atol=1e-12, trace=True)
    z, u = engine.philox_streams(seed, 10 + np.arange(C), nits - 1, dm.n_param)
    rhs = oracle_rhs(name)
    for c in range(C):
        ref = orc.mh_chain(rhs, theta0[c], tab, dm.n_param, nits=nits, z=z[c], u=u[c], rtol=1e-13, atol=1e-13)
        _, cur = _replay(np.nan, ref["chinew"], u[c])
        assert np.array_equal(out["accepted"][c].astype(bool), ref["accepted"])
        np.testing.assert_allclose(out["chinew"][c], ref["chinew"], rtol=1e-8)
        np.testing.assert_allclose(out["samples"][c][:, :3], ref["kept"][:, :3], rtol=1e-11)
        np.testing.assert_array_equal(out["samples"][c][:, -2], ref["kept"][:, -2])
        np.testing.assert_allclose(out["samples"][c][:, -1], ref["kept"][:, -1], rtol=1e-15)


def test_static_parameters_do_not_walk_and_segments_continue():
    name = "two_i"
    g = golden(name)
    dm, _ = device_model(name)
    theta0 = np.tile(g["chain_def_s0_theta0"], (64, 1))
    a = dm.mcmc(theta0, nits=200, walk=[0, 1, 2, 4], seed=5, trace=True)
    assert np.all(a["samples"][:, :, 3] == theta0[0, 3])            # lam static
    assert np.all(a["theta"][:, 3] == theta0[0, 3])
    # the same chains in 4 launches give bit-identical results (state persists in the buffers)
    b = dm.mcmc(theta0, nits=200, walk=[0, 1, 2, 4], seed=5, trace=True, segments=4)
    assert np.array_equal(a["samples"], b["samples"])
    assert np.array_equal(a["accepted"], b["accepted"])
    assert np.array_equal(a["summaries"], b["summaries"])
    # chains differ from one another (per-chain streams) and accept at a sane rate
    assert len({tuple(r) for r in a["theta"]}) > 32
    rate = a["accepted"].mean()
    assert 0.02 < rate < 0.9


def test_summaries_are_welford_of_kept_log_samples_and_rhat():
    name = "zero_i"
    g = golden(name)
    dm, _ = device_model(name)
    C = 32
    theta0 = np.tile(g["chain_def_s0_theta0"], (C, 1))
    out = dm.mcmc(theta0, nits=400, seed=3)
    P = dm.n_param
    logs = np.log(out["samples"][:, :, :P])
    np.testing.assert_array_equal(out["summaries"][:, 0], out["n_keep"])
    np.testing.assert_allclose(out["summaries"][:, 1:1 + P], logs.mean(axis=1), rtol=1e-12)
    m2 = ((logs - logs.mean(axis=1, keepdims=True)) ** 2).sum(axis=1)
    np.testing.assert_allclose(out["summaries"][:, 1 + P:], m2, rtol=1e-9, atol=1e-18)
    from odelib_b200.rhat import rhat_from_summaries
    np.testing.assert_allclose(rhat_from_summaries(out["summaries"], P), orc.rhat(logs), rtol=1e-10)


def test_edge_semantics_nan_chi_rejects_and_device_buffers():
    import torch
    name = "zero_i"
    g = golden(name)
    dm, _ = device_model(name)
    th = np.tile(g["chain_def_s0_theta0"], (8, 1))
    th[3, 0] = np.nan                           # a chain whose a-priori solve fails never accepts
    out = dm.mcmc(th, nits=60, seed=1, trace=True)
    assert out["accepted"][3].sum() == 0 and np.isnan(out["chain_state"][3, 0])
    assert out["fail_count"][3] > 0
    dev = dm.mcmc(torch.from_numpy(th).cuda(), nits=60, seed=1, trace=True, device_buffers=True)
    torch.cuda.synchronize()
    assert np.array_equal(dev["samples"].cpu().numpy(), out["samples"], equal_nan=True)


@pytest.mark.parametrize("rng_mode", ["philox", "host"])
def test_prefetching_width_does_not_change_the_chain(rng_mode):
    """odl_mcmc_opts.speculate = K lanes per chain evaluate K iterations at once along the all-rejected path:
    every K gives the same chain bit for bit -- samples, per-iteration trace, summaries, final state, and the
    count of consumed integrator steps (discarded speculative solves are not counted)."""
    dm, _ = device_model("two_i")
    rng = np.random.default_rng(5)
    center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
    C, nits = 37, 120                                           # ragged: 37 chains do not fill the last group/warp
    starts = center * np.exp(0.05 * rng.standard_normal((C, 5)))
    kw = dict(nits=nits, rng_mode=rng_mode, trace=True, seed=3)
    if rng_mode == "host":
        kw["z"] = 0.05 * rng.standard_normal((C, nits - 1, 5))
        kw["u"] = rng.random((C, nits - 1))
    base = dm.mcmc(starts, speculate=1, **kw)
    assert 0.05 < base["accepted"].mean() < 0.7
    for K in (2, 4, 8, 16, 32, 0):
        out = dm.mcmc(starts, speculate=K, **kw)
        for key in ("samples", "chinew", "accepted", "summaries", "chain_state", "theta", "step_count", "fail_count"):
            assert np.array_equal(base[key], out[key], equal_nan=True), (K, key)
    # segmented chains keep working with groups: two launches of 60 + 59 iterations
    seg = dm.mcmc(starts, speculate=8, segments=2, **kw)
    assert np.array_equal(base["samples"], seg["samples"]) and np.array_equal(base["summaries"], seg["summaries"])
    # static parameters (walk list) and burn-in bookkeeping
    a = dm.mcmc(starts, speculate=1, walk=[0, 2, 4], burnin=10, **kw)
    b = dm.mcmc(starts, speculate=16, walk=[0, 2, 4], burnin=10, **kw)
    assert np.array_equal(a["samples"], b["samples"]) and a["samples"].shape[1] == nits - 1 - 10
    assert np.all(a["samples"][:, :, 1] == starts[:, None, 1])          # a static parameter never moves


def test_best_kept_row_per_chain_matches_idxmin_of_the_samples():
    """chain_state[:, 3:5] / best_theta: the first minimum of chi over each chain's kept rows (what the reference's
    set_best_params finds with idxmin over the posterior frame, Framework.py:725-731), tracked in the kernel."""
    dm, _ = device_model("two_i")
    rng = np.random.default_rng(9)
    center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
    starts = center * np.exp(0.1 * rng.standard_normal((53, 5)))
    for spec, segments in ((1, 1), (8, 1), (4, 3)):
        out = dm.mcmc(starts, nits=160, seed=2, speculate=spec, segments=segments)
        chi = out["samples"][:, :, 5]
        rows = np.argmin(chi, axis=1)                            # first occurrence of the minimum
        c = np.arange(len(starts))
        assert np.array_equal(out["best_chi"], chi[c, rows])
        assert np.array_equal(out["best_iteration"], out["samples"][c, rows, 5 + 3])
        assert np.array_equal(out["best_theta"], out["samples"][c, rows, :5])


def test_posterior_ratio_with_device_prior_log_densities_matches_the_oracle():
    """Opt-in `prior=` (north_star: the kernel evaluates prior log-densities): acceptance on the posterior ratio with the
    Hastings term of the multiplicative walk, against the oracle's restatement with scipy's logpdf on the same streams;
    without `prior=` the chain is the reference's (priors never enter), unchanged."""
    import scipy.stats
    dm, tab = device_model("two_i")
    rng = np.random.default_rng(17)
    center = np.array([7.475e-09, 1.069e-07, 19.73, 1.934, 2.799])
    table = [("lognorm", 3.0, 0.0, 1e-8), ("lognorm", 3.0, 0.0, 1e-8), ("norm", 0.0, 20.0, 0.5), ("uniform", 0.0, 0.5, 3.0),
             ("lognorm", 0.02, 0.0, 2.8)]                        # tight priors on beta / tau: they decide acceptances
    dists = [scipy.stats.lognorm(3.0, 0.0, 1e-8), scipy.stats.lognorm(3.0, 0.0, 1e-8), scipy.stats.norm(20.0, 0.5),
             scipy.stats.uniform(0.5, 3.0), scipy.stats.lognorm(0.02, 0.0, 2.8)]
    log_prior = lambda th: float(sum(d.logpdf(x) for d, x in zip(dists, th)))
    C, nits = 6, 80
    starts = center * np.exp(0.02 * rng.standard_normal((C, 5)))
    z = 0.05 * rng.standard_normal((C, nits - 1, 5))
    u = rng.random((C, nits - 1))
    kw = dict(nits=nits, rng_mode="host", z=z, u=u, rtol=1e-11, atol=1e-11, trace=True, max_steps=2000000)
    plain = dm.mcmc(starts, **kw)
    with_prior = dm.mcmc(starts, prior=table, **kw)
    assert (plain["accepted"] != with_prior["accepted"]).sum() > 5            # the priors change decisions
    rhs = oracle_rhs("two_i")
    for c in range(3):
        ref = orc.mh_chain(rhs, starts[c], tab, 5, nits=nits, z=z[c], u=u[c], rtol=1e-12, atol=1e-12, log_prior=log_prior)
        assert np.array_equal(with_prior["accepted"][c].astype(bool), ref["accepted"])
        np.testing.assert_allclose(with_prior["samples"][c][:, :5], ref["kept"][:, :5], rtol=1e-12)
        ref0 = orc.mh_chain(rhs, starts[c], tab, 5, nits=nits, z=z[c], u=u[c], rtol=1e-12, atol=1e-12)
        assert np.array_equal(plain["accepted"][c].astype(bool), ref0["accepted"])
    # the current point's log prior rides in chain_state[5]; prefetching width does not change the chain
    lp_end = [log_prior(th) for th in with_prior["theta"]]
    np.testing.assert_allclose(with_prior["chain_state"][:, 5], lp_end, rtol=1e-12)
    for K in (1, 8):
        k = dm.mcmc(starts, prior=table, speculate=K, **kw)
        assert np.array_equal(k["samples"], with_prior["samples"]) and np.array_equal(k["accepted"], with_prior["accepted"])
    out = dm.mcmc(starts, prior=[("uniform", 0.0, 1.0, 1.0)] * 5, **kw)       # every proposal outside the support
    assert out["accepted"].sum() == 0


@pytest.mark.parametrize("C", [5, 64, 70])
def test_sample_layouts_hold_the_same_rows(C):
    """Kept rows in chain-major memory (the reference frame's order) and in iteration-major memory (a warp's rows of one
    iteration contiguous: coalesced full-sector stores when every lane of the warp keeps a row, per-lane rows otherwise
    -- ragged last warp, speculation) are the same rows; `samples` is indexed [chain, row, column] either way."""
    import torch
    name = "two_i"
    g = golden(name)
    dm, _ = device_model(name)
    theta0 = np.tile(g["chain_def_s0_theta0"], (C, 1)) * np.exp(0.02 * np.random.default_rng(C).standard_normal((C, 5)))
    a = dm.mcmc(theta0, nits=61, seed=4, sample_layout="chain", speculate=1)
    b = dm.mcmc(theta0, nits=61, seed=4, sample_layout="iteration", speculate=1)
    c = dm.mcmc(theta0, nits=61, seed=4, sample_layout="iteration", speculate=4)
    d = dm.mcmc(torch.from_numpy(theta0).cuda(), nits=61, seed=4, speculate=1, device_buffers=True)
    assert a["samples"].shape == b["samples"].shape == (C, 30, 10) and a["samples"].flags.c_contiguous
    assert b["samples"].transpose(1, 0, 2).flags.c_contiguous       # the buffer itself is [row][chain][column]
    assert np.array_equal(a["samples"], b["samples"]) and np.array_equal(a["samples"], c["samples"])
    assert np.array_equal(a["samples"], d["samples"].cpu().numpy())
    assert np.all(a["samples"][:, :, 8] == np.arange(31, 61)[None, :])   # iteration column: rows in order, none missing
    for k in ("summaries", "chain_state", "theta"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k])


def test_rhat_through_the_abi_equals_numpy():
    """odl_rhat (device reduction; with several ranks behind an ncclAllGather): R-hat and the pooled log-moments of all
    kept rows against the numpy formulas on the same summaries; rows with count 0 (padding of ragged shards) are skipped."""
    import torch
    from odelib_b200.rhat import pooled_log_stats, rhat_from_summaries
    name = "two_i"
    g = golden(name)
    dm, _ = device_model(name)
    C = 300
    theta0 = np.tile(g["chain_def_s0_theta0"], (C, 1)) * np.exp(0.05 * np.random.default_rng(5).standard_normal((C, 5)))
    out = dm.mcmc(theta0, nits=120, seed=8, keep_samples=False)
    assert dm.comm_init() == 1                                      # no process group: world of one, no NCCL
    for summ in (out["summaries"], torch.from_numpy(out["summaries"]).cuda()):
        rh, (N, mean, std), total = dm.rhat(summ)
        assert total == C and N == C * out["n_keep"]
        np.testing.assert_allclose(rh, rhat_from_summaries(out["summaries"], 5), rtol=1e-12)
        N2, mean2, std2 = pooled_log_stats(out["summaries"], 5)
        np.testing.assert_allclose(mean, mean2, rtol=1e-13)
        np.testing.assert_allclose(std, std2, rtol=1e-11)
    # ODL_RHAT_LOCAL (this rank's chains only, whatever communicator the handle has joined): the same numbers on one GPU
    rh_l, (N_l, mean_l, _), total_l = dm.rhat(out["summaries"], local=True)
    assert total_l == C and N_l == N and np.array_equal(rh_l, rh) and np.array_equal(mean_l, mean)
    padded = np.vstack([out["summaries"][:100], np.zeros((7, 11)), out["summaries"][100:]])
    rh2, _, total2 = dm.rhat(padded)
    assert total2 == C + 7
    np.testing.assert_allclose(rh2, rhat_from_summaries(out["summaries"], 5), rtol=1e-12)


def test_reference_streams_generated_on_the_device():
    """odl_reference_streams_device (MT19937 + polar gauss per chain on the GPU, for runs too large for the host
    generator) against the host generator, which is bitwise numpy's: uniforms bit-identical; gaussians through CUDA's log(),
    within 1 ulp of glibc's -- most values identical, the others a few ulp apart; a chain run on them makes the reference's
    decisions."""
    from odelib_b200 import _capi
    dm, _ = device_model("two_i")
    seeds = np.array([0, 1, 7, 299, 12345, 2 ** 32 - 1, 42, 43], dtype=np.uint32)
    for n_walk, n_prior in ((5, 5), (4, 3), (1, 0)):             # even and odd draws per iteration (the gauss cache)
        n_iter = 700                                            # > 624 words per chain: several twists
        z = np.empty((len(seeds), n_iter, n_walk)); u = np.empty((len(seeds), n_iter))
        _capi.check(_capi.lib().odl_reference_streams(seeds.ctypes.data, len(seeds), n_iter, n_walk, n_prior, 0.05,
                                                      z.ctypes.data, u.ctypes.data))
        zd, ud = dm.reference_streams(seeds, n_iter, n_walk, n_prior, 0.05)
        zd, ud = zd.cpu().numpy(), ud.cpu().numpy()
        assert np.array_equal(ud, u)
        assert (zd == z).mean() > 0.9
        rel = np.abs(zd - z) / np.abs(z)
        print("device streams: identical", (zd == z).mean(), "max rel diff", rel.max())
        np.testing.assert_allclose(zd, z, rtol=2e-15, atol=0)    # a few ulp: 1 ulp of log() through a division and a root
    rs = np.random.RandomState(7)                               # and numpy itself, for one chain
    g = rs.standard_normal(10); u0 = rs.random_sample()
    z7, u7 = dm.reference_streams(np.array([7], np.uint32), 3, 5, 5, 0.05)
    np.testing.assert_allclose(z7.cpu().numpy()[0, 0], 0.05 * g[:5], rtol=2e-15)
    assert u7.cpu().numpy()[0, 0] == u0
    # the golden chain of the reference (seed 0) on device-generated streams: the reference's decisions
    g2 = golden("two_i")
    pre = "chain_tight_s0_"
    nits = int(g2[pre + "nits"])
    zr, ur = dm.reference_streams(np.array([0], np.uint32), nits - 1, dm.n_param, dm.n_param, 0.05)
    import torch
    out = dm.mcmc(torch.from_numpy(g2[pre + "theta0"][None, :]).cuda(), nits=nits, rng_mode="host", z=zr, u=ur, rtol=1e-13,
                  atol=1e-13, trace=True, pnum=int(g2["pnum"]), max_steps=2000000, device_buffers=True)
    assert np.array_equal(out["accepted"][0].cpu().numpy().astype(bool), g2[pre + "accepted"])
    np.testing.assert_allclose(out["samples"][0].cpu().numpy()[:, :dm.n_param], g2[pre + "kept"][:, :dm.n_param], rtol=1e-12)
