"""CPU oracle for the ODElib hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``odelib_b200``) never does; it fails loudly when its CUDA library is missing.

What it is
----------
A numpy/scipy restatement of the reference's algorithm for the path
"integrate -> pick predictions on the output grid -> chi / R^2 / AIC -> one
Metropolis-Hastings chain", written from the behaviour of /root/reference
(citations are ``file:line`` into that tree):

=====================  ====================================================
function here          reference it follows
=====================  ====================================================
``build_tables``       Framework.py:281-329 (_formatdf, _df_fitsetup),
                       :332-381 (_get_summation_index), :234 (times),
                       :246-249 (inits from the t==0 rows)
``integrate_grid``     Framework.py:656 (odeint call, scipy defaults)
``observe``            Framework.py:659-664 (summation), :677-682 (pick)
``chi``                Statistics/stats.py:22-41, Framework.py:685-697
``rsqrd`` / ``aic``    Statistics/stats.py:44-56, Framework.py:699-706
``solve_unit``         Framework.py:41-48 (_Fit_worker body)
``reference_streams``  Samplers.py:70,108,119-121,127 + Framework.py:103,119
``mh_chain``           Statistics/Samplers.py:53-174
``cutchi``             Framework.py:1004-1005
``rawstats``           Framework.py:11-17
=====================  ====================================================

Where the arithmetic really lives: the ODE solve is third-party --
``scipy.integrate.odeint`` (ODEPACK LSODA), a floor-only dependency of the
reference (requirements.txt:4 ``scipy >= 1.5.1``; nothing pinned).  This image
has scipy 1.18.1 / numpy 2.3.5 and the oracle calls that very function, exactly
as the reference does, instead of re-deriving LSODA.  Random numbers: numpy's
legacy ``RandomState`` (MT19937 + polar Box-Muller), numpy >= 1.19 floor.

Parity pin: the reference ships NO tests and no golden vectors,
so there is nothing of its own to pin against.  Instead ``tests/golden/`` holds
vectors produced by running the UNMODIFIED reference in the build container
(``tests/golden/make_golden.py``, three import shims, no source edits) and
``tests/test_oracle_golden.py`` checks this restatement against them; the three
printed posterior rows of the demo notebook (Demo_InfectionStates.ipynb
:2297-2307) are checked as weak known-answer vectors as well.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.integrate import odeint

SCIPY_ODEINT_TOL = 1.49012e-8  # scipy's default rtol = atol for odeint (LSODA)
RWALK_SD = 0.05                # Framework.py:107


# ----------------------------------------------------------------------------
# data tables
# ----------------------------------------------------------------------------
@dataclass
class Tables:
    """Everything the reference ctor derives from the DataFrame (Framework.py:168-263)."""
    state_names: tuple
    out_names: tuple               # names after summation (Framework.py:370-381)
    sum_index: dict                # first index -> tuple of indices summed into it
    keep: tuple                    # columns kept after summation
    times: np.ndarray              # output grid, np.linspace(0, tmax, t_steps)
    tindex: dict                   # organism -> int array of grid indices (nearest grid point)
    ln_obs: dict                   # organism -> ln(abundance)
    log_sigma: dict                # organism -> log_sigma
    y0: np.ndarray                 # initial state, state_names order
    n_obs: int = 0
    obs_order: tuple = field(default_factory=tuple)  # organisms in the order chi concatenates them

    def out_columns(self):
        return {name: i for i, name in enumerate(self.out_names)}


def _summation_maps(state_names, state_summations):
    """Framework.py:332-381: sums are stored in the lowest member index."""
    pos = {s: i for i, s in enumerate(state_names)}
    sum_index, renamed, used = {}, {}, set()
    for new_name, members in (state_summations or {}).items():
        idx = []
        for mname in members:
            if mname in used:
                raise ValueError(f"{mname} state varaiable cannot be used in two summations")
            if mname not in pos:
                raise ValueError(f"{mname} state varaiable is not a valid state name")
            used.add(mname)
            idx.append(pos[mname])
        idx.sort()
        sum_index[idx[0]] = tuple(idx)
        renamed[idx[0]] = new_name
    out_names, keep = [], []
    for i, s in enumerate(state_names):
        if i in renamed:
            out_names.append(renamed[i]); keep.append(i)
        elif s not in used:
            out_names.append(s); keep.append(i)
    if not state_summations:
        return {}, tuple(state_names), tuple(range(len(state_names)))
    return sum_index, tuple(out_names), tuple(keep)


def build_tables(df, state_names, state_summations=None, t_steps=1000, inits=None):
    """Restates _formatdf/_df_fitsetup/ctor init handling for the (organism,time,abundance,log_sigma) layout.

    ``df`` is a pandas DataFrame; rows are ordered by (organism, time) with a stable sort
    exactly like ``DataFrame.sort_values(by=['organism','time'])`` (Framework.py:286).
    """
    state_names = tuple(state_names)
    sum_index, out_names, keep = _summation_maps(state_names, state_summations)
    d = df.sort_values(by=["organism", "time"])
    organism = d["organism"].to_numpy()
    tt = d["time"].to_numpy(dtype=float)
    ab = d["abundance"].to_numpy(dtype=float)
    ln_ab = np.log(ab)                                                     # Framework.py:302
    sig = d["log_sigma"].to_numpy(dtype=float) if "log_sigma" in d else np.ones_like(ab)  # :303-304
    times = np.linspace(0, tt.max(), t_steps)                              # Framework.py:234
    tindex, ln_obs, log_sigma = {}, {}, {}
    for org in dict.fromkeys(organism):                                    # first-seen order
        rows = organism == org
        # nearest grid index, first minimum (Framework.py:316)
        tindex[org] = np.array([int(np.where(np.abs(a - times) == np.abs(a - times).min())[0][0])
                                for a in tt[rows]], dtype=np.int64)
        ln_obs[org] = ln_ab[rows]
        log_sigma[org] = sig[rows]
    # initial states: default 0, data rows with time == 0 (first seen wins), then explicit inits
    y0 = {s: 0 for s in state_names}                                       # Framework.py:216
    seen = set()
    for org, t, a in zip(organism, tt, d["abundance"].to_numpy()):
        if t == 0 and org not in seen:                                     # Framework.py:246-249
            seen.add(org)
            if org in y0:
                y0[org] = a
            # a summed name (e.g. 'H') is accepted and ignored (Framework.py:476-477, check disabled)
    for k, v in (inits or {}).items():
        if k in y0:
            y0[k] = v
    # chi concatenates in post-summation state order (Framework.py:679-681, :690)
    obs_order = tuple(n for n in out_names if n in tindex)
    return Tables(state_names, out_names, sum_index, keep, times, tindex, ln_obs, log_sigma,
                  np.array([y0[s] for s in state_names], dtype=float), int(len(d)), obs_order)


# ----------------------------------------------------------------------------
# integrate + score
# ----------------------------------------------------------------------------
def integrate_grid(rhs, y0, times, theta, rtol=None, atol=None, mxstep=0):
    """Framework.py:656 -- ``odeint(func, y0=initials, t=self.times, args=ps)``; ps is a 1-tuple of a list."""
    kw = {}
    if rtol is not None:
        kw["rtol"] = rtol
    if atol is not None:
        kw["atol"] = atol
    if mxstep:
        kw["mxstep"] = mxstep
    return odeint(rhs, y0=list(y0), t=times, args=(list(theta),), **kw)


def observe(mod, tab: Tables):
    """Summation into the lowest member column, column selection, pick at grid indices."""
    if tab.sum_index:
        mod = mod.copy()
        for first, members in tab.sum_index.items():
            mod[:, first] = mod[:, list(members)].sum(axis=1)              # Framework.py:662
        mod = mod[:, list(tab.keep)]                                       # :664
    return {name: mod[:, i][tab.tindex[name]] for i, name in enumerate(tab.out_names) if name in tab.tindex}


def chi(O, C, S):
    """stats.py:41 -- masked arithmetic drops every non-finite term; all-masked sums to np.ma.masked."""
    return (((np.ma.masked_invalid(O) - C) ** 2) / (2 * (S ** 2))).sum()


def chi_of(pred, tab: Tables):
    """Framework.py:685-697."""
    O = np.concatenate([tab.ln_obs[s] for s in pred])
    with np.errstate(all="ignore"):
        C = np.concatenate([np.log(pred[s]) for s in pred])
    S = np.concatenate([tab.log_sigma[s] for s in pred])
    with np.errstate(all="ignore"):
        return chi(O, C, S)


def rsqrd_of(pred, tab: Tables):
    """stats.py:49-56 via Framework.py:699-702: linear space, nansum, population variance."""
    ssres = sstot = 0.0
    for s in pred:
        obs = np.exp(tab.ln_obs[s])
        ssres += np.nansum((pred[s] - obs) ** 2)
        sstot += pred[s].shape[0] * np.var(obs)
    return 1 - ssres / sstot


def aic_of(chi_value, n_params):
    """stats.py:44-47."""
    return -2 * (-chi_value) + 2 * n_params


def solve_unit(rhs, theta, tab: Tables, rtol=None, atol=None, y0=None, mxstep=0):
    """One unit of work: theta -> predictions at the 37 observation rows -> chi, R^2.

    Returns (pred_vector[n_obs] in chi's concatenation order, chi, r2)."""
    mod = integrate_grid(rhs, tab.y0 if y0 is None else y0, tab.times, theta, rtol, atol, mxstep)
    pred = observe(mod, tab)
    c = chi_of(pred, tab)
    with np.errstate(all="ignore"):
        r2 = rsqrd_of(pred, tab)
    vec = np.concatenate([pred[s] for s in pred])
    return vec, (np.nan if c is np.ma.masked else float(c)), float(r2)


def cutchi(tab: Tables, sd_fitdistance):
    """Framework.py:1004-1005: chi of a prediction sd log-sigmas above every observation."""
    calc = {s: np.exp(tab.ln_obs[s] + sd_fitdistance * tab.log_sigma[s]) for s in tab.ln_obs}
    return float(chi_of(calc, tab))


def rawstats(x):
    """Framework.py:11-17 (pandas ``std`` => ddof=1)."""
    lx = np.log(np.asarray(x, dtype=float))
    log_mean = lx.mean()
    log_std = lx.std(ddof=1)
    return math.exp(log_mean), ((math.exp(log_std ** 2) - 1) * math.exp(2 * log_mean + log_std ** 2.0)) ** 0.5


# ----------------------------------------------------------------------------
# Metropolis-Hastings chain
# ----------------------------------------------------------------------------
def reference_streams(seed, n_walk, n_iter, n_prior_draws=None):
    """The random numbers one reference chain consumes, without running it.

    Per iteration (Samplers.py:104-127): one ``np.random.normal(0, 0.05)`` per walking
    parameter (Framework.py:119/:122), then one ``dist.rvs`` per walking parameter that has
    a prior (the never-used ``pdf()`` of Framework.py:103; for ``lognorm`` that is one
    ``standard_normal``), then one ``np.random.rand()``.  All from the global legacy
    RandomState seeded with the chain seed (Samplers.py:70).  Valid when every prior draw
    costs exactly one gaussian (true for scipy.stats.lognorm, the only prior the demo uses).
    """
    n_prior = n_walk if n_prior_draws is None else n_prior_draws
    rs = np.random.RandomState(seed)
    z = np.empty((n_iter, n_walk))
    u = np.empty(n_iter)
    for i in range(n_iter):
        for p in range(n_walk):
            z[i, p] = rs.normal(0, np.full((), RWALK_SD))
        for _ in range(n_prior):
            rs.standard_normal()
        u[i] = rs.rand()
    return z, u


def mh_chain(rhs, theta0, tab: Tables, n_params_total, nits=1000, burnin=None, walk=None,
             z=None, u=None, seed=0, rtol=None, atol=None, y0_from_param=None, log_prior=None):
    """Samplers.py:53-174 restated for scalar parameters.

    theta0: start values in parameter_names order.  walk: boolean mask of walking parameters
    (static_parameters are the False entries).  z[nits-1, n_walk], u[nits-1]: host streams
    (generated with ``reference_streams(seed, ...)`` when omitted).
    y0_from_param: optional {state_index: param_index} for the '<state>0' convention: the proposal
    loop copies such a parameter into the state's initial value (Samplers.py:110-114) and the reject
    branch restores it (:139-143).  The a-priori solve (Samplers.py:88) runs BEFORE any of that, from
    the model's istates (tab.y0), whatever the '<state>0' parameter holds.

    log_prior: None = the reference's chain (prior densities are evaluated but never enter the ratio,
    Samplers.py:118-127 -- SURVEY.md A1, A3).  A callable theta -> log prior density switches to the
    Metropolis-Hastings ratio of the posterior: acc = exp((chi - chinew) + (lp' - lp) + sum(ln theta' - ln theta)),
    the last term being the Hastings correction of the multiplicative log-normal walk (the `facs` the reference
    accumulates at Samplers.py:105-109 and drops).  Not reference behaviour: the product's opt-in `use_priors`.

    Returns dict with the per-iteration proposals/chinew/decisions and the kept rows
    (theta, chi, rsquared, aic, iteration, acceptance_ratio) exactly as the reference frame.
    """
    theta = np.array(theta0, dtype=float)
    P = theta.size
    walk = np.ones(P, bool) if walk is None else np.asarray(walk, bool)
    widx = np.flatnonzero(walk)
    n_iter = nits - 1                                                      # arange(1, nits)  :84
    if not burnin:
        burnin = int(nits / 2)                                             # :85-86
    if z is None or u is None:
        z, u = reference_streams(seed, widx.size, n_iter)
    y0 = tab.y0.copy()

    def apply_y0(th):
        if y0_from_param:
            for si, pi in y0_from_param.items():
                y0[si] = th[pi]

    _, chi_cur, r2_cur = solve_unit(rhs, theta, tab, rtol, atol, y0)        # :88-90: istates, not '<state>0'
    aic_cur = aic_of(chi_cur, n_params_total)
    lp_cur = log_prior(theta) if log_prior is not None else 0.0
    old = theta.copy()
    accepts = 0
    props = np.empty((n_iter, P)); chinews = np.empty(n_iter); decisions = np.zeros(n_iter, bool)
    rows = []
    for k in range(n_iter):
        it = k + 1
        with np.errstate(all="ignore"):
            for j, p in enumerate(widx):
                theta[p] = np.exp(np.log(theta[p]) + z[k, j])             # Framework.py:119
        apply_y0(theta)
        props[k] = theta
        _, chinew, r2new = solve_unit(rhs, theta, tab, rtol, atol, y0)     # :115-116
        chinews[k] = chinew
        with np.errstate(all="ignore"):
            if log_prior is None:
                acc = np.exp(np.log(np.exp(chi_cur - chinew)))             # :124-125
            else:
                lp_new = log_prior(theta)
                hast = float(np.sum(np.log(theta) - np.log(old)))
                acc = np.exp((chi_cur - chinew) + (lp_new - lp_cur) + hast)
        if acc > u[k]:                                                     # :127  (NaN -> reject)
            if log_prior is not None:
                lp_cur = lp_new
            chi_cur, r2_cur = chinew, r2new
            aic_cur = aic_of(chi_cur, n_params_total)
            old = theta.copy()
            accepts += 1
            decisions[k] = True
        else:
            theta = old.copy()                                             # :138
            apply_y0(theta)
        if it > burnin:                                                    # :147
            rows.append(np.concatenate([theta, [chi_cur, r2_cur, aic_cur, it, accepts / it]]))
    kept = np.array(rows) if rows else np.empty((0, P + 5))
    return {"proposals": props, "chinew": chinews, "accepted": decisions, "kept": kept,
            "z": z, "u": u}


# ----------------------------------------------------------------------------
# Gelman-Rubin R-hat (new functionality -- SURVEY.md §8e pins the definition)
# ----------------------------------------------------------------------------
def rhat(log_samples):
    """log_samples[m_chains, n_kept, P] -> R-hat[P] (classic, no splitting)."""
    x = np.asarray(log_samples, dtype=float)
    m, n, _ = x.shape
    means = x.mean(axis=1)
    W = x.var(axis=1, ddof=1).mean(axis=0)
    B = n * means.var(axis=0, ddof=1)
    var_plus = (n - 1) / n * W + B / n
    return np.sqrt(var_plus / W)


# ----------------------------------------------------------------------------
# the demo models (Demo_InfectionStates.ipynb:60-128) -- workload definitions
# ----------------------------------------------------------------------------
def zero_i(y, t, ps):
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    return np.array([mu * S - phi * S * V, beta * phi * S * V - phi * S * V])


def one_i(y, t, ps):
    mu, phi, beta, lam = ps[0], ps[1], ps[2], ps[3]
    S, I1, V = y[0], y[1], y[2]
    return np.array([mu * S - phi * S * V, phi * S * V - lam * I1, beta * lam * I1 - phi * S * V])


def two_i(y, t, ps):
    mu, phi, beta, lam, tau = ps[0], ps[1], ps[2], ps[3], ps[4]
    S, I1, I2, V = y[0], y[1], y[2], y[3]
    return np.array([mu * S - phi * S * V, phi * S * V - tau * I1, tau * I1 - lam * I2,
                     beta * lam * I2 - phi * S * V])
