"""Samplers with the reference's names (ODElib/Statistics/Samplers.py): sample_lhs and MetropolisHastings.

MetropolisHastings keeps the reference signature and return frame but runs the chain inside the CUDA
kernel (odl_mcmc); its random numbers are, by default, exactly the ones the reference would have drawn
from numpy's legacy global RandomState seeded with ``modelframework.random_seed`` (Samplers.py:70), so a
chain started from the same point makes the same decisions as the reference chain (up to integrator
tolerance at near-tie acceptances).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

RWALK_SD = 0.05   # parameter.rwalk default (Framework.py:107)


def lhs(n, samples, random_state=None):
    """Classic Latin-hypercube design in the unit cube: one uniform point per stratum and per dimension,
    strata shuffled independently per dimension.  Stands in for pyDOE2.lhs (Samplers.py:3,:33), which is
    not a dependency here; draws from numpy's global RandomState like pyDOE2 does when no state is given."""
    rs = np.random if random_state is None else random_state
    edges = np.linspace(0, 1, samples + 1)
    pts = rs.rand(samples, n) * (edges[1:] - edges[:-1])[:, None] + edges[:-1][:, None]
    design = np.empty_like(pts)
    for j in range(n):
        design[:, j] = pts[rs.permutation(samples), j]
    return design


def element_names(name, shape):
    """Column names of a parameter's elements in the flat layout: ``name`` for a scalar, ``name[i]`` / ``name[i,j]``
    (row-major) for an array-valued one."""
    if tuple(shape) == ():
        return [name]
    return ["{}[{}]".format(name, ",".join(str(k) for k in idx)) for idx in np.ndindex(*shape)]


def sample_lhs(parameter_dict, samples):
    """LHS in the unit cube pushed through each prior's ppf (Samplers.py:6-51).

    Array-valued parameters as the reference intends them (:26-32): one LHS dimension per NON-ZERO element, zeros are
    structural and stay zero.  The reference's own array branch cannot run (``parameter_dict[p][0]`` at :45 subscripts a
    parameter object); here the elements come back as one column each (``name[i,j]``) rather than as
    arrays stored in the rows, which is the layout every batched call works on."""
    names = list(parameter_dict)
    masks = {p: np.ravel(np.asarray(parameter_dict[p].val) != 0) if np.ndim(parameter_dict[p].val) else np.array([True])
             for p in names}
    cube = lhs(int(sum(m.sum() for m in masks.values())), samples)
    cols = {}
    j = 0
    for p in names:
        par = parameter_dict[p]
        for name, live in zip(element_names(p, np.shape(par.val)), masks[p]):
            if live:
                cols[name] = np.asarray(par.dist.ppf(cube[:, j], **par.hp), dtype=np.float64)
                j += 1
            else:
                cols[name] = np.zeros(samples)
    return pd.DataFrame(cols)


def _one_gaussian_rvs(par):
    """True when ``par.dist.rvs(**par.hp)`` consumes exactly one standard_normal of the legacy stream."""
    name = getattr(par.dist, "name", None)
    return name in ("lognorm", "norm")


def reference_streams(seed, walking, n_iter):
    """Proposal increments z[n_iter, n_walk] and uniforms u[n_iter] exactly as the reference chain with this
    seed consumes them (Samplers.py:70, :108, :119-121, :127; Framework.py:103, :119).

    walking: list of parameter objects in proposal order.  Per iteration the reference draws one
    N(0, 0.05) per walking parameter (per element of an array-valued one), then one ``dist.rvs`` per walking
    parameter that has a prior (the unused ``pdf()`` evaluation), then one uniform."""
    rs = np.random.RandomState(seed)
    dims = [int(np.size(p.val)) for p in walking]                 # rwalk draws one normal per ELEMENT (Framework.py:108,:119)
    nw = sum(dims)
    z = np.empty((n_iter, nw))
    u = np.empty(n_iter)
    with_prior = [p for p in walking if p.dist]
    if all(_one_gaussian_rvs(p) for p in with_prior):
        npd = len(with_prior)
        if (nw + npd) % 2 == 0:
            # legacy gauss caches the second value of each polar pair: with an even count per iteration the
            # cache is empty whenever the uniform is drawn, so blocks can be drawn vectorised
            for i in range(n_iter):
                g = rs.standard_normal(nw + npd)
                z[i] = 0.0 + RWALK_SD * g[:nw]
                u[i] = rs.random_sample()
            return z, u
        for i in range(n_iter):
            for j in range(nw):
                z[i, j] = rs.normal(0, RWALK_SD)
            for _ in range(npd):
                rs.standard_normal()
            u[i] = rs.random_sample()
        return z, u
    for i in range(n_iter):      # arbitrary priors: consume the stream through scipy itself
        for j in range(nw):
            z[i, j] = rs.normal(0, RWALK_SD)
        for p in with_prior:
            p.dist.rvs(random_state=rs, **p.hp)
        u[i] = rs.random_sample()
    return z, u


def device_streams_possible(seeds, walking):
    """True when the library's own generators reproduce the reference's stream consumption: every prior's ``rvs`` is one
    gaussian (lognorm, norm) and the seeds are 32-bit (numpy's init_genrand path)."""
    return all(_one_gaussian_rvs(p) for p in walking if p.dist) and all(0 <= int(s) < 2 ** 32 for s in seeds)


def reference_streams_batch(seeds, walking, n_iter):
    """reference_streams for many chains at once: z[C, n_iter, n_walk], u[C, n_iter].  Priors whose ``rvs`` is one
    gaussian (lognorm, norm) -- the demo's case -- are regenerated by the library's own MT19937 / polar-gauss
    restatement of numpy's legacy RandomState (odl_reference_streams, host code, ~10 ns per draw); anything else
    goes through numpy chain by chain."""
    seeds = [int(s) for s in seeds]
    nw = sum(int(np.size(p.val)) for p in walking)                # one proposal increment per element
    with_prior = [p for p in walking if p.dist]
    if all(_one_gaussian_rvs(p) for p in with_prior) and all(0 <= s < 2 ** 32 for s in seeds):
        from .. import _capi
        z = np.empty((len(seeds), n_iter, nw))
        u = np.empty((len(seeds), n_iter))
        sd = np.asarray(seeds, dtype=np.uint32)
        _capi.check(_capi.lib().odl_reference_streams(sd.ctypes.data, len(seeds), int(n_iter), nw, len(with_prior),
                                                      RWALK_SD, z.ctypes.data, u.ctypes.data))
        return z, u
    z = np.empty((len(seeds), n_iter, nw))
    u = np.empty((len(seeds), n_iter))
    for c, seed in enumerate(seeds):
        z[c], u[c] = reference_streams(seed, walking, n_iter)
    return z, u


def MetropolisHastings(modelframework, nits=1000, burnin=None, static_parameters=set(), print_progress=True,
                       rng="reference", rtol=None, atol=None):
    """One Metropolis-Hastings chain (Samplers.py:53-174) on the GPU.

    Returns the reference's frame: parameter columns, chi, rsquared, aic, iteration, acceptance_ratio
    (rows for iterations > burnin).  The model's parameters are left at the chain's last point, as the
    reference leaves them."""
    frames = modelframework._run_chains([modelframework._current_theta()], [modelframework.random_seed], nits,
                                        burnin, static_parameters, rng=rng, rtol=rtol, atol=atol,
                                        update_model=True)
    df = frames[0]
    if print_progress:
        print("chain finished: {} kept rows, acceptance ratio {:.3f}".format(
            len(df), df["acceptance_ratio"].iloc[-1] if len(df) else float("nan")))
    return df
