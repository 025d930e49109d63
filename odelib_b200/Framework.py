"""ModelFramework / parameter: ODElib's Python surface (ODElib/Framework.py) over the B200 hot path.

Same constructor, ``integrate`` / ``get_chi`` / ``get_residuals`` / ``fit_survey`` / ``MCMC`` signatures and
return frames as the reference, so a notebook written against ODElib keeps working; what changed is what
happens underneath:

* the ODE callable is traced once and compiled by NVRTC (tracer.py, engine.DeviceModel) instead of being
  called back from scipy's LSODA (Framework.py:656);
* ``fit_survey`` / the ``_Fit_worker`` loop (Framework.py:41-48, :800-816) is one launch of the batched
  DOPRI5 kernel with chi / R^2 fused;
* ``MCMC`` / ``Samplers.MetropolisHastings`` (Framework.py:946-1061, Samplers.py:53-174) run all chains in
  one device-resident kernel; ``cpu_cores`` is accepted and ignored;
* there is no CPU fallback: without the CUDA library / a B200 every compute call raises.

Documented deviations from the reference: ``get_residuals`` returns the aligned
observation-row residuals (A19, the reference's pandas alignment is a cartesian product); plain numeric
parameter kwargs in the ctor are accepted (A15, broken in the reference); no debug prints (A4).
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd

from .engine import SCIPY_TOL, DeviceModel, ObsTables
from .rhat import (allgather_rows, broadcast_rows, ess_from_summaries, pooled_log_stats, rhat_from_summaries,
                   shard_bounds)
from .Statistics import Samplers, stats


def _rawstats_from_logmoments(log_mean, log_std):
    return np.exp(log_mean), ((np.exp(log_std ** 2) - 1) * np.exp(2 * log_mean + log_std ** 2.0)) ** 0.5


def rawstats(pdseries):
    """Log-space median and log-normal standard deviation of a posterior column (Framework.py:11-17)."""
    lx = np.log(pdseries)
    return _rawstats_from_logmoments(lx.mean(), lx.std())


class PosteriorSummary:
    """What ``MCMC(..., posterior="summary")`` returns instead of the posterior frame: everything the reference's
    fitting report and ``set_best_params`` take from that frame (Framework.py:1047-1060, :725-731), reduced on
    the device -- per-parameter median / standard deviation (rawstats), the best kept row, R-hat -- without
    materialising chains x rows x columns on the host (65,536 chains x 500 rows are 2.9 GB as a frame)."""

    def __init__(self, pnames, n_chains, n_rows, stats, best, best_chi, rhat, acceptance_ratio, ess=None):
        self.parameter_names, self.n_chains, self.n_rows = list(pnames), int(n_chains), int(n_rows)
        self.stats, self.best, self.best_chi, self.rhat = stats, best, float(best_chi), rhat
        self.ess = ess                                            # effective sample size per parameter (rhat.py)
        self.acceptance_ratio = float(acceptance_ratio)

    def __repr__(self):
        rows = ["PosteriorSummary: {} chains, {} kept rows".format(self.n_chains, self.n_rows)]
        for p, (med, sd) in self.stats.items():
            rows.append("  {}: median = {:0.3e}, sd = {:0.3e}, best = {:0.3e}".format(p, med, sd, self.best[p]))
        rows.append("  best chi = {:0.3e}".format(self.best_chi))
        return "\n".join(rows)


class parameter:
    """A parameter value with an optional scipy.stats prior and its hyper-parameters (Framework.py:50-163)."""

    def __init__(self, stats_gen=None, hyperparameters=None, init_value=None, name=None):
        self.dist = stats_gen
        self.hp = hyperparameters
        self.name = name
        if init_value is not None and np.any(init_value):         # 0 counts as "not given", as in the reference
            self.val = np.array(init_value)
        else:
            if not self.dist:
                raise ValueError("You must specify a scipy distribution if not passing a value")
            self.val = np.array(self.dist.rvs(**self.hp))
        self._dim = self.val.shape

    def pdf(self, val=None):
        """Prior density at ``val``; without an argument, at a fresh prior draw (Framework.py:97-105)."""
        if not self.dist:
            return 1.0
        if val is not None and np.any(val):                      # `if val:` in the reference (raises for arrays, :99)
            return self.dist.pdf(val, **self.hp)
        return self.dist.pdf(self.dist.rvs(**self.hp), **self.hp)

    def rwalk(self, std=.05):
        """Multiplicative log-normal random-walk step (Framework.py:107-122)."""
        self.val = np.exp(np.log(self.val) + np.random.normal(0, np.full(np.shape(self.val), std)))

    def has_distribution(self):
        return bool(self.dist)

    def copy(self):
        return parameter(init_value=self.val, stats_gen=self.dist, hyperparameters=self.hp, name=self.name)

    def __repr__(self):
        out = [str(self.val) + '  ']
        if self.dist:
            out.append("(distribution:{}, ".format(self.dist.name))
            out.append("hyperparameters:{})".format(str(self.hp)))
        return ' '.join(out)

    __str__ = __repr__


class ModelFramework:
    """Drop-in for ODElib.ModelFramework (Framework.py:166-1165) for the integrate / score / sweep / MCMC path."""

    def __init__(self, ODE, parameter_names, state_names, dataframe=None, state_summations=None,
                 t_end=5, t_steps=1000, random_seed=0, **kwargs):
        self._pnames = tuple(parameter_names)
        self._snames = tuple(state_names)
        self._model = ODE
        self.parameters = {p: None for p in self._pnames}
        self.istates = {s: 0 for s in self._snames}
        self.random_seed = random_seed
        # device-side options (not in the reference): tolerances default to scipy odeint's
        self.rtol = kwargs.pop("rtol", None)
        self.atol = kwargs.pop("atol", None)
        self.device = kwargs.pop("device", None)
        # stepper of the chains: "auto" probes the chain starts with the capped DOPRI5 pass and, when a quarter of them
        # do not finish (a stiff posterior region -- where LSODA itself switches to BDF), runs the chains on the
        # variable-order BDF kernel; "dopri5" / "bdf" / "radau5" / "ros23" force one
        self.solver = kwargs.pop("solver", "auto")
        # distributed=True: under torchrun (one process per GPU, torch.distributed initialised) fit_survey and MCMC shard
        # their rows / chains over the ranks in contiguous blocks and every rank returns the complete result -- what
        # `cpu_cores` is to the reference (Framework.py:755-785).  Every rank must then make the same calls.
        self.distributed = bool(kwargs.pop("distributed", False))
        # where the library keeps the compiled cubins of this model ("default": odelib_b200/_cubin_cache; None: no cache)
        self._cache_dir = kwargs.pop("cache_dir", "default")
        self._dm = None
        if state_summations:
            (self._summations_index, self._summation_snames, self._sumkeep,
             self._suminds) = self._get_summation_index(state_summations)
        else:
            self._summations_index, self._summation_snames, self._sumkeep, self._suminds = {}, tuple(), tuple(), tuple()
        self._obs_logabundance, self._obs_logsigma, self._obs_abundance = {}, {}, {}
        self._pred_tindex = {}
        if isinstance(dataframe, pd.DataFrame):
            self.df = self._formatdf(dataframe.copy())
            self.times = np.linspace(0, max(self.df['time']), t_steps)
            self._samples = len(self.df)
            self._pred_tindex, self._obs_logabundance, self._obs_logsigma = self._df_fitsetup()
        else:
            self.df = None
            self._samples = None
            self.times = np.linspace(0, t_end, t_steps)
        inits, params = {}, {}
        if self.df is not None:
            first = self.df[self.df['time'] == 0]['abundance']
            for org, abundance in first.items():
                inits.setdefault(org, abundance)
        for k, v in kwargs.items():
            if k in self._pnames:
                params[k] = v
            if k in self._snames:
                inits[k] = v
        self.set_parameters(**params)
        self.set_inits(**inits)
        self._pnum = 0
        for p in self.parameters:
            if self.parameters[p] is not None:
                self._pnum += np.count_nonzero(self.parameters[p].val)

    # ------------------------------------------------------------------ data setup (Framework.py:266-381)
    def reset_dataframe(self, df):
        self.df = self._formatdf(df.copy())
        self.times = np.linspace(0, max(self.df['time']), len(self.times))
        self._pred_tindex, self._obs_logabundance, self._obs_logsigma = self._df_fitsetup()
        self._samples = len(self.df)
        inits = {}
        for org, abundance in self.df[self.df['time'] == 0]['abundance'].items():
            inits.setdefault(org, abundance)
        self.set_inits(**inits)

    def _formatdf(self, df):
        """(organism, time, abundance[, log_sigma | replicate]) -> frame indexed by organism, time-sorted."""
        df = df.sort_values(by=['organism', 'time'])
        if 'replicate' in df:
            work = df[['organism', 'time', 'abundance']].copy()
            work['log_abundance'] = np.log(work['abundance'])
            agg = work.groupby(by=['time', 'organism']).mean()
            agg['log_sigma'] = work.groupby(by=['time', 'organism']).std()['log_abundance']
            df = agg.reset_index(level='time').sort_values(by='time', kind='stable').sort_index(kind='stable')
        else:
            df = df.set_index('organism')
            if 'abundance' in df and 'log_abundance' not in df:
                df['log_abundance'] = np.log(df['abundance'].to_numpy())
            if 'log_sigma' not in df:
                df['log_sigma'] = 1
                warnings.warn("log_sigma not found, setting log variance to 1")
        return df

    def _df_fitsetup(self):
        """Nearest output-grid index of every observation row (first minimum) + per-organism arrays."""
        tindex, logab, logsig = {}, {}, {}
        for org in dict.fromkeys(self.df.index):
            rows = self.df.loc[[org]]
            tt = rows['time'].to_numpy(dtype=float)
            gap = np.abs(tt[:, None] - self.times[None, :])
            tindex[org] = np.argmax(gap == gap.min(axis=1, keepdims=True), axis=1).astype(np.int64)
            logab[org] = rows['log_abundance'].to_numpy(dtype=float)
            logsig[org] = rows['log_sigma'].to_numpy(dtype=float)
        return tindex, logab, logsig

    def _get_summation_index(self, summation_mapping):
        """{first member index: member indices}, names after summation, kept columns, renamed indices."""
        where = {s: i for i, s in enumerate(self._snames)}
        sums, renamed, used = {}, {}, set()
        for new_name, members in summation_mapping.items():
            idx = []
            for member in members:
                if member in used:
                    raise ValueError("{} state varaiable cannot be used in two summations".format(member))
                if member not in where:
                    raise ValueError("{} state varaiable is not a valid state name".format(member))
                used.add(member)
                idx.append(where[member])
            if len(idx) < 1:
                raise ValueError("Summation of {} needs two or more state variables".format(new_name))
            idx.sort()
            sums[idx[0]] = tuple(idx)
            renamed[idx[0]] = new_name
        names, keep = [], []
        for i, s in enumerate(self._snames):
            if i in renamed:
                names.append(renamed[i]); keep.append(i)
            elif s not in used:
                names.append(s); keep.append(i)
        return sums, tuple(names), tuple(keep), renamed

    # ------------------------------------------------------------------ names / values
    def get_pnames(self):
        return list(self._pnames)

    def get_snames(self, after_summation=True, predict_obs=False):
        if after_summation and self._summations_index:
            return list(self._summation_snames)
        if predict_obs:
            return list(self._pred_tindex.keys())
        return list(self._snames)

    def get_numstatevar(self):
        return len(self._snames)

    def get_model(self):
        return self._model

    def __repr__(self):
        out = ["Current Model = {}".format(str(self._model.__module__) + '.' + str(self._model.__name__)), "Parameters:"]
        out += ["\t{} = {}".format(p, self.parameters[p]) for p in self.get_pnames()]
        out.append("Initial States:")
        out += ["\t{} = {}".format(s, self.istates[s]) for s in self._snames]
        if self._summations_index:
            out.append("Current State Summations")
            for i, members in self._summations_index.items():
                out.append("\t{}={}".format(self._suminds[i], '+'.join(self._snames[j] for j in members)))
        return '\n'.join(out)

    __str__ = __repr__

    def set_parameters(self, **kwargs):
        for p, v in kwargs.items():
            if p not in self.parameters and p in self._flat_slots():
                owner, idx = self._flat_slots()[p]                # one element of an array-valued parameter: "phi[0,1]"
                val = np.array(self.parameters[owner].val, dtype=np.float64)
                val[idx] = v
                self.parameters[owner].val = val
                continue
            if p not in self.parameters:
                raise Exception("{} is an unknown parameter. Acceptable parameters are: {}".format(p, ', '.join(self._pnames)))
            if isinstance(v, parameter):
                self.parameters[p] = v
                if not v.name:
                    v.name = p
            elif self.parameters[p] is not None:
                self.parameters[p].val = v
            else:
                self.parameters[p] = parameter(init_value=v, name=p)

    def set_inits(self, **kwargs):
        summed = set(self._summation_snames)
        for s, v in kwargs.items():
            if s in self.istates:
                self.istates[s] = v
            elif s in summed:
                pass                     # initial value of a summed observable: accepted, not enforced (:476-493)
            else:
                raise Exception("{} is an unknown state variable. Acceptable parameters are: {}".format(s, ', '.join(self._snames)))

    def get_inits(self, as_dict=False):
        if as_dict:
            return self.istates
        return np.array([self.istates[s] for s in self._snames])

    def get_parameters(self, as_dict=False, **kwargs):
        vals = [kwargs[p] if p in kwargs else self.parameters[p].val for p in self._pnames]
        if as_dict:
            return dict(zip(self._pnames, vals))
        return tuple([vals])

    # -- flat layout: one slot per parameter ELEMENT (identical to parameter_names when every parameter is a scalar)
    def _pshapes(self):
        return tuple(() if self.parameters[p] is None else tuple(np.shape(self.parameters[p].val)) for p in self._pnames)

    @property
    def _flat_names(self):
        shapes = self._pshapes()
        if getattr(self, "_flat_cache", (None,))[0] != shapes:
            names, slots = [], {}
            for p, shape in zip(self._pnames, shapes):
                for name, idx in zip(Samplers.element_names(p, shape), np.ndindex(*shape) if shape else [()]):
                    names.append(name)
                    slots[name] = (p, idx)
            self._flat_cache = (shapes, tuple(names), slots)
        return self._flat_cache[1]

    def _flat_slots(self):
        """element name -> (parameter name, index tuple)."""
        self._flat_names
        return self._flat_cache[2]

    def _flat_owner(self):
        slots = self._flat_slots()
        return [slots[f][0] for f in self._flat_names]

    def _flatten(self, values):
        """One value (scalar or array) per parameter name, in parameter_names order -> flat theta."""
        return np.concatenate([np.ravel(np.asarray(v, dtype=np.float64)) for v in values]) if len(values) else np.empty(0)

    def _current_theta(self):
        return self._flatten([self.parameters[p].val for p in self._pnames])

    def _set_theta(self, theta):
        """Flat theta -> parameter values (arrays keep their shape)."""
        k = 0
        for p, shape in zip(self._pnames, self._pshapes()):
            size = int(np.prod(shape)) if shape else 1
            self.parameters[p].val = np.array(theta[k]) if shape == () else np.array(theta[k:k + size], dtype=np.float64).reshape(shape)
            k += size

    def _device_ode(self):
        """The user's RHS as the tracer calls it: ``ps`` is the flat slot vector; array-valued parameters are handed to
        the user's function as arrays of that shape (views of the slot vector), as odeint hands them over at
        Framework.py:656."""
        shapes = self._pshapes()
        if all(sh == () for sh in shapes):
            return self._model
        ode = self._model

        def flat_ode(y, t, ps):
            vals, k = [], 0
            for shape in shapes:
                size = int(np.prod(shape)) if shape else 1
                vals.append(ps[k] if shape == () else np.asarray(ps[k:k + size], dtype=object).reshape(shape))
                k += size
            return ode(y, t, vals)
        return flat_ode

    # ------------------------------------------------------------------ device model
    def _y0_map(self):
        """'<state>0' parameters double as that state's initial value (Samplers.py:110-114)."""
        fn = self._flat_names
        return np.array([fn.index(s + '0') if (s + '0') in fn else -1 for s in self._snames], np.int32)

    def _observe_groups(self):
        if not self._summations_index:
            return [(i,) for i in range(len(self._snames))]
        return [self._summations_index.get(i, (i,)) for i in self._sumkeep]

    def _tables_stamp(self, y0):
        """Content key of everything `_device` uploads: output grid, initial states, '<state>0' map, observation rows."""
        parts = [self.times.tobytes(), y0.tobytes(), self._y0_map().tobytes()]
        if self.df is not None:
            for s in self._pred_tindex:
                parts += [str(s).encode(), self._pred_tindex[s].tobytes(), self._obs_logabundance[s].tobytes(),
                          self._obs_logsigma[s].tobytes()]
        return hash(tuple(parts))

    def _device(self):
        """Compile (once) and (re)load tables when data / initial states changed.

        ``copy()`` shares the compiled DeviceModel between instances, so "what is loaded" is recorded on the
        DeviceModel itself (``_loaded_stamp``) and compared by content: a copy with other initial states or data
        re-uploads its own tables (a few KB) before it computes, and so does the original after it."""
        if self._dm is None or getattr(self, "_dm_shapes", None) != self._pshapes():
            self._dm_shapes = self._pshapes()                     # the slot layout is part of the compiled model
            extra = {} if getattr(self, "_cache_dir", "default") == "default" else {"cache_dir": self._cache_dir}
            self._dm = DeviceModel(self._device_ode(), len(self._snames), len(self._flat_names), self._observe_groups(),
                                   device=self.device, y0_from_param=bool((self._y0_map() >= 0).any()), **extra)
        y0 = np.asarray(self.get_inits(), dtype=np.float64)
        stamp = self._tables_stamp(y0)
        if stamp != getattr(self._dm, "_loaded_stamp", None):
            self._dm._loaded_stamp = None                         # stays invalid if an upload below raises
            if self.df is not None:
                out_names = self.get_snames(after_summation=True)
                cols = [(i, self._pred_tindex[s], self._obs_logabundance[s], self._obs_logsigma[s])
                        for i, s in enumerate(out_names) if s in self._pred_tindex]
                self._dm.set_data(ObsTables(self.times, cols), y0, self._y0_map())
            self._dm.set_grid(self.times, y0, self._y0_map())
            self._dm._loaded_stamp = stamp
        return self._dm

    # ------------------------------------------------------------------ integrate + score (Framework.py:617-722)
    def integrate(self, inits=None, parameters=None, predict_obs=False, as_dataframe=True, sum_subpopulations=True):
        """Solve on ``self.times``; same four return shapes as the reference (Framework.py:622-683)."""
        dm = self._device()
        theta = self._flatten(parameters[0] if parameters else self.get_parameters()[0])
        y0 = None if inits is None else np.asarray(inits, dtype=np.float64)
        traj, status, _ = dm.trajectory(theta[None, :], y0=y0, rtol=self.rtol, atol=self.atol)
        if status[0] != 0:
            warnings.warn("integration stopped early (status {}); remaining rows are NaN".format(int(status[0])))
        mod = traj[0]
        if sum_subpopulations and self._summations_index:
            for first, members in self._summations_index.items():
                mod[:, first] = mod[:, list(members)].sum(axis=1)
            mod = mod[:, list(self._sumkeep)]
        if as_dataframe:
            df = pd.DataFrame(mod, columns=self.get_snames(after_summation=sum_subpopulations))
            df['time'] = self.times
            if predict_obs:
                parts = []
                for s in self.get_snames(predict_obs=True):
                    idx = self._pred_tindex[s]
                    parts.append(pd.DataFrame({'time': self.times[idx], 'abundance': df[s].to_numpy()[idx]},
                                              index=pd.Index([s] * len(idx), name='organism')))
                return pd.concat(parts)
            return df
        if predict_obs:
            return {s: mod[:, i][self._pred_tindex[s]]
                    for i, s in enumerate(self.get_snames(after_summation=sum_subpopulations)) if s in self._pred_tindex}
        return mod

    def get_chi(self, mod_dict):
        O, Cc, S = [], [], []
        with np.errstate(all="ignore"):
            for s in mod_dict:
                O.append(self._obs_logabundance[s]); Cc.append(np.log(mod_dict[s])); S.append(self._obs_logsigma[s])
        return stats.chi(np.concatenate(O), np.concatenate(Cc), np.concatenate(S))

    def get_Rsqrd(self, mod_dict):
        return stats.Rsqrd(mod_dict, {s: np.exp(self._obs_logabundance[s]) for s in self._obs_logabundance if s in mod_dict})

    def get_AIC(self, chi):
        return stats.AIC(chi, self._pnum)

    def get_adjRsqrd(self, mod_dict, Rsqrd=None):
        if not Rsqrd:
            Rsqrd = self.get_Rsqrd(mod_dict)
        return stats.get_adjusted_rsquared(Rsqrd, self._samples, self._pnum)

    def get_fitstats(self, prediction_dict=dict()):
        if not prediction_dict:
            prediction_dict = self.integrate(predict_obs=True, as_dataframe=False)
        fs = {'Chi': self.get_chi(prediction_dict), 'R^2': self.get_Rsqrd(prediction_dict)}
        fs['AIC'] = self.get_AIC(fs['Chi'])
        return fs

    def get_residuals(self):
        """Linear-space residual (model - data) per observation row, index = organism."""
        mod = self.integrate(predict_obs=True)
        parts = [self.df.loc[[s]]['abundance'].to_numpy() for s in self.get_snames(predict_obs=True)]
        return pd.Series(mod['abundance'].to_numpy() - np.concatenate(parts), index=mod.index, name='abundance')

    # ------------------------------------------------------------------ batch seam: _Fit_worker (Framework.py:41-48)
    def sweep(self, parameter_array, rtol=None, atol=None, solver="auto", as_dataframe=False, out=None,
              outputs=("chi", "r2", "status", "nsteps")):
        """chi (and R^2, status, steps) for every row of ``parameter_array`` [n, P] (parameter_names order).

        numpy in -> numpy out (host buffers, copies inside the call); torch CUDA tensor in -> tensors out.
        ``outputs``: what to produce besides chi -- ``("chi",)`` is what the reference's `_Fit_worker` returns
        (Framework.py:41-48) and leaves 8 instead of 24 bytes per row to copy back."""
        dm = self._device()
        res = dm.sweep(parameter_array, rtol=self.rtol if rtol is None else rtol,
                       atol=self.atol if atol is None else atol, solver=solver, out=out, outputs=outputs)
        if as_dataframe:
            df = pd.DataFrame(np.asarray(parameter_array), columns=list(self._flat_names))
            df['chi'] = res['chi']
            return df
        return res

    def _lhs_samples(self, samples=100, **kwargs):
        pdists, pstatic = {}, {}
        for p in self.parameters:
            if p in kwargs:
                pdists[p] = kwargs[p]
            elif self.parameters[p].has_distribution():
                pdists[p] = self.parameters[p]
            else:
                pstatic[p] = self.parameters[p].val
        df = Samplers.sample_lhs(parameter_dict=pdists, samples=samples)
        for p in pstatic:
            for name, v in zip(Samplers.element_names(p, np.shape(pstatic[p])), np.ravel(pstatic[p])):
                df[name] = float(v)
        return df

    DEVICE_SAMPLING_FROM = 65536       # surveys at least this large are sampled on the device (sampler="auto")
    # solver="auto": DOPRI5 attempts per solve before the chain is handed to BDF.  The chains of a launch wait for its
    # slowest one (measured before chains stopped at their first failed solve), and the BDF re-run costs about the same whether it holds 24 chains or 300 (it is latency-bound):
    # 4096 chains x 200 iterations from a wide survey take 0.29 s at 4096, 0.20 s at 2048, 0.15 s at 1024, 0.13 s at 512
    EXPLICIT_STEP_BUDGET = 1024
    PROBE_STEPS = 2048                 # solver="auto": DOPRI5 attempts the probe of the chain starts may take
    STEP_BUDGET_FACTOR = 8             # ... and the budget of a solve in multiples of the probe's mean step count

    def _prior_table(self):
        """(kind, a, b, c) per parameter for the device sampler (engine.DeviceModel.sample_lhs), or None when a prior
        is not one of lognorm / norm / uniform in scipy.stats' (s, loc, scale) parameterisation."""
        table = []
        for f, value in zip(self._flat_owner(), self._current_theta()):
            par = self.parameters[f]
            if par is None or not par.has_distribution() or (np.ndim(par.val) and value == 0.0):
                table.append(("const", float(value), 0.0, 0.0))   # no prior, or a structural zero of an array
                continue
            name, hp = getattr(par.dist, "name", None), dict(par.hp or {})
            extra = set(hp) - {"s", "loc", "scale"}
            if name not in ("lognorm", "norm", "uniform") or extra or (name == "lognorm" and "s" not in hp):
                return None
            table.append((name, float(hp.get("s", 0.0)), float(hp.get("loc", 0.0)), float(hp.get("scale", 1.0))))
        return table

    def _lhs_samples_device(self, samples, sampler="auto"):
        """Latin-hypercube sample of the priors as a CUDA tensor [samples, P], or None when the host sampler is to be
        used (small survey, unsupported prior, sampler="host").  The seed is one draw from numpy's global RandomState,
        so `np.random.seed(k)` keeps a survey reproducible -- as it does for the reference, whose pyDOE2.lhs draws
        from that state too (Samplers.py:33)."""
        if sampler == "host" or (sampler == "auto" and samples < self.DEVICE_SAMPLING_FROM):
            return None
        table = self._prior_table()
        if table is None:
            if sampler == "device":
                raise NotImplementedError("device sampling supports lognorm / norm / uniform priors (s, loc, scale)")
            return None
        seed = int(np.random.randint(0, 2 ** 62))
        return self._device().sample_lhs(table, samples, seed)

    def fit_survey(self, samples=1000, cpu_cores=1, sampler="auto"):
        """LHS sample of the priors, chi of every sample (Framework.py:800-816).  ``cpu_cores`` is ignored.
        sampler: "host" = numpy (Samplers.sample_lhs), "device" = odl_sample_lhs, "auto" = device for large surveys."""
        if self._world()[0] > 1:
            return self._fit_survey_sharded(samples, sampler)
        theta_dev = self._lhs_samples_device(samples, sampler)
        if theta_dev is not None:
            res = self._device().sweep(theta_dev, rtol=self.rtol, atol=self.atol, solver="auto", outputs=("chi",))
            # the host copy of the table is fresh memory nobody else holds: the frame adopts it (no second 40 MB copy)
            out = pd.DataFrame(theta_dev.cpu().numpy(), columns=list(self._flat_names), copy=False)
            out['chi'] = res['chi'].cpu().numpy()
            return out
        ps = self._lhs_samples(samples)[list(self._flat_names)]
        res = self.sweep(ps.to_numpy(dtype=np.float64), outputs=("chi",))
        out = ps.reset_index(drop=True)
        out['chi'] = res['chi']
        return out

    # ------------------------------------------------------------------ several GPUs
    def _world(self):
        """(world_size, rank) of the default process group for a model built with distributed=True, else (1, 0)."""
        if not self.distributed:
            return 1, 0
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1, 0
        return dist.get_world_size(), dist.get_rank()

    def _unfinished_fraction(self, n_bad, n, steps_ok=0.0):
        """Share of the probed chain starts the capped DOPRI5 pass did not finish, and the mean attempted-step count of
        the ones it did -- over ALL ranks when the chains are sharded, so that the stepper and the step budget (and with
        them every chain) do not depend on the number of GPUs."""
        if self._world()[0] > 1:
            import torch
            import torch.distributed as dist
            from .rhat import _collective_device
            t = torch.tensor([float(n_bad), float(n), float(steps_ok)], dtype=torch.float64, device=_collective_device())
            dist.all_reduce(t)
            n_bad, n, steps_ok = t[0].item(), t[1].item(), t[2].item()
        return n_bad / max(1.0, float(n)), steps_ok / max(1.0, float(n) - float(n_bad))

    def _fit_survey_sharded(self, samples, sampler):
        """fit_survey over all ranks: rank 0 draws the design (its numpy / device stream alone decides it) and
        broadcasts it, every rank solves a contiguous block of rows, chi is all-gathered; no collective on the data
        path itself.  Every rank returns the whole frame, rows in the design's order."""
        import torch
        ws, rank = self._world()
        dm = self._device()
        theta = None
        if rank == 0:
            theta = self._lhs_samples_device(samples, sampler)
            if theta is None:
                theta = np.ascontiguousarray(self._lhs_samples(samples)[list(self._flat_names)].to_numpy(dtype=np.float64))
        theta = broadcast_rows(theta, (int(samples), len(self._flat_names))).to(torch.device("cuda", dm.device))
        lo, hi = shard_bounds(int(samples), ws, rank)
        if hi > lo:
            chi = dm.sweep(theta[lo:hi], rtol=self.rtol, atol=self.atol, solver="auto", outputs=("chi",))["chi"]
        else:
            chi = torch.empty(0, dtype=torch.float64, device=theta.device)
        chi = allgather_rows(chi)
        out = pd.DataFrame(theta.cpu().numpy(), columns=list(self._flat_names), copy=False)
        out['chi'] = chi.cpu().numpy()
        return out

    def _mcmc_sharded(self, starts, seeds, nits, static_parameters, rng, want_frame, use_priors):
        """The chains of one MCMC call over all ranks: contiguous blocks, chain seed / Philox key = GLOBAL chain index
        (so the chains do not depend on the number of GPUs), then one all-gather of the per-chain results (summaries
        for R-hat and the report; the kept rows too when the frame is wanted)."""
        ws, rank = self._world()
        C = len(starts)
        lo, hi = shard_bounds(C, ws, rank)
        P = len(self._flat_names)
        burnin = int(nits / 2)
        if rng == "auto":                                         # decided on the whole run, not on this rank's block
            n_walk = sum(1 for owner in self._flat_owner() if owner not in set(static_parameters or ()))
            rng = self._auto_rng(C, nits - 1, n_walk)
        if hi > lo:
            local = self._run_chains(starts[lo:hi], seeds[lo:hi], nits, burnin, static_parameters, rng=rng,
                                     return_raw=True, keep_samples=want_frame, use_priors=use_priors)
        else:                                                     # fewer chains than ranks: this rank only takes part in the collectives
            if self.solver == "auto":
                self._unfinished_fraction(0, 0)
            n_keep = max(0, nits - 1 - burnin)
            local = {"theta": np.empty((0, P)), "chain_state": np.empty((0, 8)), "summaries": np.empty((0, 1 + 2 * P)),
                     "fail_count": np.empty(0, np.int32), "step_count": np.empty(0, np.int64), "best_theta": np.empty((0, P)),
                     "samples": np.empty((0, n_keep, P + 5)) if want_frame else None, "n_keep": n_keep, "burnin": burnin}
        out = dict(local)
        if C > 1 and local["n_keep"] > 1:                         # R-hat: odl_rhat (ncclAllGather + device reduction)
            dm = self._device()
            dm.comm_init()
            out["rhat_device"] = dm.rhat(np.asarray(local["summaries"]))
        for key in ("theta", "chain_state", "summaries", "fail_count", "step_count", "best_theta"):
            out[key] = allgather_rows(np.asarray(local[key]))
        out["best_chi"], out["best_iteration"] = out["chain_state"][:, 3], out["chain_state"][:, 4]
        if want_frame:
            smp = local["samples"]
            smp = allgather_rows(smp)
            out["samples"] = smp.cpu().numpy() if hasattr(smp, "is_cuda") else smp
        self._last_mcmc = out
        return self._frame_from_samples(out["samples"], set(static_parameters or ())) if want_frame else out

    def explore_equilibriums(self, samples=1000, cpu_cores=1, **parameter_mapping):
        """Final state of every LHS sample (Framework.py:819-854, `_Equilibrium_worker` :24-38) -- one batched launch
        of the trajectory kernel on a two-point output grid (t0, t_end): only the final states leave the device."""
        ps = self._lhs_samples(samples, **parameter_mapping)[list(self._flat_names)]
        dm = self._device()
        y0 = np.asarray(self.get_inits(), dtype=np.float64)
        dm._loaded_stamp = None                                   # another grid is loaded for the duration of this call
        dm.set_grid(np.array([self.times[0], self.times[-1]]), y0, self._y0_map())
        try:
            traj, _, _ = dm.trajectory(ps.to_numpy(dtype=np.float64), rtol=self.rtol, atol=self.atol)
        finally:
            self._device()                                        # integrate() expects the full grid
        df = pd.DataFrame(traj[:, -1, :], columns=self.get_snames(after_summation=False))
        for p in self._flat_names:
            df[p] = ps[p].to_numpy()
        return df

    def gradient(self, parameter_name, p_range, intialstates=None, seed_equilibrium=True, aggregate_enpoints=False,
                 print_status=True):
        """One simulation per value of ``parameter_name`` in ``p_range`` (Framework.py:1063-1127, as its docstring and
        comments intend it: the reference's body cannot run -- it tests ``intialstates`` for None the wrong way round
        (:1083-1086) and overwrites the ``parameter`` object with a number (:1095)).

        seed_equilibrium=True: a continuation, each run starts from the previous run's final state floored at 0.001
        (:1099-1101) -- sequential by construction, one single-system launch per value.  seed_equilibrium=False:
        every run starts from the same state, so all of ``p_range`` is ONE batched launch.  Returns a frame with one
        column per state variable (sub-populations not summed, :1097) plus ``parameter_name``: the final states
        (``aggregate_enpoints=True``, [len(p_range), n+1]) or every grid row of every run ([len(p_range)*T, n+1])."""
        if parameter_name not in self._flat_names:
            raise ValueError("{} is not a (scalar or element) parameter of this model".format(parameter_name))
        p_range = np.asarray(p_range, dtype=np.float64).ravel()
        num_sim = len(p_range)
        init = np.asarray(self.get_inits() if intialstates is None else intialstates, dtype=np.float64)
        if init.shape != (len(self._snames),):
            raise ValueError("intialstates must hold one value per state variable")
        col = self.get_snames(after_summation=False) + [parameter_name]
        if print_status and num_sim:
            print("Preparing to run {} simulations between {} and {}".format(num_sim, p_range.min(), p_range.max()))
        if num_sim == 0:
            return pd.DataFrame(np.empty((0, len(col))), columns=col)
        dm = self._device()
        theta = np.tile(self._current_theta(), (num_sim, 1))
        theta[:, self._flat_names.index(parameter_name)] = p_range
        no_map = np.full(len(self._snames), -1, np.int32)          # explicit initial states win, as in integrate(inits=)
        y0 = np.asarray(self.get_inits(), dtype=np.float64)
        grid = np.array([self.times[0], self.times[-1]]) if aggregate_enpoints else self.times
        dm._loaded_stamp = None                                   # another grid is loaded for the duration of this call
        dm.set_grid(grid, y0, no_map)
        try:
            if seed_equilibrium:
                traj = np.empty((num_sim, len(grid), len(self._snames)))
                for i in range(num_sim):
                    if print_status:
                        print("{:.2f}% Complete".format(i / num_sim * 100), end='\r')
                    traj[i] = dm.trajectory(theta[i:i + 1], y0=init, rtol=self.rtol, atol=self.atol)[0][0]
                    init = np.clip(traj[i, -1, :], a_min=.001, a_max=None)
            else:
                traj = dm.trajectory(theta, y0=init, rtol=self.rtol, atol=self.atol)[0]
        finally:
            self._device()
        if print_status:
            print("100.00% Complete")
        if aggregate_enpoints:
            rows = np.column_stack([traj[:, -1, :], p_range])
        else:
            rows = np.column_stack([traj.reshape(-1, traj.shape[2]), np.repeat(p_range, traj.shape[1])])
        return pd.DataFrame(rows, columns=col)

    def copy(self, overwrite=dict()):
        """Independent copy sharing the compiled device model (Framework.py:901-943)."""
        new = ModelFramework.__new__(ModelFramework)
        for k, v in self.__dict__.items():
            if k == 'parameters':
                new.parameters = {p: (q.copy() if q is not None else None) for p, q in v.items()}
            elif isinstance(v, (list, dict, pd.DataFrame, np.ndarray)):
                new.__dict__[k] = v.copy()
            else:
                new.__dict__[k] = v
        ps = {k: v for k, v in overwrite.items() if k in new._pnames}
        st = {k: v for k, v in overwrite.items() if k in new._snames}
        if ps:
            new.set_parameters(**ps)
        if st:
            new.set_inits(**st)
        return new

    def set_best_params(self, posteriors):
        im = posteriors['chi'].idxmin()
        best = posteriors.loc[im][list(self._flat_names)].to_dict()
        self.set_parameters(**best)
        if self._snames[0] + '0' in self.get_pnames():
            self.set_inits(**{s: best[s + '0'] for s in self._snames})

    # ------------------------------------------------------------------ MCMC (Framework.py:946-1061)
    def _run_chains(self, starts, seeds, nits, burnin, static_parameters, rng="auto", rtol=None, atol=None,
                    update_model=False, return_raw=False, return_frame=False, keep_samples=True, use_priors=False):
        """All chains in one kernel.  starts: list of theta vectors, or a CUDA tensor [C, P] (chain starts chosen on
        the device); seeds: per-chain seeds (chain index)."""
        dm = self._device()
        static = set(static_parameters or ())
        walk_names = [p for p in self._pnames if p not in static]
        walk = [i for i, owner in enumerate(self._flat_owner()) if owner not in static]
        on_device = hasattr(starts, "is_cuda")
        theta0 = starts if on_device else np.array(starts, dtype=np.float64)
        C = theta0.shape[0]
        n_iter = nits - 1
        if not burnin:
            burnin = int(nits / 2)
        if rng == "auto":
            # the reference's own numpy streams (bit-for-bit the reference chain), regenerated on the host by the
            # library (odl_reference_streams) while the streams stay small next to the run (160 MB of host memory);
            # Philox on the device beyond that
            rng = self._auto_rng(C, n_iter, len(walk))
        solver = self.solver
        if solver == "auto":
            probe = dm.sweep(theta0, rtol=self.rtol if rtol is None else rtol, atol=self.atol if atol is None else atol,
                             solver="dopri5", max_steps=self.PROBE_STEPS, stiff_check=True, early_check_steps=-1)
            st_, ns_ = probe["status"], probe["nsteps"]
            if on_device:
                n_bad, steps_ok = int((st_ != 0).sum().item()), float(ns_[st_ == 0].sum().item())
            else:
                n_bad, steps_ok = int(np.sum(np.asarray(st_) != 0)), float(np.asarray(ns_)[np.asarray(st_) == 0].sum())
            frac, typical = self._unfinished_fraction(n_bad, C, steps_ok)
            solver = "bdf" if frac > 0.25 else "dopri5"
            # the budget of an explicit solve: well above what this model's solves take (a fixed 1024 sent a third of
            # the chains of a 12-state model, whose ordinary solves take 350 attempts, to the stiff re-run)
            budget = int(max(self.EXPLICIT_STEP_BUDGET, self.STEP_BUDGET_FACTOR * typical))
        else:
            budget = self.EXPLICIT_STEP_BUDGET
        self._last_solver = solver
        # "auto" that settled on DOPRI5: every solve gets a bounded step budget, and a chain that ever exhausts it (a
        # proposal in a stiff corner -- the reference's LSODA would switch to BDF there) is re-run, whole, on the BDF
        # kernel.  A chain's result depends on that chain alone: all DOPRI5, or all BDF.  In the first run such a chain
        # stops at the failed solve (stop_failed): a chain that sits in a stiff corner would otherwise burn the whole
        # budget on every later proposal while the chains of its warp -- and the launch -- wait for it.
        retry = self.solver == "auto" and solver == "dopri5"
        kw = dict(nits=nits, burnin=burnin, walk=walk, pnum=self._pnum, rtol=self.rtol if rtol is None else rtol,
                  atol=self.atol if atol is None else atol, keep_samples=keep_samples, device_buffers=on_device,
                  solver=solver, max_steps=budget if retry else 2000000, stop_failed=retry)
        if use_priors:
            # not the reference's chain: the reference evaluates the priors and never uses them (Samplers.py:118-127)
            table = self._prior_table()
            if table is None:
                raise NotImplementedError("use_priors supports lognorm / norm / uniform priors (s, loc, scale)")
            kw["prior"] = [("const", 0, 0, 0) if (p in static or k == "const") else (k, a, b, c)
                           for p, (k, a, b, c) in zip(self._flat_owner(), table)]
        if rng == "reference":
            walking = [self.parameters[p] for p in walk_names]
            big = C * n_iter * (len(walk) + 1) > self.HOST_STREAM_DOUBLES
            if big and Samplers.device_streams_possible(seeds, walking):
                # too large for the host generator (~10 ns per draw): MT19937 per chain on the device; everything of
                # this run then lives in device buffers
                if not on_device:
                    import torch
                    theta0 = torch.from_numpy(np.ascontiguousarray(theta0)).to(torch.device("cuda", dm.device))
                    on_device = True
                    kw["device_buffers"] = True
                z, u = dm.reference_streams(seeds, n_iter, sum(int(np.size(p.val)) for p in walking),
                                            len([p for p in walking if p.dist]), Samplers.RWALK_SD)
            else:
                z, u = Samplers.reference_streams_batch(seeds, walking, n_iter)
            streams = dict(rng_mode="host", z=z, u=u)
        elif rng == "philox":
            streams = dict(rng_mode="philox", seed=int(self.random_seed), chain_ids=np.asarray(seeds, dtype=np.int64))
        else:
            raise ValueError("rng must be 'auto', 'reference' or 'philox'")
        out = dm.mcmc(theta0, **streams, **kw)
        self._last_rerun = 0
        if retry:
            fails = out["fail_count"]
            bad = np.flatnonzero((fails.cpu().numpy() if on_device else np.asarray(fails)) > 0)
            if bad.size:
                self._last_rerun = int(bad.size)
                # n <= 8: the BDF chain kernel.  Larger systems have no cooperative stiff stepper yet: their re-run is the
                # cooperative DOPRI5 kernel WITHOUT a budget -- exact as well (the plain DOPRI5 chain), slow only where
                # the chain really sits in a stiff corner; the thread-per-system BDF kernel keeps such systems in local
                # memory and took 13 ms per iteration for 12 states
                redo = "bdf" if dm.n_state <= 8 else "dopri5"
                sub = dict(streams)
                if rng == "reference":
                    if hasattr(z, "is_cuda"):                     # streams generated on the device
                        import torch
                        pick = torch.as_tensor(bad, device=z.device)
                        sub["z"], sub["u"] = z[pick].contiguous(), u[pick].contiguous()
                    else:
                        sub["z"], sub["u"] = z[bad], u[bad]
                else:
                    sub["chain_ids"] = np.asarray(seeds, dtype=np.int64)[bad]
                if on_device:
                    import torch
                    sel = torch.as_tensor(bad, device=theta0.device)
                    again = dm.mcmc(theta0[sel].contiguous(), **sub, **dict(kw, solver=redo, max_steps=2000000, stop_failed=False))
                else:
                    sel = bad
                    again = dm.mcmc(theta0[bad], **sub, **dict(kw, solver=redo, max_steps=2000000, stop_failed=False))
                for key in ("theta", "chain_state", "samples", "summaries", "fail_count", "step_count", "best_theta"):
                    if out.get(key) is not None:
                        out[key][sel] = again[key]
        if on_device:                                             # small per-chain results to the host; samples on demand
            out = {k: (v.cpu().numpy() if hasattr(v, "is_cuda") and k != "samples" else v) for k, v in out.items()}
            if out["samples"] is not None and (return_frame or not return_raw):
                out["samples"] = out["samples"].cpu().numpy()
        if C > 1 and out["n_keep"] > 1 and out.get("summaries") is not None and self._world()[0] == 1:
            out["rhat_device"] = dm.rhat(np.asarray(out["summaries"]), local=True)   # odl_rhat: reduction on the device
        self._last_mcmc = out
        if return_raw:
            return out
        cols = list(self._flat_names) + ['chi', 'rsquared', 'aic', 'iteration', 'acceptance_ratio']
        static_cols = [(f, owner) for f, owner in zip(self._flat_names, self._flat_owner()) if owner in static]
        samples = out["samples"]                                  # [C, n_keep, P+5], rows = the reference frame's columns
        if return_frame:
            frames = self._frame_from_samples(samples, static)
        else:
            frames = []
            for c in range(C):
                df = pd.DataFrame(samples[c], columns=cols)
                df['iteration'] = df['iteration'].astype(np.int64)
                for f, p in static_cols:
                    df[f] = self.parameters[p].hp['scale']
                if df.empty:
                    df = pd.DataFrame([[np.nan] * (len(self._flat_names) + 3)])
                frames.append(df)
        if update_model:
            self._set_theta(out["theta"][0])
            if any(m >= 0 for m in self._y0_map()):
                self.set_inits(**{s: out["theta"][0][m] for s, m in zip(self._snames, self._y0_map()) if m >= 0})
        return frames

    # reference streams up to this size are generated on the host (2 GB; bitwise numpy's, 16 threads: 4096 chains x 10,000
    # iterations in 0.5 s -- the device generator, one thread per chain with its key array in global memory, needs 2 s for
    # the same and only wins from ~65,536 chains on)
    HOST_STREAM_DOUBLES = 250_000_000
    DEVICE_STREAM_DOUBLES = 4_000_000_000    # ... up to this size (32 GB of HBM) on the device (odl_reference_streams_device)

    @classmethod
    def _auto_rng(cls, n_chains, n_iter, n_walk):
        """'reference': the reference chain's own numpy random numbers (the library regenerates them: on the host while the
        streams are small, on the device beyond that -- there the gaussians agree with numpy's to 1 ulp); 'philox' only for
        runs whose streams would not fit in device memory either."""
        return "reference" if n_chains * n_iter * (n_walk + 1) <= cls.DEVICE_STREAM_DOUBLES else "philox"

    def _frame_from_samples(self, samples, static):
        """One frame for all chains from the kernel's kept rows [C, n_keep, P+5], assembled without per-chain pandas
        work (Framework.py:1035-1038 equivalent)."""
        cols = list(self._flat_names) + ['chi', 'rsquared', 'aic', 'iteration', 'acceptance_ratio']
        df = pd.DataFrame(samples.reshape(-1, samples.shape[-1]), columns=cols, copy=False)   # adopts the kernel's rows
        df['iteration'] = df['iteration'].astype(np.int64)
        for f, p in zip(self._flat_names, self._flat_owner()):
            if p in static:            # reference quirk A13: static columns report the prior's scale (Samplers.py:166-170)
                df[f] = self.parameters[p].hp['scale']
        df['chain#'] = np.repeat(np.arange(samples.shape[0]), samples.shape[1])
        return df

    def _survey_starts_on_device(self, chain_inits, fitsurvey_samples, sd_fitdistance):
        """Chain starts for ``MCMC(chain_inits=int)`` (Framework.py:993-1016) without bringing the survey back:
        LHS sample -> device, one sweep, threshold filter + ordered compaction on the device, the reference's own
        random picks (DataFrame.sample(n, replace=True) == np.random.choice(len, n, replace=True)) drawn on the
        host from the COUNT alone, gather on the device.  Returns a CUDA tensor [chain_inits, P], or None when
        the survey has no finite chi (the reference then warns and starts every chain from the current values)."""
        import torch
        dm = self._device()
        theta = self._lhs_samples_device(fitsurvey_samples)
        if theta is None:
            ps = self._lhs_samples(fitsurvey_samples)[list(self._flat_names)]
            theta = torch.from_numpy(np.ascontiguousarray(ps.to_numpy(dtype=np.float64))).to(torch.device("cuda", dm.device))
        res = dm.sweep(theta, rtol=self.rtol, atol=self.atol, solver="auto", outputs=("chi",))
        calc = {s: np.exp(self._obs_logabundance[s] + sd_fitdistance * self._obs_logsigma[s]) for s in self._obs_logabundance}
        cutchi = self.get_chi(calc)                              # = n_obs * sd^2 / 2
        _, n_finite = dm.select_below(res["chi"], np.inf)
        if n_finite == 0:
            warnings.warn("Pre-sampling of Multidimentional space failed")
            return None
        index, count = dm.select_below(res["chi"], float(cutchi))
        if count == 0:
            raise ValueError("Preliminary sampling found no parameter sets which meet the minimal threshold \n"
                             "  Try: \n   1. Increasing sd_fitdistance \n   2. Increasing fitsurvey_samples \n"
                             "   3. Different priors and / or different parameter guesses")
        picks = np.random.choice(count, size=chain_inits, replace=True)
        return dm.gather_rows(theta, picks, index=index)

    def posterior_summary(self, static_parameters=()):
        """PosteriorSummary of the last chains, from the device-side reductions alone."""
        out = self._last_mcmc
        fn = self._flat_names
        P = len(fn)
        N, log_mean, log_std = pooled_log_stats(out["summaries"], P)
        stats_ = {p: _rawstats_from_logmoments(log_mean[i], log_std[i]) for i, p in enumerate(fn)}
        best_chi = np.asarray(out["best_chi"], dtype=np.float64)
        have = np.asarray(out["best_iteration"]) > 0
        if have.any() and np.isfinite(best_chi[have]).any():
            c = int(np.nanargmin(np.where(have, best_chi, np.nan)))       # first chain holding the minimum, as idxmin
            best = dict(zip(fn, np.asarray(out["best_theta"])[c]))
            bchi = best_chi[c]
        else:
            best, bchi = dict(zip(fn, self._current_theta())), np.nan
        for f, p in zip(fn, self._flat_owner()):                  # quirk A13: static columns hold the prior's scale
            if p in (static_parameters or ()):
                best[f] = self.parameters[p].hp['scale']
        C = len(best_chi)
        if C > 1 and out["n_keep"] > 1:                           # from odl_rhat when the chains ran through it
            rh = dict(zip(fn, out["rhat_device"][0] if "rhat_device" in out else rhat_from_summaries(out["summaries"], P)))
        else:
            rh = None
        ess = dict(zip(fn, ess_from_summaries(out["summaries"], P))) if rh is not None else None
        n_iter = max(1, int(out["n_keep"]) + int(out["burnin"]))
        acc = float(np.asarray(out["chain_state"])[:, 2].mean()) / n_iter
        return PosteriorSummary(fn, C, N, stats_, best, bchi, rh, acc, ess)

    def MCMC(self, chain_inits=1, iterations_per_chain=1000, cpu_cores=1, static_parameters=list(), print_report=True,
             fitsurvey_samples=1000, sd_fitdistance=3.0, rng="auto", posterior="frame", use_priors=False):
        """Many Metropolis-Hastings chains (Framework.py:946-1061); returns the concatenated posterior frame
        with a ``chain#`` column.  ``cpu_cores`` is ignored: every chain runs concurrently on the GPU.

        The steps either side of the chains stay on the device too: with ``chain_inits=int`` the survey is filtered
        and the starts are gathered there (only the count of acceptable rows comes back), and the fitting report and
        ``set_best_params`` read the kernel's own reductions (pooled log-moments, best kept row, R-hat) instead of
        scanning the frame.  ``posterior="summary"`` skips the frame altogether and returns a PosteriorSummary.

        ``use_priors=True`` (not reference behaviour): the acceptance ratio becomes the posterior's -- prior
        log-densities of the walking parameters and the Hastings term of the multiplicative walk, evaluated in the
        kernel.  The reference computes the prior densities every iteration and drops them (Samplers.py:118-127)."""
        if posterior not in ("frame", "summary"):
            raise ValueError("posterior must be 'frame' or 'summary'")
        if isinstance(chain_inits, pd.DataFrame):
            chain_inits = [row.to_dict() for _, row in chain_inits[list(self._flat_names)].iterrows()]
        base = self._current_theta()
        ws, rank = self._world()
        if isinstance(chain_inits, (int, np.integer)) and ws > 1:
            # rank 0 surveys and picks (its numpy stream alone decides the picks), everyone gets the starts
            import torch
            import torch.distributed as dist
            meta, starts = [None, False], None
            if rank == 0:
                try:
                    starts = self._survey_starts_on_device(int(chain_inits), fitsurvey_samples, sd_fitdistance)
                    meta[1] = starts is not None
                except ValueError as exc:
                    meta[0] = str(exc)
            dist.broadcast_object_list(meta, src=0)
            if meta[0]:
                raise ValueError(meta[0])
            if meta[1]:
                starts = broadcast_rows(starts, (int(chain_inits), len(base))).to(torch.device("cuda", self._device().device))
            else:
                starts = [base.copy() for _ in range(chain_inits)]
        elif isinstance(chain_inits, (int, np.integer)):
            starts = self._survey_starts_on_device(int(chain_inits), fitsurvey_samples, sd_fitdistance)
            if starts is None:
                starts = [base.copy() for _ in range(chain_inits)]
        else:
            starts = []
            for d in chain_inits:
                th = base.copy()
                fn = self._flat_names
                for k, v in d.items():
                    if k in fn:
                        th[fn.index(k)] = float(v)
                    elif k in self._pnames:                       # a whole array-valued parameter
                        lo = self._flat_owner().index(k)
                        vals = np.ravel(np.asarray(v, dtype=np.float64))
                        th[lo:lo + len(vals)] = vals
                starts.append(th)
        n_chains = len(starts)
        seeds = list(range(n_chains))                             # chain seed = chain index (:1015, :1020)
        want_frame = posterior == "frame"
        if ws > 1:
            result = self._mcmc_sharded(starts, seeds, iterations_per_chain, static_parameters, rng, want_frame, use_priors)
        else:
            result = self._run_chains(starts, seeds, iterations_per_chain, int(iterations_per_chain / 2), static_parameters,
                                      rng=rng, return_frame=want_frame, return_raw=not want_frame, keep_samples=want_frame,
                                      use_priors=use_priors)
        summary = self.posterior_summary(static_parameters)
        self.rhat, self.ess = summary.rhat, summary.ess
        if print_report:
            report = ["\nFitting Report\n==============="]
            for col, owner in zip(self._flat_names, self._flat_owner()):
                median, std = summary.stats[col]
                if owner in (static_parameters or ()):
                    continue                                      # constant column: std == 0 in the reference's frame
                if (median != 0.0) and (std != 0.0):
                    report.append("parameter: {}\n\tmedian = {:0.3e}, Standard deviation = {:0.3e}".format(col, median, std))
            self.set_parameters(**summary.best)                   # set_best_params (Framework.py:725-731)
            if self._snames[0] + '0' in self.get_pnames():
                self.set_inits(**{s_: summary.best[s_ + '0'] for s_ in self._snames})
            fs = self.get_fitstats(self.integrate(predict_obs=True, as_dataframe=False))
            report.append("\nMedian parameter fit stats:")
            report.append("\tChi = {:0.3e}\n\tR-squared = {:0.3e}\n\tAIC = {:0.3e}".format(fs['Chi'], fs['R^2'], fs['AIC']))
            if self.rhat:
                report.append("\nGelman-Rubin R-hat (log-parameters): " +
                              ", ".join("{}={:.3f}".format(k, v) for k, v in self.rhat.items()))
                report.append("Effective sample size (between/within chains): " +
                              ", ".join("{}={:.0f}".format(k, v) for k, v in self.ess.items()))
            print('\n'.join(report))
        return result if want_frame else summary
