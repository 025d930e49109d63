"""DeviceModel: one traced ODE model living on one B200, driven through the C ABI.

Host arrays (numpy) take the ODL_MEM_HOST path (the library stages them; copies are inside the call);
torch CUDA tensors take the ODL_MEM_DEVICE path (pointers are handed over as-is, the launch goes on
torch's current stream).  There is no CPU implementation behind this class.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .tracer import trace

SCIPY_TOL = 1.49012e-8   # scipy.integrate.odeint default rtol/atol -- what Framework.py:656 runs with


def _ptr(a):
    """Data pointer of a numpy array or torch tensor (or None)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _is_torch_cuda(a):
    return (a is not None) and (not isinstance(a, np.ndarray)) and hasattr(a, "data_ptr") and a.is_cuda


SOLVERS = {"dopri5": 0, "ros23": 1, "auto": 2, "radau5": 3, "bdf": 4}
# defaults of the ODL_SOLVER_AUTO sweep (odl_capi.cu): the DOPRI5 pass's step cap and its projection check
AUTO_CAP, AUTO_EARLY_CHECK = 704, 384


class ObsTables:
    """Observation tables in the layout odl_model_set_data wants.

    Built from what the reference ctor derives (Framework.py:234, :309-329): per observed output column
    the grid indices of its rows, ln(abundance) and log_sigma, concatenated in chi's order
    (post-summation state order, Framework.py:679-681 / :690-693).
    """

    def __init__(self, times, columns):
        """columns: list of (out_col, tindex[int], ln_obs[float], log_sigma[float]) in concatenation order."""
        times = np.asarray(times, dtype=np.float64)
        idx_all = np.concatenate([np.asarray(c[1], dtype=np.int64) for c in columns]) if columns else np.zeros(0, np.int64)
        self.grid_index = np.unique(idx_all)                       # distinct observation slots, ascending
        self.slot_time = np.ascontiguousarray(times[self.grid_index])
        self.obs_slot = np.searchsorted(self.grid_index, idx_all).astype(np.int32)
        self.obs_col = np.concatenate([np.full(len(c[1]), c[0], np.int32) for c in columns]).astype(np.int32)
        self.ln_obs = np.ascontiguousarray(np.concatenate([np.asarray(c[2], np.float64) for c in columns]))
        self.log_sigma = np.ascontiguousarray(np.concatenate([np.asarray(c[3], np.float64) for c in columns]))
        # sum_s n_s * var(O_s)  (stats.py:55, population variance)
        self.sstot = float(sum(len(c[1]) * np.var(np.exp(np.asarray(c[2], np.float64))) for c in columns))
        self.n_obs = int(self.obs_slot.size)
        self.n_slot = int(self.slot_time.size)
        self.t0 = float(times[0])


class DeviceModel:
    def __init__(self, ode, n_state, n_param, observe_groups=None, device=None, block_threads=0, min_blocks=0,
                 dense_output=True, fmad=True, compile_only=False, cache_dir=_capi.CACHE_DIR, y0_from_param=False,
                 coop_lanes=0, sliced_rhs=None):
        self.traced = trace(ode, n_state, n_param)
        self.n_state, self.n_param = n_state, n_param
        self.groups = [tuple(g) for g in observe_groups] if observe_groups is not None else [(i,) for i in range(n_state)]
        self.n_out = len(self.groups)
        # n > 8: the cooperative kernels spread a system over `lanes` lanes; the tracer lays the outputs out class by class
        # and emits the per-lane right-hand side for exactly that many lanes (tracer.slice_plan)
        lanes = 0
        if n_state > 8:
            lanes = int(coop_lanes) or (2 if n_state <= 18 else 4 if n_state <= 36 else 8 if n_state <= 72 else 16 if n_state <= 144 else 32)
        # sliced_rhs: the per-lane right-hand side (each lane evaluates only its own outputs, leaves through an index table).
        # Measured on B200 and NOT the default: it halves the FP64 work of the 35-state network (FP64 pipe 25 % -> 12 %) but the
        # index arithmetic adds more instructions than the flops it saves (LDG/PRMT/IMAD/LEA = 47 % of the executed
        # instructions), and the kernels are bound by shared-memory latency at 8 warps per SM either way: 2.45 against
        # 2.56 M chain-steps/s for the network, 6.9 against 9.9 for the 12-state chain (profiles/r2j_*)
        if sliced_rhs is None:
            sliced_rhs = os.environ.get("ODL_COOP_SLICED", "0") == "1"
        self.source = self.traced.cuda_source(fmad=fmad, observe_groups=self.groups, coop_lanes=lanes if sliced_rhs else 0)
        coop_lanes = lanes
        self.coop_lanes = lanes                                   # 0 for n <= 8 (thread per system)
        self.rhs_flops = self.traced.flops()
        L = _capi.lib()
        if cache_dir:
            os.makedirs(cache_dir, exist_ok=True)
        bo = _capi.BuildOpts()
        bo.device = -1 if device is None else int(device)
        if not min_blocks:      # registers per thread follow from the state count: 8n doubles of stages + state
            min_blocks = 4 if n_state <= 4 else 3 if n_state <= 6 else 2 if n_state <= 8 else 1
        if not block_threads and n_state > 8:
            block_threads = 64
        bo.block_threads, bo.min_blocks = int(block_threads), int(min_blocks)
        bo.dense_output = 1 if dense_output else 0
        bo.compile_only = int(compile_only)                       # 1 / True: every kernel unit; 2: the default paths' units
        bo.y0_from_param = 1 if y0_from_param else 0
        bo.coop_lanes = int(coop_lanes)
        bo.cache_dir = cache_dir.encode() if cache_dir else None
        h = C.c_void_p()
        _capi.check(L.odl_model_create(self.source.encode(), n_state, n_param, self.n_out, C.byref(bo), C.byref(h)))
        self._h = h
        self._L = L
        self.compile_only = compile_only
        self.device = None
        if not compile_only:
            import torch
            self.device = int(device) if device is not None else int(torch.cuda.current_device())
        self.n_obs = 0
        self.n_slot = 0
        self.n_grid = 0

    # -- lifetime ----------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.odl_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def build_log(self):
        return self._L.odl_model_build_log(self._h).decode(errors="replace")

    def kernel_info(self, kernel="sweep"):
        r, l, b = C.c_int(), C.c_int(), C.c_int()
        _capi.check(self._L.odl_model_kernel_info(self._h, kernel.encode(), C.byref(r), C.byref(l), C.byref(b)))
        return {"regs": r.value, "local_bytes": l.value, "blocks_per_sm": b.value}

    def unit_seconds(self, unit):
        """(NVRTC seconds, cache hit) of a kernel unit ("sweep", "sweep_bdf", "mcmc", ...); (-1, False) = not compiled."""
        sec, hit = C.c_double(), C.c_int()
        _capi.check(self._L.odl_model_unit_seconds(self._h, unit.encode(), C.byref(sec), C.byref(hit)))
        return sec.value, bool(hit.value)

    def last_kernel_ms(self):
        ms = C.c_float()
        _capi.check(self._L.odl_model_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def last_pass_ms(self):
        ms = (C.c_float * 3)()
        _capi.check(self._L.odl_model_last_pass_ms(self._h, ms))
        return [float(x) for x in ms]

    # -- tables ------------------------------------------------------------------------------------
    def set_data(self, tables: ObsTables, y0, y0_from_param=None):
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        y0p = np.ascontiguousarray(y0_from_param if y0_from_param is not None else -np.ones(self.n_state), dtype=np.int32)
        _capi.check(self._L.odl_model_set_data(self._h, tables.n_slot, tables.slot_time.ctypes.data, tables.n_obs,
                                               tables.obs_slot.ctypes.data, tables.obs_col.ctypes.data,
                                               tables.ln_obs.ctypes.data, tables.log_sigma.ctypes.data, tables.sstot,
                                               y0.ctypes.data, y0p.ctypes.data, tables.t0))
        self.n_obs, self.n_slot = tables.n_obs, tables.n_slot
        self.tables = tables

    def set_grid(self, times, y0, y0_from_param=None):
        times = np.ascontiguousarray(times, dtype=np.float64)
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        y0p = np.ascontiguousarray(y0_from_param if y0_from_param is not None else -np.ones(self.n_state), dtype=np.int32)
        _capi.check(self._L.odl_model_set_grid(self._h, times.size, times.ctypes.data, y0.ctypes.data, y0p.ctypes.data))
        self.n_grid = int(times.size)

    # -- helpers -----------------------------------------------------------------------------------
    @staticmethod
    def _solver_opts(rtol, atol, max_steps, solver, stiff_check, h0=0.0, hmax=0.0, stiff_min_steps=0, pass_caps=(0, 0),
                     tail_solver=None, early_check_steps=0, tail_lanes=0, tail_warps=0, auto_flags=0):
        so = _capi.SolverOpts()
        so.rtol = SCIPY_TOL if rtol is None else float(rtol)
        so.atol = SCIPY_TOL if atol is None else float(atol)
        so.h0, so.hmax = float(h0), float(hmax)
        so.max_steps = int(max_steps)
        so.solver = SOLVERS[solver] if isinstance(solver, str) else int(solver)
        so.tail_solver = 0 if tail_solver is None else SOLVERS[tail_solver]
        so.early_check_steps = int(early_check_steps)
        so.tail_lanes = int(tail_lanes)
        so.stiff_check = 1 if stiff_check else 0
        so.stiff_min_steps = int(stiff_min_steps)
        so.pass_cap0 = int(pass_caps[0]) if isinstance(pass_caps, (tuple, list)) else int(pass_caps)
        so.tail_warps, so.auto_flags = int(tail_warps), int(auto_flags)
        return so

    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    # -- forward sweep: _Fit_worker (Framework.py:41-48) -------------------------------------------
    def sweep(self, theta, rtol=None, atol=None, max_steps=500000, solver="dopri5", stiff_check=False,
              return_pred=False, out=None, stiff_min_steps=0, pass_caps=(0, 0), tail_solver=None, early_check_steps=0,
              tail_lanes=0, tail_warps=0, auto_flags=0, outputs=("chi", "r2", "status", "nsteps")):
        """theta [n, P] (numpy -> host path, torch cuda tensor -> device path).

        Returns dict(chi, r2, status, nsteps[, pred]) of the same kind as the input.  ``outputs``: which of r2 / status /
        nsteps to produce besides chi (the others come back as None and are neither written nor copied)."""
        want = set(outputs) | {"chi"}
        so = self._solver_opts(rtol, atol, max_steps, solver, stiff_check, stiff_min_steps=stiff_min_steps,
                               pass_caps=pass_caps, tail_solver=tail_solver, early_check_steps=early_check_steps, tail_lanes=tail_lanes,
                               tail_warps=tail_warps, auto_flags=auto_flags)
        if _is_torch_cuda(theta):
            import torch
            th = theta.contiguous()
            assert th.dtype == torch.float64 and th.dim() == 2 and th.shape[1] == self.n_param
            n = th.shape[0]
            o = out or {}
            chi = o.get("chi") if "chi" in o else torch.empty(n, dtype=torch.float64, device=th.device)
            r2 = (o.get("r2") if "r2" in o else torch.empty(n, dtype=torch.float64, device=th.device)) if "r2" in want else None
            status = (o.get("status") if "status" in o else torch.empty(n, dtype=torch.int32, device=th.device)) if "status" in want else None
            nsteps = (o.get("nsteps") if "nsteps" in o else torch.empty(n, dtype=torch.int32, device=th.device)) if "nsteps" in want else None
            pred = (o.get("pred") if "pred" in o else torch.empty((n, self.n_obs), dtype=torch.float64, device=th.device)) if return_pred else None
            mem, stream = _capi.MEM_DEVICE, self._stream()
        else:
            th = np.ascontiguousarray(theta, dtype=np.float64)
            if th.ndim != 2 or th.shape[1] != self.n_param:
                raise ValueError(f"theta must be [n, {self.n_param}]")
            n = th.shape[0]
            o = out or {}
            chi = o.get("chi") if "chi" in o else np.empty(n, np.float64)
            r2 = (o.get("r2") if "r2" in o else np.empty(n, np.float64)) if "r2" in want else None
            status = (o.get("status") if "status" in o else np.empty(n, np.int32)) if "status" in want else None
            nsteps = (o.get("nsteps") if "nsteps" in o else np.empty(n, np.int32)) if "nsteps" in want else None
            pred = (o.get("pred") if "pred" in o else np.empty((n, self.n_obs), np.float64)) if return_pred else None
            mem, stream = _capi.MEM_HOST, None
        _capi.check(self._L.odl_sweep(self._h, C.byref(so), n, _ptr(th), mem, _ptr(chi), _ptr(r2), _ptr(status),
                                      _ptr(nsteps), _ptr(pred), stream))
        res = {"chi": chi, "r2": r2, "status": status, "nsteps": nsteps}
        if return_pred:
            res["pred"] = pred
        return res

    # -- Latin-hypercube sample of the priors on the device (Samplers.py:6-51) ----------------------
    PRIOR_KINDS = {"const": 0, "lognorm": 1, "norm": 2, "uniform": 3}

    def sample_lhs(self, priors, n, seed=0):
        """priors: one (kind, a, b, c) per parameter (kind in PRIOR_KINDS; scipy.stats parameterisation: lognorm
        (s, loc, scale), norm (-, loc, scale), uniform (-, loc, scale), const (value, -, -)) -> CUDA tensor [n, P]."""
        import torch
        kind = np.array([self.PRIOR_KINDS[k] if isinstance(k, str) else int(k) for k, _, _, _ in priors], np.int32)
        a, b, c = (np.array([float(p[i]) for p in priors], np.float64) for i in (1, 2, 3))
        theta = torch.empty((int(n), len(priors)), dtype=torch.float64, device=torch.device("cuda", self.device))
        _capi.check(self._L.odl_sample_lhs(self._h, int(n), len(priors), kind.ctypes.data, a.ctypes.data, b.ctypes.data,
                                           c.ctypes.data, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(theta), self._stream()))
        return theta

    # -- the reference chain's numpy random numbers, generated on the device (Samplers.py:70, :108, :127) -------
    def reference_streams(self, seeds, n_iter, n_walk, n_prior_draws, step_sd=0.05):
        """z [C, n_iter, n_walk], u [C, n_iter] as CUDA tensors: numpy's legacy RandomState(seed) per chain, consumed as
        MetropolisHastings consumes it (odl_reference_streams_device; uniforms bit-identical, gaussians to 1 ulp)."""
        import torch
        sd = np.ascontiguousarray(seeds, dtype=np.uint32)
        dev = torch.device("cuda", self.device)
        z = torch.empty((len(sd), int(n_iter), int(n_walk)), dtype=torch.float64, device=dev)
        u = torch.empty((len(sd), int(n_iter)), dtype=torch.float64, device=dev)
        _capi.check(self._L.odl_reference_streams_device(self._h, sd.ctypes.data, len(sd), int(n_iter), int(n_walk),
                                                         int(n_prior_draws), float(step_sd), _ptr(z), _ptr(u), self._stream()))
        return z, u

    # -- chain-start selection on the device (Framework.py:1004-1012) -------------------------------
    def select_below(self, chi, cut):
        """Rows of the CUDA tensor `chi` with chi < cut, ascending -> (index tensor [n] int32, count)."""
        import torch
        assert _is_torch_cuda(chi) and chi.dtype == torch.float64 and chi.is_contiguous()
        index = torch.empty(chi.numel(), dtype=torch.int32, device=chi.device)
        count = C.c_longlong(0)
        _capi.check(self._L.odl_select_below(self._h, _ptr(chi), chi.numel(), float(cut), _ptr(index), C.byref(count),
                                             self._stream()))
        return index, int(count.value)

    def gather_rows(self, src, picks, index=None):
        """dst[r] = src[index[picks[r]]] (or src[picks[r]]) for a CUDA matrix `src`; picks: host int64 array."""
        import torch
        assert _is_torch_cuda(src) and src.dtype == torch.float64 and src.dim() == 2
        src = src.contiguous()
        picks = np.ascontiguousarray(picks, dtype=np.int64)
        dst = torch.empty((picks.size, src.shape[1]), dtype=torch.float64, device=src.device)
        _capi.check(self._L.odl_gather_rows(self._h, _ptr(src), int(src.shape[1]), _ptr(index), picks.ctypes.data,
                                            picks.size, _ptr(dst), self._stream()))
        return dst

    # -- the one collective: R-hat over the chains of every rank ----------------------
    def comm_init(self, group=None):
        """Join the NCCL communicator of this model's library handle: rank 0 draws the id (odl_comm_unique_id), the
        process group that torchrun set up carries its 128 bytes to the other ranks (plumbing), every rank calls
        odl_comm_init.  Without an initialised process group (one GPU): world 1, no NCCL."""
        import torch
        import torch.distributed as dist
        if getattr(self, "_comm_world", None) is not None:
            return self._comm_world
        world, rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        idbuf = np.zeros(_capi.COMM_ID_BYTES, np.uint8)
        if world > 1:
            if rank == 0:
                _capi.check(self._L.odl_comm_unique_id(idbuf.ctypes.data))
            dev = torch.device("cuda", self.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.from_numpy(idbuf).to(dev)
            dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            idbuf = t.cpu().numpy()
        _capi.check(self._L.odl_comm_init(self._h, idbuf.ctypes.data, world, rank))
        self._comm_world = world
        return world

    def rhat(self, summaries, local=False):
        """Gelman-Rubin R-hat [P] over the chains of all ranks from this rank's per-chain summaries [m_local, 1+2P]
        (numpy or CUDA tensor): odl_rhat -- ncclAllGather + reduction on the device.  Also returns the pooled
        (count, log_mean[P], log_std[P]) of all kept rows and the total number of chains."""
        P = self.n_param
        if _is_torch_cuda(summaries):
            sm, mem = summaries.contiguous(), _capi.MEM_DEVICE
        else:
            sm, mem = np.ascontiguousarray(summaries, dtype=np.float64), _capi.MEM_HOST
        rh = np.empty(P)
        pooled = np.empty(1 + 2 * P)
        total = C.c_longlong(0)
        if local:                                                 # this rank's chains only, no collective (see odl_rhat)
            mem |= _capi.RHAT_LOCAL
        _capi.check(self._L.odl_rhat(self._h, _ptr(sm), int(sm.shape[0]), P, mem, rh.ctypes.data, pooled.ctypes.data,
                                     C.byref(total), self._stream() if mem == _capi.MEM_DEVICE else None))
        N = pooled[0]
        with np.errstate(all="ignore"):
            std = np.sqrt(pooled[1 + P:] / (N - 1.0)) if N > 1 else np.full(P, np.nan)
        return rh, (float(N), pooled[1:1 + P].copy(), std), int(total.value)

    # -- full-grid trajectories: ModelFramework.integrate (Framework.py:656) -----------------------
    def trajectory(self, theta, y0=None, rtol=None, atol=None, max_steps=500000):
        so = self._solver_opts(rtol, atol, max_steps, "dopri5", False)
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        n = th.shape[0]
        y0a = None if y0 is None else np.ascontiguousarray(np.broadcast_to(np.asarray(y0, np.float64), (n, self.n_state)))
        traj = np.empty((n, self.n_grid, self.n_state), np.float64)
        status = np.empty(n, np.int32)
        nsteps = np.empty(n, np.int32)
        _capi.check(self._L.odl_trajectory(self._h, C.byref(so), n, th.ctypes.data, _ptr(y0a), _capi.MEM_HOST,
                                           traj.ctypes.data, status.ctypes.data, nsteps.ctypes.data, None))
        return traj, status, nsteps

    # -- Metropolis-Hastings: Samplers.py:53-174 for many chains -----------------------------------
    def mcmc(self, theta0, nits=1000, burnin=None, walk=None, pnum=None, rng_mode="philox", seed=0, chain_offset=0,
             z=None, u=None, forced=None, rtol=None, atol=None, max_steps=500000, solver="dopri5", step_sd=0.05,
             trace=False, keep_samples=True, summaries=True, segments=1, device_buffers=False, speculate=0,
             chain_ids=None, prior=None, sample_layout="iteration", explicit_budget=0, stop_failed=False):
        """Run len(theta0) independent chains.  Returns dict with numpy arrays (or torch tensors when
        device_buffers=True): theta (final points), samples [C, nits-1-burnin, P+5], summaries
        [C, 1+2P], chain_state [C,8] (chi, r2, accepts, best_chi, best_iteration, ...), best_theta [C, P] (the
        chain's first minimum of chi over its kept rows), and with trace=True chinew/accepted [C, nits-1].

        sample_layout: how the kernel lays the kept rows out in memory -- "iteration" (default: [n_keep][C][P+5], the rows
        a warp keeps in one iteration are contiguous and leave as coalesced stores) or "chain" ([C][n_keep][P+5]).  Either
        way ``samples`` comes back indexed [chain, row, column]; for "iteration" that is a transposed VIEW of the buffer
        (nothing is copied until somebody asks for chain-major memory, e.g. the frame builder's reshape).

        prior: None = the reference's chain (prior densities never enter the acceptance ratio, Samplers.py:118-127);
        a list of (kind, a, b, c) per parameter (as sample_lhs) = Metropolis-Hastings on the posterior, the prior
        log-densities and the Hastings term of the multiplicative walk evaluated in the kernel.

        solver="auto": per solve, DOPRI5 within ``explicit_budget`` attempted steps (0: until Hairer's test calls the solve
        stiff), else the same solve again on BDF (odl_mcmc_auto_kernel).  stop_failed: a chain stops at its first failed
        solve (the caller re-runs it with another stepper; fail_count > 0 marks it)."""
        so = self._solver_opts(rtol, atol, max_steps, solver, False, pass_caps=int(explicit_budget))
        P = self.n_param
        walk = list(range(P)) if walk is None else [int(w) for w in walk]
        if not burnin:
            burnin = int(nits / 2)                                  # Samplers.py:85-86
        n_iter = nits - 1
        n_keep = max(0, n_iter - burnin)
        mode = {"philox": 0, "host": 1, "forced": 2}[rng_mode] if isinstance(rng_mode, str) else int(rng_mode)
        mo = _capi.McmcOpts()
        walk_arr = (C.c_int * max(1, len(walk)))(*walk)
        mo.n_chain = 0
        mo.chain_offset, mo.nits, mo.burnin = int(chain_offset), int(nits), int(burnin)
        mo.rng_mode, mo.n_walk, mo.walk = mode, len(walk), C.cast(walk_arr, C.POINTER(C.c_int))
        mo.pnum = int(P if pnum is None else pnum)
        mo.row_stride = P + 5
        mo.step_sd, mo.seed = float(step_sd), int(seed) & 0xFFFFFFFFFFFFFFFF
        mo.speculate = int(speculate)
        mo.stop_failed_chains = 1 if stop_failed else 0
        it_major = {"iteration": True, "chain": False}[sample_layout]
        mo.sample_layout = _capi.SAMPLES_ITERATION_MAJOR if it_major else _capi.SAMPLES_CHAIN_MAJOR

        if device_buffers:
            import torch
            dev = theta0.device if _is_torch_cuda(theta0) else torch.device("cuda", torch.cuda.current_device())
            f64 = dict(dtype=torch.float64, device=dev)
            theta = torch.as_tensor(theta0, **f64).contiguous().clone()
            Cn = theta.shape[0]
            new = lambda shape, dt=torch.float64: torch.zeros(shape, dtype=dt, device=dev)
            conv = lambda a: None if a is None else torch.as_tensor(a, **f64).contiguous()
            mem, stream = _capi.MEM_DEVICE, self._stream()
            u8, i32, i64 = torch.uint8, torch.int32, torch.int64
        else:
            theta = np.array(theta0, dtype=np.float64, order="C", ndmin=2)
            Cn = theta.shape[0]
            new = lambda shape, dt=np.float64: np.zeros(shape, dt)
            conv = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
            mem, stream = _capi.MEM_HOST, None
            u8, i32, i64 = np.uint8, np.int32, np.int64
        if theta.shape[1] != P:
            raise ValueError(f"theta0 must be [C, {P}]")
        mo.n_chain = Cn
        state = new((Cn, 8))
        best = new((Cn, P))
        samples = (new((n_keep, Cn, P + 5)) if it_major else new((Cn, n_keep, P + 5))) if keep_samples else None
        summ = new((Cn, 1 + 2 * P)) if summaries else None
        tr_chi = new((Cn, n_iter)) if trace else None
        tr_acc = new((Cn, n_iter), u8) if trace else None
        fails = new((Cn,), i32)
        steps = new((Cn,), i64)
        z, u, forced = conv(z), conv(u), conv(forced)
        ptab = None
        if prior is not None:                                   # (kind, a, b, c) per parameter: posterior-ratio chains
            rows = [[float(self.PRIOR_KINDS[k] if isinstance(k, str) else k), float(a), float(b), float(c)]
                    for k, a, b, c in prior]
            assert len(rows) == P
            ptab = conv(np.array(rows, dtype=np.float64))
        ids = None
        if chain_ids is not None:                               # global chain index per chain (Philox key)
            if device_buffers:
                ids = torch.as_tensor(np.asarray(chain_ids, dtype=np.int64), device=dev).contiguous()
            else:
                ids = np.ascontiguousarray(chain_ids, dtype=np.int64)
            assert ids.shape[0] == Cn
        io = _capi.McmcIO(_ptr(theta), _ptr(state), _ptr(samples), _ptr(summ), _ptr(z), _ptr(u), _ptr(forced),
                          _ptr(tr_chi), _ptr(tr_acc), _ptr(fails), _ptr(steps), _ptr(best), _ptr(ids), _ptr(ptab))
        # optional segmentation of long chains into several launches (state persists in the buffers)
        bounds = np.linspace(1, nits, int(segments) + 1).astype(int)
        ms = 0.0
        for a, b in zip(bounds[:-1], bounds[1:]):
            if b <= a:
                continue
            mo.it_begin, mo.it_end = int(a), int(b)
            _capi.check(self._L.odl_mcmc(self._h, C.byref(so), C.byref(mo), C.byref(io), mem, stream))
            if not device_buffers:
                ms += self.last_kernel_ms()
        if samples is not None and it_major:                   # [chain, row, column] view of the iteration-major buffer
            samples = samples.permute(1, 0, 2) if device_buffers else samples.transpose(1, 0, 2)
        res = {"theta": theta, "chain_state": state, "samples": samples, "summaries": summ, "fail_count": fails,
               "step_count": steps, "kernel_ms": ms, "n_keep": n_keep, "burnin": burnin,
               "best_theta": best, "best_chi": state[:, 3], "best_iteration": state[:, 4]}
        if trace:
            res["chinew"], res["accepted"] = tr_chi, tr_acc
        return res


def fp64_peak(device=-1, repeats=5):
    """Measured DFMA throughput (TFLOP/s, FMA = 2 flop) -- the FP64 roofline denominator."""
    t, ms = C.c_double(), C.c_float()
    _capi.check(_capi.lib().odl_fp64_peak(int(device), int(repeats), C.byref(t), C.byref(ms)))
    return t.value, ms.value


def philox4x32_10(counter, key):
    """Host copy of the device generator (odl_kernels.cuh: odl_philox) for reproducing device streams.

    counter: uint32 array [..., 4]; key: uint32 array [..., 2] -> uint32 [..., 4]."""
    c = np.array(counter, dtype=np.uint64, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint64), c.shape[:-1] + (2,)), copy=True)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[..., 0]
        p1 = M1 * c[..., 2]
        n0 = (p1 >> np.uint64(32)) ^ c[..., 1] ^ k[..., 0]
        n2 = (p0 >> np.uint64(32)) ^ c[..., 3] ^ k[..., 1]
        c = np.stack([n0 & mask, p1 & mask, n2 & mask, p0 & mask], axis=-1)
        k[..., 0] = (k[..., 0] + np.uint64(0x9E3779B9)) & mask
        k[..., 1] = (k[..., 1] + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def philox_streams(seed, chains, n_iter, n_walk, step_sd=0.05):
    """The (z, u) streams the device draws in ODL_RNG_PHILOX mode, recomputed on the host.

    chains: global chain indices.  Returns z [C, n_iter, n_walk], u [C, n_iter]."""
    chains = np.asarray(chains, dtype=np.uint64)
    Cn = chains.size
    it = np.arange(1, n_iter + 1, dtype=np.uint64)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)

    def u53(a, b):
        v = ((a.astype(np.uint64) << np.uint64(21)) ^ (b.astype(np.uint64) >> np.uint64(11))) & np.uint64((1 << 53) - 1)
        return v.astype(np.float64) * 2.0 ** -53

    def draw(block):
        ctr = np.empty((Cn, n_iter, 4), dtype=np.uint64)
        ctr[..., 0] = it[None, :]
        ctr[..., 1] = block
        ctr[..., 2] = (chains & np.uint64(0xFFFFFFFF))[:, None]
        ctr[..., 3] = (chains >> np.uint64(32))[:, None]
        return philox4x32_10(ctr, key)

    r = draw(0)
    u = u53(r[..., 0], r[..., 1])
    z = np.empty((Cn, n_iter, n_walk))
    for j in range(0, n_walk, 2):
        r = draw(1 + j // 2)
        u1 = 1.0 - u53(r[..., 0], r[..., 1])
        u2 = u53(r[..., 2], r[..., 3])
        rad = np.sqrt(-2.0 * np.log(u1))
        z[..., j] = step_sd * (rad * np.cos(2.0 * np.pi * u2))
        if j + 1 < n_walk:
            z[..., j + 1] = step_sd * (rad * np.sin(2.0 * np.pi * u2))
    return z, u
