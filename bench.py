#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200 (BASELINE.json: ODE solves/s + MCMC chain-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU under torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   the reference's own CPU path (oracle/_ref)

One "step" = one pass of the forward sweep (BASELINE.json configs[1]): `--sets` (default 1,048,576) parameter
sets per GPU drawn from the two_i priors of the demo, each integrated to the demo's observation grid and
scored (chi, R^2) -- one odl_sweep call (ODL_SOLVER_AUTO: ordering, DOPRI5 bulk pass, BDF stiff pass beside it).
`value` = solves/s with theta resident in HBM, timed with CUDA events on the launching stream; `e2e` = the same
through ModelFramework.sweep with pinned HOST buffers (H2D + D2H inside the timed call).  In the same JSON line:
"mcmc" (chain-steps/s, the reference's CPU sampler timed beside it), "configs" (BASELINE configs 3-5), "cold_start_s",
"cpu_baseline" (the unmodified reference's fit_survey on the host cores, oracle/_ref).  Prints exactly one JSON line
on rank 0 (the last line of stdout).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "two_i"
TSTEPS = 1000
CENTER = {"two_i": [7.475e-09, 1.069e-07, 19.73, 1.934, 2.799]}   # posterior medians of the demo notebook (:15120-15128)


def demo_frame():
    import pandas as pd
    df = pd.read_csv(os.path.join(ROOT, "tests", "golden", "demodata.csv"))
    return df.replace({"virus": "V", "host": "H"})


def prior_draws(n, seed, offset=0):
    """iid draws from the demo's lognorm priors, reproducible per global row index block."""
    from odelib_b200 import demo_models
    rng = np.random.default_rng([seed, offset])
    pri = demo_models.PRIORS[MODEL]
    names = demo_models.PARAMETER_NAMES[MODEL]
    return np.column_stack([pri[p][1] * np.exp(pri[p][0] * rng.standard_normal(n)) for p in names])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU side: the reference's algorithm (oracle port: scipy odeint + numpy masked chi) on host cores
# ----------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init():
    from oracle import odelib_oracle as orc
    from odelib_b200 import demo_models
    _CPU["orc"] = orc
    _CPU["tab"] = orc.build_tables(demo_frame(), demo_models.STATE_NAMES[MODEL], {"H": ["S", "I1", "I2"]}, TSTEPS,
                                   {"S": 5236900})


def _cpu_chunk(theta):
    import warnings
    warnings.filterwarnings("ignore")
    orc, tab = _CPU["orc"], _CPU["tab"]
    out = np.empty(len(theta))
    for i, th in enumerate(theta):
        out[i] = orc.solve_unit(orc.two_i, th, tab)[1]
    return out


def cpu_sweep_rate(theta, cores, pool=None):
    """solves/s of the oracle port over `theta` with `cores` processes (fork; like the reference's Pool)."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
        pool.map(_cpu_chunk, [theta[:2]] * cores)            # spin the workers up outside the timing
    chunks = np.array_split(theta, cores * 4)
    t0 = time.perf_counter()
    pool.map(_cpu_chunk, chunks)
    dt = time.perf_counter() - t0
    if own:
        pool.close(); pool.join()
    return len(theta) / dt, dt


def reference_model():
    """The UNMODIFIED reference's ModelFramework for the demo two_i model (oracle/_ref), or None when the install is absent."""
    from oracle import ref_runner
    if not ref_runner.available():
        return None, ref_runner
    return ref_runner.demo_model(MODEL, os.path.join(ROOT, "tests", "golden", "demodata.csv")), ref_runner


def cpu_reference_sweep(n, cores, repeats=1, warmup=0):
    """solves/s of the reference's own `fit_survey(samples=n, cpu_cores=cores)` (Framework.py:800-816): LHS of the two_i
    priors, multiprocessing.Pool fan-out of `_Fit_worker`, frame concat -- stock code path, pool start-up included, as a
    user of the reference pays it.  -> (rate, seconds per call) or None without oracle/_ref."""
    model, rr = reference_model()
    if model is None:
        return None
    for k in range(warmup):
        rr.time_fit_survey(model, n, cores, seed=100 + k)
    total = 0.0
    for k in range(repeats):
        total += rr.time_fit_survey(model, n, cores, seed=k)[1]
    return n * repeats / total, total / repeats


def cpu_reference_chains(cores):
    """CPU chain-steps/s of the reference's sampler: `Samplers.MetropolisHastings` 1 chain x
    1000 iterations on one core, and `ModelFramework.MCMC(32 chains x 1000, cpu_cores=all)` from explicit starts."""
    from oracle import ref_runner as rr
    if not rr.available():
        return None
    names = [p[0] for p in rr.PRIORS[MODEL]]
    model = rr.demo_model(MODEL, os.path.join(ROOT, "tests", "golden", "demodata.csv"), init=CENTER[MODEL])
    r1, t1, f1 = rr.time_single_chain(model, 1000, seed=0)
    rng = np.random.default_rng(1)
    starts = [dict(zip(names, np.array(CENTER[MODEL]) * np.exp(0.05 * rng.standard_normal(len(names))))) for _ in range(32)]
    r32, t32, f32 = rr.time_mcmc(model, starts, 1000, cores)
    return {"kind": "reference", "cores": cores,
            "single_chain": {"chain_steps_per_s": r1, "seconds": t1, "iterations": 1000, "cores": 1,
                             "api": "ODElib.Statistics.Samplers.MetropolisHastings(model, nits=1000)"},
            "mcmc_32x1000": {"chain_steps_per_s": r32, "seconds": t32, "chains": 32, "iterations": 1000, "cores": cores,
                             "rows": int(len(f32)),
                             "api": "ODElib.ModelFramework.MCMC(chain_inits=[32 dicts], iterations_per_chain=1000, cpu_cores=all)"}}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path -- `ModelFramework.fit_survey` of the
    unmodified package (oracle/_ref) with all host cores -- on a bounded sample of the workload per step; the oracle
    port (scipy odeint + masked chi, bare) is timed beside it as a second figure."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n = cores * 1500
    ref = cpu_reference_sweep(n, cores, repeats=args.steps, warmup=min(args.warmup, 2))
    theta = prior_draws(n, 0)
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
    pool.map(_cpu_chunk, [theta[:2]] * cores)
    port_rate, port_dt = cpu_sweep_rate(theta, cores, pool)
    pool.close(); pool.join()
    port = {"value": port_rate, "unit": "solves/s", "cores": cores, "kind": "port",
            "sample": f"{n} prior draws, scipy odeint + numpy masked chi in {cores} forked processes (no frame, no pool start-up)"}
    if ref is not None:
        value, per_step = ref
        kind = "reference"
        sample = (f"ODElib.ModelFramework.fit_survey(samples={n}, cpu_cores={cores}) of the unmodified reference (oracle/_ref): "
                  f"LHS of the two_i priors + Pool fan-out of _Fit_worker + concat, {per_step:.1f} s per step")
    else:
        value, per_step, kind, sample = port_rate, port_dt, "port", port["sample"]
    emit_record({
        "impl": "reference", "metric": "ode_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": kind, "sample": sample, "port": port},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def workload_config(args, n_gpus):
    return {"workload": f"forward sweep: {args.sets} prior-sampled parameter sets per GPU of the demo {MODEL} "
                        f"infection model (4 states, 5 parameters) integrated to the 37 demo observations "
                        f"(19 grid times of t_steps={TSTEPS}) + chi/R^2 [BASELINE.json configs[1]]",
            "sets_per_gpu": args.sets, "sets_total": args.sets * n_gpus, "rtol": 1.49012e-8, "atol": 1.49012e-8,
            "solver": "auto: rows cost-ordered on the device (|J(y0)| key), dopri5(4) with dense output <=704 attempted "
                      "steps (projection check at 384) on ~90 % of the SMs, variable-order BDF for what is left (~0.9 %), "
                      "continuing from where DOPRI5 stopped, BESIDE it on SMs of its own (14 CTAs of 8 warps, clusters of 2; "
                      "a second consumer on the SMs the DOPRI5 pass frees when it ends)",
            "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": f"shard{n_gpus}"}



# ----------------------------------------------------------------------------------------------------
# BASELINE.json configs 3-5 and the cold start: extra legs of the same JSON line (key "configs" / "cold_start_s")
# ----------------------------------------------------------------------------------------------------
def config_legs(world, rank, local, dev, peak_tflops, barrier, max_over_ranks, sum_over_ranks):
    """Config 3 (N-class chains, 4096 chains x 10,000 iterations per GPU), config 4 (stiff variant: 65,536-set sweep +
    1024 chains x 2000 iterations on the BDF kernels, per GPU), config 5 (5x5 network: 65,536 chains SHARED by the
    GPUs -- strong scaling -- with the R-hat all-gather timed).  All through ModelFramework-built models on synthetic
    data of each config's shape (odelib_b200/workloads.py)."""
    import torch
    from odelib_b200 import workloads

    def flops_per_step(dm):
        return 6 * dm.rhs_flops + 71 * dm.n_state + 10

    def run_chains(dm, starts, nits, **kw):
        dm.mcmc(starts, nits=min(nits, 40), rng_mode="philox", seed=1, device_buffers=True, keep_samples=False, **kw)   # warm-up
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        res = dm.mcmc(starts, nits=nits, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False, **kw)
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b) * 1e-3), res

    out = {}
    legs = os.environ.get("ODL_BENCH_LEGS", "c3,c4,c5").split(",")    # development: run a subset
    note = lambda msg: print("[bench] " + msg, file=sys.stderr, flush=True) if rank == 0 else None
    # ---- config 3 ----
    c3 = {}
    for N in ((1, 4, 10) if "c3" in legs else ()):
        note(f"config 3, N={N}")
        m, center = workloads.nclass(N, device=local)
        dm = m._device()
        # N = 10 (12 states, cooperative kernels): 2,000 of the 10,000 iterations.  Over long runs a few chains drift along
        # the unidentifiable direction tau -> infinity into stiff territory (the reference's LSODA switches to BDF there);
        # systems beyond 8 states have no cooperative stiff stepper yet, their re-run is the cooperative DOPRI5 kernel
        # without a step budget, and 44 such chains took 108 s of a 116 s run at 10,000 iterations ("next")
        C, nits, P = 4096, (10000 if dm.n_state <= 8 else 2000), dm.n_param
        rng = np.random.default_rng([3, N, rank])
        starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C, P)))).to(dev)
        seeds = list(range(rank * C, rank * C + C))
        # through the facade's chain runner (what ModelFramework.MCMC calls): every solve has a budget of 1024 DOPRI5
        # attempts, and a chain that ever exhausts it -- a proposal in a stiff corner, where the reference's LSODA switches
        # to BDF -- stops there and is re-run, whole, on the BDF kernel with the same random stream.  (Unbounded, one such proposal in
        # 4e7 holds its warp for seconds: 10,000 iterations took 81 s instead of 2 s.)
        kw = dict(rng="philox", return_raw=True, keep_samples=False)
        m._run_chains(starts, seeds, 60, 30, (), **kw)                          # warm-up (kernels, buffers)
        barrier(); t0 = time.perf_counter()
        res = m._run_chains(starts, seeds, nits, nits // 2, (), **kw)
        torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        steps = sum_over_ranks(float(np.asarray(res["step_count"]).sum()))
        coop = dm.n_state > 8
        c3[f"N={N}"] = {"states": dm.n_state, "parameters": P, "chains_per_gpu": C, "iterations": nits, "seconds": t,
                        "chain_steps_per_s": C * world * (nits - 1) / t,
                        "fp64_tflops_per_gpu": steps * flops_per_step(dm) / t / 1e12 / world,
                        "frac_of_fp64_peak": steps * flops_per_step(dm) / t / 1e12 / world / peak_tflops,
                        "accept_rate": float(np.asarray(res["chain_state"])[:, 2].mean()) / (nits - 1),
                        "chains_rerun_on_bdf": int(getattr(m, "_last_rerun", 0)),
                        "api": "ModelFramework._run_chains (the body of ModelFramework.MCMC): solver='auto', Philox streams",
                        "kernel": "odl_mcmc_coop_kernel (%d lanes per system)" % dm.coop_lanes if coop
                                  else "odl_mcmc_kernel (thread per system, prefetching MH)"}
    out["c3_nclass_chains"] = c3
    if "c4" not in legs and "c5" not in legs:
        return out
    # ---- config 4 ----
    note("config 4")
    m, center = workloads.stiff(device=local)
    dm = m._device()
    n4 = 65536
    theta = torch.from_numpy(workloads.stiff_thetas(n4, seed=rank)).to(dev)
    for _ in range(2):
        r4 = dm.sweep(theta, solver="auto", max_steps=2000000)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
    barrier()
    for a, b in ev:
        a.record(); r4 = dm.sweep(theta, solver="auto", max_steps=2000000); b.record()
    barrier()
    t4 = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) * 1e-3 / 3)
    starts = theta[:1024].contiguous()
    t4c, res4 = run_chains(dm, starts, 2000, solver="bdf", max_steps=2000000, chain_offset=rank * 1024)
    out["c4_stiff"] = {"sweep": {"sets_per_gpu": n4, "seconds": t4, "solves_per_s": n4 * world / t4,
                                 "ok_fraction": float((r4["status"] == 0).double().mean().item()),
                                 "mean_steps": float(r4["nsteps"].double().mean().item()),
                                 "path": "odl_sweep(auto): capped DOPRI5 pass, every row finished by the variable-order BDF pass"},
                       "chains": {"chains_per_gpu": 1024, "iterations": 2000, "seconds": t4c,
                                  "chain_steps_per_s": 1024 * world * 1999 / t4c, "kernel": "odl_mcmc_bdf_kernel",
                                  "accept_rate": float(res4["chain_state"][:, 2].mean().item()) / 1999}}
    # ---- config 5: strong scaling ----
    note("config 5")
    m, center = workloads.network(device=local)
    dm = m._device()
    total = 65536
    C5 = total // world
    nits5 = {1: 100, 2: 200, 4: 400}.get(world, 1000)
    rng = np.random.default_rng([5, rank])
    starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C5, dm.n_param)))).to(dev)
    t5, res5 = run_chains(dm, starts, nits5, chain_offset=rank * C5, max_steps=2048)      # bounded solves (see config 3)
    steps5 = sum_over_ranks(float(res5["step_count"].sum().item()))
    dm.comm_init()
    dm.rhat(res5["summaries"])                                   # first call: NCCL sets its channels up
    barrier(); t0 = time.perf_counter()
    rh, _, seen = dm.rhat(res5["summaries"])
    torch.cuda.synchronize(); t_rh = time.perf_counter() - t0
    out["c5_network_5x5"] = {"states": dm.n_state, "parameters": dm.n_param, "chains_total": total, "chains_per_gpu": C5,
                             "iterations": nits5, "scaling": "strong", "seconds": t5, "chain_steps_per_s": total * (nits5 - 1) / t5,
                             "fp64_tflops_per_gpu": steps5 * flops_per_step(dm) / t5 / 1e12 / world,
                             "frac_of_fp64_peak": steps5 * flops_per_step(dm) / t5 / 1e12 / world / peak_tflops,
                             "rhat": {"seconds": t_rh, "chains_gathered": int(seen), "bytes_gathered": int(seen) * (1 + 2 * dm.n_param) * 8,
                                      "max": float(np.nanmax(rh)),
                                      "collective": "odl_rhat: ncclAllGather over %d ranks + device reduction" % world if world > 1
                                                    else "odl_rhat: device reduction (1 GPU, no collective)"},
                             "kernel": "odl_mcmc_coop_kernel (%d lanes per system)" % dm.coop_lanes,
                             "proposals_over_the_step_budget": int(sum_over_ranks(float(res5["fail_count"].sum().item()))),
                             "note": "iterations shortened below 8 GPUs (100 / 200 / 400 / 1000 at 1 / 2 / 4 / 8) so that the leg stays "
                                     "within seconds; the rate is per chain-step"}
    return out


def cold_start_leg(local, reference_single_chain_s):
    """Seconds from `ModelFramework(...)` of a model the library has never seen (empty cubin cache: NVRTC runs) to the first
    results, and the same with the cache warm (a new process on a machine that has run the model before)."""
    import shutil
    import tempfile
    import scipy.stats
    import odelib_b200 as ODElib
    from odelib_b200 import demo_models, workloads
    from odelib_b200.Statistics import Samplers
    out = {}
    tmp = tempfile.mkdtemp(prefix="odl_cold_")
    try:
        for label in ("cold", "warm_cache"):
            pri = demo_models.PRIORS[MODEL]
            t0 = time.perf_counter()
            pobj = {p: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=c)
                    for (p, (s, sc)), c in zip(pri.items(), CENTER[MODEL])}
            m = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES[MODEL],
                                      state_names=demo_models.STATE_NAMES[MODEL], dataframe=demo_frame(),
                                      state_summations={"H": ["S", "I1", "I2"]}, S=5236900, t_steps=TSTEPS, device=local,
                                      cache_dir=tmp, **pobj)
            m.sweep(prior_draws(4096, 7))
            t_sweep = time.perf_counter() - t0
            frame = Samplers.MetropolisHastings(m, nits=1000, print_progress=False)
            t_chain = time.perf_counter() - t0
            out[label] = {"two_i_first_sweep_s": t_sweep, "two_i_sweep_then_single_chain_1000_s": t_chain, "kept_rows": int(len(frame))}
            t0 = time.perf_counter()
            m2 = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES[MODEL],
                                       state_names=demo_models.STATE_NAMES[MODEL], dataframe=demo_frame(),
                                       state_summations={"H": ["S", "I1", "I2"]}, S=5236900, t_steps=TSTEPS, device=local,
                                       cache_dir=(tmp + "/c1_" + label if label == "cold" else tmp + "/c1_cold"), **pobj)
            os.makedirs(m2._cache_dir, exist_ok=True)
            Samplers.MetropolisHastings(m2, nits=1000, print_progress=False)
            out[label]["config1_single_chain_1000_s"] = time.perf_counter() - t0
        rhs, n, P, groups = workloads.network(spec_only=True)
        from odelib_b200 import engine
        for label in ("cold", "warm_cache"):
            t0 = time.perf_counter()
            dmn = engine.DeviceModel(rhs, n, P, groups, device=local, cache_dir=tmp)
            dmn.kernel_info("mcmc_coop"); dmn.kernel_info("sweep_coop")          # waits for the units a first call needs
            out[label]["network_5x5_kernels_ready_s"] = time.perf_counter() - t0
            dmn.close()
        out["reference_config1_single_chain_1000_s"] = reference_single_chain_s
        out["note"] = ("cold = empty cubin cache (NVRTC compiles the kernels a call needs, side by side on host threads); warm_cache = "
                       "what every later process pays.  config1 = Samplers.MetropolisHastings(model, nits=1000) from a fresh "
                       "ModelFramework (BASELINE config 1), the reference's time for the same call beside it")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out

# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import odelib_b200 as ODElib
    from odelib_b200 import _capi, demo_models, engine
    import scipy.stats

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=5))   # a rank that is missing fails the run soon

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- the model, through the reference-facing surface -------------------------------------------
    pri = demo_models.PRIORS[MODEL]
    pobj = {p: ODElib.parameter(stats_gen=scipy.stats.lognorm, hyperparameters={"s": s, "scale": sc}, init_value=sc)
            for p, (s, sc) in pri.items()}
    model = ODElib.ModelFramework(ODE=demo_models.two_i, parameter_names=demo_models.PARAMETER_NAMES[MODEL],
                                  state_names=demo_models.STATE_NAMES[MODEL], dataframe=demo_frame(),
                                  state_summations={"H": ["S", "I1", "I2"]}, S=5236900, t_steps=TSTEPS, device=local,
                                  **pobj)
    dm = model._device()
    n, P, ns = args.sets, dm.n_param, dm.n_state
    theta_host = torch.from_numpy(prior_draws(n, 0, offset=rank)).pin_memory()      # weak scaling: own block per rank
    theta_dev = theta_host.to(dev)
    out = {"chi": torch.empty(n, dtype=torch.float64, device=dev), "r2": torch.empty(n, dtype=torch.float64, device=dev),
           "status": torch.empty(n, dtype=torch.int32, device=dev), "nsteps": torch.empty(n, dtype=torch.int32, device=dev)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    launches0 = _capi.lib().odl_launch_count()

    peak_tflops, _ = engine.fp64_peak(local)

    # ---- device-resident sweep: `value` -------------------------------------------------------------
    SW = dict(solver="auto", max_steps=500000)      # every system is solved: ordering + DOPRI5 bulk pass + BDF pass
    for _ in range(max(args.warmup, 3)):
        dm.sweep(theta_dev, out=out, **SW)
    clocks = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    clocks.start()
    l0 = _capi.lib().odl_launch_count()
    for a, b in ev:
        flush.fill_(1)                      # L2 flush between timed steps (outside the event pair)
        a.record()
        dm.sweep(theta_dev, out=out, **SW)
        b.record()
    barrier()
    timed_launches = _capi.lib().odl_launch_count() - l0
    pass_ms = dm.last_pass_ms()
    ms = [a.elapsed_time(b) for a, b in ev]
    t_total = max_over_ranks(sum(ms) * 1e-3)
    clk = clocks.stop()
    value = n * world * args.steps / t_total
    nsteps = out["nsteps"].to(torch.int64)
    status = out["status"]
    flops_step = 6 * dm.rhs_flops + 71 * ns + 10                     # SURVEY.md §8d flop model (DOPRI5)
    flops_bdf_step = dm.rhs_flops + 2 * ns * ns + 37 * ns            # one Newton iteration; LU / change_D not counted
    flops_solve = dm.n_slot * (30 + 12 * ns) + dm.n_obs * 8
    avg_ms = float(np.mean(ms))
    bytes_launch = n * (P * 8 + 8 + 8 + 4 + 4)
    ok_frac = float((status == 0).float().mean().item())
    mean_steps = float(nsteps.double().mean().item())
    # the bulk kernel on its own: the DOPRI5 pass of the sweep (<= 704 attempted steps, projection check at 384) in
    # input order.  A solve depends on nothing but its own row, so this is also the exact set the sweep's DOPRI5 pass
    # finishes; the rest was finished by the BDF pass.
    bulk_out = {k: torch.empty_like(v) for k, v in out.items()}
    for _ in range(2):
        dm.sweep(theta_dev, out=bulk_out, solver="dopri5", max_steps=engine.AUTO_CAP, stiff_check=True,
                 early_check_steps=engine.AUTO_EARLY_CHECK)
    torch.cuda.synchronize()
    bulk_ms = dm.last_kernel_ms()
    bulk_ok = bulk_out["status"] == 0
    bulk_flops = float(bulk_out["nsteps"].to(torch.int64)[bulk_ok].sum().item()) * flops_step + \
        float(bulk_ok.sum().item()) * flops_solve
    # algorithmic flops of the whole sweep: DOPRI5-finished systems at the DOPRI5 rate, BDF-finished ones at the
    # (lower) BDF rate; the DOPRI5 attempts the deferred systems burned before leaving are not counted
    stiff_steps = float(nsteps[~bulk_ok].sum().item())
    # the rows the stiff pass finishes CONTINUE from where the DOPRI5 pass stopped: the DOPRI5 attempts they took up to
    # there are part of their solution (before the hand-over existed they were thrown away, and were not counted)
    handed_dopri_steps = float(bulk_out["nsteps"].to(torch.int64)[~bulk_ok].sum().item())
    flops_launch = bulk_flops + handed_dopri_steps * flops_step + stiff_steps * flops_bdf_step + \
        float((~bulk_ok).sum().item()) * flops_solve
    achieved = flops_launch / (avg_ms * 1e-3) / 1e12
    # the same launch as it runs inside the sweep -- rows in cost order -- on all SMs: the sweep with the stiff pass AFTER the
    # bulk pass (ODL_AUTO_SEQUENTIAL; same kernels, same results), middle entry of its pass times
    for _ in range(2):
        dm.sweep(theta_dev, out=bulk_out, auto_flags=_capi.AUTO_SEQUENTIAL, **SW)
    torch.cuda.synchronize()
    seq_pass = dm.last_pass_ms()
    bulk_ordered = {"kernel": "odl_sweep_kernel inside the sweep run with the stiff pass AFTER it (rows in cost order, all SMs)",
                    "ms": seq_pass[1], "achieved_TFLOPs": bulk_flops / (seq_pass[1] * 1e-3) / 1e12,
                    "frac_of_fp64_peak": bulk_flops / (seq_pass[1] * 1e-3) / 1e12 / peak_tflops,
                    "sequential_sweep_passes_ms": {"ordering": seq_pass[0], "dopri5_bulk": seq_pass[1], "bdf_stiff": seq_pass[2]}}
    bulk = {"kernel": "odl_sweep_kernel (DOPRI5 pass: <= 704 attempted steps, projection check at 384), input order",
            "ms": bulk_ms, "finished_fraction": float(bulk_ok.float().mean().item()),
            "achieved_TFLOPs": bulk_flops / (bulk_ms * 1e-3) / 1e12,
            "frac_of_fp64_peak": bulk_flops / (bulk_ms * 1e-3) / 1e12 / peak_tflops, "in_cost_order": bulk_ordered}

    # ---- end to end through the facade with pinned host buffers: `e2e` ---------------------------------
    th_np = theta_host.numpy()
    host_out = {k: torch.empty(n, dtype=(torch.float64 if k in ("chi", "r2") else torch.int32)).pin_memory().numpy()
                for k in ("chi", "r2", "status", "nsteps")}
    def e2e_leg(outputs):
        sub = {k: host_out[k] for k in outputs}
        for _ in range(2):
            model.sweep(th_np, out=sub, outputs=outputs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            model.sweep(th_np, out=sub, outputs=outputs)      # H2D of theta, 5 kernels, D2H of the outputs, sync
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    # what the reference's batch seam returns is chi per row (_Fit_worker, Framework.py:41-48); the status word rides
    # along (12 bytes per row).  The same call with every diagnostic output (R^2, step counts: 24 bytes per row) is
    # timed next to it.
    e2e_all_t = e2e_leg(("chi", "r2", "status", "nsteps"))
    assert np.array_equal(host_out["nsteps"], out["nsteps"].cpu().numpy())
    host_out["chi"][:] = 0.0
    e2e_t = e2e_leg(("chi", "status"))
    chi_dev = out["chi"].cpu().numpy()
    assert np.array_equal(host_out["chi"], chi_dev, equal_nan=True) and np.array_equal(host_out["status"], out["status"].cpu().numpy())
    e2e = {"value": n * world * args.steps / e2e_t, "unit": "solves/s", "h2d_bytes_per_step": n * P * 8,
           "d2h_bytes_per_step": n * 12,
           "api": "ModelFramework.sweep(numpy pinned, outputs=('chi','status')) -> odl_sweep(ODL_MEM_HOST)",
           "all_outputs": {"value": n * world * args.steps / e2e_all_t, "d2h_bytes_per_step": n * 24,
                           "api": "the same call returning chi, R^2, status and step counts"}}

    # ---- MCMC leg: chain-steps/s ---------------------------------------------------------------------
    mcmc = None
    if args.chains > 0:
        C, nits = args.chains, args.nits
        rng = np.random.default_rng([1, rank])
        starts = torch.from_numpy(np.array(CENTER[MODEL]) * np.exp(0.05 * rng.standard_normal((C, P)))).to(dev)
        kw = dict(nits=nits, rng_mode="philox", seed=0, chain_offset=rank * C, pnum=P, device_buffers=True)
        dm.mcmc(starts, **kw)                                   # warm-up at full size: buffers come from torch's cache
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        res = dm.mcmc(starts, **kw)
        b.record()
        barrier()
        t_m = max_over_ranks(a.elapsed_time(b) * 1e-3)
        # the one collective of the path: odl_rhat = ncclAllGather of the chain summaries over NVLink + reduction on the
        # device, behind the C ABI (torch.distributed only hands the communicator's id to the ranks)
        dm.comm_init()
        torch.cuda.synchronize(); t_r0 = time.perf_counter()
        rh, _, chains_seen = dm.rhat(res["summaries"])
        t_rhat = time.perf_counter() - t_r0
        assert chains_seen == C * world
        steps_total = sum_over_ranks(float(res["step_count"].sum().item()))
        kept_bytes = C * res["n_keep"] * (P + 5) * 8
        # a chain is the same chain wherever and beside whatever it runs (Philox key = global chain index): every rank
        # runs its LAST chain again alone (another launch shape, another prefetching width) and compares bit for bit;
        # global chain 0 has the same start on every world size, so its hash must be the same in every line of a
        # 1/2/4/8-GPU scaling run
        import hashlib
        alone = dm.mcmc(starts[C - 1:C].contiguous(), **dict(kw, chain_offset=rank * C + C - 1))
        same_alone = bool(torch.equal(alone["samples"][0].contiguous().view(torch.int64),
                                      res["samples"][C - 1].contiguous().view(torch.int64)))   # bits (NaN-safe)
        assert sum_over_ranks(0.0 if same_alone else 1.0) == 0.0, "a chain re-run alone differs from the same chain in the sharded launch"
        chain0_sha1 = hashlib.sha1(res["samples"][0].contiguous().cpu().numpy().tobytes()).hexdigest()
        mcmc = {"chain_steps_per_s": C * world * (nits - 1) / t_m, "chains_per_gpu": C, "chains_total": C * world,
                "iterations": nits, "seconds": t_m, "solves_per_s": C * world * nits / t_m,
                "fp64_tflops": (steps_total * flops_step) / t_m / 1e12 / world,
                "sample_stream_GBps": kept_bytes / t_m / 1e9,
                "accept_rate": float(res["chain_state"][:, 2].mean().item()) / (nits - 1),
                "rhat_max": float(np.nanmax(rh)), "rhat_seconds": t_rhat,
                "rhat_note": "500 iterations from scattered starts do not converge (R-hat >> 1): this leg times the kernel",
                "rhat_collective": "odl_rhat: ncclAllGather over %d ranks + device reduction" % world if world > 1
                                   else "odl_rhat: device reduction (1 GPU, no collective)",
                "identity": {"last_chain_of_every_rank_rerun_alone_bit_identical": same_alone,
                             "global_chain0_samples_sha1": chain0_sha1,
                             "note": "the sha1 must be the same on every world size (same start, Philox key = global chain index)"}}

        # the same kernel with the GPU filled (BASELINE config 5's chain count per GPU x 8): throughput regime
        if args.chains_large > 0:
            CL, nl = args.chains_large, args.nits_large
            st2 = torch.from_numpy(np.array(CENTER[MODEL]) * np.exp(0.05 * rng.standard_normal((CL, P)))).to(dev)
            kw2 = dict(nits=nl, rng_mode="philox", seed=0, chain_offset=rank * CL, pnum=P, device_buffers=True, keep_samples=False)
            dm.mcmc(st2, **kw2)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record()
            r2 = dm.mcmc(st2, **kw2)
            b.record()
            barrier()
            t_l = max_over_ranks(a.elapsed_time(b) * 1e-3)
            steps_l = sum_over_ranks(float(r2["step_count"].sum().item()))
            mcmc["filled_gpu"] = {"chains_per_gpu": CL, "iterations": nl, "seconds": t_l,
                                  "chain_steps_per_s": CL * world * (nl - 1) / t_l,
                                  "solves_per_s": CL * world * nl / t_l,
                                  "fp64_tflops_per_gpu": steps_l * flops_step / t_l / 1e12 / world,
                                  "frac_of_fp64_peak": steps_l * flops_step / t_l / 1e12 / world / peak_tflops}

    # ---- the callers either side of the path, through the reference's own API (rank 0 only; not part of `value`) ----
    facade = None
    if rank == 0 and not args.no_facade:
        np.random.seed(0)
        model.fit_survey(samples=n)                               # warm-up at full size (buffers, page faults of the pools)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sv = model.fit_survey(samples=n)                          # device LHS + sweep + frame
        torch.cuda.synchronize(); t_sv = time.perf_counter() - t0
        mc = dict(chain_inits=4096, iterations_per_chain=200, fitsurvey_samples=n, sd_fitdistance=6.0, print_report=False,
                  posterior="summary")
        model.MCMC(**mc)                                          # warm-up
        t0 = time.perf_counter()
        summ = model.MCMC(**mc)
        t_mc = time.perf_counter() - t0
        facade = {"fit_survey": {"samples": n, "seconds": t_sv, "rows_below_chi_666": int((sv["chi"] < 666).sum()),
                                 "api": "ModelFramework.fit_survey(samples) -> frame [samples, P+1] (device LHS, odl_sweep)"},
                  "mcmc_from_survey": {"chains": 4096, "iterations_per_chain": 200, "fitsurvey_samples": n, "seconds": t_mc,
                                       "chains_rerun_on_bdf": int(getattr(model, "_last_rerun", 0)),
                                       "best_chi": summ.best_chi,
                                       "api": "ModelFramework.MCMC(chain_inits=4096, posterior='summary'): survey, start "
                                              "selection, chains, report reductions on the device"}}

    configs = None
    if not args.no_configs:
        configs = config_legs(world, rank, local, dev, peak_tflops, barrier, max_over_ranks, sum_over_ranks)
    launches = _capi.lib().odl_launch_count() - launches0

    # ---- CPU baseline (rank 0, N=1 only): the reference itself on the host cores, the bare port beside it ----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ncpu = min(cores * 4000, n)
        rate, dt = cpu_sweep_rate(theta_host.numpy()[:ncpu], cores)
        port = {"value": rate, "unit": "solves/s", "cores": cores, "kind": "port",
                "sample": f"first {ncpu} parameter sets of the same sweep, scipy odeint (LSODA, default tol) + numpy "
                          f"masked chi in {cores} forked processes, {dt:.1f} s"}
        nref = cores * 1500
        ref = cpu_reference_sweep(nref, cores, repeats=2, warmup=1)
        if ref is not None:
            cpu = {"value": ref[0], "unit": "solves/s", "cores": cores, "kind": "reference",
                   "sample": f"ODElib.ModelFramework.fit_survey(samples={nref}, cpu_cores={cores}) of the unmodified reference "
                             f"(oracle/_ref, stock code path: LHS of the two_i priors, Pool fan-out of _Fit_worker, concat), "
                             f"{ref[1]:.1f} s per call", "port": port}
        else:
            cpu = port
        if mcmc is not None:
            mcmc["cpu_reference"] = cpu_reference_chains(cores)
    cold = None
    if rank == 0 and world == 1 and not args.no_cold:
        ref1 = (mcmc or {}).get("cpu_reference") or {}
        cold = cold_start_leg(local, (ref1.get("single_chain") or {}).get("seconds"))

    traffic = ncu_traffic()
    if rank == 0:
        line = {
            "metric": "ode_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops,
                         "traffic": traffic.get("bytes") if n == (1 << 20) else None,
                         "traffic_note": traffic.get("note"),
                         "kernel": "odl_sweep (9 launches: 3 ordering kernels, odl_sweep_bdf_kernel (consumer, beside), odl_gate_kernel, "
                                   "odl_sweep_kernel, odl_feed_done_kernel, odl_sweep_bdf_kernel (second consumer on the SMs the bulk "
                                   "pass frees), odl_sweep_bdf_kernel (pick-up of what the consumers left: normally nothing))",
                         "avg_launch_ms": avg_ms,
                         "peak_source": "measured live: odl_fp64_peak DFMA chains (MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_launch": flops_launch, "flops_per_step_attempt": flops_step,
                         "mean_steps_per_solve": mean_steps,
                         "flop_model": "per attempted DOPRI5 step 6*F_rhs+71n+10 (F_rhs=11, n=4 -> 360), + 19*(30+12n) + 37*8 per "
                                       "solve; the systems the BDF pass finishes: their DOPRI5 attempts up to the hand-over at 360, their BDF steps at "
                                       "F_rhs+2n^2+37n (= 191) per attempt; ordering, LU and change_D work are not counted",
                         "passes_ms": {"ordering": pass_ms[0], "dopri5_bulk": pass_ms[1], "bdf_stiff_after_bulk_ended": pass_ms[2]},
                         "bulk_kernel": bulk,
                         "hbm": {"algorithmic_bytes_per_launch": bytes_launch,
                                 "achieved_GBps": bytes_launch / (avg_ms * 1e-3) / 1e9, "peak_GBps": hbm_peak(),
                                 "frac": bytes_launch / (avg_ms * 1e-3) / 1e9 / hbm_peak()}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(timed_launches), "gpu_launches_total": int(launches),
            "clocks": clk, "ok_fraction": ok_frac, "mcmc": mcmc, "facade": facade, "configs": configs, "cold_start_s": cold,
        }
        emit_record(line)
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes of the dominant launch from the committed ncu capture (profiles/traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:  # noqa: BLE001
        return {}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6650.0   # fallback stated in B200_PROFILING.md


_RECORD_FD = None


def emit_record(obj):
    """The one JSON line of the run, to the process's ORIGINAL stdout (see main)."""
    data = (json.dumps(obj) + "\n").encode()
    if _RECORD_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RECORD_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets", type=int, default=1 << 20, help="parameter sets per GPU per step")
    ap.add_argument("--chains", type=int, default=4096, help="MCMC chains per GPU (0 = skip the MCMC leg)")
    ap.add_argument("--nits", type=int, default=500, help="iterations per chain in the MCMC leg")
    ap.add_argument("--chains-large", type=int, default=65536, help="chains per GPU for the filled-GPU MCMC measurement (0 = skip)")
    ap.add_argument("--nits-large", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-facade", action="store_true", help="skip the fit_survey / MCMC-from-survey timings")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs for BASELINE configs 3-5")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold-start leg")
    args = ap.parse_args()
    # stdout carries ONE line, the JSON record: whatever libraries print there meanwhile (NCCL's version banner under
    # NCCL_DEBUG, the reference's own per-iteration prints) goes to stderr -- at the descriptor level, so that C code is
    # covered too; the record itself is written to the saved descriptor
    global _RECORD_FD
    sys.stdout.flush()
    _RECORD_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
