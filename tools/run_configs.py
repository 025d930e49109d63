"""Measure BASELINE.json's five configurations on one GPU, with the oracle port timed on a
bounded sample of the same workload beside each.  Writes gpurun_out/configs.json (copied to profiles/).

    python tools/run_configs.py [--quick]
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from odelib_b200 import demo_models, engine  # noqa: E402
from oracle import odelib_oracle as orc  # noqa: E402
from tests.helpers import device_model, golden, oracle_rhs, prior_draws, synthetic_problem  # noqa: E402
from tests.test_gpu_models import nclass_problem  # noqa: E402
from tests.test_gpu_stiff import stiff_thetas  # noqa: E402

QUICK = "--quick" in sys.argv
out = {}
peak, _ = engine.fp64_peak(0)
out["fp64_peak_tflops_measured"] = peak


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, r


def cpu_rate(rhs, thetas, tab, y0=None):
    t0 = time.perf_counter()
    for th in thetas:
        orc.solve_unit(rhs, th, tab, y0=y0)
    return len(thetas) / (time.perf_counter() - t0)


def flops_per_step(dm):
    return 6 * dm.rhs_flops + 71 * dm.n_state + 10


# ---- C1: the demo fit, single chain of 1000 iterations per model + the notebook's 32 x 1000 -------------------
c1 = {}
for name in ("zero_i", "one_i", "two_i"):
    dm, tab = device_model(name)
    g = golden(name)
    th0 = g["chain_def_s0_theta0"]
    z, u = orc.reference_streams(0, dm.n_param, 999)
    t1, r1 = timed(lambda: dm.mcmc(th0[None], nits=1000, rng_mode="host", z=z[None], u=u[None], pnum=int(g["pnum"])))
    starts = np.tile(th0, (32, 1))
    t32, r32 = timed(lambda: dm.mcmc(starts, nits=1000, rng_mode="philox", seed=0))
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        orc.mh_chain(oracle_rhs(name), th0, tab, int(g["pnum"]), nits=120, z=z[:119], u=u[:119])
    cpu = 119 / (time.perf_counter() - t0)
    c1[name] = {"single_chain_steps_per_s": 999 / t1, "single_chain_seconds": t1, "chains32_steps_per_s": 32 * 999 / t32,
                "cpu_oracle_chain_steps_per_s_1core": cpu,
                "note": "oracle chain = bare integrate+chi per step; the reference spends a further ~60 % on unused prior pdfs"}
out["C1_demo_fit"] = c1

# ---- C2: forward sweep, 1M prior draws, all three models ---------------------------------------------------------
c2 = {}
n = (1 << 18) if QUICK else (1 << 20)
for name in ("zero_i", "one_i", "two_i"):
    dm, tab = device_model(name)
    theta = torch.from_numpy(prior_draws(name, n, seed=0)).cuda()
    t, r = timed(lambda: dm.sweep(theta, solver="auto"))
    ns = r["nsteps"].double()
    ok = float(((r["status"] & 7) == 0).double().mean().item())
    fl = float(ns.sum().item()) * flops_per_step(dm)
    c2[name] = {"sets": n, "seconds": t, "solves_per_s": n / t, "ok_fraction": ok, "mean_steps": float(ns.mean().item()),
                "pass_ms": dm.last_pass_ms(), "fp64_tflops": fl / t / 1e12, "frac_fp64_peak": fl / t / 1e12 / peak,
                "cpu_oracle_solves_per_s_1core": cpu_rate(oracle_rhs(name), theta[:300].cpu().numpy(), tab)}
out["C2_forward_sweep"] = c2

# ---- C3: N-class chain, 4096 chains ------------------------------------------------------------------------------
c3 = {}
nits = 300 if QUICK else 1000
for N in ((2, 6, 10) if QUICK else (1, 2, 4, 6, 8, 10)):
    if N == 1:
        dm, tab = device_model("one_i"); rhs = oracle_rhs("one_i")
        center = np.array([1.238e-08, 3.550e-08, 19.40, 1.835]); P = 4
    else:
        rhs, names, sums, center, y0, orgs = nclass_problem(N)
        dm, tab = synthetic_problem(rhs, names, sums, center, y0, orgs, seed=N); P = 5
    rng = np.random.default_rng(N)
    starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((4096, P)))).cuda()
    t, r = timed(lambda: dm.mcmc(starts, nits=nits, rng_mode="philox", seed=1, device_buffers=True), reps=2)
    steps = float(r["step_count"].sum().item())
    c3[f"N={N}"] = {"states": dm.n_state, "chains": 4096, "iterations": nits, "seconds": t,
                    "chain_steps_per_s": 4096 * (nits - 1) / t, "mean_integrator_steps_per_solve": steps / (4096 * nits),
                    "fp64_tflops": steps * flops_per_step(dm) / t / 1e12, "accept_rate": float(r["chain_state"][:, 2].mean().item()) / (nits - 1),
                    "kernel": dm.kernel_info("mcmc_coop" if dm.n_state > 8 else "mcmc"),
                    "mapping": "cooperative, lanes per system by state count" if dm.n_state > 8 else "thread per system, prefetching MH",
                    "cpu_oracle_solves_per_s_1core": cpu_rate(rhs, starts[:40].cpu().numpy(), tab)}
out["C3_nclass_chains"] = c3

# ---- C4: stiff variant -------------------------------------------------------------------------------------------
dm, tab = device_model("two_i")
ns4 = 16384 if QUICK else 65536
theta = torch.from_numpy(stiff_thetas(ns4, seed=0)).cuda()
c4 = {}
for solver in ("bdf", "radau5", "ros23", "auto"):
    t, r = timed(lambda: dm.sweep(theta, solver=solver, max_steps=2000000), reps=2)
    c4[f"sweep_{solver}"] = {"sets": ns4, "seconds": t, "solves_per_s": ns4 / t, "mean_steps": float(r["nsteps"].double().mean().item()),
                             "ok_fraction": float((r["status"] == 0).double().mean().item())}
starts = theta[:1024]
for solver in ("bdf", "radau5", "ros23"):
    its = 100 if QUICK else 400
    t, r = timed(lambda: dm.mcmc(starts, nits=its, solver=solver, seed=2, device_buffers=True, max_steps=2000000), reps=1)
    c4[f"mcmc_{solver}"] = {"chains": 1024, "iterations": its, "seconds": t, "chain_steps_per_s": 1024 * (its - 1) / t}
c4["cpu_oracle_solves_per_s_1core"] = cpu_rate(oracle_rhs("two_i"), theta[:100].cpu().numpy(), tab)
out["C4_stiff"] = c4

# ---- C5: 5 x 5 network, 35 states ----------------------------------------------------------------------------------
rhs, nst, P, groups = demo_models.network(5, 5)
H, V = 5, 5
names = [f"S{i}" for i in range(H)] + [f"I{i}{j}" for i in range(H) for j in range(V)] + [f"V{j}" for j in range(V)]
sums = {f"H{i}": [f"S{i}"] + [f"I{i}{j}" for j in range(V)] for i in range(H)}
rng = np.random.default_rng(1)
center = np.concatenate([0.3 * np.exp(0.2 * rng.standard_normal(H)), 2e-8 * np.exp(0.5 * rng.standard_normal(H * V)),
                         20 * np.exp(0.1 * rng.standard_normal(V)), 2.0 * np.exp(0.2 * rng.standard_normal(V))])
y0 = [1e6 * (1 + i) for i in range(H)] + [0.0] * (H * V) + [2e6 * (1 + j) for j in range(V)]
dm, tab = synthetic_problem(rhs, names, sums, center, y0, [f"H{i}" for i in range(H)] + [f"V{j}" for j in range(V)], seed=1)
C5 = 2048 if QUICK else 8192           # 65,536 chains over 8 GPUs = 8,192 per GPU
its = 60 if QUICK else 200
starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C5, P)))).cuda()
t, r = timed(lambda: dm.mcmc(starts, nits=its, rng_mode="philox", seed=1, device_buffers=True), reps=1)
steps = float(r["step_count"].sum().item())
out["C5_network_5x5"] = {"states": 35, "parameters": 40, "chains_per_gpu": C5, "iterations": its, "seconds": t,
                         "chain_steps_per_s": C5 * (its - 1) / t, "mean_integrator_steps_per_solve": steps / (C5 * its),
                         "fp64_tflops": steps * flops_per_step(dm) / t / 1e12, "kernel": dm.kernel_info("mcmc_coop"),
                         "cpu_oracle_solves_per_s_1core": cpu_rate(rhs, starts[:20].cpu().numpy(), tab),
                         "note": "cooperative kernel: 8 lanes per system, state slices in registers, RHS from a shared row (DESIGN.md 4)"}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
