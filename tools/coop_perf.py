"""Dev tool (GPU): 5x5 network (35 states) and N-class chains: cooperative mapping vs thread-per-system."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import demo_models
from tests.helpers import synthetic_problem
from tests.test_gpu_models import nclass_problem

def timed(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r

LANES = int(sys.argv[1]) if len(sys.argv) > 1 else 0


def network():
    rhs, n, P, groups = demo_models.network(5, 5)
    H, V = 5, 5
    names = [f"S{i}" for i in range(H)] + [f"I{i}{j}" for i in range(H) for j in range(V)] + [f"V{j}" for j in range(V)]
    sums = {f"H{i}": [f"S{i}"] + [f"I{i}{j}" for j in range(V)] for i in range(H)}
    rng = np.random.default_rng(1)
    center = np.concatenate([0.3 * np.exp(0.2 * rng.standard_normal(H)), 2e-8 * np.exp(0.5 * rng.standard_normal(H * V)),
                             20 * np.exp(0.1 * rng.standard_normal(V)), 2.0 * np.exp(0.2 * rng.standard_normal(V))])
    y0 = [1e6 * (1 + i) for i in range(H)] + [0.0] * (H * V) + [2e6 * (1 + j) for j in range(V)]
    dm, tab = synthetic_problem(rhs, names, sums, center, y0, [f"H{i}" for i in range(H)] + [f"V{j}" for j in range(V)], seed=1,
                                device_kw=dict(coop_lanes=LANES))
    return dm, center, P

if __name__ == '__main__':
    for label, (dm, center, P) in (("network_5x5", network()),
                                   ("n_class_10", (lambda q: (synthetic_problem(q[0], q[1], q[2], q[3], q[4], q[5], seed=10, device_kw=dict(coop_lanes=LANES))[0], q[3], 5))(nclass_problem(10)))):
        flops_step = 6 * dm.rhs_flops + 71 * dm.n_state + 10
        print(label, "kernel info coop", dm.kernel_info("sweep_coop"), dm.kernel_info("mcmc_coop"), "tps", dm.kernel_info("mcmc"), flush=True)
        rng = np.random.default_rng(2)
        for C, its in ((2048, 40), (4096, 40), (8192, 40)):
            starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((C, P)))).cuda()
            for spec in (0, -1, -2, -4):
                t, r = timed(lambda: dm.mcmc(starts, nits=its, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False, speculate=spec))
                steps = float(r["step_count"].sum().item())
                print(f"  mcmc chains {C:6d} speculate {spec} ({'coop' if spec <= 0 else 'thread-per-system'}): {C * (its - 1) / t / 1e3:9.1f} k chain-steps/s, "
                      f"{steps * flops_step / t / 1e12:6.3f} TFLOP/s (consumed)", flush=True)
        theta = torch.from_numpy(center * np.exp(0.05 * rng.standard_normal((65536, P)))).cuda()
        for kw, name in ((dict(), "coop dopri5"),):
            t, r = timed(lambda: dm.sweep(theta, max_steps=200000, **kw))
            steps = float(r["nsteps"].double().sum().item())
            print(f"  sweep 65536 {name:26s}: {65536 / t / 1e6:7.3f} M solves/s, {steps * flops_step / t / 1e12:6.3f} TFLOP/s, ok {float((r['status'] == 0).double().mean()):.4f}", flush=True)
