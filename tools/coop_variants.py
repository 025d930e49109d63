"""Dev tool (GPU): the cooperative kernels of the 5x5 network and the 12-state chain under (lanes per system, sliced RHS)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from odelib_b200 import workloads
from odelib_b200.engine import DeviceModel

def timed(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return time.perf_counter() - t0, r

for label, maker in (("network_5x5", lambda: workloads.network(device=0)), ("n_class_10", lambda: workloads.nclass(10, device=0))):
    m, center = maker()
    base = m._device()
    P = base.n_param
    rng = np.random.default_rng(2)
    starts = torch.from_numpy(center * np.exp(0.02 * rng.standard_normal((8192, P)))).cuda()
    theta = torch.from_numpy(center * np.exp(0.05 * rng.standard_normal((65536, P)))).cuda()
    for lanes, sliced in ((0, False), (0, True), (16 if base.n_state > 16 else 8, True), (16 if base.n_state > 16 else 8, False)):
        dm = DeviceModel(m._device_ode(), base.n_state, P, m._observe_groups(), device=0, coop_lanes=lanes, sliced_rhs=sliced)
        dm.set_data(base.tables, np.asarray(m.get_inits(), float))
        flops_step = 6 * dm.rhs_flops + 71 * dm.n_state + 10
        t, r = timed(lambda: dm.mcmc(starts, nits=40, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False))
        steps = float(r["step_count"].sum().item())
        t2, r2 = timed(lambda: dm.sweep(theta, max_steps=200000))
        print(f"{label} lanes {lanes or 'default'} sliced {int(sliced)}: mcmc 8192 chains {8192 * 39 / t / 1e6:6.2f} M chain-steps/s "
              f"({steps * flops_step / t / 1e12:5.2f} TFLOP/s), sweep 65536 {65536 / t2 / 1e6:6.2f} M solves/s, regs {dm.kernel_info('mcmc_coop')['regs']}", flush=True)
        dm.close()
