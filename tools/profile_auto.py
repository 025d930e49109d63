"""Dev tool (GPU): ncu-friendly runs.  mode auto: three ODL_SOLVER_AUTO sweeps of 1M two_i prior draws (ordering +
odl_sweep_kernel + odl_sweep_bdf_kernel per call); mode mcmc: 4096 chains x 300 iterations (prefetching MH);
mode coop: the 5x5 network, 8192 chains x 20 iterations on the cooperative kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
if mode == "auto":
    dm, tab = device_model("two_i")
    theta = torch.from_numpy(prior_draws("two_i", 1 << 20, seed=0)).cuda()
    flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0       # 8 = stiff pass after the bulk pass (ncu serialises kernels)
    for _ in range(3):
        out = dm.sweep(theta, solver="auto", max_steps=500000, auto_flags=flags)
    torch.cuda.synchronize()
    print("auto kernel_ms", dm.last_kernel_ms(), dm.last_pass_ms(), dm.kernel_info("sweep"), dm.kernel_info("sweep_bdf"))
elif mode in ("stream_iteration", "stream_chain"):
    # the sample stream of a filled GPU: 65,536 chains, one lane per chain, every kept row written
    dm, tab = device_model("two_i")
    C = 65536
    starts = torch.from_numpy(np.array(bench.CENTER["two_i"]) * np.exp(0.05 * np.random.default_rng(1).standard_normal((C, 5)))).cuda()
    for _ in range(2):
        res = dm.mcmc(starts, nits=41, seed=0, device_buffers=True, sample_layout=mode.split("_")[1])
    torch.cuda.synchronize()
    print(mode, "kernel_ms", dm.last_kernel_ms(), "kept bytes", C * 20 * 80, dm.kernel_info("mcmc"))
elif mode == "mcmc":
    dm, tab = device_model("two_i")
    C = 4096
    starts = torch.from_numpy(np.array(bench.CENTER["two_i"]) * np.exp(0.05 * np.random.default_rng(1).standard_normal((C, 5)))).cuda()
    for _ in range(2):
        res = dm.mcmc(starts, nits=300, seed=0, device_buffers=True)
    torch.cuda.synchronize()
    print("mcmc kernel_ms", dm.last_kernel_ms(), dm.kernel_info("mcmc"))
else:
    sys.argv = sys.argv[:1]
    from tools.coop_perf import network  # noqa: E402  (prints its own table first)
    dm, center, P = network()
    starts = torch.from_numpy(center * np.exp(0.02 * np.random.default_rng(2).standard_normal((8192, P)))).cuda()
    for _ in range(2):
        res = dm.mcmc(starts, nits=20, rng_mode="philox", seed=1, device_buffers=True, keep_samples=False)
    torch.cuda.synchronize()
    print("coop mcmc kernel_ms", dm.last_kernel_ms(), dm.kernel_info("mcmc_coop"))
