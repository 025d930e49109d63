"""Dev tool (GPU): a short, ncu-friendly run of the sweep kernel (and optionally the MCMC kernel)."""
import sys

import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
mode = sys.argv[3] if len(sys.argv) > 3 else "sweep"
dm, tab = device_model("two_i")
theta = torch.from_numpy(bench.prior_draws(n, 0, 0)).cuda()
if mode == "sweep":
    for _ in range(3):
        out = dm.sweep(theta, solver="dopri5", max_steps=cap)
    torch.cuda.synchronize()
    print("kernel_ms", dm.last_kernel_ms(), "mean steps", out["nsteps"].double().mean().item(), dm.kernel_info("sweep"))
else:
    import numpy as np
    starts = torch.from_numpy(np.array(bench.CENTER["two_i"]) * np.exp(0.05 * np.random.default_rng(1).standard_normal((n, 5)))).cuda()
    for _ in range(2):
        res = dm.mcmc(starts, nits=cap, seed=0, device_buffers=True)
    torch.cuda.synchronize()
    print("kernel_ms", dm.last_kernel_ms(), dm.kernel_info("mcmc"))
