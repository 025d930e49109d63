"""Dev tool (GPU): ncu-friendly run of the stiff tail pass alone -- the systems the capped DOPRI5 pass of a
two_i prior sweep did not finish, integrated by `solver` (bdf | radau5)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
solver = sys.argv[2] if len(sys.argv) > 2 else "bdf"
dm, tab = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).cuda()
plain = dm.sweep(theta, solver="dopri5", stiff_check=True, max_steps=512)
hard = theta[plain["status"] != 0].contiguous()
for _ in range(3):
    r = dm.sweep(hard, solver=solver, max_steps=200000)
torch.cuda.synchronize()
print(solver, "n", hard.shape[0], "kernel_ms", dm.last_kernel_ms(), "max steps", int(r["nsteps"].max()), dm.kernel_info("sweep_" + solver))
