"""Dev tool (GPU): host-memory sweep (H2D + kernels + D2H inside the call), theta uploaded in two pieces vs one."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws
n = 1 << 20
dm, tab = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).pin_memory().numpy()
out = {k: torch.empty(n, dtype=(torch.float64 if k in ("chi", "r2") else torch.int32)).pin_memory().numpy()
       for k in ("chi", "r2", "status", "nsteps")}
for flags in (0, 4, 0, 4):
    for _ in range(2):
        dm.sweep(theta, solver="auto", max_steps=500000, out=out, auto_flags=flags)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        dm.sweep(theta, solver="auto", max_steps=500000, out=out, auto_flags=flags)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("flags", flags, "(two pieces)" if flags == 0 else "(one piece)", "ms per call %.3f" % (dt * 1e3), "kernel_ms %.3f" % dm.last_kernel_ms(),
          [round(x, 3) for x in dm.last_pass_ms()], flush=True)
