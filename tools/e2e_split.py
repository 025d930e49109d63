"""Dev tool (GPU): host-memory AUTO sweep of 1M rows, share of the rows in the first of the two upload pieces."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws
n = 1 << 20
dm, tab = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).pin_memory().numpy()
outs = ("chi", "status")
out = {"chi": torch.empty(n, dtype=torch.float64).pin_memory().numpy(), "status": torch.empty(n, dtype=torch.int32).pin_memory().numpy()}
for first, flags in ((0.5, 0), (0.25, 0), (0.33, 0), (0.2, 0), (0.15, 0), (0.25, 8), (0.33, 8), (0.5, 8)):
    os.environ["ODL_FIRST_PIECE"] = str(first)
    for _ in range(2):
        dm.sweep(theta, solver="auto", max_steps=500000, out=out, auto_flags=flags, outputs=outs)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        dm.sweep(theta, solver="auto", max_steps=500000, out=out, auto_flags=flags, outputs=outs)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("first piece %.2f %s ms per call %.3f kernel_ms %.3f" % (first, "sequential" if flags else "beside", dt * 1e3, dm.last_kernel_ms()),
          [round(x, 3) for x in dm.last_pass_ms()], flush=True)
