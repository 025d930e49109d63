"""Dev tool (GPU): small ODL_SOLVER_AUTO sweeps under every flag combination, with the device counter block."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import device_model, prior_draws
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dm, tab = device_model("two_i")
theta = torch.from_numpy(prior_draws("two_i", n, seed=0)).cuda()
ref = dm.sweep(theta, solver="dopri5", max_steps=512)
torch.cuda.synchronize()
print("dopri5 ok", int((ref["status"] != 0).sum()), "unfinished", flush=True)
for flags in (1, 0, 3, 2):
    t0 = time.time()
    out = dm.sweep(theta, solver="auto", max_steps=200000, auto_flags=flags)
    torch.cuda.synchronize()
    c = np.zeros(96, np.int32)
    dm._L.odl_debug_counters(dm._h, c.ctypes.data, 96)
    print("flags", flags, "sec %.3f" % (time.time() - t0), "work", c[0], "feed", c[16], "ticket", c[32], "entered", c[48], "left", c[64],
          "watchdog", c[80], "status!=0", int((out["status"] != 0).sum()), "pass_ms", dm.last_pass_ms(), flush=True)
